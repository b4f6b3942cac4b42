// Earth-Mover distance of generate_obj_mesh_wnf (reference src/common.py:45-51):
//     d = scipy.spatial.distance.cdist(points1, points2)          # float64 Euclidean
//     emd = d[scipy.optimize.linear_sum_assignment(d)].sum() / len(d)
// i.e. the cost of a minimum-cost perfect matching between 2048 ground-truth points and 2048 mesh
// vertices.  The reference solves it on the CPU with a Hungarian / Jonker-Volgenant solver; here
// the matching is found on the GPU with Bertsekas' forward AUCTION algorithm (Jacobi version: all
// unassigned rows bid in parallel) with epsilon-scaling, in float64.  The optimal COST is unique
// (the matching need not be); an auction that ends with slack epsilon is within n*epsilon of it,
// so emd (= cost / n) is within eps_final (default 1e-9) of scipy's value.
// A rectangular problem (n1 != n2) is squared with zero-cost dummy rows / columns, which leaves the
// optimal cost of the real rows unchanged (linear_sum_assignment matches min(n1,n2) pairs).
#include "common.cuh"
#include <math_constants.h>

namespace vtaco {

struct EmdState {
  const double* cost;   // [n][n] row-major, padded
  double* price;        // [n]
  int* owner;           // [n] row owning column j, or -1
  int* assigned;        // [n] column of row i, or -1
  double* bid;          // [n] per row: the price it offers
  int* bid_col;         // [n]
  unsigned long long* best_bid;   // [n] per column: max offer (bit pattern of a non-negative double)
  int* winner;          // [n] per column
  int* unassigned;      // [1]
  int n;
};

__global__ void __launch_bounds__(256) emd_cost_kernel(const float* __restrict__ p1, int n1, const float* __restrict__ p2,
                                                       int n2, int n, double* __restrict__ cost) {
  const long long total = (long long)n * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t % n);
    double d = 0.0;                          // dummy row / column
    if (i < n1 && j < n2) {
      const double dx = (double)p1[3 * i] - (double)p2[3 * j], dy = (double)p1[3 * i + 1] - (double)p2[3 * j + 1],
                   dz = (double)p1[3 * i + 2] - (double)p2[3 * j + 2];
      d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    }
    cost[t] = d;
  }
}

__global__ void emd_reset_kernel(EmdState S, int reset_prices) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < S.n) {
    S.owner[j] = -1;
    S.assigned[j] = -1;
    S.best_bid[j] = 0ull;
    S.winner[j] = 0x7fffffff;
    if (reset_prices) S.price[j] = 0.0;
  }
  if (j == 0) *S.unassigned = S.n;
}

// one block per row: value of column j is -cost - price; bid = price[j1] + (v1 - v2) + eps
__global__ void __launch_bounds__(128) emd_bid_kernel(EmdState S, double eps) {
  const int i = blockIdx.x;
  if (S.assigned[i] >= 0) return;
  const double* row = S.cost + (size_t)i * S.n;
  double v1 = -CUDART_INF, v2 = -CUDART_INF;
  int j1 = -1;
  for (int j = threadIdx.x; j < S.n; j += blockDim.x) {
    const double v = -row[j] - S.price[j];
    if (v > v1) { v2 = v1; v1 = v; j1 = j; }
    else if (v > v2) v2 = v;
  }
  // block reduction of (v1, j1, v2): smaller column index wins ties (deterministic)
  __shared__ double s1[128], s2[128];
  __shared__ int sj[128];
  s1[threadIdx.x] = v1; s2[threadIdx.x] = v2; sj[threadIdx.x] = j1;
  __syncthreads();
  for (int d = 64; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) {
      const double a1 = s1[threadIdx.x], a2 = s2[threadIdx.x], b1 = s1[threadIdx.x + d], b2 = s2[threadIdx.x + d];
      const int aj = sj[threadIdx.x], bj = sj[threadIdx.x + d];
      const bool b_wins = (b1 > a1) || (b1 == a1 && bj >= 0 && (aj < 0 || bj < aj));
      const double top = b_wins ? b1 : a1, loser = b_wins ? a1 : b1;
      s1[threadIdx.x] = top;
      sj[threadIdx.x] = b_wins ? bj : aj;
      s2[threadIdx.x] = fmax(fmax(a2, b2), loser);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int j = sj[0];
    double second = s2[0];
    if (!(second > -CUDART_INF)) second = s1[0];     // n == 1
    const double b = S.price[j] + (s1[0] - second) + eps;
    S.bid[i] = b;
    S.bid_col[i] = j;
    atomicMax(S.best_bid + j, (unsigned long long)__double_as_longlong(b));   // b > 0: bit patterns order like values
  }
}

__global__ void emd_resolve_kernel(EmdState S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n || S.assigned[i] >= 0) return;
  const int j = S.bid_col[i];
  if ((unsigned long long)__double_as_longlong(S.bid[i]) == S.best_bid[j]) atomicMin(S.winner + j, i);
}

__global__ void emd_assign_kernel(EmdState S) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S.n) return;
  const int w = S.winner[j];
  if (w == 0x7fffffff) return;
  const int prev = S.owner[j];
  if (prev >= 0) S.assigned[prev] = -1;
  else atomicSub(S.unassigned, 1);
  S.owner[j] = w;
  S.assigned[w] = j;
  S.price[j] = S.bid[w];
  S.best_bid[j] = 0ull;
  S.winner[j] = 0x7fffffff;
}

__global__ void __launch_bounds__(256) emd_total_kernel(EmdState S, int n1, int n2, double* total, int32_t* assignment) {
  __shared__ double part[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < S.n; i += blockDim.x) {
    const int j = S.assigned[i];
    const bool real = i < n1 && j >= 0 && j < n2;
    if (real) acc += S.cost[(size_t)i * S.n + j];
    if (assignment && i < n1) assignment[i] = real ? j : -1;
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) part[threadIdx.x] += part[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = part[0];
}

static long long emd_align(long long v) { return (v + 255) / 256 * 256; }

}  // namespace vtaco

using namespace vtaco;

extern "C" int64_t vtaco_emd_workspace_bytes(int64_t n1, int64_t n2) {
  if (n1 < 1 || n2 < 1) return VTACO_ERR_INVALID_ARG;
  const long long n = n1 > n2 ? n1 : n2;
  if (n > 8192) return VTACO_ERR_UNSUPPORTED;
  return emd_align(8 * n * n) + 3 * emd_align(8 * n) + 4 * emd_align(4 * n) + 512;
}

extern "C" int vtaco_emd(const float* p1, int64_t n1, const float* p2, int64_t n2, void* workspace,
                         int64_t workspace_bytes, double eps_final, double* emd_host, int32_t* assignment,
                         int64_t* iterations_host, void* stream) {
  if (!p1 || !p2 || !workspace || !emd_host || n1 < 1 || n2 < 1) return VTACO_ERR_INVALID_ARG;
  const int64_t need = vtaco_emd_workspace_bytes(n1, n2);
  if (need < 0) return (int)need;
  if (workspace_bytes < need) return VTACO_ERR_CAPACITY;
  if (!(eps_final > 0.0)) eps_final = 1e-9;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)(n1 > n2 ? n1 : n2);
  char* w = reinterpret_cast<char*>(workspace);
  EmdState S;
  S.n = n;
  double* cost = reinterpret_cast<double*>(w); w += emd_align(8ll * n * n);
  S.cost = cost;
  S.price = reinterpret_cast<double*>(w); w += emd_align(8ll * n);
  S.bid = reinterpret_cast<double*>(w); w += emd_align(8ll * n);
  S.best_bid = reinterpret_cast<unsigned long long*>(w); w += emd_align(8ll * n);
  S.owner = reinterpret_cast<int*>(w); w += emd_align(4ll * n);
  S.assigned = reinterpret_cast<int*>(w); w += emd_align(4ll * n);
  S.bid_col = reinterpret_cast<int*>(w); w += emd_align(4ll * n);
  S.winner = reinterpret_cast<int*>(w); w += emd_align(4ll * n);
  S.unassigned = reinterpret_cast<int*>(w);
  double* total = reinterpret_cast<double*>(w + 256);

  emd_cost_kernel<<<num_sms() * 8, 256, 0, st>>>(p1, (int)n1, p2, (int)n2, n, cost);
  const int nb = (n + 255) / 256;
  // epsilon scaling: the costs are Euclidean distances of points in a box of a few units, so start
  // at a slack of 0.5 and divide by 5 per phase down to eps_final; prices carry over between phases.
  double eps = 0.5;
  long long iters = 0;
  bool first = true;
  for (;;) {
    if (eps < eps_final) eps = eps_final;
    emd_reset_kernel<<<nb, 256, 0, st>>>(S, first ? 1 : 0);
    first = false;
    int left = n;
    while (left > 0) {
      for (int k = 0; k < 16; ++k) {
        emd_bid_kernel<<<n, 128, 0, st>>>(S, eps);
        emd_resolve_kernel<<<nb, 256, 0, st>>>(S);
        emd_assign_kernel<<<nb, 256, 0, st>>>(S);
      }
      iters += 16;
      VTACO_CUDA_CHECK(cudaMemcpyAsync(&left, S.unassigned, sizeof(int), cudaMemcpyDeviceToHost, st));
      VTACO_CUDA_CHECK(cudaStreamSynchronize(st));
      if (iters > 4000000) return VTACO_ERR_UNSUPPORTED;   // no convergence (NaN input?)
    }
    if (eps <= eps_final) break;
    eps /= 5.0;
  }
  emd_total_kernel<<<1, 256, 0, st>>>(S, (int)n1, (int)n2, total, assignment);
  double tot = 0.0;
  VTACO_CUDA_CHECK(cudaMemcpyAsync(&tot, total, sizeof(double), cudaMemcpyDeviceToHost, st));
  VTACO_CUDA_CHECK(cudaStreamSynchronize(st));
  VTACO_LAUNCH_CHECK();
  *emd_host = tot / (double)n1;       // `/ len(d)`: the number of rows of cdist(points1, points2)
  if (iterations_host) *iterations_host = iters;
  return VTACO_OK;
}
