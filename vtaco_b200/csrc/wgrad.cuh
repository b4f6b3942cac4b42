// Weight / bias gradients of the per-query (decoder) and per-point (encoder) linear layers:
//   dW[o][k] += sum_q G[q][o] * A[q][k];   db[o] += sum_q G[q][o]
// G = gradients at the layer outputs, A = layer inputs, both stored as rows per query.
// One CTA per (product, query chunk): 128-row tiles staged in shared memory, the 32x32
// outer-product accumulators spread over 256 threads (4 per thread), one atomicAdd per
// element and CTA.  Matrices wider than 32 are split into several products by the caller
// (pointer offsets + leading dimensions).
#pragma once
#include "common.cuh"
#include <algorithm>

namespace vtaco {

constexpr int kMaxProd = 32;
constexpr int kWTile = 128;

struct WProd {
  const float* G;   // rows of g_ld floats, n_out (<= 32) used
  const float* A;   // rows of a_ld floats, n_in (<= 32) used
  float* dW;        // [n_out][w_ld], accumulated into
  float* db;        // [n_out] or NULL
  int g_ld, n_out, a_ld, n_in, w_ld;
  int a_relu;       // use max(A, 0) (layer input was relu(A))
};
struct WParams {
  WProd prod[kMaxProd];
  long long Q;
  int q_per_cta;
};

static __global__ void __launch_bounds__(256) linear_wgrad_kernel(const __grid_constant__ WParams P) {
  __shared__ __align__(16) float sG[kWTile][32];
  __shared__ __align__(16) float sA[kWTile][32];
  const WProd& pr = P.prod[blockIdx.y];
  const long long q_begin = (long long)blockIdx.x * P.q_per_cta;
  const long long q_end = min(P.Q, q_begin + P.q_per_cta);
  const int o = threadIdx.x >> 3, k4 = threadIdx.x & 7;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bsum = 0.f;
  for (long long q0 = q_begin; q0 < q_end; q0 += kWTile) {
    const int nq = (int)min((long long)kWTile, q_end - q0);
    __syncthreads();
#pragma unroll 4
    for (int idx = threadIdx.x; idx < kWTile * 32; idx += 256) {
      const int r = idx >> 5, c = idx & 31;
      float g = 0.f, a = 0.f;
      if (r < nq) {
        if (c < pr.n_out) g = __ldg(pr.G + (size_t)(q0 + r) * pr.g_ld + c);
        if (c < pr.n_in) a = __ldg(pr.A + (size_t)(q0 + r) * pr.a_ld + c);
        if (pr.a_relu) a = fmaxf(a, 0.f);
      }
      sG[r][c] = g;
      sA[r][c] = a;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kWTile; ++r) {   // rows beyond nq are zero
      const float g = sG[r][o];
      const float4 a = *reinterpret_cast<const float4*>(&sA[r][4 * k4]);
      acc[0] = fmaf(g, a.x, acc[0]);
      acc[1] = fmaf(g, a.y, acc[1]);
      acc[2] = fmaf(g, a.z, acc[2]);
      acc[3] = fmaf(g, a.w, acc[3]);
      bsum += g;
    }
  }
  if (o < pr.n_out) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * k4 + e;
      if (k < pr.n_in) atomicAdd(pr.dW + (size_t)o * pr.w_ld + k, acc[e]);
    }
    if (pr.db && k4 == 0) atomicAdd(pr.db + o, bsum);
  }
}

static inline void wgrad_add(WParams& W, int& np, const float* G, int g_ld, int n_out, const float* A, int a_ld,
                             int n_in, float* dW, int w_ld, float* db, int a_relu = 0) {
  WProd& r = W.prod[np++];
  r.G = G; r.g_ld = g_ld; r.n_out = n_out; r.A = A; r.a_ld = a_ld; r.n_in = n_in;
  r.dW = dW; r.w_ld = w_ld; r.db = db; r.a_relu = a_relu;
}

// launch `np` products over W.Q rows, ~8 CTAs per SM over all products
static inline int launch_wgrad(WParams& W, int np, cudaStream_t stream) {
  if (np <= 0 || W.Q <= 0) return VTACO_OK;
  if (np > kMaxProd) return VTACO_ERR_UNSUPPORTED;
  const long long Q = W.Q;
  const long long want = (long long)num_sms() * 8 / np + 1;
  long long chunks = std::min<long long>((Q + kWTile - 1) / kWTile, want);
  if (chunks < 1) chunks = 1;
  long long per = (Q + chunks - 1) / chunks;
  per = (per + kWTile - 1) / kWTile * kWTile;
  W.q_per_cta = (int)per;
  chunks = (Q + per - 1) / per;
  linear_wgrad_kernel<<<dim3((unsigned)chunks, (unsigned)np), 256, 0, stream>>>(W);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

}  // namespace vtaco
