// Backward of the PointNet part of LocalPoolPointnet — included by encoder.cu.
//
// Gradients of reference src/encoder/pointnet.py:135-172 (forward up to the UNets) with respect
// to fc_pos, the ResnetBlockFC stack and fc_c, given the gradients of the scatter_mean feature
// planes / grid (what torch autograd hands back from the UNet / UNet3D or from the decoder's
// feature sampling).  torch_scatter semantics: scatter_mean divides the cell gradient by the
// cell count; scatter_max routes the cell gradient to ONE arg-max point per (cell, channel) —
// here the lowest point index among exact ties.
//
// Pipeline (all on the caller's stream, buffers from the caller's workspace):
//   1. forward recompute with every level kept (the inference forward rotates three pooling
//      buffers and keeps nothing): net_k for every block, the pooling buffer read by block k;
//   2. head:  dc[t] = sum_keys dfeat[key][cell(t)] / count;  dnet_last = fc_c^T dc;
//   3. per block, last to first: one thread per point recomputes the block input x (64), the
//      hidden ReLU mask, and back-propagates  gh = (W1^T g) * [h>0],
//      dx = Ws^T g + (W0^T gh) * [x>0];  dx[:32] is the gradient of the previous net, dx[32:]
//      the gradient of the pooled feature, which is summed per cell (warp-aggregated row
//      atomics, shared with the forward) and routed back to the arg-max points;
//   4. weight gradients by linear_wgrad_kernel (wgrad.cuh) from the stored rows.

namespace vtaco {

struct EncBwdParams {
  EncParams F;              // geometry, indices, weights
  const float* dfeat[4];    // channels-last feature gradients [B][cells][32]
  const float* net_prev;    // net_{k-1} rows (block kernel, !FIRST)
  const float* pool_k[4];   // pooling buffer read by block k
  const float* gnet;        // dL/dnet_k rows
  float* gnet_prev;         // dL/dnet_{k-1} rows (net part of dx)
  float* dpooled;           // dx[32:64] rows
  float* xrow;              // block input rows [n][64]
  float* rh;                // relu(h) rows
  float* gh;                // masked hidden gradient rows
  float* dx0;               // FIRST: dL/d fc_pos output rows [n][64]
  float* dc;                // head: dL/dc rows
  const float* net_last;
};

__device__ __forceinline__ void row_store32(float* __restrict__ row, const float (&v)[32]) {
  float4* r4 = reinterpret_cast<float4*>(row);
#pragma unroll
  for (int j = 0; j < 8; ++j) r4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void row_load32(float (&v)[32], const float* __restrict__ row) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = r4[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
// dot of a K-major weight row (32 outputs) with g
__device__ __forceinline__ float row_dot32(const float* __restrict__ Wk, const float (&g)[32]) {
  const float4* w4 = reinterpret_cast<const float4*>(Wk);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 w = w4[j];
    s0 = fmaf(w.x, g[4 * j], s0);
    s1 = fmaf(w.y, g[4 * j + 1], s1);
    s2 = fmaf(w.z, g[4 * j + 2], s2);
    s3 = fmaf(w.w, g[4 * j + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// dc = sum_keys dfeat[cell]/count (scatter_mean backward, pointnet.py:93,108); dnet_last = fc_c^T dc
__global__ void __launch_bounds__(kET) encb_head_kernel(const __grid_constant__ EncBwdParams Q) {
  extern __shared__ __align__(16) float esm[];
  float* sW = esm;                   // fc_c (1056)
  float* sX = sW + ENC_FCC_FLOATS;   // [32][kES]
  const EncParams& P = Q.F;
  const float* Wg = P.W + ENC_OFF_BLOCKS + (long long)P.n_blocks * ENC_BLOCK_STRIDE;
  for (int i = threadIdx.x; i < ENC_FCC_FLOATS / 4; i += kET)
    reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(Wg) + i);
  __syncthreads();
  const long long n = (long long)blockIdx.x * kET + threadIdx.x;
  if (n >= P.n) return;
  const long long b = n / P.T;
  float dc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) dc[j] = 0.f;
  for (int k = 0; k < P.nkeys; ++k) {
    if (!Q.dfeat[k]) continue;
    const float inv = 1.0f / (float)P.count[k][P.slot[k][n]];
    const float4* r4 = reinterpret_cast<const float4*>(Q.dfeat[k] + (b * P.cells[k] + P.idx[k][n]) * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(r4 + j);
      dc[4 * j] = fmaf(t.x, inv, dc[4 * j]);
      dc[4 * j + 1] = fmaf(t.y, inv, dc[4 * j + 1]);
      dc[4 * j + 2] = fmaf(t.z, inv, dc[4 * j + 2]);
      dc[4 * j + 3] = fmaf(t.w, inv, dc[4 * j + 3]);
    }
  }
  row_store32(Q.dc + n * 32, dc);
  float* xcol = sX + threadIdx.x;
#pragma unroll 4
  for (int k = 0; k < 32; ++k) xcol[k * kES] = row_dot32(sW + k * 32, dc);
  float g[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) g[k] = xcol[k * kES];
  row_store32(Q.gnet_prev + n * 32, g);
}

// Backward of one ResnetBlockFC(64 -> 32) (layers.py:41-50) including its input assembly.
template <bool FIRST>
__global__ void __launch_bounds__(kET) encb_block_kernel(const __grid_constant__ EncBwdParams Q, int blk) {
  extern __shared__ __align__(16) float esm[];
  float* sW = esm;                          // block weights (5184) + fc_pos (256)
  float* sX = sW + ENC_BLOCK_STRIDE + 256;  // [64][kES] block input x
  float* sD = sX + 64 * kES;                // [64][kES] dx
  const EncParams& P = Q.F;
  const float* Wg = P.W + ENC_OFF_BLOCKS + (long long)blk * ENC_BLOCK_STRIDE;
  for (int i = threadIdx.x; i < ENC_BLOCK_STRIDE / 4; i += kET)
    reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(Wg) + i);
  if (FIRST)
    for (int i = threadIdx.x; i < 256 / 4; i += kET)
      reinterpret_cast<float4*>(sW + ENC_BLOCK_STRIDE)[i] = __ldg(reinterpret_cast<const float4*>(P.W) + i);
  __syncthreads();

  const int tid = threadIdx.x, lane = tid & 31, col0 = tid & ~31;
  const long long nb0 = (long long)blockIdx.x * kET;
  const long long n = nb0 + tid;
  const bool valid = n < P.n;
  int myslot[4] = {0, 0, 0, 0};
  for (int k = 0; k < P.nkeys; ++k) myslot[k] = valid ? P.slot[k][n] : 0;

  // ---- block input x, exactly as the forward assembles it ----
  float* xcol = sX + tid;
  float* dcol = sD + tid;
  if (FIRST) {
    const float* Wp = sW + ENC_BLOCK_STRIDE;
    const float px = valid ? P.p[n * 3] : 0.f, py = valid ? P.p[n * 3 + 1] : 0.f, pz = valid ? P.p[n * 3 + 2] : 0.f;
#pragma unroll 8
    for (int j = 0; j < 64; ++j)
      xcol[j * kES] = fmaf(Wp[128 + j], pz, fmaf(Wp[64 + j], py, fmaf(Wp[j], px, Wp[ENC_OFF_BPOS + j])));
  } else {
    for (int i = 0; i < 32; ++i) {
      const long long ni = nb0 + col0 + i;
      if (ni >= P.n) break;
      sX[lane * kES + col0 + i] = Q.net_prev[ni * 32 + lane];
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < P.nkeys) {
          const int sl = __shfl_sync(kFullMask, myslot[k], i);
          float v = Q.pool_k[k][(long long)sl * 32 + lane];
          if (P.pool_mean) v = v / (float)P.count[k][sl];
          s += v;
        }
      }
      sX[(32 + lane) * kES + col0 + i] = s;
    }
  }
  __syncwarp();
  if (!valid) return;   // no block-level barrier below

  // ---- hidden layer recompute: h = fc_0(relu(x)) ----
  float h[32], g[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) h[j] = sW[ENC_B_B0 + j];
#pragma unroll 2
  for (int k = 0; k < 64; ++k) axpy32(h, sW + ENC_B_W0 + k * 32, fmaxf(xcol[k * kES], 0.f));
  uint32_t mh = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    mh |= (h[j] > 0.f ? 1u : 0u) << j;
    h[j] = fmaxf(h[j], 0.f);
  }
  row_store32(Q.rh + n * 32, h);

  // ---- gh = (W1^T g) * [h > 0] ----
  row_load32(g, Q.gnet + n * 32);
#pragma unroll 4
  for (int k = 0; k < 32; ++k) dcol[k * kES] = row_dot32(sW + ENC_B_W1 + k * 32, g);
#pragma unroll
  for (int k = 0; k < 32; ++k) h[k] = ((mh >> k) & 1u) ? dcol[k * kES] : 0.f;
  row_store32(Q.gh + n * 32, h);

  // ---- dx = Ws^T g + (W0^T gh) * [x > 0] ----
#pragma unroll 2
  for (int k = 0; k < 64; ++k) {
    const float a = row_dot32(sW + ENC_B_WS + k * 32, g);
    const float c = row_dot32(sW + ENC_B_W0 + k * 32, h);
    dcol[k * kES] = a + (xcol[k * kES] > 0.f ? c : 0.f);
  }

  // ---- rows out: x (64) for the weight gradients, dx split into net / pooled parts ----
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int k = 0; k < 32; ++k) h[k] = xcol[(32 * half + k) * kES];
    row_store32(Q.xrow + n * 64 + 32 * half, h);
#pragma unroll
    for (int k = 0; k < 32; ++k) h[k] = dcol[(32 * half + k) * kES];
    if (FIRST) row_store32(Q.dx0 + n * 64 + 32 * half, h);
    else row_store32((half == 0 ? Q.gnet_prev : Q.dpooled) + n * 32, h);
  }
}

// arg-max election: lowest point index whose value equals the pooled maximum (lane == channel)
__global__ void __launch_bounds__(256) encb_arg_kernel(const float* __restrict__ net, long long n, int nkeys,
                                                       const int32_t* s0, const int32_t* s1, const int32_t* s2,
                                                       const int32_t* s3, const float* p0, const float* p1,
                                                       const float* p2, const float* p3, int32_t* a0, int32_t* a1,
                                                       int32_t* a2, int32_t* a3) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ni = t >> 5;
  const int lane = threadIdx.x & 31;
  if (ni >= n) return;
  const int32_t* sl[4] = {s0, s1, s2, s3};
  const float* pool[4] = {p0, p1, p2, p3};
  int32_t* arg[4] = {a0, a1, a2, a3};
  const float v = net[ni * 32 + lane];
  for (int k = 0; k < nkeys; ++k) {
    const long long r = (long long)sl[k][ni] * 32 + lane;
    if (v == pool[k][r]) atomicMin(arg[k] + r, (int)ni);
  }
}

// gnet[t] += sum_keys routed cell gradient (max: only the elected arg-max point; mean: / count)
__global__ void __launch_bounds__(256) encb_route_kernel(float* __restrict__ gnet, long long n, int nkeys, int mean,
                                                         const int32_t* s0, const int32_t* s1, const int32_t* s2,
                                                         const int32_t* s3, const float* c0, const float* c1,
                                                         const float* c2, const float* c3, const int32_t* a0,
                                                         const int32_t* a1, const int32_t* a2, const int32_t* a3,
                                                         const int32_t* n0, const int32_t* n1, const int32_t* n2,
                                                         const int32_t* n3) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ni = t >> 5;
  const int lane = threadIdx.x & 31;
  if (ni >= n) return;
  const int32_t* sl[4] = {s0, s1, s2, s3};
  const float* cg[4] = {c0, c1, c2, c3};
  const int32_t* arg[4] = {a0, a1, a2, a3};
  const int32_t* cnt[4] = {n0, n1, n2, n3};
  float acc = 0.f;
  for (int k = 0; k < nkeys; ++k) {
    const int s = sl[k][ni];
    const long long r = (long long)s * 32 + lane;
    if (mean) acc += cg[k][r] / (float)cnt[k][s];
    else if (arg[k][r] == (int)ni) acc += cg[k][r];
  }
  gnet[ni * 32 + lane] += acc;
}

}  // namespace vtaco

using namespace vtaco;

namespace {
struct EncBwdLayout {
  int32_t *idx[4], *slot[4], *count[4], *map[4], *arg[4];
  float* pool[4][8];   // [key][level 1..nb-1] pooling buffer read by block `level`
  float* cellgrad[4];
  float* net[8];
  float *gnet[2], *dc, *dpooled, *xrow, *rh, *gh, *dx0;
  long long bytes;
};

EncBwdLayout carve_enc_bwd(char* base, int B, long long T, int nkeys, const int32_t* kind, const int32_t* reso, int nb) {
  EncBwdLayout L{};
  const long long n = (long long)B * T;
  Carver c{base, 0};
  for (int k = 0; k < nkeys; ++k) {
    L.idx[k] = c.take<int32_t>(n); L.slot[k] = c.take<int32_t>(n); L.count[k] = c.take<int32_t>(n);
    L.map[k] = c.take<int32_t>((long long)B * cells_of(kind[k], reso[k]));
    L.arg[k] = c.take<int32_t>(n * 32);
    for (int l = 1; l < nb; ++l) L.pool[k][l] = c.take<float>(n * 32);
    L.cellgrad[k] = c.take<float>(n * 32);
  }
  for (int l = 0; l < nb; ++l) L.net[l] = c.take<float>(n * 32);
  L.gnet[0] = c.take<float>(n * 32); L.gnet[1] = c.take<float>(n * 32);
  L.dc = c.take<float>(n * 32); L.dpooled = c.take<float>(n * 32);
  L.xrow = c.take<float>(n * 64); L.rh = c.take<float>(n * 32); L.gh = c.take<float>(n * 32);
  L.dx0 = c.take<float>(n * 64);
  L.bytes = c.off;
  return L;
}
}  // namespace

extern "C" int64_t vtaco_encoder_backward_workspace_bytes(int32_t B, int64_t T, int32_t n_keys, const int32_t* kind,
                                                          const int32_t* reso, int32_t n_blocks) {
  if (B <= 0 || T <= 0 || n_keys <= 0 || n_keys > 4 || !kind || !reso || n_blocks < 1 || n_blocks > 8)
    return VTACO_ERR_INVALID_ARG;
  return carve_enc_bwd(nullptr, B, T, n_keys, kind, reso, n_blocks).bytes;
}

extern "C" int vtaco_encoder_backward(const vtaco_encoder_bwd_args* a, void* stream) {
  if (!a || !a->p || !a->weights || !a->workspace || !a->d_params) return VTACO_ERR_INVALID_ARG;
  if (a->B <= 0 || a->T <= 0 || a->n_keys <= 0 || a->n_keys > 4 || a->n_blocks < 1 || a->n_blocks > 8)
    return VTACO_ERR_INVALID_ARG;
  const long long n = (long long)a->B * a->T;
  if (n >= (1ll << 26)) return VTACO_ERR_UNSUPPORTED;   // arg / row offsets stay in int32 range
  for (int k = 0; k < a->n_keys; ++k) {
    if (a->kind[k] < 0 || a->kind[k] > 3 || a->reso[k] < 1) return VTACO_ERR_INVALID_ARG;
    if (a->kind[k] == VTACO_GRID ? a->reso[k] > 1290 : a->reso[k] > 46340) return VTACO_ERR_UNSUPPORTED;
  }
  const int nb = a->n_blocks, nk = a->n_keys;
  EncBwdLayout L = carve_enc_bwd(reinterpret_cast<char*>(a->workspace), a->B, a->T, nk, a->kind, a->reso, nb);
  if (L.bytes > a->workspace_bytes) return VTACO_ERR_CAPACITY;
  cudaStream_t st = (cudaStream_t)stream;

  EncParams P = {};
  P.p = a->p; P.n = n; P.T = a->T; P.B = a->B;
  P.nc = make_norm_const(a->padding, a->div_mode);
  P.nkeys = nk; P.pool_mean = a->pool_mean ? 1 : 0; P.n_blocks = nb; P.W = a->weights;
  for (int k = 0; k < nk; ++k) {
    P.kind[k] = a->kind[k]; P.reso[k] = a->reso[k]; P.cells[k] = cells_of(a->kind[k], a->reso[k]);
    P.idx[k] = L.idx[k]; P.slot[k] = L.slot[k]; P.count[k] = L.count[k]; P.map[k] = L.map[k];
    VTACO_CUDA_CHECK(cudaMemsetAsync(P.map[k], 0x7f, sizeof(int32_t) * a->B * P.cells[k], st));
    VTACO_CUDA_CHECK(cudaMemsetAsync(P.count[k], 0, sizeof(int32_t) * n, st));
  }
  const unsigned g256 = (unsigned)((n + 255) / 256), gE = (unsigned)((n + kET - 1) / kET);
  const unsigned gRow = (unsigned)((n * 32 + 255) / 256);
  const long long n_groups = (n + kEG - 1) / kEG;
  const unsigned gB = enc_block_grid(n_groups);
  const size_t smem_bwd = (ENC_BLOCK_STRIDE + 256 + 128 * kES) * sizeof(float);
  const size_t smem_fin = (ENC_FCC_FLOATS + 32 * kES) * sizeof(float);
  static std::atomic<bool> configured[64];
  int dev = 0;
  VTACO_CUDA_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63].load(std::memory_order_relaxed)) {
    VTACO_CUDA_CHECK(cudaFuncSetAttribute(encb_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd));
    VTACO_CUDA_CHECK(cudaFuncSetAttribute(encb_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd));
    configured[dev & 63].store(true, std::memory_order_relaxed);
  }

  // ---- 1. forward recompute, every level kept ----
  for (int k = 0; k < nk; ++k) P.pool[0][k] = (nb > 1) ? L.pool[k][1] : nullptr;
  enc_index_kernel<<<g256, 256, 0, st>>>(P);
  enc_slot_kernel<<<g256, 256, 0, st>>>(P, nb > 1 ? 1 : 0);
  for (int k = 0; k < nk; ++k) {   // block 0 scatters into pool level 1 and initialises level 2
    P.pool[0][k] = nb > 1 ? L.pool[k][1] : nullptr;
    P.pool[1][k] = nb > 2 ? L.pool[k][2] : nullptr;
  }
  P.net[0] = L.net[0];
  enc_block_kernel<true><<<gB, 128, kEncBlockSmem, st>>>(P, 0, -1, nb > 1 ? 0 : -1, nb > 2 ? 1 : -1, 0, 0, n_groups);
  for (int i = 1; i < nb; ++i) {
    const bool last = (i == nb - 1);
    for (int k = 0; k < nk; ++k) {
      P.pool[0][k] = L.pool[k][i];
      P.pool[1][k] = last ? nullptr : L.pool[k][i + 1];
      P.pool[2][k] = (i + 2 < nb) ? L.pool[k][i + 2] : nullptr;
    }
    P.net[0] = L.net[i - 1];
    P.net[1] = L.net[i];
    enc_block_kernel<false><<<gB, 128, kEncBlockSmem, st>>>(P, i, 0, last ? -1 : 1, (i + 2 < nb) ? 2 : -1, 0, 1, n_groups);
  }
  VTACO_LAUNCH_CHECK();

  // ---- 2. head ----
  EncBwdParams Q = {};
  Q.F = P;
  for (int k = 0; k < nk; ++k) Q.dfeat[k] = a->d_out_cl[k];
  Q.dc = L.dc;
  Q.gnet_prev = L.gnet[(nb - 1) & 1];
  encb_head_kernel<<<gE, kET, smem_fin, st>>>(Q);
  VTACO_LAUNCH_CHECK();
  float* dp = a->d_params;
  {
    WParams W{};
    W.Q = n;
    int np = 0;
    float* d = dp + ENC_OFF_BLOCKS + (long long)nb * ENC_BLOCK_STRIDE;
    wgrad_add(W, np, L.dc, 32, 32, L.net[nb - 1], 32, 32, d, 32, d + 1024);
    const int r = launch_wgrad(W, np, st);
    if (r != VTACO_OK) return r;
  }

  // ---- 3. blocks, last to first ----
  for (int i = nb - 1; i >= 0; --i) {
    Q.gnet = L.gnet[i & 1];
    Q.gnet_prev = L.gnet[(i + 1) & 1];
    Q.dpooled = L.dpooled; Q.xrow = L.xrow; Q.rh = L.rh; Q.gh = L.gh; Q.dx0 = L.dx0;
    if (i > 0) {
      Q.net_prev = L.net[i - 1];
      for (int k = 0; k < nk; ++k) Q.pool_k[k] = L.pool[k][i];
      encb_block_kernel<false><<<gE, kET, smem_bwd, st>>>(Q, i);
      // pooled-feature gradient: per-cell sum, then routing to the arg-max points
      for (int k = 0; k < nk; ++k) {
        VTACO_CUDA_CHECK(cudaMemsetAsync(L.cellgrad[k], 0, sizeof(float) * n * 32, st));
        if (!P.pool_mean) VTACO_CUDA_CHECK(cudaMemsetAsync(L.arg[k], 0x7f, sizeof(int32_t) * n * 32, st));
      }
      rows_scatter_kernel<true><<<gE, kET, 0, st>>>(L.dpooled, n, nk, L.slot[0], L.slot[1], L.slot[2], L.slot[3],
                                                    L.cellgrad[0], L.cellgrad[1], L.cellgrad[2], L.cellgrad[3]);
      if (!P.pool_mean)
        encb_arg_kernel<<<gRow, 256, 0, st>>>(L.net[i - 1], n, nk, L.slot[0], L.slot[1], L.slot[2], L.slot[3],
                                              L.pool[0][i], L.pool[1][i], L.pool[2][i], L.pool[3][i], L.arg[0],
                                              L.arg[1], L.arg[2], L.arg[3]);
      encb_route_kernel<<<gRow, 256, 0, st>>>(L.gnet[(i + 1) & 1], n, nk, P.pool_mean, L.slot[0], L.slot[1],
                                              L.slot[2], L.slot[3], L.cellgrad[0], L.cellgrad[1], L.cellgrad[2],
                                              L.cellgrad[3], L.arg[0], L.arg[1], L.arg[2], L.arg[3], L.count[0],
                                              L.count[1], L.count[2], L.count[3]);
    } else {
      encb_block_kernel<true><<<gE, kET, smem_bwd, st>>>(Q, 0);
    }
    VTACO_LAUNCH_CHECK();
    // ---- 4. weight gradients of block i (native [out][in] orientation at the packed offsets) ----
    WParams W{};
    W.Q = n;
    int np = 0;
    float* d = dp + ENC_OFF_BLOCKS + (long long)i * ENC_BLOCK_STRIDE;
    const float* g = L.gnet[i & 1];
    wgrad_add(W, np, L.gh, 32, 32, L.xrow, 64, 32, d + ENC_B_W0, 64, d + ENC_B_B0, 1);        // fc_0, inputs 0..31
    wgrad_add(W, np, L.gh, 32, 32, L.xrow + 32, 64, 32, d + ENC_B_W0 + 32, 64, nullptr, 1);   // fc_0, inputs 32..63
    wgrad_add(W, np, g, 32, 32, L.rh, 32, 32, d + ENC_B_W1, 32, d + ENC_B_B1);                // fc_1
    wgrad_add(W, np, g, 32, 32, L.xrow, 64, 32, d + ENC_B_WS, 64, nullptr);                   // shortcut
    wgrad_add(W, np, g, 32, 32, L.xrow + 32, 64, 32, d + ENC_B_WS + 32, 64, nullptr);
    if (i == 0) {                                                                             // fc_pos [64][3]
      wgrad_add(W, np, L.dx0, 64, 32, a->p, 3, 3, dp + ENC_OFF_WPOS, 3, dp + ENC_OFF_BPOS);
      wgrad_add(W, np, L.dx0 + 32, 64, 32, a->p, 3, 3, dp + ENC_OFF_WPOS + 96, 3, dp + ENC_OFF_BPOS + 32);
    }
    const int r = launch_wgrad(W, np, st);
    if (r != VTACO_OK) return r;
  }
  return VTACO_OK;
}
