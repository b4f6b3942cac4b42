// UNet3D building blocks on our own kernels (reference src/encoder/unet3d.py:
// SingleConv 'gcr' = GroupNorm -> Conv3d(3x3x3, pad 1, no bias) -> ReLU, Encoder = MaxPool3d(2) +
// DoubleConv, Decoder = nearest-upsample + concat + DoubleConv, final 1x1x1 Conv3d with bias).
//
// One 'gcr' layer = ONE kernel: implicit-GEMM convolution on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, A and B from shared memory, accumulator in TMEM) over channels-last
// activations, with
//   * GroupNorm-apply fused into the operand loader (per-channel scale / shift derived in the
//     prologue from per-channel (sum, sumsq) that the PRODUCER of the input accumulated),
//   * nearest-upsample + concat fused into the loader (channels >= C1 are read from the
//     half-resolution tensor), so the concatenated tensor is never materialised,
//   * ReLU + per-channel (sum, sumsq) of the output (the next layer's GroupNorm statistics) fused
//     into the epilogue.
// Tiling: a CTA owns 1 x 16 x 8 (z,y,x) output voxels = the 128 rows of an M128 N32 MMA and 32
// output channels.  Per 16 input channels the loader builds the halo brick (3 x 18 x 10 voxels) in
// shared memory as 4 planes [channel quad][voxel][4], the UMMA canonical K-major no-swizzle
// layout with the voxel as the row: a filter tap is then just a different START ADDRESS of the
// same brick (row-group stride = one brick row), so im2col costs no data movement at all —
// 27 taps x 2 MMAs (K = 8) per chunk.
// Arithmetic: single-pass TF32 (operands rounded to nearest TF32, fp32 accumulation) — what the
// reference runs on a GPU (cuDNN with torch.backends.cudnn.allow_tf32 = True, torch's default).
#include "common.cuh"
#include <math_constants.h>

namespace vtaco {

constexpr int kCvThreads = 256;
constexpr int kCvKC = 16;                 // input channels per chunk (4 planes of 4)
constexpr int kCvTileY = 16, kCvTileX = 8;
constexpr int kCvMaxCin = 512;
constexpr int kCvItems = (3 * 18 * 10 * 4 + kCvThreads - 1) / kCvThreads;   // brick float4s per thread (3x3x3 halo brick)
constexpr uint32_t kCvIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (4u << 17) | (8u << 24);  // F32 acc, TF32 x TF32, K-major, N=32, M=128

struct ConvParams {
  const float* x;      // [N][D][H][W][C1]
  const float* x2;     // [N][D2][H2][W2][C2] or NULL
  const float* w;      // packed: [Cout/32][Cin/16][taps][4][32][4]
  const float* bias;   // [Cout] or NULL
  const double* in_stats;   // [N][Cin][2] or NULL (no GroupNorm)
  const float* gamma;  // [Cin]
  const float* beta;
  float* y;            // [N][D][H][W][Cout]
  double* out_stats;   // [N][Cout][2] or NULL
  int N, D, H, W, C1, C2, D2, H2, W2, Cout, ksize, groups, relu;
  int tiles_x, tiles_y;
  double eps;
};

__device__ __forceinline__ uint32_t cv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t cv_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // between core matrices adjacent in K
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  return d;
}
__device__ __forceinline__ void cv_mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(kCvIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float cv_tf32(float v) {   // round to nearest (ties away), like cvt.rna.tf32
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

struct CvSmem { int brick, w, scale, shift, stat, bar, tmem, total, plane_stride, row_pitch, bvox; };
__host__ __device__ inline CvSmem cv_layout(int ksize, int cin) {
  CvSmem s;
  const int h = ksize / 2;
  const int bz = 1 + 2 * h, by = kCvTileY + 2 * h, bx = kCvTileX + 2 * h;
  s.bvox = bz * by * bx;
  s.row_pitch = bx * 16;                                   // bytes between brick rows (= 8-row groups of the MMA)
  s.plane_stride = (s.bvox * 16 + 80 + 127) / 128 * 128 + 16;   // odd multiple of 16 B: spreads the planes over the banks
  s.brick = 0;
  s.w = s.brick + 4 * s.plane_stride;
  s.w = (s.w + 127) / 128 * 128;
  const int taps = ksize * ksize * ksize;
  s.scale = s.w + taps * 2048;
  s.shift = s.scale + cin * 4;
  s.stat = s.shift + cin * 4;
  s.bar = (s.stat + 64 * 4 + 15) / 16 * 16;   // [0] MMA completion, [1] weights landed (TMA bulk copy)
  s.tmem = s.bar + 16;
  s.total = s.tmem + 16;
  return s;
}

__global__ void __launch_bounds__(kCvThreads, 2) conv3d_tc_kernel(const __grid_constant__ ConvParams P) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int Cin = P.C1 + P.C2;
  const CvSmem L = cv_layout(P.ksize, Cin);
  float* sScale = reinterpret_cast<float*>(sm + L.scale);
  float* sShift = reinterpret_cast<float*>(sm + L.shift);
  float* sStat = reinterpret_cast<float*>(sm + L.stat);
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sm + L.bar);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sm + L.tmem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int h = P.ksize / 2, taps = P.ksize * P.ksize * P.ksize;
  const int bx_ext = kCvTileX + 2 * h, by_ext = kCvTileY + 2 * h;

  // tile -> (z, y0, x0)
  int t = blockIdx.x;
  const int tx = t % P.tiles_x; t /= P.tiles_x;
  const int ty = t % P.tiles_y;
  const int z = t / P.tiles_y;
  const int y0 = ty * kCvTileY, x0 = tx * kCvTileX;
  const int ntile = blockIdx.y, n = blockIdx.z;

  // ---- prologue: GroupNorm scale / shift per input channel, barrier, TMEM ----
  if (P.in_stats) {
    const int cpg = Cin / P.groups;
    const double cnt = (double)P.D * P.H * P.W * cpg;
    for (int c = tid; c < Cin; c += kCvThreads) {
      const int g = c / cpg;
      double s = 0.0, ss = 0.0;
      for (int j = 0; j < cpg; ++j) {
        s += P.in_stats[((size_t)n * Cin + g * cpg + j) * 2];
        ss += P.in_stats[((size_t)n * Cin + g * cpg + j) * 2 + 1];
      }
      const double mean = s / cnt;
      double var = ss / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      const double rstd = 1.0 / sqrt(var + P.eps);
      const double ga = P.gamma ? (double)P.gamma[c] : 1.0, be = P.beta ? (double)P.beta[c] : 0.0;
      sScale[c] = (float)(rstd * ga);
      sShift[c] = (float)(be - mean * rstd * ga);
    }
  } else {
    for (int c = tid; c < Cin; c += kCvThreads) { sScale[c] = 1.0f; sShift[c] = 0.0f; }
  }
  if (tid < 64) sStat[tid] = 0.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cv_smem_u32(sBar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cv_smem_u32(sBar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cv_smem_u32(sTmem)), "r"(32)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *sTmem;
  const uint32_t bar = cv_smem_u32(sBar), wbar = cv_smem_u32(sBar + 1);
  uint32_t wphase = 0;
  const uint32_t brick_sm = cv_smem_u32(sm + L.brick), w_sm = cv_smem_u32(sm + L.w);
  uint32_t phase = 0;

  // this thread's brick items (i = tid + u * 256 -> voxel i / 4, channel quad i % 4 == tid % 4): global voxel
  // index in the full-resolution source and in the half-resolution one, or -1 outside the volume
  const int total_items = L.bvox * 4;
  int voxA[kCvItems], voxB[kCvItems];
#pragma unroll
  for (int u = 0; u < kCvItems; ++u) {
    const int i = tid + u * kCvThreads;
    const int v = i >> 2;
    const int bxv = v % bx_ext, r = v / bx_ext;
    const int byv = r % by_ext, bzv = r / by_ext;
    const int gz = z + bzv - h, gy = y0 + byv - h, gx = x0 + bxv - h;
    const bool inb = i < total_items && gz >= 0 && gz < P.D && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
    voxA[u] = inb ? (int)((((size_t)n * P.D + gz) * P.H + gy) * P.W + gx) : -1;
    voxB[u] = -1;
    if (inb && P.C2 > 0) {   // ATen nearest: src = min(floor(dst * in / out), in - 1)
      const int sz = min(gz * P.D2 / P.D, P.D2 - 1), sy = min(gy * P.H2 / P.H, P.H2 - 1), sx = min(gx * P.W2 / P.W, P.W2 - 1);
      voxB[u] = (int)((((size_t)n * P.D2 + sz) * P.H2 + sy) * P.W2 + sx);
    }
  }

  const int n_chunks = Cin / kCvKC;
  const float4* wsrc = reinterpret_cast<const float4*>(P.w) + (size_t)ntile * n_chunks * taps * 128;
  for (int ch = 0; ch < n_chunks; ++ch) {
    // ---- weights of this (out-channel tile, chunk): taps x 2 KB, contiguous in the packed buffer: ONE bulk
    //      async copy (TMA engine, completion on an mbarrier) instead of 13 dependent loads per thread ----
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)taps * 2048u;
      const float4* src = wsrc + (size_t)ch * taps * 128;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w_sm),
                   "l"(src), "r"(bytes), "r"(wbar)
                   : "memory");
    }
    // ---- halo brick of 16 input channels: GroupNorm-apply, TF32 rounding, zero padding.  The brick geometry
    //      is the same for every chunk, so the voxel offsets of this thread's items were computed once (voxA /
    //      voxB); the loads of a batch are all issued before the first is used. ----
    {
      const int c0 = ch * kCvKC;
      const bool second = c0 >= P.C1;                       // channels from the half-resolution tensor (upsample + concat)
      const float* src = second ? P.x2 : P.x;
      const int Cs = second ? P.C2 : P.C1, cs0 = second ? c0 - P.C1 : c0;
      const int q = tid & 3;
      const float4 sc = *reinterpret_cast<const float4*>(sScale + c0 + 4 * q);
      const float4 sh = *reinterpret_cast<const float4*>(sShift + c0 + 4 * q);
#pragma unroll
      for (int b0 = 0; b0 < kCvItems; b0 += 5) {
        float4 val[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          if (b0 + u < kCvItems) {
            const int vox = second ? voxB[b0 + u] : voxA[b0 + u];
            val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vox >= 0) val[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)vox * Cs + cs0) + q);
          }
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          if (b0 + u < kCvItems) {
            const int i = tid + (b0 + u) * kCvThreads;
            if (i < total_items) {
              float4 o = val[u];
              if ((second ? voxB[b0 + u] : voxA[b0 + u]) >= 0) {
                o.x = cv_tf32(fmaf(o.x, sc.x, sh.x));
                o.y = cv_tf32(fmaf(o.y, sc.y, sh.y));
                o.z = cv_tf32(fmaf(o.z, sc.z, sh.z));
                o.w = cv_tf32(fmaf(o.w, sc.w, sh.w));
              }
              *reinterpret_cast<float4*>(sm + L.brick + q * L.plane_stride + (i >> 2) * 16) = o;
            }
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      uint32_t pred;
      asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
      if (pred) {
        asm volatile(   // the weights have landed
            "{\n.reg .pred p;\nCW_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra CW_DONE;\nbra CW_WAIT;\nCW_DONE:\n}\n" ::"r"(wbar),
            "r"(wphase)
            : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int tap = 0; tap < taps; ++tap) {
          const int dx = tap % P.ksize, dy = (tap / P.ksize) % P.ksize, dz = tap / (P.ksize * P.ksize);
          const uint32_t a0 = brick_sm + (uint32_t)(((dz * by_ext + dy) * bx_ext + dx) * 16);
          const uint32_t b0 = w_sm + (uint32_t)tap * 2048u;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            cv_mma_ss(tmem_d, cv_desc(a0 + ks * 2 * L.plane_stride, L.plane_stride, L.row_pitch),
                      cv_desc(b0 + ks * 1024, 512, 128), (ch | tap | ks) ? 1u : 0u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      }
      __syncwarp();
    }
    wphase ^= 1;
    // the MMAs read the brick and the weights: wait before the next chunk overwrites them
    asm volatile(
        "{\n.reg .pred p;\nCV_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra CV_DONE;\nbra CV_WAIT;\nCV_DONE:\n}\n" ::"r"(bar),
        "r"(phase)
        : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // ---- epilogue: thread = (voxel row m, 16-channel half); bias, ReLU, store, per-channel statistics ----
  const int lq = warp & 3, hv = warp >> 2;
  const int m = lq * 32 + lane;
  const int gy = y0 + (m >> 3), gx = x0 + (m & 7);
  const bool valid = gy < P.H && gx < P.W;
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(tmem_d + ((uint32_t)(32 * lq) << 16) + (uint32_t)(16 * hv))
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  float o[16];
  const int co0 = ntile * 32 + 16 * hv;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = __uint_as_float(r[j]);
    if (P.bias) v += __ldg(P.bias + co0 + j);
    if (P.relu) v = fmaxf(v, 0.0f);
    o[j] = valid ? v : 0.0f;
  }
  if (valid) {
    float4* dst = reinterpret_cast<float4*>(P.y + ((((size_t)n * P.D + z) * P.H + gy) * P.W + gx) * P.Cout + co0);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  if (P.out_stats) {
    // per-channel sum / sum of squares over the warp's 32 voxels: halves fold first, then a reduce-scatter
    float s[16], q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      s[j] = o[j] + __shfl_xor_sync(0xffffffffu, o[j], 16);
      const float sq = o[j] * o[j];
      q[j] = sq + __shfl_xor_sync(0xffffffffu, sq, 16);
    }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {      // after the step with width w a lane keeps w values
      const bool upper = (lane & w) != 0;
#pragma unroll
      for (int j = 0; j < w; ++j) {
        const float send_s = upper ? s[j] : s[j + w], keep_s = upper ? s[j + w] : s[j];
        const float send_q = upper ? q[j] : q[j + w], keep_q = upper ? q[j + w] : q[j];
        s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
        q[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
      }
    }
    // lane l (l < 16) now holds channel bitrev-free index: bit w of the lane selected the upper half at width w
    if (lane < 16) {
      const int c = 16 * hv + (lane & 15);
      atomicAdd(sStat + c, s[0]);
      atomicAdd(sStat + 32 + c, q[0]);
    }
    __syncthreads();
    if (tid < 64) {
      const int c = tid & 31, which = tid >> 5;
      atomicAdd(P.out_stats + ((size_t)n * P.Cout + ntile * 32 + c) * 2 + which, (double)sStat[tid]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32) : "memory");
}

// MaxPool3d(2) over channels-last activations + per-channel (sum, sumsq) of the result.
// One thread per (output voxel, channel quad).
__global__ void __launch_bounds__(256) maxpool2_cl_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D,
                                                          int H, int W, int C, double* __restrict__ stats) {
  const int C4 = C >> 2, Do = D >> 1, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Do * Ho * Wo * C4;
  extern __shared__ float pool_sm[];      // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) pool_sm[i] = 0.f;
  __syncthreads();
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % C4);
    long long v = t / C4;
    const int xo = (int)(v % Wo); v /= Wo;
    const int yo = (int)(v % Ho); v /= Ho;
    const int zo = (int)(v % Do);
    const int n = (int)(v / Do);
    float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zi = 2 * zo + (k >> 2), yi = 2 * yo + ((k >> 1) & 1), xi = 2 * xo + (k & 1);
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + ((((size_t)n * D + zi) * H + yi) * W + xi) * C) + q);
      m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
    }
    reinterpret_cast<float4*>(y + ((((size_t)n * Do + zo) * Ho + yo) * Wo + xo) * C)[q] = m;
    if (stats) {   // N == 1 per launch when statistics are wanted (the host loops over samples)
      atomicAdd(pool_sm + 4 * q, m.x); atomicAdd(pool_sm + 4 * q + 1, m.y);
      atomicAdd(pool_sm + 4 * q + 2, m.z); atomicAdd(pool_sm + 4 * q + 3, m.w);
      atomicAdd(pool_sm + C + 4 * q, m.x * m.x); atomicAdd(pool_sm + C + 4 * q + 1, m.y * m.y);
      atomicAdd(pool_sm + C + 4 * q + 2, m.z * m.z); atomicAdd(pool_sm + C + 4 * q + 3, m.w * m.w);
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      const int c = i % C, which = i / C;
      atomicAdd(stats + (size_t)c * 2 + which, (double)pool_sm[i]);
    }
  }
}

// per-channel (sum, sumsq) of a channels-last tensor [S][C] (one sample)
__global__ void __launch_bounds__(256) channel_stats_cl_kernel(const float* __restrict__ x, long long S, int C,
                                                               double* __restrict__ stats) {
  extern __shared__ float st_sm[];        // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) st_sm[i] = 0.f;
  __syncthreads();
  const int C4 = C >> 2;
  // a thread keeps one channel quad (blockDim.x is a multiple of C4) and strides over the voxels
  const int q = threadIdx.x % C4, lanes = blockDim.x / C4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  for (long long v = (long long)blockIdx.x * lanes + threadIdx.x / C4; v < S; v += (long long)gridDim.x * lanes) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)v * C) + q);
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    ss.x = fmaf(a.x, a.x, ss.x); ss.y = fmaf(a.y, a.y, ss.y); ss.z = fmaf(a.z, a.z, ss.z); ss.w = fmaf(a.w, a.w, ss.w);
  }
  atomicAdd(st_sm + 4 * q, s.x); atomicAdd(st_sm + 4 * q + 1, s.y); atomicAdd(st_sm + 4 * q + 2, s.z); atomicAdd(st_sm + 4 * q + 3, s.w);
  atomicAdd(st_sm + C + 4 * q, ss.x); atomicAdd(st_sm + C + 4 * q + 1, ss.y); atomicAdd(st_sm + C + 4 * q + 2, ss.z);
  atomicAdd(st_sm + C + 4 * q + 3, ss.w);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int c = i % C, which = i / C;
    atomicAdd(stats + (size_t)c * 2 + which, (double)st_sm[i]);
  }
}

}  // namespace vtaco

using namespace vtaco;

extern "C" int vtaco_conv3d_cl(const vtaco_conv3d_args* a, void* stream) {
  if (!a || !a->x || !a->w_packed || !a->y) return VTACO_ERR_INVALID_ARG;
  if (a->N < 1 || a->D < 1 || a->H < 1 || a->W < 1 || a->C1 < 1 || a->C2 < 0 || a->Cout < 1) return VTACO_ERR_INVALID_ARG;
  if (a->ksize != 1 && a->ksize != 3) return VTACO_ERR_UNSUPPORTED;
  const int Cin = a->C1 + a->C2;
  if (a->C1 % kCvKC || a->C2 % kCvKC || a->Cout % 32 || Cin > kCvMaxCin) return VTACO_ERR_UNSUPPORTED;
  if (a->C2 > 0 && (!a->x2 || a->D2 < 1 || a->H2 < 1 || a->W2 < 1)) return VTACO_ERR_INVALID_ARG;
  if (a->in_stats && (a->groups < 1 || Cin % a->groups)) return VTACO_ERR_INVALID_ARG;
  ConvParams P = {};
  P.x = a->x; P.x2 = a->x2; P.w = a->w_packed; P.bias = a->bias; P.in_stats = a->in_stats; P.gamma = a->gamma; P.beta = a->beta;
  P.y = a->y; P.out_stats = a->out_stats;
  P.N = a->N; P.D = a->D; P.H = a->H; P.W = a->W; P.C1 = a->C1; P.C2 = a->C2; P.D2 = a->D2; P.H2 = a->H2; P.W2 = a->W2;
  P.Cout = a->Cout; P.ksize = a->ksize; P.groups = a->groups; P.relu = a->relu ? 1 : 0; P.eps = a->eps;
  P.tiles_x = (a->W + kCvTileX - 1) / kCvTileX;
  P.tiles_y = (a->H + kCvTileY - 1) / kCvTileY;
  const CvSmem L = cv_layout(a->ksize, Cin);
  static std::atomic<int> configured[64];
  int dev = 0;
  VTACO_CUDA_CHECK(cudaGetDevice(&dev));
  if (configured[dev & 63].load(std::memory_order_relaxed) < L.total) {
    VTACO_CUDA_CHECK(cudaFuncSetAttribute(conv3d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    configured[dev & 63].store(113 * 1024, std::memory_order_relaxed);
  }
  if (L.total > 113 * 1024) return VTACO_ERR_UNSUPPORTED;
  const long long tiles = (long long)P.tiles_x * P.tiles_y * a->D;
  if (tiles > 0x7fffffffll || a->Cout / 32 > 65535 || a->N > 65535) return VTACO_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (unsigned)(a->Cout / 32), (unsigned)a->N);
  conv3d_tc_kernel<<<grid, kCvThreads, L.total, (cudaStream_t)stream>>>(P);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

extern "C" int vtaco_maxpool2_cl(const float* x, float* y, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                                 double* stats, void* stream) {
  if (!x || !y || N < 1 || D < 2 || H < 2 || W < 2 || C < 4 || (C & 3)) return VTACO_ERR_INVALID_ARG;
  if ((D | H | W) & 1) return VTACO_ERR_UNSUPPORTED;
  if (stats && N != 1) return VTACO_ERR_UNSUPPORTED;
  const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2) * (C / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  maxpool2_cl_kernel<<<(unsigned)blocks, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, y, N, D, H, W, C, stats);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

extern "C" int vtaco_channel_stats_cl(const float* x, int64_t S, int32_t C, double* stats, void* stream) {
  if (!x || !stats || S < 1 || C < 4 || (C & 3) || 256 % (C / 4)) return VTACO_ERR_INVALID_ARG;
  long long blocks = (S * (C / 4) + 255) / 256;
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  channel_stats_cl_kernel<<<(unsigned)blocks, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, S, C, stats);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}
