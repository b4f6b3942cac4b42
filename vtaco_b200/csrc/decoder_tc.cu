// Fused LocalDecoder on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as decoder.cu (reference src/conv_onet/models/decoder.py:71-161); selected
// with vtaco_decoder_args.variant == 2.  Why: ncu on the SIMT kernel shows the 32-wide
// contraction is pipe-bound (shared-memory wavefronts 71 %, FMA pipe 62 %), not latency
// bound (profiles/decoder_ncu_summary.json) — the condition the design brief sets for moving
// it to the tensor pipe.
//
// fp32 fidelity on a TF32 pipe: every operand is split x = hi + lo with hi = tf32(x) and the
// product is evaluated as hi*hi + lo*hi + hi*lo (3xTF32, fp32 accumulation in TMEM); the
// dropped lo*lo term and the truncation of lo are ~2^-22 relative.
//
// Structure (one persistent CTA per SM, 384 threads = 3 independent groups of 4 warps):
//   * a group owns a tile of 128 queries = the 128 TMEM lanes; thread t <-> query t <-> lane t;
//   * all 3*n_blocks weight matrices (hi and lo, UMMA canonical K-major, no swizzle) stay in
//     shared memory for the life of the CTA (120 KB); they are the B operands;
//   * activations are the A operands and live in TMEM (tcgen05.st from registers), so a layer
//     costs per thread: tcgen05.ld of 32 accumulator columns, bias/residual/ReLU/split
//     (~4 ALU ops per element), tcgen05.st of hi and lo — no shared-memory traffic at all;
//   * biases ride along as one more K-block: a constant (1,1,0,..) A block times a B block
//     whose rows 0/1 hold the bias hi/lo; fc_c[i+1](c) accumulates into the same TMEM
//     accumulator as fc_1 of block i, so a block costs 2 accumulator reads, not 3;
//   * one elected thread per group (the role rotates over the 4 warps) issues the tcgen05.mma
//     (M128 N32 K8, kind::tf32) of a step and commits to the group's mbarrier; the groups
//     interleave on the tensor pipe, so one group's ALU phase overlaps the others' MMA phases;
//   * the residual stream stays in registers; feature gather, fc_p, tips and fc_out run on the
//     CUDA cores exactly as in the SIMT kernel.
#include "decoder_tc_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace vtaco {


struct TcSmem {
  // byte offsets into dynamic shared memory
  int w, bias, small, tips, stage, bars, tmem_ptr, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int n_blocks) {
  TcSmem s;
  s.w = 0;                                          // 3*nb matrices x (hi 4 KB | lo 4 KB)
  s.bias = s.w + 3 * n_blocks * 8192;               // (2*nb+1) bias K-blocks of 1 KB
  s.small = s.bias + (2 * n_blocks + 1) * 1024 + 8192;   // (+ fc_p_img.weight[:, 3:] hi|lo) Wp[3][32], bp[32], Wout[64], bout[2]+pad
  s.tips = s.small + (128 + 68) * 4;
  s.stage = s.tips + VTACO_MAX_TIPS * 32 * 4;
  s.stage = (s.stage + 15) / 16 * 16;
  s.bars = s.stage + (kTcThreads / 32) * 32 * kStageStride * 4;
  s.tmem_ptr = s.bars + 64;
  s.total = s.tmem_ptr + 16;
  return s;
}

// MMAs of one accumulation step, issued by one thread.  D = [A_x * W] + ones * Bias [+ C * Wc]
__device__ __forceinline__ void issue_product(uint32_t d, uint32_t a_hi, uint32_t w_smem, uint32_t accumulate_first,
                                              int n_products = 3) {
  const uint32_t a_lo = a_hi + 32;
  if (n_products == 2) {   // mixed: BF16 correction (K = 64) then the TF32 main product
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) tc_mma_ts_bf16(d, a_lo + 8 * kk, make_bdesc(w_smem + 4096 + kk * 1024), kk > 0 ? 1u : accumulate_first);
    accumulate_first = 1;
  }
  if (n_products >= 3) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) tc_mma_ts(d, a_lo + 8 * kk, make_bdesc(w_smem + kk * 1024), kk > 0 ? 1u : accumulate_first);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) tc_mma_ts(d, a_hi + 8 * kk, make_bdesc(w_smem + 4096 + kk * 1024), 1);
    accumulate_first = 1;
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) tc_mma_ts(d, a_hi + 8 * kk, make_bdesc(w_smem + kk * 1024), kk > 0 ? 1u : accumulate_first);
}

// Debug-only phase tracing (VTACO_TC_TRACE=1): clock64 stamps of one warp of block 0, group 0.
#define TC_STAMP(slot)                                                            \
  do {                                                                            \
    if (trace && blockIdx.x == 0 && warp == 0 && lane == 0 && trace_n < 4096)     \
      trace[trace_n++] = ((long long)(slot) << 56) | (clock64() & 0x00ffffffffffffffll); \
  } while (0)

template <bool DENSE, bool MIXED>
__global__ void __launch_bounds__(kTcThreads, 1) decoder_tc_kernel(const __grid_constant__ DecParams P,
                                                                   const float* __restrict__ wtc,
                                                                   long long* __restrict__ trace) {
  int trace_n = 0;
  extern __shared__ __align__(1024) unsigned char tsm[];
  const TcSmem L = tc_smem_layout(P.n_blocks);
  float* sWtc = reinterpret_cast<float*>(tsm + L.w);
  float* sSmall = reinterpret_cast<float*>(tsm + L.small);
  float* sTip = reinterpret_cast<float*>(tsm + L.tips);
  float* sStage = reinterpret_cast<float*>(tsm + L.stage);
  uint64_t* sBars = reinterpret_cast<uint64_t*>(tsm + L.bars);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(tsm + L.tmem_ptr);

  // warp index through a shuffle: the compiler can then prove g / wg and everything derived
  // from them warp-uniform (uniform registers for the MMA descriptors, uniform branches)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(kFull, tid >> 5, 0);
  const int g = warp >> 2, wg = warp & 3, tg = tid & 127;
  const int nb = P.n_blocks;

  // ---- one-time setup: weights + bias blocks (contiguous in wtc), small vectors, barriers, TMEM ----
  const bool cimg = P.use_img && P.c_img;   // per-query tactile feature tensor (decoder.py:83-85)
  const int wtc_floats = 3 * nb * 2048 + (2 * nb + 1) * 256 + (cimg ? 2048 : 0);
  for (int i = tid; i < wtc_floats / 4; i += kTcThreads)
    reinterpret_cast<float4*>(sWtc)[i] = __ldg(reinterpret_cast<const float4*>(wtc) + i);
  for (int i = tid; i < 128; i += kTcThreads) sSmall[i] = P.weights[(P.use_img ? VTACO_DEC_OFF_WPI : VTACO_DEC_OFF_WP) + i];
  for (int i = tid; i < 68; i += kTcThreads)
    sSmall[128 + i] = P.weights[VTACO_DEC_OFF_BLOCKS + nb * VTACO_DEC_BLOCK_STRIDE + i];
  if (P.n_tips > 0) {
    for (int o = tid; o < P.n_tips * 32; o += kTcThreads) {
      const int f = o >> 5, j = o & 31;
      float a = 0.f;
      for (int k = 0; k < 32; ++k)
        a = fmaf(__ldg(P.weights + VTACO_DEC_OFF_WIMG + k * 32 + j), __ldg(P.tip_feat + f * 32 + k), a);
      sTip[o] = a;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < kTcGroups; ++i) mbar_init(smem_u32(sBars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *sTmem, 0);
  // group columns: C_hi 0, C_lo 32, X_hi 64, X_lo 96, ones 128 (8), D 136 (32)
  const uint32_t tbase = tmem_base + ((uint32_t)(32 * wg) << 16) + (uint32_t)(g * kColsPerGroup);
  const uint32_t tC = tbase, tX = tbase + 64, tOnes = tbase + 128, tD = tbase + 136;
  const uint32_t mbase = tmem_base + (uint32_t)(g * kColsPerGroup);
  const uint32_t mC = mbase, mX = mbase + 64, mOnes = mbase + 128, mD = mbase + 136;
  const uint32_t bar = smem_u32(sBars + g);
  const uint32_t wsm = smem_u32(sWtc), bsm = smem_u32(tsm + L.bias);
  const uint32_t wimg_sm = bsm + (uint32_t)(2 * nb + 1) * 1024u;
  uint32_t ph = 0;
  {  // the constant A block that multiplies the bias rows: columns (1, 1, 0, 0, 0, 0, 0, 0)
    const uint32_t one = __float_as_uint(1.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tOnes), "r"(one),
                 "r"(one), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                 : "memory");
    tc_wait_st();
  }
  float* stage = sStage + warp * 32 * kStageStride;
  const int grp = lane >> 3, sub = lane & 7;
  const int nx = P.nx;
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;
  int step = 0;  // accumulation steps issued so far by this group (rotates the issuing warp)
  constexpr bool mixed = MIXED;   // compile-time: the 3xTF32 kernel keeps its register budget

  for (long long tile = (long long)blockIdx.x * kTcGroups + g; tile < P.n_tiles;
       tile += (long long)gridDim.x * kTcGroups) {
    // ---------------- this thread's query ----------------
    float px, py, pz;
    long long oidx;
    int qb;
    bool valid;
    if (DENSE) {
      unsigned t = (unsigned)tile;                        // n_tiles < 2^31 (checked at launch)
      const int bz = (int)(t % (unsigned)P.t_nbz); t /= (unsigned)P.t_nbz;   // bricks: 2 (x) x 2 (y) x 32 (z): a warp = one
      const int by = (int)(t % (unsigned)P.t_nby); t /= (unsigned)P.t_nby;   // z-run, its 32 logits are one 128-byte store
      const int bx = (int)(t % (unsigned)P.t_nbx);
      const int b = (int)(t / (unsigned)P.t_nbx);
      const int ix = P.x0 + bx * 2 + (tg >> 6), iy = by * 2 + ((tg >> 5) & 1), iz = bz * 32 + (tg & 31);
      valid = (ix < P.t_xend) && (iy < nx) && (iz < nx);
      px = __ldg(P.axis + min(ix, nx - 1));
      py = __ldg(P.axis + min(iy, nx - 1));
      pz = __ldg(P.axis + min(iz, nx - 1));
      qb = b;
      oidx = (((long long)b * nx + ix) * nx + iy) * nx + iz;
    } else {
      const long long n = tile * kTcTile + tg;
      valid = n < P.total;
      const long long nn = valid ? n : 0;
      px = __ldg(P.p + nn * 3 + 0);
      py = __ldg(P.p + nn * 3 + 1);
      pz = __ldg(P.p + nn * 3 + 2);
      qb = (int)(nn / P.N);
      oidx = nn;
    }
    if (!valid) oidx = 0;
    TC_STAMP(13);  // tile start: coordinates loaded

    // ---------------- gather (+ the query's c_img row): operands of step 0 ----------------
    if (P.has_c || cimg) {
     if (P.has_c) {
      float cv[32];
      bool sep_done = false;
      if (DENSE && P.grid && !P.nearest && !(P.plane[0] || P.plane[1] || P.plane[2])) {
        // Separable dense gather: a warp is one z-run of 32 lattice points at fixed (x, y), so the
        // (x,y)-bilinear part of the trilinear sample is shared by the whole run.  Reduce the 4
        // (x,y) corners once per needed z-row into shared memory (4 z-rows x 8 channel quads per
        // load step), then every thread lerps its two z-rows: ~40 tap loads per run instead of 256.
        const int R = P.Rg;
        const float tx = unnormalize(norm3d(px, P.nc), R), ty = unnormalize(norm3d(py, P.nc), R);
        const float tz = unnormalize(norm3d(pz, P.nc), R);
        const float flx = floorf(tx), fly = floorf(ty), flz = floorf(tz);
        const int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
        const float fx1 = tx - flx, fx0 = (flx + 1.0f) - tx, fy1 = ty - fly, fy0 = (fly + 1.0f) - ty;
        const float fz1 = tz - flz, fz0 = (flz + 1.0f) - tz;
        const int zmin = __shfl_sync(kFull, z0, 0);
        const int zmax = min(__shfl_sync(kFull, z0, 31) + 1, R - 1);
        const int nz = zmax - zmin + 1;
        if (nz <= 32) {   // warp-uniform
          const int dx = (x0 + 1 < R) ? 8 : 0, dy = (y0 + 1 < R) ? R * 8 : 0;   // clamped corners carry weight 0
          const float w00 = fx0 * fy0, w01 = fx1 * fy0, w10 = fx0 * fy1, w11 = fx1 * fy1;
          const int zr = lane >> 3;
          const float4* col = reinterpret_cast<const float4*>(P.grid) + (size_t)qb * R * R * R * 8 +
                              ((size_t)y0 * R + x0) * 8 + sub;
#pragma unroll 2
          for (int zb = 0; zb < nz; zb += 4) {
            const int zrow = zb + zr;
            if (zrow < nz) {
              const float4* p = col + (size_t)(zmin + zrow) * R * R * 8;
              const float4 v00 = __ldg(p), v01 = __ldg(p + dx), v10 = __ldg(p + dy), v11 = __ldg(p + dy + dx);
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              a = f4_fma(w00, v00, a);
              a = f4_fma(w01, v01, a);
              a = f4_fma(w10, v10, a);
              a = f4_fma(w11, v11, a);
              *reinterpret_cast<float4*>(stage + zrow * kStageStride + 4 * sub) = a;
            }
          }
          __syncwarp();
          const float* r0 = stage + (z0 - zmin) * kStageStride;
          const float* r1 = stage + (min(z0 + 1, R - 1) - zmin) * kStageStride;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * j);
            const float4 b = *reinterpret_cast<const float4*>(r1 + 4 * j);
            cv[4 * j + 0] = fmaf(b.x, fz1, a.x * fz0);
            cv[4 * j + 1] = fmaf(b.y, fz1, a.y * fz0);
            cv[4 * j + 2] = fmaf(b.z, fz1, a.z * fz0);
            cv[4 * j + 3] = fmaf(b.w, fz1, a.w * fz0);
          }
          __syncwarp();
          sep_done = true;
        }
      }
      if (!sep_done) {
      // generic gather: the owner computes the taps, 8 lanes fetch them
      TapInfo tv, tp0, tp1, tp2;
      if (P.grid) tv = tap_volume(norm3d(px, P.nc), norm3d(py, P.nc), norm3d(pz, P.nc), P.Rg, P.nearest);
      const bool planes = P.plane[0] || P.plane[1] || P.plane[2];
      if (planes) {
        const float ux = norm2d(px, P.nc), uy = norm2d(py, P.nc), uz = norm2d(pz, P.nc);
        if (P.plane[0]) tp0 = tap_plane(ux, uz, P.Rp, P.nearest);
        if (P.plane[1]) tp1 = tap_plane(ux, uy, P.Rp, P.nearest);
        if (P.plane[2]) tp2 = tap_plane(uy, uz, P.Rp, P.nearest);
      }
      TC_STAMP(14);  // taps computed
#pragma unroll 2
      for (int it = 0; it < 8; ++it) {
        const int src = it * 4 + grp;
        const int b = __shfl_sync(kFull, qb, src);
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.grid) {
          const int R = P.Rg;
          const float4* vol = reinterpret_cast<const float4*>(P.grid) + (size_t)b * R * R * R * 8 + sub;
          c = fetch_volume(vol, R, tap_bcast(tv, src), P.nearest);
        }
        if (planes) {
          const int R = P.Rp;
          const size_t boff = (size_t)b * R * R * 8 + sub;
          if (P.plane[0]) c = f4_add(c, fetch_plane(reinterpret_cast<const float4*>(P.plane[0]) + boff, R, tap_bcast(tp0, src), P.nearest));
          if (P.plane[1]) c = f4_add(c, fetch_plane(reinterpret_cast<const float4*>(P.plane[1]) + boff, R, tap_bcast(tp1, src), P.nearest));
          if (P.plane[2]) c = f4_add(c, fetch_plane(reinterpret_cast<const float4*>(P.plane[2]) + boff, R, tap_bcast(tp2, src), P.nearest));
        }
        *reinterpret_cast<float4*>(stage + src * kStageStride + 4 * sub) = c;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(stage + lane * kStageStride + 4 * j);
        cv[4 * j] = v.x; cv[4 * j + 1] = v.y; cv[4 * j + 2] = v.z; cv[4 * j + 3] = v.w;
      }
      __syncwarp();
      }  // generic gather
      TC_STAMP(15);  // features of the thread's query in registers
      split_store(tC, cv, mixed);
     }
      if (cimg) {   // fc_p_img(cat[p, c_img]) = fc_p_img[:, :3] p + b + W_img c_img: the last term rides in step 0
        float xv[32];
        const float4* row = reinterpret_cast<const float4*>(P.c_img + (size_t)oidx * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = __ldg(row + j);
          xv[4 * j] = v.x; xv[4 * j + 1] = v.y; xv[4 * j + 2] = v.z; xv[4 * j + 3] = v.w;
        }
        split_store(tX, xv, mixed);
      }
      tc_wait_st();
      tc_fence_before();
      group_sync(g);
      TC_STAMP(16);  // C in TMEM, group synced
      if (wg == (step & 3) && elect_one()) {     // step 0: D = C*Wc_0 + ones*bc_0 [+ c_img*W_img]
        tc_fence_after();
        uint32_t acc = 0;
        if (P.has_c) {
          issue_product(mD, mC, wsm, 0, P.tc_products);
          tc_mma_ts(mD, mOnes, make_bdesc(bsm), 1);
          acc = 1;
        }
        if (cimg) issue_product(mD, mX, wimg_sm, acc, P.tc_products);
        tc_commit(bar);
      }
      ++step;
    }

    // ---------------- net = fc_p(p) | fc_p_img(p, tip feature) on the CUDA cores ----------------
    float net[32];
    {
      const float4* w0 = reinterpret_cast<const float4*>(sSmall);
      const float4* w1 = reinterpret_cast<const float4*>(sSmall + 32);
      const float4* w2 = reinterpret_cast<const float4*>(sSmall + 64);
      const float4* bp = reinterpret_cast<const float4*>(sSmall + 96);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a0 = w0[j], a1 = w1[j], a2 = w2[j], bb = bp[j];
        net[4 * j + 0] = fmaf(a2.x, pz, fmaf(a1.x, py, fmaf(a0.x, px, bb.x)));
        net[4 * j + 1] = fmaf(a2.y, pz, fmaf(a1.y, py, fmaf(a0.y, px, bb.y)));
        net[4 * j + 2] = fmaf(a2.z, pz, fmaf(a1.z, py, fmaf(a0.z, px, bb.z)));
        net[4 * j + 3] = fmaf(a2.w, pz, fmaf(a1.w, py, fmaf(a0.w, px, bb.w)));
      }
    }
    if (P.use_img && P.n_tips > 0) {
      const int f = tip_of_query(P, oidx, px, py, pz);
      if (f >= 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) net[j] += sTip[f * 32 + j];
      }
    }
    TC_STAMP(17);    // fc_p / tips done
    uint32_t r[32];
    float x[32];
    if (P.has_c || cimg) {  // net += fc_c[0](c) [+ W_img c_img]
      mbar_wait(bar, ph); ph ^= 1;
      tc_fence_after();
      tmem_ld32(tD, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) net[j] += __uint_as_float(r[j]);
    }

    TC_STAMP(18);    // step 0 consumed
    // ---------------- residual blocks: 2 accumulation steps each ----------------
    for (int i = 0; i < nb; ++i) {
      TC_STAMP(1);   // ALU phase starts (accumulator already read)
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaxf(net[j], 0.f);
      split_store(tX, x, mixed);
      TC_STAMP(2);   // operands computed, tcgen05.st issued
      tc_wait_st();
      tc_fence_before();
      TC_STAMP(3);   // stores complete
      group_sync(g);
      TC_STAMP(4);   // group barrier passed
      if (wg == (step & 3) && elect_one()) {     // D = relu(net)*W0_i + ones*b0_i
        tc_fence_after();
        issue_product(mD, mX, wsm + (3 * i + 1) * 8192, 0, P.tc_products);
        tc_mma_ts(mD, mOnes, make_bdesc(bsm + (2 * i + 1) * 1024), 1);
        tc_commit(bar);
      }
      ++step;
      TC_STAMP(5);   // (issuing warp: MMAs issued)
      mbar_wait(bar, ph); ph ^= 1;
      TC_STAMP(6);   // MMAs complete
      tc_fence_after();
      tmem_ld32(tD, r);
      tc_wait_ld();
      TC_STAMP(7);   // accumulator in registers
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaxf(__uint_as_float(r[j]), 0.f);
      split_store(tX, x, mixed);
      TC_STAMP(8 + 16 * (i & 1));    // W1 step: operands computed (+16: warp 0 is this step's issuer)
      tc_wait_st();
      tc_fence_before();
      group_sync(g);
      TC_STAMP(9 + 16 * (i & 1));
      if (wg == (step & 3) && elect_one()) {     // D = relu(h)*W1_i + ones*(b1_i + bc_{i+1}) [+ C*Wc_{i+1}]
        tc_fence_after();
        issue_product(mD, mX, wsm + (3 * i + 2) * 8192, 0, P.tc_products);
        tc_mma_ts(mD, mOnes, make_bdesc(bsm + (2 * i + 2) * 1024), 1);
        if (P.has_c && i + 1 < nb) issue_product(mD, mC, wsm + (3 * (i + 1)) * 8192, 1, P.tc_products);
        tc_commit(bar);
      }
      ++step;
      TC_STAMP(10 + 16 * (i & 1));
      mbar_wait(bar, ph); ph ^= 1;
      TC_STAMP(11 + 16 * (i & 1));
      tc_fence_after();
      tmem_ld32(tD, r);
      tc_wait_ld();
      TC_STAMP(12 + 16 * (i & 1));
#pragma unroll
      for (int j = 0; j < 32; ++j) net[j] += __uint_as_float(r[j]);
    }

    TC_STAMP(19);    // blocks done
    // ---------------- heads ----------------
    {
      const float* Wo = sSmall + 128;
      const float slope = P.leaky ? 0.2f : 0.0f;
      float o = Wo[64], oc = Wo[65];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        net[j] = net[j] > 0.f ? net[j] : net[j] * slope;
        o = fmaf(Wo[j], net[j], o);
      }
      if (P.contact) {
#pragma unroll
        for (int j = 0; j < 32; ++j) oc = fmaf(Wo[32 + j], net[j], oc);
      }
      if (valid) {
        store_logit(P, oidx, o);
        if (P.contact) P.contact[oidx] = oc;
        vmin = fminf(vmin, o);
        vmax = fmaxf(vmax, o);
      }
    }
  }

  if (P.minmax_key) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(kFull, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, d));
    }
    if (lane == 0 && vmin <= vmax) {
      atomicMin(P.minmax_key, float_to_key(vmin));
      atomicMax(P.minmax_key + 1, float_to_key(vmax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// =======================================================================================
// Two threads per query (variants 5 / 6): the same algorithm with 8 warps per 128-query tile.
// Warps w and w+4 of a group share a TMEM lane quarter and split the 32 channels 16 / 16, so
// every per-step phase (TMEM load, bias/residual/ReLU/split, TMEM store) is half as long per
// warp and each scheduler has 6 warps instead of 3 to interleave.  TMEM capacity (3 tiles per SM)
// is what rules out simply adding tiles.  768 threads -> 85 registers per thread.
// =======================================================================================
constexpr int kTc2Threads = 768;
constexpr int kTcMaxAxis = 2048;  // dense mode: lattice axes up to this length are kept in shared memory
constexpr int kTc2Stage = 20;   // floats per staged row (16 channels + pad)

// channels [16*hv, 16*hv+16) of x -> operand columns of the block at `tblk` (hi at +0, lo / correction at +32)
template <bool MIXED>
__device__ __forceinline__ void split_store16(uint32_t tblk, int hv, const float (&x)[16]) {
  uint32_t hi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) hi[j] = trunc_tf32(x[j]);
  tmem_st16(tblk + 16 * hv, hi);
  if (MIXED) {
    uint32_t lo8[8], hi8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      lo8[c] = pack_bf16(x[2 * c] - __uint_as_float(hi[2 * c]), x[2 * c + 1] - __uint_as_float(hi[2 * c + 1]));
      hi8[c] = pack_bf16(__uint_as_float(hi[2 * c]), __uint_as_float(hi[2 * c + 1]));
    }
    tmem_st8(tblk + 32 + 8 * hv, lo8);        // bf16 lo, k = 16*hv .. +15
    tmem_st8(tblk + 48 + 8 * hv, hi8);        // bf16 hi, k = 32 + 16*hv .. +15
  } else {
    uint32_t lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) lo[j] = __float_as_uint(x[j] - __uint_as_float(hi[j]));
    tmem_st16(tblk + 32 + 16 * hv, lo);
  }
}

struct Tc2Smem { int w, bias, small, tips, stage, head, axis, bars, tmem_ptr, total; };
__host__ __device__ inline Tc2Smem tc2_smem_layout(int n_blocks) {
  Tc2Smem s;
  s.w = 0;
  s.bias = s.w + 3 * n_blocks * 8192;
  s.small = s.bias + (2 * n_blocks + 1) * 1024 + 8192;   // bias blocks, then fc_p_img.weight[:, 3:] hi|lo
  s.tips = s.small + (128 + 68) * 4;
  s.stage = s.tips + VTACO_MAX_TIPS * 32 * 4;
  s.stage = (s.stage + 15) / 16 * 16;
  s.head = s.stage + (kTc2Threads / 32) * 32 * kTc2Stage * 4;   // per warp: 32 rows x 16 channels
  s.axis = s.head + kTcGroups * 128 * 2 * 4;                     // partial head sums of the upper halves
  s.bars = s.axis + kTcMaxAxis * 4;
  s.tmem_ptr = s.bars + 64;
  s.total = s.tmem_ptr + 16;
  return s;
}

template <bool DENSE, bool MIXED>
__global__ void __launch_bounds__(kTc2Threads, 1) decoder_tc2_kernel(const __grid_constant__ DecParams P,
                                                                     const float* __restrict__ wtc,
                                                                     long long* __restrict__ trace) {
  int trace_n = 0;
  extern __shared__ __align__(1024) unsigned char tsm[];
  const Tc2Smem L = tc2_smem_layout(P.n_blocks);
  float* sWtc = reinterpret_cast<float*>(tsm + L.w);
  float* sSmall = reinterpret_cast<float*>(tsm + L.small);
  float* sTip = reinterpret_cast<float*>(tsm + L.tips);
  float* sStage = reinterpret_cast<float*>(tsm + L.stage);
  float* sHead = reinterpret_cast<float*>(tsm + L.head);
  float* sAxis = reinterpret_cast<float*>(tsm + L.axis);
  uint64_t* sBars = reinterpret_cast<uint64_t*>(tsm + L.bars);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(tsm + L.tmem_ptr);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(kFull, tid >> 5, 0);
  const int g = warp >> 3, wq = warp & 7;   // group, warp within the group
  const int lq = wq & 3, hv = wq >> 2;      // TMEM lane quarter (== warp % 4), channel half
  const int tq = lq * 32 + lane;            // query within the tile == TMEM lane
  const int nb = P.n_blocks;
  const int nx = P.nx;

  const bool cimg = P.use_img && P.c_img;   // per-query tactile feature tensor (decoder.py:83-85)
  const int wtc_floats = 3 * nb * 2048 + (2 * nb + 1) * 256 + (cimg ? 2048 : 0);
  for (int i = tid; i < wtc_floats / 4; i += kTc2Threads)
    reinterpret_cast<float4*>(sWtc)[i] = __ldg(reinterpret_cast<const float4*>(wtc) + i);
  for (int i = tid; i < 128; i += kTc2Threads) sSmall[i] = P.weights[(P.use_img ? VTACO_DEC_OFF_WPI : VTACO_DEC_OFF_WP) + i];
  for (int i = tid; i < 68; i += kTc2Threads)
    sSmall[128 + i] = P.weights[VTACO_DEC_OFF_BLOCKS + nb * VTACO_DEC_BLOCK_STRIDE + i];
  const bool axis_sm = DENSE && nx <= kTcMaxAxis;
  if (axis_sm)
    for (int i = tid; i < nx; i += kTc2Threads) sAxis[i] = __ldg(P.axis + i);
  if (P.n_tips > 0) {
    for (int o = tid; o < P.n_tips * 32; o += kTc2Threads) {
      const int f = o >> 5, j = o & 31;
      float a = 0.f;
      for (int k = 0; k < 32; ++k)
        a = fmaf(__ldg(P.weights + VTACO_DEC_OFF_WIMG + k * 32 + j), __ldg(P.tip_feat + f * 32 + k), a);
      sTip[o] = a;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < kTcGroups; ++i) mbar_init(smem_u32(sBars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *sTmem, 0);
  const uint32_t tbase = tmem_base + ((uint32_t)(32 * lq) << 16) + (uint32_t)(g * kColsPerGroup);
  const uint32_t tC = tbase, tX = tbase + 64, tOnes = tbase + 128, tD = tbase + 136;
  const uint32_t mbase = tmem_base + (uint32_t)(g * kColsPerGroup);
  const uint32_t mC = mbase, mX = mbase + 64, mOnes = mbase + 128, mD = mbase + 136;
  const uint32_t bar = smem_u32(sBars + g);
  const uint32_t wsm = smem_u32(sWtc), bsm = smem_u32(tsm + L.bias);
  const uint32_t wimg_sm = bsm + (uint32_t)(2 * nb + 1) * 1024u;
  const int gsync_id = g + 1;
  auto group_sync2 = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(gsync_id) : "memory"); };
  uint32_t ph = 0;
  if (hv == 0) {
    const uint32_t one = __float_as_uint(1.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tOnes), "r"(one),
                 "r"(one), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                 : "memory");
    tc_wait_st();
  }
  float* stage = sStage + warp * 32 * kTc2Stage;
  float* head = sHead + g * 256;
  const int sub = lane & 3;          // float4 within the 16-channel half
  const int zr = lane >> 2;          // 8 rows / queries per load step
  const int ch0 = 16 * hv;           // first channel of this thread
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;
  int step = 0;
  const bool sep_cfg = DENSE && P.has_c && P.grid && !P.nearest && !(P.plane[0] || P.plane[1] || P.plane[2]);
  const int n_prod = MIXED ? 2 : 3;

  for (long long tile = (long long)blockIdx.x * kTcGroups + g; tile < P.n_tiles;
       tile += (long long)gridDim.x * kTcGroups) {
    float px, py, pz;
    long long oidx;
    int qb;
    bool valid;
    if (DENSE) {
      unsigned t = (unsigned)tile;
      const int bz = (int)(t % (unsigned)P.t_nbz); t /= (unsigned)P.t_nbz;
      const int by = (int)(t % (unsigned)P.t_nby); t /= (unsigned)P.t_nby;
      const int bx = (int)(t % (unsigned)P.t_nbx);
      qb = (int)(t / (unsigned)P.t_nbx);
      const int ix = P.x0 + bx * 2 + (tq >> 6), iy = by * 2 + ((tq >> 5) & 1), iz = bz * 32 + (tq & 31);
      valid = (ix < P.t_xend) && (iy < nx) && (iz < nx);
      const int cx = min(ix, nx - 1), cy = min(iy, nx - 1), cz = min(iz, nx - 1);
      px = axis_sm ? sAxis[cx] : __ldg(P.axis + cx);
      py = axis_sm ? sAxis[cy] : __ldg(P.axis + cy);
      pz = axis_sm ? sAxis[cz] : __ldg(P.axis + cz);
      oidx = (((long long)qb * nx + ix) * nx + iy) * nx + iz;
    } else {
      const long long n = tile * kTcTile + tq;
      valid = n < P.total;
      const long long nn = valid ? n : 0;
      px = __ldg(P.p + nn * 3 + 0);
      py = __ldg(P.p + nn * 3 + 1);
      pz = __ldg(P.p + nn * 3 + 2);
      qb = (int)(nn / P.N);
      oidx = nn;
    }
    if (!valid) oidx = 0;
    TC_STAMP(13);  // tile start: coordinates loaded

    // ---------------- gather: this thread's 16 channels of its query (+ of its c_img row): operands of step 0 ----------------
    if (P.has_c || cimg) {
     if (P.has_c) {
      float cv[16];
      bool sep_done = false;
      if (sep_cfg) {
        const int R = P.Rg;
        const float tx = unnormalize(norm3d(px, P.nc), R), ty = unnormalize(norm3d(py, P.nc), R);
        const float tz = unnormalize(norm3d(pz, P.nc), R);
        const float flx = floorf(tx), fly = floorf(ty), flz = floorf(tz);
        const int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
        const float fx1 = tx - flx, fx0 = (flx + 1.0f) - tx, fy1 = ty - fly, fy0 = (fly + 1.0f) - ty;
        const float fz1 = tz - flz, fz0 = (flz + 1.0f) - tz;
        const int zmin = __shfl_sync(kFull, z0, 0);
        const int zmax = min(__shfl_sync(kFull, z0, 31) + 1, R - 1);
        const int nz = zmax - zmin + 1;
        if (nz <= 32) {
          const int dx = (x0 + 1 < R) ? 8 : 0, dy = (y0 + 1 < R) ? R * 8 : 0;
          const float w00 = fx0 * fy0, w01 = fx1 * fy0, w10 = fx0 * fy1, w11 = fx1 * fy1;
          const float4* col = reinterpret_cast<const float4*>(P.grid) + (size_t)qb * R * R * R * 8 +
                              ((size_t)y0 * R + x0) * 8 + 4 * hv + sub;
#pragma unroll 2
          for (int zb = 0; zb < nz; zb += 8) {
            const int zrow = zb + zr;
            if (zrow < nz) {
              const float4* p = col + (size_t)(zmin + zrow) * R * R * 8;
              const float4 v00 = __ldg(p), v01 = __ldg(p + dx), v10 = __ldg(p + dy), v11 = __ldg(p + dy + dx);
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              a = f4_fma(w00, v00, a);
              a = f4_fma(w01, v01, a);
              a = f4_fma(w10, v10, a);
              a = f4_fma(w11, v11, a);
              *reinterpret_cast<float4*>(stage + zrow * kTc2Stage + 4 * sub) = a;
            }
          }
          __syncwarp();
          const float* r0 = stage + (z0 - zmin) * kTc2Stage;
          const float* r1 = stage + (min(z0 + 1, R - 1) - zmin) * kTc2Stage;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * j);
            const float4 b = *reinterpret_cast<const float4*>(r1 + 4 * j);
            cv[4 * j + 0] = fmaf(b.x, fz1, a.x * fz0);
            cv[4 * j + 1] = fmaf(b.y, fz1, a.y * fz0);
            cv[4 * j + 2] = fmaf(b.z, fz1, a.z * fz0);
            cv[4 * j + 3] = fmaf(b.w, fz1, a.w * fz0);
          }
          __syncwarp();
          sep_done = true;
        }
      }
      if (!sep_done) {   // generic gather: the owner computes the taps, 4 lanes fetch a query's 16 channels
        TapInfo tv, tp0, tp1, tp2;
        if (P.grid) tv = tap_volume(norm3d(px, P.nc), norm3d(py, P.nc), norm3d(pz, P.nc), P.Rg, P.nearest);
        const bool planes = P.plane[0] || P.plane[1] || P.plane[2];
        if (planes) {
          const float ux = norm2d(px, P.nc), uy = norm2d(py, P.nc), uz = norm2d(pz, P.nc);
          if (P.plane[0]) tp0 = tap_plane(ux, uz, P.Rp, P.nearest);
          if (P.plane[1]) tp1 = tap_plane(ux, uy, P.Rp, P.nearest);
          if (P.plane[2]) tp2 = tap_plane(uy, uz, P.Rp, P.nearest);
        }
#pragma unroll 2
        for (int it = 0; it < 4; ++it) {
          const int src = it * 8 + zr;
          const int b = __shfl_sync(kFull, qb, src);
          float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.grid) {
            const int R = P.Rg;
            const float4* vol = reinterpret_cast<const float4*>(P.grid) + (size_t)b * R * R * R * 8 + 4 * hv + sub;
            c = fetch_volume(vol, R, tap_bcast(tv, src), P.nearest);
          }
          if (planes) {
            const int R = P.Rp;
            const size_t boff = (size_t)b * R * R * 8 + 4 * hv + sub;
            if (P.plane[0]) c = f4_add(c, fetch_plane(reinterpret_cast<const float4*>(P.plane[0]) + boff, R, tap_bcast(tp0, src), P.nearest));
            if (P.plane[1]) c = f4_add(c, fetch_plane(reinterpret_cast<const float4*>(P.plane[1]) + boff, R, tap_bcast(tp1, src), P.nearest));
            if (P.plane[2]) c = f4_add(c, fetch_plane(reinterpret_cast<const float4*>(P.plane[2]) + boff, R, tap_bcast(tp2, src), P.nearest));
          }
          *reinterpret_cast<float4*>(stage + src * kTc2Stage + 4 * sub) = c;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(stage + lane * kTc2Stage + 4 * j);
          cv[4 * j] = v.x; cv[4 * j + 1] = v.y; cv[4 * j + 2] = v.z; cv[4 * j + 3] = v.w;
        }
        __syncwarp();
      }
      TC_STAMP(15);  // features of the thread's query in registers
      split_store16<MIXED>(tC, hv, cv);
     }
      if (cimg) {   // fc_p_img(cat[p, c_img]) = fc_p_img[:, :3] p + b + W_img c_img: the last term rides in step 0
        float xv[16];
        const float4* row = reinterpret_cast<const float4*>(P.c_img + (size_t)oidx * 32 + ch0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = __ldg(row + j);
          xv[4 * j] = v.x; xv[4 * j + 1] = v.y; xv[4 * j + 2] = v.z; xv[4 * j + 3] = v.w;
        }
        split_store16<MIXED>(tX, hv, xv);
      }
      tc_wait_st();
      tc_fence_before();
      group_sync2();
      if (wq == (step & 7) && elect_one()) {     // step 0: D = C*Wc_0 + ones*bc_0 [+ c_img*W_img]
        tc_fence_after();
        uint32_t acc = 0;
        if (P.has_c) {
          issue_product(mD, mC, wsm, 0, n_prod);
          tc_mma_ts(mD, mOnes, make_bdesc(bsm), 1);
          acc = 1;
        }
        if (cimg) issue_product(mD, mX, wimg_sm, acc, n_prod);
        tc_commit(bar);
      }
      ++step;
    }

    // ---------------- net = fc_p(p) | fc_p_img(p, tip feature): this thread's 16 channels ----------------
    float net[16];
    {
      const float4* w0 = reinterpret_cast<const float4*>(sSmall + ch0);
      const float4* w1 = reinterpret_cast<const float4*>(sSmall + 32 + ch0);
      const float4* w2 = reinterpret_cast<const float4*>(sSmall + 64 + ch0);
      const float4* bp = reinterpret_cast<const float4*>(sSmall + 96 + ch0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a0 = w0[j], a1 = w1[j], a2 = w2[j], bb = bp[j];
        net[4 * j + 0] = fmaf(a2.x, pz, fmaf(a1.x, py, fmaf(a0.x, px, bb.x)));
        net[4 * j + 1] = fmaf(a2.y, pz, fmaf(a1.y, py, fmaf(a0.y, px, bb.y)));
        net[4 * j + 2] = fmaf(a2.z, pz, fmaf(a1.z, py, fmaf(a0.z, px, bb.z)));
        net[4 * j + 3] = fmaf(a2.w, pz, fmaf(a1.w, py, fmaf(a0.w, px, bb.w)));
      }
    }
    if (P.use_img && P.n_tips > 0) {
      const int f = tip_of_query(P, oidx, px, py, pz);
      if (f >= 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) net[j] += sTip[f * 32 + ch0 + j];
      }
    }
    uint32_t r[16];
    float x[16];
    if (P.has_c || cimg) {  // net += fc_c[0](c) [+ W_img c_img]
      mbar_wait(bar, ph); ph ^= 1;
      tc_fence_after();
      tmem_ld16(tD + ch0, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) net[j] += __uint_as_float(r[j]);
    }

    // ---------------- residual blocks ----------------
    for (int i = 0; i < nb; ++i) {
      TC_STAMP(1);   // ALU phase starts (accumulator already read)
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = fmaxf(net[j], 0.f);
      split_store16<MIXED>(tX, hv, x);
      TC_STAMP(2);   // operands computed, tcgen05.st issued
      tc_wait_st();
      tc_fence_before();
      TC_STAMP(3);   // stores complete
      group_sync2();
      TC_STAMP(4);   // group barrier passed
      if (wq == (step & 7) && elect_one()) {     // D = relu(net)*W0_i + ones*b0_i
        tc_fence_after();
        issue_product(mD, mX, wsm + (3 * i + 1) * 8192, 0, n_prod);
        tc_mma_ts(mD, mOnes, make_bdesc(bsm + (2 * i + 1) * 1024), 1);
        tc_commit(bar);
        TC_STAMP(20);  // issuer only: 13 MMAs + commit issued
      }
      ++step;
      TC_STAMP(5);   // (issuing warp: MMAs issued)
      mbar_wait(bar, ph); ph ^= 1;
      TC_STAMP(6);   // MMAs complete
      tc_fence_after();
      tmem_ld16(tD + ch0, r);
      tc_wait_ld();
      TC_STAMP(7);   // accumulator in registers
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = fmaxf(__uint_as_float(r[j]), 0.f);
      split_store16<MIXED>(tX, hv, x);
      tc_wait_st();
      tc_fence_before();
      group_sync2();
      if (wq == (step & 7) && elect_one()) {     // D = relu(h)*W1_i + ones*(b1_i + bc_{i+1}) [+ C*Wc_{i+1}]
        tc_fence_after();
        issue_product(mD, mX, wsm + (3 * i + 2) * 8192, 0, n_prod);
        tc_mma_ts(mD, mOnes, make_bdesc(bsm + (2 * i + 2) * 1024), 1);
        if (P.has_c && i + 1 < nb) issue_product(mD, mC, wsm + (3 * (i + 1)) * 8192, 1, n_prod);
        tc_commit(bar);
      }
      ++step;
      mbar_wait(bar, ph); ph ^= 1;
      tc_fence_after();
      tmem_ld16(tD + ch0, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) net[j] += __uint_as_float(r[j]);
    }

    TC_STAMP(19);    // blocks done
    // ---------------- heads: partial dot products of the two halves, combined through shared memory ----------------
    {
      const float* Wo = sSmall + 128;
      const float slope = P.leaky ? 0.2f : 0.0f;
      float o = 0.f, oc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        net[j] = net[j] > 0.f ? net[j] : net[j] * slope;
        o = fmaf(Wo[ch0 + j], net[j], o);
      }
      if (P.contact) {
#pragma unroll
        for (int j = 0; j < 16; ++j) oc = fmaf(Wo[32 + ch0 + j], net[j], oc);
      }
      if (hv == 1) { head[tq] = o; head[128 + tq] = oc; }
      group_sync2();
      if (hv == 0 && valid) {
        o = (Wo[64] + o) + head[tq];
        store_logit(P, oidx, o);
        if (P.contact) P.contact[oidx] = (Wo[65] + oc) + head[128 + tq];
        vmin = fminf(vmin, o);
        vmax = fmaxf(vmax, o);
      }
    }
  }

  if (P.minmax_key) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(kFull, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(kFull, vmax, d));
    }
    if (lane == 0 && vmin <= vmax) {
      atomicMin(P.minmax_key, float_to_key(vmin));
      atomicMax(P.minmax_key + 1, float_to_key(vmax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int launch_decoder_tc(DecParams P, bool dense, const float* wtc, cudaStream_t stream) {
  if (!wtc) return VTACO_ERR_INVALID_ARG;
  if (dense) {
    P.t_xend = P.x1;
    P.t_nbz = (P.nx + 31) / 32;
    P.t_nby = (P.nx + 1) / 2;
    P.t_nbx = (P.x1 - P.x0 + 1) / 2;
    P.n_tiles = (long long)P.B * P.t_nbx * P.t_nby * P.t_nbz;
  } else {
    P.n_tiles = (P.total + kTcTile - 1) / kTcTile;
  }
  if (P.n_tiles >= (1ll << 31)) return VTACO_ERR_UNSUPPORTED;
  const bool mixed = (P.tc_products == 2);
  if (P.tc_split == 2) {   // two threads per query
    const Tc2Smem L2 = tc2_smem_layout(P.n_blocks);
    if (L2.total > 227 * 1024) return VTACO_ERR_UNSUPPORTED;
    using Kernel2 = void (*)(DecParams, const float*, long long*);
    const Kernel2 k2 = dense ? (mixed ? (Kernel2)decoder_tc2_kernel<true, true> : (Kernel2)decoder_tc2_kernel<true, false>)
                             : (mixed ? (Kernel2)decoder_tc2_kernel<false, true> : (Kernel2)decoder_tc2_kernel<false, false>);
    static std::atomic<size_t> configured2[4][64];   // idempotent opt-in cache, safe across host threads
    int dev2 = 0;
    VTACO_CUDA_CHECK(cudaGetDevice(&dev2));
    const int ki2 = (dense ? 2 : 0) + (mixed ? 1 : 0);
    if (configured2[ki2][dev2 & 63].load(std::memory_order_relaxed) < (size_t)L2.total) {
      VTACO_CUDA_CHECK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, L2.total));
      configured2[ki2][dev2 & 63].store(L2.total, std::memory_order_relaxed);
    }
    long long grid2 = (P.n_tiles + kTcGroups - 1) / kTcGroups;
    if (grid2 > num_sms()) grid2 = num_sms();
    long long* trace2 = nullptr;
    static const bool want_trace2 = getenv("VTACO_TC_TRACE") != nullptr;
    if (want_trace2) {
      VTACO_CUDA_CHECK(cudaMalloc(&trace2, 4096 * sizeof(long long)));
      VTACO_CUDA_CHECK(cudaMemsetAsync(trace2, 0, 4096 * sizeof(long long), stream));
    }
    k2<<<(unsigned)grid2, kTc2Threads, L2.total, stream>>>(P, wtc, trace2);
    VTACO_LAUNCH_CHECK();
    if (trace2) {   // debug: average cycles between consecutive stamps of warp 0 / block 0, per (from -> to) slot pair
      static long long h2[4096];
      VTACO_CUDA_CHECK(cudaStreamSynchronize(stream));
      VTACO_CUDA_CHECK(cudaMemcpy(h2, trace2, sizeof(h2), cudaMemcpyDeviceToHost));
      cudaFree(trace2);
      static double sum2[32][32];
      static long long cnt2[32][32];
      for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) { sum2[a][b] = 0; cnt2[a][b] = 0; }
      for (int i = 1; i < 4096 && h2[i]; ++i) {
        const int a = (int)(h2[i - 1] >> 56) & 31, b = (int)(h2[i] >> 56) & 31;
        const long long d = (h2[i] & 0x00ffffffffffffffll) - (h2[i - 1] & 0x00ffffffffffffffll);
        if (d >= 0 && d < 1000000) { sum2[a][b] += (double)d; cnt2[a][b]++; }
      }
      for (int a = 0; a < 32; ++a)
        for (int b = 0; b < 32; ++b)
          if (cnt2[a][b]) fprintf(stderr, "[vtaco tc2 trace] %d -> %d : %8.0f cycles (n=%lld)\n", a, b, sum2[a][b] / cnt2[a][b], cnt2[a][b]);
    }
    return VTACO_OK;
  }
  const TcSmem L = tc_smem_layout(P.n_blocks);
  if (L.total > 227 * 1024) return VTACO_ERR_UNSUPPORTED;
  using Kernel = void (*)(DecParams, const float*, long long*);
  const Kernel kernel = dense ? (mixed ? (Kernel)decoder_tc_kernel<true, true> : (Kernel)decoder_tc_kernel<true, false>)
                              : (mixed ? (Kernel)decoder_tc_kernel<false, true> : (Kernel)decoder_tc_kernel<false, false>);
  static std::atomic<size_t> configured[4][64];
  int dev = 0;
  VTACO_CUDA_CHECK(cudaGetDevice(&dev));
  const int ki = (dense ? 2 : 0) + (mixed ? 1 : 0);
  if (configured[ki][dev & 63].load(std::memory_order_relaxed) < (size_t)L.total) {
    VTACO_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured[ki][dev & 63].store(L.total, std::memory_order_relaxed);
  }
  long long grid = (P.n_tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  long long* trace = nullptr;
  static const bool want_trace = getenv("VTACO_TC_TRACE") != nullptr;
  if (want_trace) {
    VTACO_CUDA_CHECK(cudaMalloc(&trace, 4096 * sizeof(long long)));
    VTACO_CUDA_CHECK(cudaMemsetAsync(trace, 0, 4096 * sizeof(long long), stream));
  }
  kernel<<<(unsigned)grid, kTcThreads, L.total, stream>>>(P, wtc, trace);
  VTACO_LAUNCH_CHECK();
  if (trace) {   // debug: average cycles between consecutive stamps, per (from -> to) slot pair
    static long long h[4096];
    VTACO_CUDA_CHECK(cudaStreamSynchronize(stream));
    VTACO_CUDA_CHECK(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    static double sum[32][32];
    static long long cnt[32][32];
    for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) { sum[a][b] = 0; cnt[a][b] = 0; }
    for (int i = 1; i < 4096 && h[i]; ++i) {
      const int a = (int)(h[i - 1] >> 56) & 31, b = (int)(h[i] >> 56) & 31;
      const long long d = (h[i] & 0x00ffffffffffffffll) - (h[i - 1] & 0x00ffffffffffffffll);
      if (d >= 0 && d < 1000000) { sum[a][b] += (double)d; cnt[a][b]++; }
    }
    for (int a = 0; a < 32; ++a)
      for (int b = 0; b < 32; ++b)
        if (cnt[a][b]) fprintf(stderr, "[vtaco tc trace] %d -> %d : %8.0f cycles (n=%lld)\n", a, b, sum[a][b] / cnt[a][b], cnt[a][b]);
  }
  return VTACO_OK;
}

}  // namespace vtaco
