// tcgen05 / TMEM PTX wrappers and the gather helpers shared by the tensor-core decoder kernels
// (decoder_tc.cu: 3 tiles per SM, 3xTF32; decoder_tc4.cu: 4 tiles per SM, TF32 + BF16 residual).
#pragma once
#include "decoder_common.cuh"
#include <cuda_bf16.h>

namespace vtaco {

constexpr int kTcThreads = 384;
constexpr int kTcGroups = 3;
constexpr int kTcTile = 128;
constexpr int kColsPerGroup = 168;   // C_hi 0, C_lo 32, X_hi 64, X_lo 96, ones 128 (8), D 136 (32)
constexpr int kStageStride = 36;     // floats per staged query row (conflict-free LDS.128)
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (4u << 17) | (8u << 24);  // F32 acc, TF32 x TF32, K-major, N=32, M=128
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | (4u << 17) | (8u << 24);  // F32 acc, BF16 x BF16 (kind::f16, K=16)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// exactly one lane of a converged warp; unlike `lane == 0` the compiler knows a single lane is
// active, so tcgen05.mma sequences compile to back-to-back UTCHMMA without per-instruction
// ELECT / retry loops (measured: ~45 -> ~16 cycles of issue per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(kIdescTf32), "r"(accumulate)
      : "memory");
}
// same with BF16 operands (two per 32-bit TMEM column, K = 16 per instruction)
__device__ __forceinline__ void tc_mma_ts_bf16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(kIdescBf16), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; K-chunk stride 512 B, 8-row-group stride 128 B
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(512 >> 4) << 16;   // leading (K) byte offset
  d |= (uint64_t)(128 >> 4) << 32;   // stride (N) byte offset
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  return d;
}

#define TC_R32(r) r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], \
                  r[16], r[17], r[18], r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}

// Activation split: hi = x with the 13 low mantissa bits cleared (exactly what the tensor core
// keeps of an fp32 container), lo = x - hi (exact).  cvt.rna.tf32 would cost 4 SASS ops per
// element (it is emulated); truncation is 1 LOP3 and leaves |lo| <= 2^-10 |x|, whose own
// truncation to TF32 is a 2^-21 relative error — far inside the 1e-4 parity bar.
__device__ __forceinline__ uint32_t trunc_tf32(float x) { return __float_as_uint(x) & 0xffffe000u; }

__device__ __forceinline__ uint32_t pack_bf16(float k_even, float k_odd) {   // element 2c in the low half
  const __nv_bfloat162 v = __floats2bfloat162_rn(k_even, k_odd);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// x[32] (fp32) -> operands stored to TMEM columns [col, col+32) and [col+32, col+64):
//   3xTF32 : hi | lo (both TF32 in fp32 containers)
//   mixed  : hi (TF32) | correction operand in BF16, K = 64: bf16(lo[0..31]) then bf16(hi[0..31]),
//            two elements per column.  The main product hi*W_hi stays TF32; lo*W and hi*W_lo are
//            ~2^-11 of it, so BF16's 2^-9 relative rounding leaves a ~2^-19 relative error.
__device__ __forceinline__ void split_store(uint32_t taddr, const float (&x)[32], bool mixed) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) hi[j] = trunc_tf32(x[j]);
  if (mixed) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      lo[c] = pack_bf16(x[2 * c] - __uint_as_float(hi[2 * c]), x[2 * c + 1] - __uint_as_float(hi[2 * c + 1]));
      lo[16 + c] = pack_bf16(__uint_as_float(hi[2 * c]), __uint_as_float(hi[2 * c + 1]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) lo[j] = __float_as_uint(x[j] - __uint_as_float(hi[j]));
  }
  tmem_st32(taddr, hi);
  tmem_st32(taddr + 32, lo);
}

// ---------------------------------------------------------------------------------------
// Interpolation set-up computed ONCE per query by its owner thread and broadcast to the 8
// lanes that fetch the taps (the SIMT kernel recomputes it in every lane).
// ---------------------------------------------------------------------------------------
struct TapInfo {
  int base;       // index (in taps) of corner (x0,y0[,z0])
  int step;       // bit0: x0+1 < R, bit1: y0+1 < R, bit2: z0+1 < R   (else the weight is 0 and the address is clamped)
  float fx, fy, fz;
};
__device__ __forceinline__ TapInfo tap_volume(float ux, float uy, float uz, int R, bool nearest) {
  const float tx = unnormalize(ux, R), ty = unnormalize(uy, R), tz = unnormalize(uz, R);
  TapInfo t;
  if (nearest) {
    const int x = (int)nearbyintf(tx), y = (int)nearbyintf(ty), z = (int)nearbyintf(tz);
    t.base = (z * R + y) * R + x; t.step = 0; t.fx = t.fy = t.fz = 0.f;
    return t;
  }
  const float flx = floorf(tx), fly = floorf(ty), flz = floorf(tz);
  const int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
  t.base = (z0 * R + y0) * R + x0;
  t.step = (int)(x0 + 1 < R) | ((int)(y0 + 1 < R) << 1) | ((int)(z0 + 1 < R) << 2);
  t.fx = tx - flx; t.fy = ty - fly; t.fz = tz - flz;
  return t;
}
__device__ __forceinline__ TapInfo tap_plane(float ua, float ub, int R, bool nearest) {
  const float tx = unnormalize(ua, R), ty = unnormalize(ub, R);
  TapInfo t;
  t.fz = 0.f;
  if (nearest) {
    const int x = (int)nearbyintf(tx), y = (int)nearbyintf(ty);
    t.base = y * R + x; t.step = 0; t.fx = t.fy = 0.f;
    return t;
  }
  const float flx = floorf(tx), fly = floorf(ty);
  const int x0 = (int)flx, y0 = (int)fly;
  t.base = y0 * R + x0;
  t.step = (int)(x0 + 1 < R) | ((int)(y0 + 1 < R) << 1);
  t.fx = tx - flx; t.fy = ty - fly;
  return t;
}
__device__ __forceinline__ TapInfo tap_bcast(const TapInfo& t, int src) {
  TapInfo r;
  r.base = __shfl_sync(kFull, t.base, src);
  r.step = __shfl_sync(kFull, t.step, src);
  r.fx = __shfl_sync(kFull, t.fx, src);
  r.fy = __shfl_sync(kFull, t.fy, src);
  r.fz = __shfl_sync(kFull, t.fz, src);
  return r;
}
// ATen corner order / weights (see sample_volume in decoder_common.cuh); the weight of a
// clamped corner is exactly 0 because its fraction is 0.
__device__ __forceinline__ float4 fetch_volume(const float4* __restrict__ vol, int R, const TapInfo& t, bool nearest) {
  if (nearest) return __ldg(vol + (size_t)t.base * 8);
  const int dx = (t.step & 1) ? 8 : 0, dy = (t.step & 2) ? R * 8 : 0, dz = (t.step & 4) ? R * R * 8 : 0;
  const float4* p = vol + (size_t)t.base * 8;
  const float4 v000 = __ldg(p), v001 = __ldg(p + dx), v010 = __ldg(p + dy), v011 = __ldg(p + dy + dx);
  const float4 v100 = __ldg(p + dz), v101 = __ldg(p + dz + dx), v110 = __ldg(p + dz + dy), v111 = __ldg(p + dz + dy + dx);
  const float fx1 = (t.step & 1) ? t.fx : 0.f, fy1 = (t.step & 2) ? t.fy : 0.f, fz1 = (t.step & 4) ? t.fz : 0.f;
  const float fx0 = 1.0f - t.fx, fy0 = 1.0f - t.fy, fz0 = 1.0f - t.fz;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  a = f4_fma(fx0 * fy0 * fz0, v000, a);
  a = f4_fma(fx1 * fy0 * fz0, v001, a);
  a = f4_fma(fx0 * fy1 * fz0, v010, a);
  a = f4_fma(fx1 * fy1 * fz0, v011, a);
  a = f4_fma(fx0 * fy0 * fz1, v100, a);
  a = f4_fma(fx1 * fy0 * fz1, v101, a);
  a = f4_fma(fx0 * fy1 * fz1, v110, a);
  a = f4_fma(fx1 * fy1 * fz1, v111, a);
  return a;
}
__device__ __forceinline__ float4 fetch_plane(const float4* __restrict__ pl, int R, const TapInfo& t, bool nearest) {
  if (nearest) return __ldg(pl + (size_t)t.base * 8);
  const int dx = (t.step & 1) ? 8 : 0, dy = (t.step & 2) ? R * 8 : 0;
  const float4* p = pl + (size_t)t.base * 8;
  const float4 v00 = __ldg(p), v01 = __ldg(p + dx), v10 = __ldg(p + dy), v11 = __ldg(p + dy + dx);
  const float fx1 = (t.step & 1) ? t.fx : 0.f, fy1 = (t.step & 2) ? t.fy : 0.f;
  const float fx0 = 1.0f - t.fx, fy0 = 1.0f - t.fy;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  a = f4_fma(fx0 * fy0, v00, a);
  a = f4_fma(fx1 * fy0, v01, a);
  a = f4_fma(fx0 * fy1, v10, a);
  a = f4_fma(fx1 * fy1, v11, a);
  return a;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(taddr)
               : "memory");
}

}  // namespace vtaco
