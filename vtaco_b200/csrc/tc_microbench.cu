// Debug micro-benchmark (not part of the product path): latency / throughput of small
// tcgen05.mma kind::tf32 instructions (M=128, K=8) as used by decoder_tc.cu.
#include "common.cuh"
#include <cstdio>

namespace vtaco {
__device__ __forceinline__ uint32_t mb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mb_bdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(512 >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) tc_microbench_kernel(long long* out, int n_rounds) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  float* w = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < 8192; i += 128) w[i] = 0.001f * (i & 63);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mb_smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  uint32_t phase = 0;
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    long long t0 = 0, t1 = 0, t2 = 0;
    uint32_t elected = 0;
    if (warp == 0) {
      asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(elected));
    }
    if (warp == 0 && elected) {
      t0 = clock64();
      const uint32_t ws = mb_smem_u32(w);
      for (int r = 0; r < n_rounds; ++r) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const uint32_t d = tb + 256 + (uint32_t)((i % NACC) * N);   // accumulators at columns 256..
          const uint32_t a = tb + (uint32_t)((i & 7) * 8);            // A operand columns 0..63 (garbage values)
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
              "l"(mb_bdesc(ws + (i & 3) * 1024)), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb_smem_u32(&bar)) : "memory");
      asm volatile(
          "{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(mb_smem_u32(&bar)), "r"(phase)
          : "memory");
      t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
}  // namespace vtaco

template <int N, int NACC>
static int run_mb(long long* d, int n_rounds) {
  VTACO_CUDA_CHECK(cudaFuncSetAttribute(vtaco::tc_microbench_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  vtaco::tc_microbench_kernel<N, NACC><<<1, 128, 65536>>>(d, n_rounds);
  return VTACO_OK;
}

extern "C" int vtaco_tc_microbench(int n_rounds, int n_acc, int N, long long* out_host) {
  long long* d = nullptr;
  VTACO_CUDA_CHECK(cudaMalloc(&d, 6 * sizeof(long long)));
  int st = VTACO_ERR_INVALID_ARG;
  if (N == 32 && n_acc == 1) st = run_mb<32, 1>(d, n_rounds);
  else if (N == 32 && n_acc == 4) st = run_mb<32, 4>(d, n_rounds);
  else if (N == 64 && n_acc == 1) st = run_mb<64, 1>(d, n_rounds);
  else if (N == 256 && n_acc == 1) st = run_mb<256, 1>(d, n_rounds);
  if (st != VTACO_OK) return st;
  VTACO_CUDA_CHECK(cudaDeviceSynchronize());
  VTACO_CUDA_CHECK(cudaMemcpy(out_host, d, 6 * sizeof(long long), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return VTACO_OK;
}
