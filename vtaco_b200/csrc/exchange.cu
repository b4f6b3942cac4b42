// Multi-GPU exchange steps of the sharded extraction, as device-side signalling over
// peer-mapped memory (NVLink / NVSwitch): no host synchronisation, no collective launch, CUDA-graph
// capturable.  The lattice is split into x-slabs; each rank decodes its slab (+ 2 halo rows) and
// extracts its piece of the mesh; the only data that crosses the fabric is
//   (1) one (min,max) key pair per rank  -> the global iso-level 0.5*(min+max) (generation.py:270,
//       skimage level=None),
//   (2) one (V,F) pair per rank           -> exclusive scan = vertex / face base of each piece,
//   (3) the mesh pieces themselves (12 B per vertex / face, faces rebased on the way out).
// Every rank owns a control block in symmetric memory; writers store data, fence, then store a
// sequence number; readers spin on the sequence number with acquire loads.  Tables are double-
// buffered by sequence parity: a rank can be at most one step ahead of any peer because each
// step's waits need every peer's contribution to that step.
#include "common.cuh"

namespace vtaco {

struct ExCtrl {
  int32_t keys[2][8][2];
  uint32_t key_flag[2][8];
  long long cnt[2][8][2];
  uint32_t cnt_flag[2][8];
  uint32_t done_flag[8];
  // ---- local state (never written by peers) ----
  uint32_t seq_keys, seq_cnt, seq_done;
  uint32_t err;          // 1: a wait timed out (a peer never arrived)
  float level;
  float pad_;
  long long base[2];     // vertex / face base of this rank's piece
  long long total[2];    // mesh totals
  uint32_t ticket;       // blocks of the mesh-exchange kernel that have finished (the last one signals, then resets it)
};
static_assert(sizeof(ExCtrl) <= VTACO_EXCHANGE_CTRL_BYTES, "control block too large");
static_assert(offsetof(ExCtrl, level) == VTACO_EXCHANGE_LEVEL_OFFSET, "level offset");
static_assert(offsetof(ExCtrl, err) == VTACO_EXCHANGE_ERR_OFFSET, "err offset");
static_assert(offsetof(ExCtrl, base) == VTACO_EXCHANGE_BASE_OFFSET, "base offset");

struct ExPeers { ExCtrl* c[8]; int world, rank; };

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq (wrap-safe); false after ~2 s
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t seq) {
  const unsigned long long t0 = global_ns();
  while ((int32_t)(ld_acquire_sys(flag) - seq) < 0) {
    if (global_ns() - t0 > 2000000000ull) return false;
    __nanosleep(40);
  }
  return true;
}

// (1) publish this rank's (min,max) keys to every rank, wait for everyone's, write the level.
__global__ void __launch_bounds__(32) exchange_level_kernel(ExPeers X, int32_t* keys) {
  ExCtrl* me = X.c[X.rank];
  const int lane = threadIdx.x;
  const uint32_t seq = me->seq_keys + 1;
  const int par = seq & 1;
  const int32_t lo = keys[0], hi = keys[1];
  __syncwarp();
  if (lane < X.world) {
    ExCtrl* peer = X.c[lane];
    volatile int32_t* k = peer->keys[par][X.rank];
    k[0] = lo;
    k[1] = hi;
    __threadfence_system();
    st_release_sys(&peer->key_flag[par][X.rank], seq);
  }
  bool ok = true;
  int32_t mlo = 0x7fffffff, mhi = (int32_t)0x80000000;
  if (lane < X.world) {
    ok = wait_flag(&me->key_flag[par][lane], seq);
    const volatile int32_t* k = me->keys[par][lane];
    mlo = k[0];
    mhi = k[1];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mlo = min(mlo, __shfl_xor_sync(0xffffffffu, mlo, d));
    mhi = max(mhi, __shfl_xor_sync(0xffffffffu, mhi, d));
  }
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    me->level = 0.5f * (key_to_float(mlo) + key_to_float(mhi));
    me->seq_keys = seq;
    if (!all_ok) me->err = 1;
    keys[0] = 0x7fffffff;           // reset the decoder's accumulator for the next step
    keys[1] = (int32_t)0x80000000;
  }
}

struct ExDest { float* v[8]; int32_t* f[8]; long long vcap, fcap; };

// ---- (2)+(3) in ONE kernel: count exchange, copy of the piece with BULK stores, completion signal ----
// A mesh piece is ~1-4 MB.  Fine-grained stores from the SMs sustain only ~140 GB/s of NVLink ingress per GPU
// (measured, round 1), i.e. ~45 us for the seven pieces that arrive at rank 0 on an 8-GPU node — as long as a
// slab's marching cubes.  Here a block stages 16 KB of the piece in shared memory (faces are rebased on the way)
// and one thread sends it with a TMA bulk store (cp.async.bulk.global.shared::cta), the access size the copy
// engines use; three launches and their dependency latency become one.
constexpr int kPushChunk = 4096;               // 4-byte elements per bulk store (16 KB)

__device__ __forceinline__ void bulk_store(void* dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes)
               : "memory");
}

// elements [0, n) of src -> dst (+ add) at every destination d[r] + off (element offsets); block-cooperative.
template <typename T, bool ADD>
__device__ __forceinline__ void push_bulk(T* const* dsts, int n_dst, long long off, const T* __restrict__ src,
                                          long long n, T add, T* sbuf) {
  if (n <= 0) return;
  auto put = [&](T v) { return ADD ? (T)(v + add) : v; };      // vertices are copied bit for bit (no "+ 0.0f": -0.0f)
  // all destinations share the alignment of `off` (their bases are allocation-aligned): scalar head up to 16 bytes
  long long head = (long long)((16 - ((reinterpret_cast<uintptr_t>(dsts[0] + off)) & 15)) & 15) / 4;
  if (head > n) head = n;
  const long long body = (n - head) / 4 * 4;
  if (blockIdx.x == 0) {
    for (int r = 0; r < n_dst; ++r) {
      T* d = dsts[r] + off;
      for (long long i = threadIdx.x; i < head; i += blockDim.x) d[i] = put(src[i]);
      for (long long i = head + body + threadIdx.x; i < n; i += blockDim.x) d[i] = put(src[i]);
    }
  }
  for (long long c0 = (long long)blockIdx.x * kPushChunk; c0 < body; c0 += (long long)gridDim.x * kPushChunk) {
    const int len = (int)min((long long)kPushChunk, body - c0);            // multiple of 4 elements
    for (int i = threadIdx.x; i < len; i += blockDim.x) sbuf[i] = put(src[head + c0 + i]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy writes -> visible to the bulk copy
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int r = 0; r < n_dst; ++r) bulk_store(dsts[r] + off + head + c0, sbuf, (uint32_t)len * 4u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the staging buffer may be overwritten
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) exchange_mesh_kernel(ExPeers X, ExDest D, const long long* counts,
                                                            const float* __restrict__ verts,
                                                            const int32_t* __restrict__ faces, long long* total_out) {
  __shared__ __align__(128) int32_t sbuf[kPushChunk];
  __shared__ long long s_base[4];
  __shared__ int s_ok, s_last;
  ExCtrl* me = X.c[X.rank];
  const int lane = threadIdx.x & 31;
  const uint32_t seq = me->seq_cnt + 1;        // local state: advanced by the last block only, after every block has read it
  const int par = seq & 1;
  // ---- (2) publish (V,F) (block 0), wait for everyone's (every block, on its own copy of the table), exclusive scan ----
  if (threadIdx.x < 32) {
    const long long v = counts[0], f = counts[1];
    if (blockIdx.x == 0 && lane < X.world) {
      ExCtrl* peer = X.c[lane];
      volatile long long* c = peer->cnt[par][X.rank];
      c[0] = v;
      c[1] = f;
      __threadfence_system();
      st_release_sys(&peer->cnt_flag[par][X.rank], seq);
    }
    bool ok = true;
    long long pv = 0, pf = 0;
    if (lane < X.world) {
      ok = wait_flag(&me->cnt_flag[par][lane], seq);
      const volatile long long* c = me->cnt[par][lane];
      pv = c[0];
      pf = c[1];
    }
    long long bv = 0, bf = 0, tv = 0, tf = 0;
    for (int r = 0; r < X.world; ++r) {
      const long long rv = __shfl_sync(0xffffffffu, pv, r), rf = __shfl_sync(0xffffffffu, pf, r);
      if (r < X.rank) { bv += rv; bf += rf; }
      tv += rv; tf += rf;
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) { s_base[0] = bv; s_base[1] = bf; s_base[2] = tv; s_base[3] = tf; s_ok = all_ok ? 1 : 0; }
  }
  __syncthreads();
  const long long bv = s_base[0], bf = s_base[1];
  // ---- (3a) the piece -> every destination at its base (capacity overflow: truncated, the totals tell) ----
  float* dv[8];
  int32_t* df[8];
  int n_dst = 0;
  for (int r = 0; r < X.world; ++r)
    if (D.v[r] && D.f[r]) { dv[n_dst] = D.v[r]; df[n_dst] = D.f[r]; ++n_dst; }
  long long nv3 = counts[0] * 3, nf3 = counts[1] * 3;
  if (nv3 > (D.vcap - bv) * 3) nv3 = (D.vcap - bv) * 3;
  if (nf3 > (D.fcap - bf) * 3) nf3 = (D.fcap - bf) * 3;
  push_bulk<float, false>(dv, n_dst, bv * 3, verts, nv3, 0.0f, reinterpret_cast<float*>(sbuf));
  push_bulk<int32_t, true>(df, n_dst, bf * 3, faces, nf3, (int32_t)bv, sbuf);
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this block's bulk stores have been performed
  __threadfence_system();
  __syncthreads();
  // ---- (3b) the last block tells every destination that the piece has landed; a destination waits for all pieces ----
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&me->ticket, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  __threadfence_system();
  const uint32_t dseq = me->seq_done + 1;
  if (lane < X.world && D.v[lane]) st_release_sys(&X.c[lane]->done_flag[X.rank], dseq);
  bool ok = true;
  if (D.v[X.rank] && lane < X.world) ok = wait_flag(&me->done_flag[lane], dseq);
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    me->base[0] = bv; me->base[1] = bf;
    me->total[0] = s_base[2]; me->total[1] = s_base[3];
    me->seq_cnt = seq;
    me->seq_done = dseq;
    me->ticket = 0;
    if (!all_ok || !s_ok) me->err = 1;
    if (total_out) { total_out[0] = s_base[2]; total_out[1] = s_base[3]; }
  }
}

static int make_peers(const vtaco_exchange* ex, ExPeers& X) {
  if (!ex || ex->world < 1 || ex->world > 8 || ex->rank < 0 || ex->rank >= ex->world) return VTACO_ERR_INVALID_ARG;
  for (int r = 0; r < 8; ++r) {
    X.c[r] = r < ex->world ? reinterpret_cast<ExCtrl*>(ex->ctrl[r]) : nullptr;
    if (r < ex->world && !X.c[r]) return VTACO_ERR_INVALID_ARG;
  }
  X.world = ex->world;
  X.rank = ex->rank;
  return VTACO_OK;
}

}  // namespace vtaco

using namespace vtaco;

extern "C" int vtaco_exchange_level(const vtaco_exchange* ex, int32_t* keys, void* stream) {
  ExPeers X;
  const int st = make_peers(ex, X);
  if (st != VTACO_OK) return st;
  if (!keys) return VTACO_ERR_INVALID_ARG;
  exchange_level_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(X, keys);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

extern "C" int vtaco_exchange_mesh(const vtaco_exchange* ex, const vtaco_mesh_piece* m, void* stream) {
  ExPeers X;
  const int st = make_peers(ex, X);
  if (st != VTACO_OK) return st;
  if (!m || !m->counts || !m->vertices || !m->faces) return VTACO_ERR_INVALID_ARG;
  ExDest D;
  bool any = false;
  for (int r = 0; r < 8; ++r) {
    D.v[r] = r < ex->world ? m->dst_vertices[r] : nullptr;
    D.f[r] = r < ex->world ? m->dst_faces[r] : nullptr;
    if ((D.v[r] == nullptr) != (D.f[r] == nullptr)) return VTACO_ERR_INVALID_ARG;
    any |= D.v[r] != nullptr;
  }
  if (!any || m->vertex_capacity < 0 || m->face_capacity < 0) return VTACO_ERR_INVALID_ARG;
  D.vcap = m->vertex_capacity;
  D.fcap = m->face_capacity;
  cudaStream_t s = (cudaStream_t)stream;
  // all blocks spin on the count flags, so they must be co-resident: one block per SM at most
  exchange_mesh_kernel<<<num_sms(), 256, 0, s>>>(X, D, reinterpret_cast<const long long*>(m->counts), m->vertices,
                                                 m->faces, reinterpret_cast<long long*>(m->total_counts));
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}
