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

// (2) publish (V,F), wait for everyone's, exclusive scan -> this rank's bases and the totals.
__global__ void __launch_bounds__(32) exchange_counts_kernel(ExPeers X, const long long* counts) {
  ExCtrl* me = X.c[X.rank];
  const int lane = threadIdx.x;
  const uint32_t seq = me->seq_cnt + 1;
  const int par = seq & 1;
  const long long v = counts[0], f = counts[1];
  if (lane < X.world) {
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
  if (lane == 0) {
    me->base[0] = bv; me->base[1] = bf;
    me->total[0] = tv; me->total[1] = tf;
    me->seq_cnt = seq;
    if (!all_ok) me->err = 1;
  }
}

struct ExDest { float* v[8]; int32_t* f[8]; long long vcap, fcap; };

// (3a) copy this rank's piece into every destination at its base; faces are rebased.  Remote stores are
// 16 bytes wide (fine-grained 4-byte stores sustain only ~40 GB/s over NVLink): the destination is aligned up
// to 16 B with a scalar head, the local source is read with scalar loads (it is misaligned by then).
template <typename T, bool ADD>
__device__ __forceinline__ void push_range(T* __restrict__ dst, const T* __restrict__ src, long long n, T add_,
                                           long long t0, long long stride) {
  if (n <= 0) return;
  const T add = ADD ? add_ : T(0);
  auto put = [&](T v) { return ADD ? (T)(v + add) : v; };      // vertices are copied bit for bit (no "+ 0.0f": -0.0f)
  long long head = (long long)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) / 4;
  if (head > n) head = n;
  const long long n4 = (n - head) / 4;
  for (long long i = t0; i < head; i += stride) dst[i] = put(src[i]);
  struct alignas(16) V4 { T a, b, c, d; };
  V4* d4 = reinterpret_cast<V4*>(dst + head);
  for (long long i = t0; i < n4; i += stride) {
    const T* q = src + head + 4 * i;
    V4 v;
    v.a = put(q[0]); v.b = put(q[1]); v.c = put(q[2]); v.d = put(q[3]);
    d4[i] = v;
  }
  for (long long i = head + 4 * n4 + t0; i < n; i += stride) dst[i] = put(src[i]);
}

__global__ void __launch_bounds__(256) exchange_push_kernel(ExPeers X, ExDest D, const long long* counts,
                                                            const float* __restrict__ verts,
                                                            const int32_t* __restrict__ faces) {
  const ExCtrl* me = X.c[X.rank];
  const long long bv = me->base[0], bf = me->base[1];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int r = 0; r < X.world; ++r) {
    float* dv = D.v[r];
    int32_t* df = D.f[r];
    if (!dv || !df) continue;
    // capacity overflow: truncated, the totals tell
    long long nv3 = counts[0] * 3, nf3 = counts[1] * 3;
    if (nv3 > (D.vcap - bv) * 3) nv3 = (D.vcap - bv) * 3;
    if (nf3 > (D.fcap - bf) * 3) nf3 = (D.fcap - bf) * 3;
    push_range<float, false>(dv + bv * 3, verts, nv3, 0.0f, t0, stride);
    push_range<int32_t, true>(df + bf * 3, faces, nf3, (int32_t)bv, t0, stride);
  }
  __threadfence_system();
}

// (3b) tell every destination that this rank's piece has landed; a destination waits for all
// pieces and publishes the totals.
__global__ void __launch_bounds__(32) exchange_done_kernel(ExPeers X, ExDest D, long long* total_out) {
  ExCtrl* me = X.c[X.rank];
  const int lane = threadIdx.x;
  const uint32_t seq = me->seq_done + 1;
  if (lane < X.world && D.v[lane]) st_release_sys(&X.c[lane]->done_flag[X.rank], seq);
  bool ok = true;
  if (D.v[X.rank] && lane < X.world) ok = wait_flag(&me->done_flag[lane], seq);
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    me->seq_done = seq;
    if (!all_ok) me->err = 1;
    if (total_out) { total_out[0] = me->total[0]; total_out[1] = me->total[1]; }
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
  exchange_counts_kernel<<<1, 32, 0, s>>>(X, reinterpret_cast<const long long*>(m->counts));
  exchange_push_kernel<<<num_sms(), 256, 0, s>>>(X, D, reinterpret_cast<const long long*>(m->counts), m->vertices,
                                                 m->faces);
  exchange_done_kernel<<<1, 32, 0, s>>>(X, D, reinterpret_cast<long long*>(m->total_counts));
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}
