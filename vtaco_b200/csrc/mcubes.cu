// GPU marching cubes over the decoded logit grid (sm_100a).
//
// Replaces the skimage.measure.marching_cubes call + vertex rescale of
// Generator3D.generate_obj_mesh_wnf (reference src/conv_onet/generation.py:268-272).
// Contract and conventions: oracle/marching_cubes.py / oracle/mc_tables.py (the
// triangulation table is generated from the rule stated there; skimage's Lewiner tables
// are not available offline — parity with skimage is UNPINNED, parity with the oracle is
// bit-exact on case indices / faces and to rounding on vertices).
//
// HBM-bound stream compaction in TWO kernels that read the grid once:
//   mc_fused_kernel : 1 thread / run of 8 consecutive z points — four rows of 9 samples as float4
//              loads -> "above" bit rows -> per point 3-bit own-edge mask + triangle count; block
//              scan; the block's exclusive prefix comes from a DECOUPLED LOOK-BACK over the
//              blocks before it (single pass, blocks take their index from a ticket so that
//              every predecessor has started); with its vertex base known the thread emits the
//              vertices of its own cut edges at once (inverse-distance weighting in double, like
//              Lewiner's code), stores the run's codes + vertex base for the face pass and
//              appends runs that own triangles to a work list (with their triangle base)
//   mc_faces_kernel : grid-stride over the work list only (~5 % of the runs): vertex ids are looked
//              up from the owners' codes / bases; output order = lattice order, then table order
//              (the order of the WORK is arbitrary, every item carries its output position).
// Slab mode (x_emit < nx; multi-GPU extraction, SURVEY 8e "gather of mesh pieces"): the volume
// holds the rank's x-rows plus two halo rows; vertex ids are numbered over the whole volume (so a
// halo-row vertex gets the id it has as the NEXT rank's first vertices), but only vertices owned
// by rows < x_emit and faces of cells in rows < x_emit are emitted and counted.  The pieces of
// consecutive slabs then concatenate to exactly the single-volume mesh (ids + the vertex base of
// the slab).
// Algorithmic bytes: 4*n (grid) + 12*V + 12*F; scratch traffic adds the active blocks' codes
// and base ids.
#include "common.cuh"
#include "mc_tables.h"
#include <float.h>

namespace vtaco {

constexpr int kMcThreads = 256;   // threads per block
constexpr int kMcRun = 8;         // consecutive z points per thread
constexpr int kMcBlockPts = kMcThreads * kMcRun;
constexpr int kMcSuper = 256;     // blocks per super-block of the two-level prefix

struct McParams {
  const float* grid;
  int nx, ny, nz, nzc;            // nzc = ceil(nz / kMcRun) runs per (x,y) row
  int x_emit;                     // rows [0, x_emit) emit vertices / faces (== nx: whole volume)
  int x_origin;                   // lattice row of the sub-volume's row 0 (vertex coordinates are global)
  long long npts, nruns;
  float level;
  const int32_t* level_keys;
  int n_level_keys;
  const float* level_ptr;
  uint8_t* code;                  // [nruns][8]: bits 0-2 own-edge mask (x,y,z), bits 3-5 triangle count (active runs only)
  uint32_t* vbase;                // [nruns]: id of the first vertex owned by the run (active runs only)
  unsigned long long* state;      // [nblocks] look-back words: flag << 62 | triangles << 31 | vertices (zeroed per call)
  uint2* work;                    // [nruns] (run, triangle base) of the runs that own triangles
  unsigned* ctrl;                 // [0] block ticket, [1] work-list length (zeroed per call)
  int nblocks, emit;
  long long* counts;
  float* verts;
  long long vcap;
  int32_t* faces;
  long long fcap;
  float voffset, vscale;
};

__device__ __forceinline__ float mc_level(const McParams& P) {
  if (P.level_ptr) return *P.level_ptr;
  if (P.level_keys) {
    int32_t lo = P.level_keys[0], hi = P.level_keys[1];
    for (int r = 1; r < P.n_level_keys; ++r) {
      lo = min(lo, P.level_keys[2 * r]);
      hi = max(hi, P.level_keys[2 * r + 1]);
    }
    return 0.5f * (key_to_float(lo) + key_to_float(hi));
  }
  return P.level;
}

// exclusive scan of a packed (hi16 | lo16) value across the block; returns the exclusive
// prefix and the block total (all threads).
__device__ __forceinline__ unsigned block_scan_excl(unsigned v, unsigned& total) {
  constexpr int kWarps = kMcThreads / 32;
  __shared__ unsigned warp_excl[kWarps];
  __shared__ unsigned tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_excl[w] = inc;  // warp totals
  __syncthreads();
  if (w == 0) {
    const unsigned s = lane < kWarps ? warp_excl[lane] : 0u;
    unsigned si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) si += t;
    }
    if (lane < kWarps) warp_excl[lane] = si - s;
    if (lane == kWarps - 1) tot = si;
  }
  __syncthreads();
  total = tot;
  return warp_excl[w] + inc - v;
}

// run -> lattice coordinates of its first point
__device__ __forceinline__ void run_coords(const McParams& P, long long run, int& i, int& j, int& k0) {
  const unsigned ru = (unsigned)run;           // nruns < 2^31 (checked by the host wrapper)
  const unsigned r = ru / (unsigned)P.nzc;
  const int kc = (int)(ru - r * (unsigned)P.nzc);
  i = (int)(r / (unsigned)P.ny);
  j = (int)(r - (unsigned)i * (unsigned)P.ny);
  k0 = kc * kMcRun;
}

// "above" bits of kMcRun+1 consecutive z samples of row (i,j), bit t <-> k0+t; rows or samples
// outside the lattice replicate the last valid sample (their cells/edges are masked out later).
__device__ __forceinline__ unsigned row_bits(const McParams& P, int i, int j, int k0, float level) {
  i = min(i, P.nx - 1);
  j = min(j, P.ny - 1);
  const float* row = P.grid + ((long long)i * P.ny + j) * P.nz;
  unsigned bits = 0;
  if ((P.nz & 7) == 0) {   // uniform over the grid: the last run of a row takes this path too (a per-lane
                           // test sent one lane of EVERY warp through the scalar path below: 750 instructions per thread)
    const float4 a = __ldg(reinterpret_cast<const float4*>(row + k0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(row + k0 + 4));
    const float c = __ldg(row + min(k0 + 8, P.nz - 1));
    bits = (unsigned)(a.x > level) | ((unsigned)(a.y > level) << 1) | ((unsigned)(a.z > level) << 2) |
           ((unsigned)(a.w > level) << 3) | ((unsigned)(b.x > level) << 4) | ((unsigned)(b.y > level) << 5) |
           ((unsigned)(b.z > level) << 6) | ((unsigned)(b.w > level) << 7) | ((unsigned)(c > level) << 8);
  } else {
#pragma unroll
    for (int t = 0; t <= kMcRun; ++t) bits |= (unsigned)(__ldg(row + min(k0 + t, P.nz - 1)) > level) << t;
  }
  return bits;
}

// codes of the 8 points of a run (see McParams::code); returns packed (ntris << 16 | nverts).
// Bit-parallel over the run: own-edge masks are XORs of the "above" rows and a cell is cut iff
// its 8 corner bits are neither all 0 nor all 1, so the ~95 % of runs that the surface does not
// touch cost a handful of logic ops; only the set bits take the per-point path.
__device__ __forceinline__ unsigned run_codes(const McParams& P, int i, int j, int k0, float level,
                                               unsigned long long& codes, const uint8_t* __restrict__ tri_count) {
  const unsigned r00 = row_bits(P, i, j, k0, level), r01 = row_bits(P, i, j + 1, k0, level);
  const unsigned r10 = row_bits(P, i + 1, j, k0, level), r11 = row_bits(P, i + 1, j + 1, k0, level);
  const bool hx = i + 1 < P.nx, hy = j + 1 < P.ny;
  const int left = P.nz - k0;                                     // points of this run inside the lattice
  const unsigned m8 = left >= kMcRun ? 0xffu : ((1u << left) - 1u);
  const unsigned mz = left - 1 >= kMcRun ? 0xffu : ((1u << max(left - 1, 0)) - 1u);   // points with k+1 < nz
  const unsigned fx = hx ? ((r00 ^ r10) & m8) : 0u;
  const unsigned fy = hy ? ((r00 ^ r01) & m8) : 0u;
  const unsigned fz = (r00 ^ (r00 >> 1)) & mz;
  const unsigned any = r00 | r10 | r01 | r11, all = r00 & r10 & r01 & r11;
  // halo rows (i >= x_emit) keep their vertex flags (the numbering runs through them) but own no cells
  const unsigned cut = (hx && hy && i < P.x_emit) ? (((any | (any >> 1)) & ~(all & (all >> 1))) & mz) : 0u;
  codes = 0;
  unsigned todo = fx | fy | fz | cut;
  unsigned packed = __popc(fx) + __popc(fy) + __popc(fz);
  while (todo) {
    const int t = __ffs(todo) - 1;
    todo &= todo - 1;
    const unsigned flags = ((fx >> t) & 1u) | (((fy >> t) & 1u) << 1) | (((fz >> t) & 1u) << 2);
    unsigned nt = 0;
    if ((cut >> t) & 1u) {
      // corner c = x | y<<1 | z<<2
      const unsigned cs = ((r00 >> t) & 1u) | (((r10 >> t) & 1u) << 1) | (((r01 >> t) & 1u) << 2) |
                          (((r11 >> t) & 1u) << 3) | (((r00 >> (t + 1)) & 1u) << 4) | (((r10 >> (t + 1)) & 1u) << 5) |
                          (((r01 >> (t + 1)) & 1u) << 6) | (((r11 >> (t + 1)) & 1u) << 7);
      nt = (unsigned)tri_count[cs];     // shared-memory copy: a divergent __constant__ index is serialised per lane
    }
    codes |= (unsigned long long)(flags | (nt << 3)) << (8 * t);
    packed += nt << 16;
  }
  return packed;
}

constexpr unsigned long long kMcFlagAgg = 1ull << 62, kMcFlagPrefix = 2ull << 62, kMcValMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// point (i,j,k) -> run index / slot within the run
__device__ __forceinline__ long long run_of(const McParams& P, int i, int j, int k) {
  return ((long long)i * P.ny + j) * P.nzc + (k >> 3);
}

constexpr int kMcSub = 4;                          // consecutive 256-run sub-blocks per CTA: 4x fewer look-back words
constexpr int kMcBlockRuns = kMcThreads * kMcSub;

__global__ void __launch_bounds__(kMcThreads) mc_fused_kernel(const __grid_constant__ McParams P) {
  __shared__ int s_bid;
  __shared__ unsigned long long s_prefix;
  __shared__ uint8_t s_tricount[256];
  if (threadIdx.x == 0) s_bid = (int)atomicAdd(P.ctrl, 1u);      // ticket: all blocks before this one have started
  s_tricount[threadIdx.x] = (uint8_t)kMcTriCount[threadIdx.x];
  __syncthreads();
  const int b = s_bid;
  const float level = mc_level(P);
  // ---- classify kMcSub runs per thread (run = block base + s * 256 + thread: coalesced, lattice order = (s, thread)) ----
  unsigned packed[kMcSub], excl[kMcSub];
  unsigned long long codes[kMcSub];
  unsigned sub_base = 0;                                         // packed totals of the sub-blocks before s
#pragma unroll
  for (int s = 0; s < kMcSub; ++s) {
    const long long run = (long long)b * kMcBlockRuns + s * kMcThreads + threadIdx.x;
    packed[s] = 0;
    codes[s] = 0;
    if (run < P.nruns) {
      int i, j, k0;
      run_coords(P, run, i, j, k0);
      packed[s] = run_codes(P, i, j, k0, level, codes[s], s_tricount);
    }
    unsigned total;
    excl[s] = sub_base + block_scan_excl(packed[s], total);
    sub_base += total;                                           // fields: vertices < 2^16 (4 x 6144), triangles < 2^16 (4 x 10240)
  }
  const unsigned total = sub_base;
  // ---- decoupled look-back (warp 0): exclusive prefix of this block over all blocks before it ----
  if (threadIdx.x < 32) {
    const unsigned long long agg = (unsigned long long)(total & 0xffffu) | ((unsigned long long)(total >> 16) << 31);
    unsigned long long prefix = 0;
    if (b > 0) {
      if (threadIdx.x == 0) st_state(P.state + b, kMcFlagAgg | agg);
      int base = b - 1;
      for (;;) {
        constexpr int kLb = 2;                                   // words per lane: 64 predecessors per round
        unsigned long long w[kLb];
#pragma unroll
        for (int h = 0; h < kLb; ++h) {                          // lane l reads predecessors base - 32 h - l
          const int idx = base - 32 * h - (int)threadIdx.x;
          w[h] = kMcFlagPrefix;                                  // before block 0: an empty prefix
          if (idx >= 0) w[h] = ld_state(P.state + idx);
        }
        // a word that is not published yet ends the usable range of this round: everything nearer than the
        // nearest prefix word must be an aggregate, otherwise look again
        int first = kLb * 32, hole = kLb * 32;
#pragma unroll
        for (int h = kLb - 1; h >= 0; --h) {
          const unsigned pp = __ballot_sync(0xffffffffu, (w[h] >> 62) == 2), zz = __ballot_sync(0xffffffffu, (w[h] >> 62) == 0);
          if (pp) first = 32 * h + (__ffs(pp) - 1);
          if (zz) hole = 32 * h + (__ffs(zz) - 1);
        }
        if (hole < first) continue;                              // an unpublished predecessor nearer than the prefix: poll again
        unsigned long long v = 0;
#pragma unroll
        for (int h = 0; h < kLb; ++h)
          if (32 * h + (int)threadIdx.x <= first) v += w[h] & kMcValMask;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        prefix += v;
        if (first < kLb * 32) break;
        base -= kLb * 32;
      }
    }
    if (threadIdx.x == 0) {
      st_state(P.state + b, kMcFlagPrefix | (prefix + agg));
      s_prefix = prefix;
      if (b == P.nblocks - 1) {                                  // totals (slab mode: halo cells were classified empty)
        const unsigned long long tot = prefix + agg;
        P.counts[1] = (long long)(tot >> 31);
        P.counts[2] = (long long)(tot & 0x7fffffffull);
        if (P.x_emit >= P.nx) P.counts[0] = (long long)(tot & 0x7fffffffull);
      }
    }
  }
  __syncthreads();
  const unsigned long long pre = s_prefix;
  const double dlevel = (double)level;
  const long long stride[3] = {(long long)P.ny * P.nz, (long long)P.nz, 1};
#pragma unroll
  for (int s = 0; s < kMcSub; ++s) {
    const long long run = (long long)b * kMcBlockRuns + s * kMcThreads + threadIdx.x;
    if (run >= P.nruns) continue;
    unsigned long long vb = (pre & 0x7fffffffull) + (excl[s] & 0xffffu);
    const unsigned long long tb = (pre >> 31) + (excl[s] >> 16);
    if (P.x_emit < P.nx && run == (long long)P.x_emit * P.ny * P.nzc) P.counts[0] = (long long)vb;   // first halo run
    if (!packed[s]) continue;
    reinterpret_cast<unsigned long long*>(P.code)[run] = codes[s];
    P.vbase[run] = (uint32_t)vb;
    if (!P.emit) continue;
    if (packed[s] >> 16) {                                       // the run owns triangles: face pass work item
      const unsigned slot = atomicAdd(P.ctrl + 1, 1u);
      P.work[slot] = make_uint2((unsigned)run, (unsigned)tb);
    }
    int i, j, k0;
    run_coords(P, run, i, j, k0);
    if (i >= P.x_emit) continue;                                 // halo rows: numbered, not emitted
#pragma unroll 1
    for (int t = 0; t < kMcRun; ++t) {
      const unsigned flags = (unsigned)(codes[s] >> (8 * t)) & 7u;
      if (!flags) continue;
      const int k = k0 + t;
      const long long p = ((long long)i * P.ny + j) * P.nz + k;
      const double d0 = fabs((double)P.grid[p] - dlevel);
      const float base[3] = {(float)(i + P.x_origin), (float)j, (float)k};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (flags & (1u << a)) {
          if ((long long)vb < P.vcap) {    // capacity overflow: counted, not written (the caller re-runs with larger buffers)
            const double d1 = fabs((double)P.grid[p + stride[a]] - dlevel);
            // Lewiner's weights w = 1 / (eps + d): t = w1 / (w0 + w1) = (eps + d0) / ((eps + d0) + (eps + d1)), one
            // division instead of three (equal to the last bit or two of the double, far below the float32 result)
            const double e0 = (double)FLT_EPSILON + d0, e1 = (double)FLT_EPSILON + d1;
            const double tt = e0 / (e0 + e1);
            float pos[3] = {base[0], base[1], base[2]};
            pos[a] = (float)((double)base[a] + tt);
            float* o = P.verts + vb * 3;
            o[0] = (pos[0] - P.voffset) * P.vscale;
            o[1] = (pos[1] - P.voffset) * P.vscale;
            o[2] = (pos[2] - P.voffset) * P.vscale;
          }
          ++vb;
        }
      }
    }
  }
}

// id of the vertex on axis `a` owned by point (i,j,k): the run's base + the vertices of the points before it
// in the run + the lower axes of the point itself
__device__ __forceinline__ int32_t vertex_id(const McParams& P, int i, int j, int k, int a) {
  const long long r = run_of(P, i, j, k);
  const unsigned long long codes = reinterpret_cast<const unsigned long long*>(P.code)[r];
  const int t = k & 7;
  const unsigned long long before = codes & ((1ull << (8 * t)) - 1ull) & 0x0707070707070707ull;
  const unsigned own = (unsigned)(codes >> (8 * t)) & 7u;
  return (int32_t)(P.vbase[r] + __popcll(before) + __popc(own & ((1u << a) - 1u)));
}

__global__ void __launch_bounds__(kMcThreads) mc_faces_kernel(const __grid_constant__ McParams P) {
  // case tables in shared memory: every lane indexes them with its own case (a divergent __constant__ index costs
  // one replay per distinct address — it was the top stall of this kernel); a case row is one 16-byte load
  __shared__ uint4 s_tri[256];
  __shared__ int8_t s_edge[12][4];
  {
    int8_t* dst = reinterpret_cast<int8_t*>(s_tri);
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) dst[i] = (i & 15) < 15 ? kMcTriTable[i >> 4][i & 15] : (int8_t)0;
    if (threadIdx.x < 48) s_edge[threadIdx.x >> 2][threadIdx.x & 3] = kMcEdge[threadIdx.x >> 2][threadIdx.x & 3];
  }
  __syncthreads();
  const unsigned n_work = P.ctrl[1];
  const float level = mc_level(P);
  for (unsigned wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_work; wi += gridDim.x * blockDim.x) {
    const uint2 item = P.work[wi];
    const long long run = item.x;
    unsigned long long tb = item.y;
    const unsigned long long codes = reinterpret_cast<const unsigned long long*>(P.code)[run];
    int i, j, k0;
    run_coords(P, run, i, j, k0);
    const unsigned r00 = row_bits(P, i, j, k0, level), r01 = row_bits(P, i, j + 1, k0, level);
    const unsigned r10 = row_bits(P, i + 1, j, k0, level), r11 = row_bits(P, i + 1, j + 1, k0, level);
#pragma unroll 1
    for (int t = 0; t < kMcRun; ++t) {
      const unsigned nt = ((unsigned)(codes >> (8 * t)) & 0xffu) >> 3;
      if (!nt) continue;
      const unsigned cs = ((r00 >> t) & 1u) | (((r10 >> t) & 1u) << 1) | (((r01 >> t) & 1u) << 2) |
                          (((r11 >> t) & 1u) << 3) | (((r00 >> (t + 1)) & 1u) << 4) | (((r10 >> (t + 1)) & 1u) << 5) |
                          (((r01 >> (t + 1)) & 1u) << 6) | (((r11 >> (t + 1)) & 1u) << 7);
      const int k = k0 + t;
      const uint4 row = s_tri[cs];
      const int8_t* edges = reinterpret_cast<const int8_t*>(&row);
      for (unsigned tr = 0; tr < nt; ++tr) {
        if ((long long)(tb + tr) >= P.fcap) break;
        int32_t* o = P.faces + (tb + tr) * 3;
#pragma unroll
        for (int corner = 0; corner < 3; ++corner) {
          const int e = (int)((reinterpret_cast<const unsigned*>(&row)[(3 * tr + corner) >> 2] >> (8 * ((3 * tr + corner) & 3))) & 0xffu);
          const char4 ed = *reinterpret_cast<const char4*>(s_edge[e]);
          o[corner] = vertex_id(P, i + ed.y, j + ed.z, k + ed.w, ed.x);
        }
      }
      (void)edges;
      tb += nt;
    }
  }
}

// min / max of a grid -> ordered-int keys (for level=None when the grid did not come from
// the decoder kernel, which tracks them itself).
__global__ void __launch_bounds__(256) grid_minmax_kernel(const float* __restrict__ g, long long n,
                                                          int32_t* __restrict__ keys) {
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(keys, float_to_key(lo));
    atomicMax(keys + 1, float_to_key(hi));
  }
}

__global__ void init_keys_kernel(int32_t* keys) {
  keys[0] = 0x7fffffff;
  keys[1] = (int32_t)0x80000000;
}

struct PeerTables { int32_t* t[8]; };
__global__ void publish_keys_kernel(int32_t* keys, PeerTables tabs, int n_peers, int rank) {
  const int32_t lo = keys[0], hi = keys[1];
  for (int r = 0; r < n_peers; ++r) {
    tabs.t[r][2 * rank] = lo;
    tabs.t[r][2 * rank + 1] = hi;
  }
  __threadfence_system();
  keys[0] = 0x7fffffff;
  keys[1] = (int32_t)0x80000000;
}

static long long mc_align(long long v) { return (v + 255) / 256 * 256; }

}  // namespace vtaco

using namespace vtaco;

extern "C" int64_t vtaco_mc_scratch_bytes(int32_t nx, int32_t ny, int32_t nz) {
  if (nx < 1 || ny < 1 || nz < 1) return VTACO_ERR_INVALID_ARG;
  const long long nruns = (long long)nx * ny * ((nz + kMcRun - 1) / kMcRun);
  const long long nb = (nruns + kMcThreads - 1) / kMcThreads;
  // codes (8 B / run), vertex base (4 B / run), work list (8 B / run), look-back state (8 B / block) + control words
  return mc_align(nruns * 8) + mc_align(4 * nruns) + mc_align(8 * nruns) + mc_align(8 * nb + 256);
}

extern "C" int vtaco_grid_minmax(const float* grid, int64_t n, int32_t* keys, void* stream) {
  if (!grid || !keys || n <= 0) return VTACO_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  init_keys_kernel<<<1, 1, 0, st>>>(keys);
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  grid_minmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(grid, n, keys);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

extern "C" int vtaco_marching_cubes(const vtaco_mc_args* a, void* stream) {
  if (!a || !a->grid || !a->scratch || !a->counts) return VTACO_ERR_INVALID_ARG;
  if (a->nx < 1 || a->ny < 1 || a->nz < 1) return VTACO_ERR_INVALID_ARG;
  if (a->phase < 1 || a->phase > 3) return VTACO_ERR_INVALID_ARG;
  if (a->x_emit < 0 || a->x_emit > a->nx) return VTACO_ERR_INVALID_ARG;
  const long long n = (long long)a->nx * a->ny * a->nz;
  if (n >= (1ll << 31)) return VTACO_ERR_UNSUPPORTED;
  if (vtaco_mc_scratch_bytes(a->nx, a->ny, a->nz) > a->scratch_bytes) return VTACO_ERR_CAPACITY;
  if ((a->phase & 2) && ((a->vertex_capacity > 0 && !a->vertices) || (a->face_capacity > 0 && !a->faces)))
    return VTACO_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  McParams P = {};
  P.grid = a->grid; P.nx = a->nx; P.ny = a->ny; P.nz = a->nz; P.npts = n;
  P.x_emit = a->x_emit > 0 ? a->x_emit : a->nx;
  P.x_origin = a->x_origin;
  P.level = a->level; P.level_keys = a->level_keys; P.level_ptr = a->level_ptr;
  P.n_level_keys = a->n_level_keys > 1 ? a->n_level_keys : 1;
  P.nzc = (a->nz + kMcRun - 1) / kMcRun;
  P.nruns = (long long)a->nx * a->ny * P.nzc;
  P.nblocks = (int)((P.nruns + kMcBlockRuns - 1) / kMcBlockRuns);
  char* s = reinterpret_cast<char*>(a->scratch);
  P.code = reinterpret_cast<uint8_t*>(s); s += mc_align(P.nruns * 8);
  P.vbase = reinterpret_cast<uint32_t*>(s); s += mc_align(4 * P.nruns);
  P.work = reinterpret_cast<uint2*>(s); s += mc_align(8 * P.nruns);
  P.state = reinterpret_cast<unsigned long long*>(s);
  P.ctrl = reinterpret_cast<unsigned*>(s + 8ll * P.nblocks);
  P.counts = reinterpret_cast<long long*>(a->counts);
  P.verts = a->vertices; P.vcap = a->vertex_capacity;
  P.faces = a->faces; P.fcap = a->face_capacity;
  P.voffset = a->voffset; P.vscale = a->vscale;
  // phase 1 = count only (nothing is written to the outputs); 2 and 3 = the whole extraction (classification is
  // part of the same pass, so "emit after a count" simply runs it again with the larger buffers)
  P.emit = (a->phase & 2) ? 1 : 0;
  if (!P.emit) P.vcap = 0;
  VTACO_CUDA_CHECK(cudaMemsetAsync(P.state, 0, 8ll * P.nblocks + 16, st));
  mc_fused_kernel<<<P.nblocks, kMcThreads, 0, st>>>(P);
  if (P.emit) {
    long long blocks = (long long)num_sms() * 4;
    if (blocks > (P.nruns + kMcThreads - 1) / kMcThreads) blocks = (P.nruns + kMcThreads - 1) / kMcThreads;
    mc_faces_kernel<<<(unsigned)blocks, kMcThreads, 0, st>>>(P);
  }
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

extern "C" int vtaco_publish_keys(int32_t* keys, int32_t* const* tables, int32_t n_peers, int32_t rank, void* stream) {
  if (!keys || !tables || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= 8) return VTACO_ERR_INVALID_ARG;
  PeerTables t;
  for (int r = 0; r < 8; ++r) t.t[r] = r < n_peers ? tables[r] : nullptr;
  publish_keys_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(keys, t, n_peers, rank);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}
