/* Dense occupancy-lattice decode + marching cubes straight from C through the ABI of
 * include/vtaco_b200.h (what a non-Python host — or the reference's maintainer behind a
 * ctypes/cffi stub — calls).  Build:
 *   gcc -std=c99 -Iinclude examples/dense_extract.c -Lvtaco_b200/lib -lvtaco_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -o dense_extract
 * Needs a B200 to run; feature grid and weights are random here (the layouts are what matters):
 *   features  : channels-last [1][R][R][R][32] fp32
 *   weights   : VTACO_DEC_PACKED_FLOATS(n_blocks) floats, layout documented in the header
 *   lattice   : nx^3 logits, axis = (1 + padding) * linspace(-0.5, 0.5, nx)
 * The SIMT kernel (variant 1) is used so that only the fp32 packing is needed. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "vtaco_b200.h"

/* the four CUDA runtime calls used, declared here so that the example needs no CUDA headers */
extern int cudaMalloc(void** p, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
extern int cudaFree(void* p);
enum { H2D = 1, D2H = 2 };

static float frand(void) { return (float)rand() / (float)RAND_MAX - 0.5f; }

int main(void) {
  const int nx = 64, R = 16, nb = 5;
  const size_t n_feat = (size_t)R * R * R * 32, n_w = VTACO_DEC_PACKED_FLOATS(nb), n_q = (size_t)nx * nx * nx;
  float* h_feat = malloc(n_feat * sizeof(float));
  float* h_w = malloc(n_w * sizeof(float));
  float* h_axis = malloc(nx * sizeof(float));
  for (size_t i = 0; i < n_feat; ++i) h_feat[i] = frand();
  for (size_t i = 0; i < n_w; ++i) h_w[i] = 0.3f * frand();
  for (int i = 0; i < nx; ++i) h_axis[i] = 1.1f * (-0.5f + (float)i / (float)(nx - 1));

  float *d_feat, *d_w, *d_axis, *d_logits;
  int32_t* d_keys;
  if (cudaMalloc((void**)&d_feat, n_feat * 4) || cudaMalloc((void**)&d_w, n_w * 4) || cudaMalloc((void**)&d_axis, nx * 4) ||
      cudaMalloc((void**)&d_logits, n_q * 4) || cudaMalloc((void**)&d_keys, 8)) {
    fprintf(stderr, "no CUDA device\n");
    return 2;
  }
  const int32_t keys0[2] = {INT32_MAX, INT32_MIN};
  cudaMemcpy(d_feat, h_feat, n_feat * 4, H2D);
  cudaMemcpy(d_w, h_w, n_w * 4, H2D);
  cudaMemcpy(d_axis, h_axis, nx * 4, H2D);
  cudaMemcpy(d_keys, keys0, 8, H2D);

  vtaco_decoder_args a = {0};
  a.B = 1; a.axis = d_axis; a.nx = nx; a.x0 = 0; a.x1 = nx;           /* dense lattice mode (p == NULL) */
  a.grid = d_feat; a.reso_grid = R; a.padding = 0.1; a.div_mode = VTACO_DIV_RECIPROCAL;
  a.sample_mode = VTACO_SAMPLE_BILINEAR; a.weights = d_w; a.n_blocks = nb;
  a.logits = d_logits; a.minmax_key = d_keys; a.variant = 1;
  int st = vtaco_decoder_forward(&a, NULL);
  if (st) { fprintf(stderr, "decoder: %s %s\n", vtaco_status_string(st), vtaco_last_cuda_error()); return 1; }

  /* marching cubes at 0.5 * (min + max), vertices rescaled like generation.py:271-272 */
  vtaco_mc_args m = {0};
  const int64_t cap_v = 1 << 20, cap_f = 1 << 21;
  const int64_t scratch = vtaco_mc_scratch_bytes(nx, nx, nx);
  void* d_scratch; float* d_v; int32_t* d_f; int64_t* d_counts;
  cudaMalloc(&d_scratch, (size_t)scratch); cudaMalloc((void**)&d_v, cap_v * 12); cudaMalloc((void**)&d_f, cap_f * 12);
  cudaMalloc((void**)&d_counts, 32);   /* int64[4]: V, F, numbered vertices, - */
  m.grid = d_logits; m.nx = nx; m.ny = nx; m.nz = nx; m.level_keys = d_keys; m.n_level_keys = 1;
  m.scratch = d_scratch; m.scratch_bytes = scratch; m.vertices = d_v; m.vertex_capacity = cap_v;
  m.faces = d_f; m.face_capacity = cap_f; m.counts = d_counts; m.voffset = nx / 2.0f; m.vscale = 1.1f / nx;
  m.phase = 3;                                                          /* count + emit */
  st = vtaco_marching_cubes(&m, NULL);
  if (st) { fprintf(stderr, "marching cubes: %s %s\n", vtaco_status_string(st), vtaco_last_cuda_error()); return 1; }
  int64_t counts[2];
  cudaDeviceSynchronize();
  cudaMemcpy(counts, d_counts, 16, D2H);
  printf("lattice %d^3 -> %lld vertices, %lld faces\n", nx, (long long)counts[0], (long long)counts[1]);
  cudaFree(d_feat); cudaFree(d_w); cudaFree(d_axis); cudaFree(d_logits); cudaFree(d_keys);
  cudaFree(d_scratch); cudaFree(d_v); cudaFree(d_f); cudaFree(d_counts);
  free(h_feat); free(h_w); free(h_axis);
  return 0;
}
