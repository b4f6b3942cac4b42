// 2-D U-Net glue on channels-last feature planes (reference src/encoder/unet.py:45-114):
// MaxPool2d(2) and the interleave half of ConvTranspose2d(kernel 2, stride 2).  The convolutions of that
// network run on vtaco_conv3d_cl (a plane is a volume of depth 1: the z-taps of a 3x3x3 filter meet the zero
// padding, a 2-D filter is its middle slice); a transposed 2x2 / stride-2 convolution is a 1x1 convolution to
// 4*Cout channels — one (a,b) sub-pixel per channel group — followed by the depth-to-space kernel below.
#include "common.cuh"
#include <math_constants.h>

namespace vtaco {

// y[n][yo][xo][c] = max over the 2x2 window of x[n][2yo+..][2xo+..][c]
__global__ void __launch_bounds__(256) maxpool2d_cl_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H,
                                                           int W, int C) {
  const int C4 = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * C4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % C4);
    long long v = t / C4;
    const int xo = (int)(v % Wo); v /= Wo;
    const int yo = (int)(v % Ho);
    const int n = (int)(v / Ho);
    float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yi = 2 * yo + (k >> 1), xi = 2 * xo + (k & 1);
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + yi) * W + xi) * C) + q);
      m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
    }
    reinterpret_cast<float4*>(y + (((size_t)n * Ho + yo) * Wo + xo) * C)[q] = m;
  }
}

// x [n][i][j][(a*2+b)*C + c]  ->  y [n][2i+a][2j+b][c]
__global__ void __launch_bounds__(256) depth_to_space2_cl_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                                                 int H, int W, int C) {
  const int C4 = C >> 2;
  const long long total = (long long)N * H * W * 4 * C4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t % C4);
    long long v = t / C4;
    const int ab = (int)(v & 3); v >>= 2;
    const int j = (int)(v % W); v /= W;
    const int i = (int)(v % H);
    const int n = (int)(v / H);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + ((((size_t)n * H + i) * W + j) * 4 + ab) * C) + q);
    reinterpret_cast<float4*>(y + (((size_t)n * 2 * H + 2 * i + (ab >> 1)) * 2 * W + 2 * j + (ab & 1)) * C)[q] = a;
  }
}

}  // namespace vtaco

using namespace vtaco;

extern "C" int vtaco_maxpool2d_cl(const float* x, float* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!x || !y || N < 1 || H < 2 || W < 2 || C < 4 || (C & 3)) return VTACO_ERR_INVALID_ARG;
  if ((H | W) & 1) return VTACO_ERR_UNSUPPORTED;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  maxpool2d_cl_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}

extern "C" int vtaco_depth_to_space2_cl(const float* x, float* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!x || !y || N < 1 || H < 1 || W < 1 || C < 4 || (C & 3)) return VTACO_ERR_INVALID_ARG;
  const long long total = (long long)N * H * W * C;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  depth_to_space2_cl_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C);
  VTACO_LAUNCH_CHECK();
  return VTACO_OK;
}
