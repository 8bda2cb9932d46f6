// Fused loss kernels: SSIM (forward + analytic backward) and masked L1.
//
// Reference SSIM = 5 depthwise 11x11 conv2d launches + ~15 elementwise kernels per call, and an
// autograd graph of the same size for the backward (network/ssim.py:19-39).  Here one kernel per
// direction: a 16x16 output tile with a 5-pixel halo is staged in shared memory, the Gaussian is
// applied separably (rows then columns) to the five moments at once, the SSIM map is reduced on the
// fly (warp shuffle -> one atomic per block), and - when a gradient is needed - the three adjoint
// maps are stored so that the backward is a second, identical stencil pass.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mine {

constexpr int kWin = 11;
constexpr int kHalo = 5;
constexpr int kTile = 16;
constexpr int kIn = kTile + 2 * kHalo;   // 26
constexpr float kC1 = 0.01f * 0.01f;
constexpr float kC2 = 0.03f * 0.03f;

__constant__ float c_gauss[kWin];

static void upload_gauss_once() {
  static bool done = false;
  if (done) return;
  float g[kWin];
  double s = 0.0;
  for (int i = 0; i < kWin; ++i) { double d = i - kWin / 2; g[i] = (float)exp(-d * d / (2.0 * 1.5 * 1.5)); s += g[i]; }
  // match the reference: window built in float32 (gauss / gauss.sum())
  float fs = 0.f;
  for (int i = 0; i < kWin; ++i) fs += g[i];
  for (int i = 0; i < kWin; ++i) g[i] = g[i] / fs;
  cudaMemcpyToSymbol(c_gauss, g, sizeof(g));
  done = true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void block_atomic_sum(float v, float* out) {
  __shared__ float s_part[8];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = (lane < (int)(blockDim.x >> 5)) ? s_part[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(out, t);
  }
}

// Separable blur of K maps held in s_in[K][kIn][kIn+1]; result for this thread's pixel in out[K].
template <int K>
__device__ __forceinline__ void separable_blur(float (*s_in)[kIn][kIn + 1], float (*s_mid)[kIn][kTile], float* out) {
  const int tid = threadIdx.x;
  // horizontal pass: kIn rows x kTile columns per map
  for (int idx = tid; idx < kIn * kTile; idx += blockDim.x) {
    const int r = idx / kTile, c = idx - r * kTile;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < kWin; ++j) acc += c_gauss[j] * s_in[k][r][c + j];
      s_mid[k][r][c] = acc;
    }
  }
  __syncthreads();
  const int ty = tid / kTile, tx = tid - ty * kTile;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kWin; ++j) acc += c_gauss[j] * s_mid[k][ty + j][tx];
    out[k] = acc;
  }
}

__global__ void __launch_bounds__(256) ssim_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ sum_out, float* __restrict__ partials,
                                                       int H, int W) {
  __shared__ float s_in[5][kIn][kIn + 1];
  __shared__ float s_mid[5][kIn][kTile];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const float* pa = a + (size_t)plane * H * W;
  const float* pb = b + (size_t)plane * H * W;
  for (int idx = threadIdx.x; idx < kIn * kIn; idx += blockDim.x) {
    const int r = idx / kIn, c = idx - r * kIn;
    const int y = y0 + r - kHalo, x = x0 + c - kHalo;
    float va = 0.f, vb = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) { va = pa[y * W + x]; vb = pb[y * W + x]; }
    s_in[0][r][c] = va; s_in[1][r][c] = vb; s_in[2][r][c] = va * va; s_in[3][r][c] = vb * vb; s_in[4][r][c] = va * vb;
  }
  __syncthreads();
  float m[5];
  separable_blur<5>(s_in, s_mid, m);
  const int ty = threadIdx.x / kTile, tx = threadIdx.x - ty * kTile;
  const int y = y0 + ty, x = x0 + tx;
  float val = 0.f;
  if (y < H && x < W) {
    const float mu_a = m[0], mu_b = m[1];
    const float var_a = m[2] - mu_a * mu_a, var_b = m[3] - mu_b * mu_b, cov = m[4] - mu_a * mu_b;
    const float n1 = 2.f * mu_a * mu_b + kC1, n2 = 2.f * cov + kC2;
    const float d1 = mu_a * mu_a + mu_b * mu_b + kC1, d2 = var_a + var_b + kC2;
    val = (n1 * n2) / (d1 * d2);
    if (partials) {
      const float df_dmu = (2.f * mu_b * d1 - n1 * 2.f * mu_a) / (d1 * d1) * (n2 / d2);
      const float df_dvar = -(n1 * n2) / (d1 * d2 * d2);
      const float df_dcov = 2.f * n1 / (d1 * d2);
      float* pp = partials + (size_t)plane * 3 * H * W + y * W + x;
      pp[0] = df_dmu - 2.f * mu_a * df_dvar - mu_b * df_dcov;
      pp[(size_t)H * W] = df_dvar;
      pp[(size_t)2 * H * W] = df_dcov;
    }
  }
  block_atomic_sum(val, sum_out);
}

__global__ void __launch_bounds__(256) ssim_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ partials,
                                                       const float* __restrict__ scale_ptr, float scale_mul,
                                                       float* __restrict__ grad_a, int H, int W) {
  __shared__ float s_in[3][kIn][kIn + 1];
  __shared__ float s_mid[3][kIn][kTile];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const float* pp = partials + (size_t)plane * 3 * H * W;
  for (int idx = threadIdx.x; idx < kIn * kIn; idx += blockDim.x) {
    const int r = idx / kIn, c = idx - r * kIn;
    const int y = y0 + r - kHalo, x = x0 + c - kHalo;
    const bool in = (y >= 0 && y < H && x >= 0 && x < W);
#pragma unroll
    for (int k = 0; k < 3; ++k) s_in[k][r][c] = in ? pp[(size_t)k * H * W + y * W + x] : 0.f;
  }
  __syncthreads();
  float m[3];
  separable_blur<3>(s_in, s_mid, m);
  const int ty = threadIdx.x / kTile, tx = threadIdx.x - ty * kTile;
  const int y = y0 + ty, x = x0 + tx;
  if (y < H && x < W) {
    const size_t o = (size_t)plane * H * W + y * W + x;
    const float scale = (scale_ptr ? *scale_ptr : 1.0f) * scale_mul;
    grad_a[o] = scale * (m[0] + 2.f * a[o] * m[1] + b[o] * m[2]);
  }
}

__global__ void __launch_bounds__(256) masked_l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ mask, float thr,
                                                            float* __restrict__ sum_out, float* __restrict__ grad_sign,
                                                            int C, int HW, int total) {
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int bimg = i / (C * HW);
    const int pix = i % HW;
    const float valid = mask[bimg * HW + pix] >= thr ? 1.f : 0.f;
    const float d = a[i] - b[i];
    acc += fabsf(d) * valid;
    if (grad_sign) grad_sign[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * valid;
  }
  block_atomic_sum(acc, sum_out);
}

void launch_ssim_fwd(const float* a, const float* b, float* sum_out, float* partials, int planes, int H, int W,
                     cudaStream_t stream) {
  upload_gauss_once();
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, planes);
  ssim_fwd_kernel<<<grid, 256, 0, stream>>>(a, b, sum_out, partials, H, W);
}

void launch_ssim_bwd(const float* a, const float* b, const float* partials, const float* scale_ptr, float scale_mul,
                     float* grad_a, int planes, int H, int W, cudaStream_t stream) {
  upload_gauss_once();
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, planes);
  ssim_bwd_kernel<<<grid, 256, 0, stream>>>(a, b, partials, scale_ptr, scale_mul, grad_a, H, W);
}

void launch_masked_l1_fwd(const float* a, const float* b, const float* mask, float thr, float* sum_out,
                          float* grad_sign, int B, int C, int HW, cudaStream_t stream) {
  const int total = B * C * HW;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  masked_l1_fwd_kernel<<<blocks, 256, 0, stream>>>(a, b, mask, thr, sum_out, grad_sign, C, HW, total);
}

}  // namespace mine
