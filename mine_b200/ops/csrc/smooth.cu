// Edge-aware smoothness losses as fused kernels (forward value + analytic gradient w.r.t. the disparity map).
//
// Reference: network/layers.py:54-99 - v1 = kornia Sobel (conv3d trick) + amax + instance_norm + where + mean
// (~25 ATen launches forward, ~40 backward per call), v2 = monodepth2 mean-normalised first differences (~20 + ~30).
// Both are called on B x 1 x H x W maps at 4 pyramid levels for the source and the target view, i.e. they are
// pure launch latency.  Here: v2 = 3 launches forward (+1 backward), v1 = 3 launches forward (+2 backward).
//
//   v2: d = disp / (mean_img(disp) + 1e-7);  L = mean |d_x - d_{x+1}| e^{-mean_c |I_x - I_{x+1}|} + same in y
//   v1: e = min(|sobel(I)|_1 / (max_img * ratio), 1);  a = |sobel(disp)|;  u = instance_norm(a) - gmin;
//       L = mean( relu(u_x) (1 - e_x) + relu(u_y) (1 - e_y) )
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mine {

__device__ __forceinline__ float warp_sum_s(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sums of K values -> atomics on K global addresses
template <int K>
__device__ __forceinline__ void block_atomic_add(const float (&v)[K], float* const (&dst)[K]) {
  __shared__ float s_part[K][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float r = warp_sum_s(v[k]);
    if (lane == 0) s_part[k][wid] = r;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float t = lane < (int)(blockDim.x >> 5) ? s_part[k][lane] : 0.f;
      t = warp_sum_s(t);
      if (lane == 0 && t != 0.f) atomicAdd(dst[k], t);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// v2
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) image_sum_kernel(const float* __restrict__ x, float* __restrict__ sums, int HW) {
  const int b = blockIdx.y;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) acc += x[(size_t)b * HW + i];
  const float v[1] = {acc};
  float* const d[1] = {sums + b};
  block_atomic_add<1>(v, d);
}

__device__ __forceinline__ float img_absdiff(const float* __restrict__ img, int b, int HW, int i, int j) {
  const float* p = img + (size_t)b * 3 * HW;
  return (fabsf(p[i] - p[j]) + fabsf(p[HW + i] - p[HW + j]) + fabsf(p[2 * HW + i] - p[2 * HW + j])) * (1.f / 3.f);
}
__device__ __forceinline__ float sgn(float t) { return t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f); }

// out[0] += loss; g = dL/dd (d = normalised disparity); gd[b] += sum_i g_i * disp_i
__global__ void __launch_bounds__(256) smooth_v2_fwd_kernel(const float* __restrict__ img, const float* __restrict__ disp,
                                                            const float* __restrict__ sums, float* __restrict__ out,
                                                            float* __restrict__ g, float* __restrict__ gd, int B, int H,
                                                            int W) {
  const int b = blockIdx.y, HW = H * W;
  const float inv_m = 1.f / (sums[b] / (float)HW + 1e-7f);
  const float cx = 1.f / ((float)B * H * (W - 1)), cy = 1.f / ((float)B * (H - 1) * W);
  const float* dp = disp + (size_t)b * HW;
  float loss = 0.f, gdot = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float d0 = dp[i] * inv_m;
    float gi = 0.f;
    if (x + 1 < W) {
      const float t = d0 - dp[i + 1] * inv_m, w = __expf(-img_absdiff(img, b, HW, i, i + 1));
      loss += fabsf(t) * w * cx; gi += sgn(t) * w * cx;
    }
    if (x > 0) {
      const float t = dp[i - 1] * inv_m - d0, w = __expf(-img_absdiff(img, b, HW, i - 1, i));
      gi -= sgn(t) * w * cx;
    }
    if (y + 1 < H) {
      const float t = d0 - dp[i + W] * inv_m, w = __expf(-img_absdiff(img, b, HW, i, i + W));
      loss += fabsf(t) * w * cy; gi += sgn(t) * w * cy;
    }
    if (y > 0) {
      const float t = dp[i - W] * inv_m - d0, w = __expf(-img_absdiff(img, b, HW, i - W, i));
      gi -= sgn(t) * w * cy;
    }
    if (g) { g[(size_t)b * HW + i] = gi; gdot += gi * dp[i]; }
  }
  const float v[2] = {loss, gdot};
  float* const dst[2] = {out, gd ? gd + b : out};
  if (gd) block_atomic_add<2>(v, dst);
  else { const float v1[1] = {loss}; float* const d1[1] = {out}; block_atomic_add<1>(v1, d1); }
}

// grad_disp = gout * ( g / (m + eps) - gd_b / ((m + eps)^2 HW) )
__global__ void __launch_bounds__(256) smooth_v2_bwd_kernel(const float* __restrict__ g, const float* __restrict__ sums,
                                                            const float* __restrict__ gd, const float* __restrict__ gout,
                                                            float* __restrict__ grad, int HW, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / HW;
  const float inv_m = 1.f / (sums[b] / (float)HW + 1e-7f);
  grad[i] = gout[0] * (g[i] * inv_m - gd[b] * inv_m * inv_m / (float)HW);
}

// ---------------------------------------------------------------------------------------------------------
// v1
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sobel_at(const float* __restrict__ p, int H, int W, int y, int x, float& gx, float& gy) {
  const int ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1, xm = x > 0 ? x - 1 : 0, xp = x + 1 < W ? x + 1 : W - 1;
  const float a = p[ym * W + xm], b = p[ym * W + x], c = p[ym * W + xp];
  const float d = p[y * W + xm], f = p[y * W + xp];
  const float g = p[yp * W + xm], h = p[yp * W + x], k = p[yp * W + xp];
  gx = (c - a) + 2.f * (f - d) + (k - g);
  gy = (g - a) + 2.f * (h - b) + (k - c);
}

// stats[b] = {max_x(|sobel I|), max_y, sum ax, sum ax^2, sum ay, sum ay^2}; a maps store the SIGNED sobel of disp
__global__ void __launch_bounds__(256) smooth_v1_stats_kernel(const float* __restrict__ img, const float* __restrict__ disp,
                                                              float* __restrict__ stats, float* __restrict__ sob, int H,
                                                              int W) {
  const int b = blockIdx.y, HW = H * W;
  float mx = 0.f, my = 0.f, s1x = 0.f, s2x = 0.f, s1y = 0.f, s2y = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    float ex = 0.f, ey = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gx, gy;
      sobel_at(img + ((size_t)b * 3 + c) * HW, H, W, y, x, gx, gy);
      ex += fabsf(gx) * 0.125f; ey += fabsf(gy) * 0.125f;
    }
    mx = fmaxf(mx, ex); my = fmaxf(my, ey);
    float gx, gy;
    sobel_at(disp + (size_t)b * HW, H, W, y, x, gx, gy);
    sob[((size_t)b * 2 + 0) * HW + i] = gx; sob[((size_t)b * 2 + 1) * HW + i] = gy;
    const float ax = fabsf(gx), ay = fabsf(gy);
    s1x += ax; s2x += ax * ax; s1y += ay; s2y += ay * ay;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, o)); }
  if ((threadIdx.x & 31) == 0) {           // values are >= 0: integer compare == float compare
    atomicMax(reinterpret_cast<int*>(stats + b * 6 + 0), __float_as_int(mx));
    atomicMax(reinterpret_cast<int*>(stats + b * 6 + 1), __float_as_int(my));
  }
  const float v[4] = {s1x, s2x, s1y, s2y};
  float* const dst[4] = {stats + b * 6 + 2, stats + b * 6 + 3, stats + b * 6 + 4, stats + b * 6 + 5};
  block_atomic_add<4>(v, dst);
}

// loss; when hmap != null also h = dL/du (per direction) and hs[b] = {sum hx, sum hx*uhat_x, sum hy, sum hy*uhat_y}
__global__ void __launch_bounds__(256) smooth_v1_loss_kernel(const float* __restrict__ img, const float* __restrict__ sob,
                                                             const float* __restrict__ stats, float gmin, float ratio,
                                                             float* __restrict__ out, float* __restrict__ hmap,
                                                             float* __restrict__ hs, int B, int H, int W) {
  const int b = blockIdx.y, HW = H * W;
  const float* st = stats + b * 6;
  const float n = (float)HW;
  const float mux = st[2] / n, muy = st[4] / n;
  const float isx = rsqrtf(fmaxf(st[3] / n - mux * mux, 0.f) + 1e-5f), isy = rsqrtf(fmaxf(st[5] / n - muy * muy, 0.f) + 1e-5f);
  const float kx = 1.f / (st[0] * ratio), ky = 1.f / (st[1] * ratio), cnorm = 1.f / ((float)B * HW);
  float loss = 0.f, hx1 = 0.f, hx2 = 0.f, hy1 = 0.f, hy2 = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    float ex = 0.f, ey = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gx, gy;
      sobel_at(img + ((size_t)b * 3 + c) * HW, H, W, y, x, gx, gy);
      ex += fabsf(gx) * 0.125f; ey += fabsf(gy) * 0.125f;
    }
    const float wx = 1.f - fminf(ex * kx, 1.f), wy = 1.f - fminf(ey * ky, 1.f);
    const float uhx = (fabsf(sob[((size_t)b * 2) * HW + i]) - mux) * isx, uhy = (fabsf(sob[((size_t)b * 2 + 1) * HW + i]) - muy) * isy;
    const float ux = uhx - gmin, uy = uhy - gmin;
    loss += (fmaxf(ux, 0.f) * wx + fmaxf(uy, 0.f) * wy) * cnorm;
    if (hmap) {
      const float hx = ux > 0.f ? wx * cnorm : 0.f, hy = uy > 0.f ? wy * cnorm : 0.f;
      hmap[((size_t)b * 2) * HW + i] = hx; hmap[((size_t)b * 2 + 1) * HW + i] = hy;
      hx1 += hx; hx2 += hx * uhx; hy1 += hy; hy2 += hy * uhy;
    }
  }
  if (hmap) {
    const float v[5] = {loss, hx1, hx2, hy1, hy2};
    float* const dst[5] = {out, hs + b * 4, hs + b * 4 + 1, hs + b * 4 + 2, hs + b * 4 + 3};
    block_atomic_add<5>(v, dst);
  } else {
    const float v[1] = {loss};
    float* const dst[1] = {out};
    block_atomic_add<1>(v, dst);
  }
}

// instance-norm backward + |.| + adjoint of the replicate-padded Sobel stencil (scatter with atomics into grad, pre-zeroed)
__global__ void __launch_bounds__(256) smooth_v1_bwd_kernel(const float* __restrict__ sob, const float* __restrict__ stats,
                                                            const float* __restrict__ hmap, const float* __restrict__ hs,
                                                            const float* __restrict__ gout, float* __restrict__ grad, int H,
                                                            int W) {
  const int b = blockIdx.y, HW = H * W;
  const float* st = stats + b * 6;
  const float n = (float)HW;
  const float mux = st[2] / n, muy = st[4] / n;
  const float isx = rsqrtf(fmaxf(st[3] / n - mux * mux, 0.f) + 1e-5f), isy = rsqrtf(fmaxf(st[5] / n - muy * muy, 0.f) + 1e-5f);
  const float mhx = hs[b * 4] / n, mhxu = hs[b * 4 + 1] / n, mhy = hs[b * 4 + 2] / n, mhyu = hs[b * 4 + 3] / n;
  const float go = gout[0];
  float* gp = grad + (size_t)b * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float sx = sob[((size_t)b * 2) * HW + i], sy = sob[((size_t)b * 2 + 1) * HW + i];
    const float uhx = (fabsf(sx) - mux) * isx, uhy = (fabsf(sy) - muy) * isy;
    const float dax = isx * (hmap[((size_t)b * 2) * HW + i] - mhx - uhx * mhxu);
    const float day = isy * (hmap[((size_t)b * 2 + 1) * HW + i] - mhy - uhy * mhyu);
    const float dsx = go * dax * sgn(sx), dsy = go * day * sgn(sy);
    if (dsx == 0.f && dsy == 0.f) continue;
    const int ym = y > 0 ? y - 1 : 0, yp = y + 1 < H ? y + 1 : H - 1, xm = x > 0 ? x - 1 : 0, xp = x + 1 < W ? x + 1 : W - 1;
    // gx = (c - a) + 2 (f - d) + (k - g);  gy = (g - a) + 2 (h - b) + (k - c)   (a..k = 3x3 neighbourhood, row major)
    atomicAdd(gp + ym * W + xm, -dsx - dsy);
    atomicAdd(gp + ym * W + x, -2.f * dsy);
    atomicAdd(gp + ym * W + xp, dsx - dsy);
    atomicAdd(gp + y * W + xm, -2.f * dsx);
    atomicAdd(gp + y * W + xp, 2.f * dsx);
    atomicAdd(gp + yp * W + xm, -dsx + dsy);
    atomicAdd(gp + yp * W + x, 2.f * dsy);
    atomicAdd(gp + yp * W + xp, dsx + dsy);
  }
}

static dim3 img_grid(int HW, int B) {
  int bx = (HW + 255) / 256;
  if (bx > 64) bx = 64;
  return dim3(bx, B);
}

void launch_smooth_v2_fwd(const float* img, const float* disp, float* sums, float* out, float* g, float* gd, int B, int H,
                          int W, cudaStream_t stream) {
  image_sum_kernel<<<img_grid(H * W, B), 256, 0, stream>>>(disp, sums, H * W);
  smooth_v2_fwd_kernel<<<img_grid(H * W, B), 256, 0, stream>>>(img, disp, sums, out, g, gd, B, H, W);
}
void launch_smooth_v2_bwd(const float* g, const float* sums, const float* gd, const float* gout, float* grad, int B, int HW,
                          cudaStream_t stream) {
  const int total = B * HW;
  smooth_v2_bwd_kernel<<<(total + 255) / 256, 256, 0, stream>>>(g, sums, gd, gout, grad, HW, total);
}
void launch_smooth_v1_fwd(const float* img, const float* disp, float* stats, float* sob, float* out, float* hmap, float* hs,
                          float gmin, float ratio, int B, int H, int W, cudaStream_t stream) {
  smooth_v1_stats_kernel<<<img_grid(H * W, B), 256, 0, stream>>>(img, disp, stats, sob, H, W);
  smooth_v1_loss_kernel<<<img_grid(H * W, B), 256, 0, stream>>>(img, sob, stats, gmin, ratio, out, hmap, hs, B, H, W);
}
void launch_smooth_v1_bwd(const float* sob, const float* stats, const float* hmap, const float* hs, const float* gout,
                          float* grad, int B, int H, int W, cudaStream_t stream) {
  smooth_v1_bwd_kernel<<<img_grid(H * W, B), 256, 0, stream>>>(sob, stats, hmap, hs, gout, grad, H, W);
}

}  // namespace mine
