// Sparse-point supervision in two launches (forward, backward) per view and pyramid level.
//
// Reference (synthesis_task.py:211-220, 276-323): project N COLMAP points with K, round to the nearest pixel, gather
// the synthesised disparity, calibrate the per-image scale  s = exp(mean_n(log d_n - log g_n))  (g = 1 / z of the
// point; scale 0, source view only), and take  mean_{b,n} |log(d_n / s_b) - log g_n|.  As framework ops this is ~35
// tiny launches forward plus as many backward, 8 times per step.  One CTA per image does all of it; the backward
// handles both gradient paths of the (differentiable) scale factor and scatters into the disparity map with
// atomics (points that round to the same pixel accumulate, like the gather's autograd).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mine {

namespace {

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                               // s_red may still be read from a previous call
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
  return t;
}

// disp [B,H,W]; K [B,3,3]; xyz [B,3,N]; scale_in [B] or null.  Outputs: idx [B,N], d_syn [B,N], sgn [B,N],
// scale_out [B], loss (+= sum |term| / (B*N)).
__global__ void __launch_bounds__(256) sparse_point_fwd_kernel(
    const float* __restrict__ disp, const float* __restrict__ K, const float* __restrict__ xyz,
    const float* __restrict__ scale_in, int* __restrict__ idx, float* __restrict__ d_syn, float* __restrict__ sgn,
    float* __restrict__ scale_out, float* __restrict__ loss, int B, int H, int W, int N) {
  __shared__ float s_red[8];
  const int b = blockIdx.x;
  const float* k = K + b * 9;
  const float k00 = k[0], k01 = k[1], k02 = k[2], k10 = k[3], k11 = k[4], k12 = k[5], k20 = k[6], k21 = k[7], k22 = k[8];
  const float* px = xyz + (size_t)b * 3 * N;
  const float* dmap = disp + (size_t)b * H * W;
  float acc = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float X = px[n], Y = px[N + n], Z = px[2 * N + n];
    const float pz = k20 * X + k21 * Y + k22 * Z;
    const float u = (k00 * X + k01 * Y + k02 * Z) / pz, v = (k10 * X + k11 * Y + k12 * Z) / pz;
    int ix = (int)rintf(u), iy = (int)rintf(v);                    // half-to-even, like torch.round
    ix = min(max(ix, 0), W - 1); iy = min(max(iy, 0), H - 1);
    const int id = iy * W + ix;
    const float d = dmap[id];
    idx[(size_t)b * N + n] = id;
    d_syn[(size_t)b * N + n] = d;
    acc += logf(d) - logf(1.f / Z);
  }
  float s;
  if (scale_in) {
    s = scale_in[b];
  } else {
    const float tot = block_sum(acc, s_red);
    s = expf(tot / (float)N);
  }
  if (threadIdx.x == 0) scale_out[b] = s;
  float lsum = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float d = d_syn[(size_t)b * N + n];
    const float t = logf(d / s) - logf(1.f / px[2 * N + n]);
    sgn[(size_t)b * N + n] = t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f);
    lsum += fabsf(t);
  }
  const float tot = block_sum(lsum, s_red);
  if (threadIdx.x == 0) atomicAdd(loss, tot / ((float)B * (float)N));
}

// g_loss [1]; g_scale [B] or null (gradient arriving at scale_out; only meaningful when the scale was computed here);
// grad_disp [B,H,W] zero-initialised; grad_scale_in [B] (written when the scale was an input).
__global__ void __launch_bounds__(256) sparse_point_bwd_kernel(
    const float* __restrict__ g_loss, const float* __restrict__ g_scale, const int* __restrict__ idx,
    const float* __restrict__ d_syn, const float* __restrict__ sgn, const float* __restrict__ scale, float* __restrict__ grad_disp,
    float* __restrict__ grad_scale_in, int computed_scale, int B, int H, int W, int N) {
  __shared__ float s_red[8];
  const int b = blockIdx.x;
  const float c = g_loss[0] / ((float)B * (float)N);
  const float s = scale[b];
  float ssum = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) ssum += sgn[(size_t)b * N + n];
  const float sign_total = block_sum(ssum, s_red);
  const float dL_ds = -c * sign_total / s;                         // own loss term through log(d / s)
  float through_scale = 0.f;
  if (computed_scale) {
    const float gs = (g_scale ? g_scale[b] : 0.f) + dL_ds;
    through_scale = gs * s / (float)N;                             // d s / d d_n = s / (N d_n)
  } else if (threadIdx.x == 0 && grad_scale_in) {
    grad_scale_in[b] = dL_ds;
  }
  float* gmap = grad_disp + (size_t)b * H * W;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float d = d_syn[(size_t)b * N + n];
    const float gd = (c * sgn[(size_t)b * N + n] + through_scale) / d;
    atomicAdd(gmap + idx[(size_t)b * N + n], gd);
  }
}

}  // namespace

void launch_sparse_point_fwd(const float* disp, const float* K, const float* xyz, const float* scale_in, int* idx,
                             float* d_syn, float* sgn, float* scale_out, float* loss, int B, int H, int W, int N,
                             cudaStream_t stream) {
  sparse_point_fwd_kernel<<<B, 256, 0, stream>>>(disp, K, xyz, scale_in, idx, d_syn, sgn, scale_out, loss, B, H, W, N);
}

void launch_sparse_point_bwd(const float* g_loss, const float* g_scale, const int* idx, const float* d_syn,
                             const float* sgn, const float* scale, float* grad_disp, float* grad_scale_in,
                             int computed_scale, int B, int H, int W, int N, cudaStream_t stream) {
  sparse_point_bwd_kernel<<<B, 256, 0, stream>>>(g_loss, g_scale, idx, d_syn, sgn, scale, grad_disp, grad_scale_in,
                                                 computed_scale, B, H, W, N);
}

}  // namespace mine
