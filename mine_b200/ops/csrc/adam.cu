// Fused multi-tensor Adam over a flat fp32 arena (one launch per learning-rate group).
// torch.optim.Adam semantics with L2 weight decay folded into the gradient (reference
// synthesis_task.py:87; torch 1.8 ran ~8 launches per parameter tensor x 221 tensors).
#include <cuda_runtime.h>

#include "kernels.h"

namespace mine {

// hyper = {lr, step} in device memory: bias corrections are derived in-kernel, so a captured CUDA graph
// replays with the current step / learning rate without re-recording.
__device__ __forceinline__ void adam_coefs(const float* __restrict__ hyper, float beta1, float beta2, float& lr_over_bc1,
                                           float& inv_sqrt_bc2) {
  const float lr = hyper[0], step = hyper[1];
  lr_over_bc1 = lr / (1.0f - powf(beta1, step));
  inv_sqrt_bc2 = rsqrtf(1.0f - powf(beta2, step));
}

__global__ void __launch_bounds__(256) fused_adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                         float4* __restrict__ m, float4* __restrict__ v, int64_t n4,
                                                         const float* __restrict__ hyper, float beta1, float beta2,
                                                         float eps, float wd) {
  float lr_over_bc1, inv_sqrt_bc2;
  adam_coefs(hyper, beta1, beta2, lr_over_bc1, inv_sqrt_bc2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float grad = G[k] + wd * P[k];
      M[k] = beta1 * M[k] + (1.f - beta1) * grad;
      V[k] = beta2 * V[k] + (1.f - beta2) * grad * grad;
      const float denom = sqrtf(V[k]) * inv_sqrt_bc2 + eps;
      P[k] -= lr_over_bc1 * (M[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

__global__ void fused_adam_tail_kernel(float* p, const float* g, float* m, float* v, int64_t start, int64_t n,
                                       const float* __restrict__ hyper, float beta1, float beta2, float eps, float wd) {
  float lr_over_bc1, inv_sqrt_bc2;
  adam_coefs(hyper, beta1, beta2, lr_over_bc1, inv_sqrt_bc2);
  const int64_t i = start + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float grad = g[i] + wd * p[i];
  const float mm = beta1 * m[i] + (1.f - beta1) * grad;
  const float vv = beta2 * v[i] + (1.f - beta2) * grad * grad;
  m[i] = mm; v[i] = vv;
  p[i] -= lr_over_bc1 * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
}

void launch_fused_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1,
                       float beta2, float eps, float weight_decay, cudaStream_t stream) {
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    fused_adam_kernel<<<blocks, 256, 0, stream>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n4,
                                                  hyper, beta1, beta2, eps, weight_decay);
  }
  const int64_t done = n4 * 4;
  if (done < n) {
    const int64_t rem = n - done;
    fused_adam_tail_kernel<<<(int)((rem + 255) / 256), 256, 0, stream>>>(p, g, m, v, done, n, hyper, beta1,
                                                                         beta2, eps, weight_decay);
  }
}

}  // namespace mine
