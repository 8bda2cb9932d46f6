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


// ---- gradient gather: many autograd-owned gradient tensors -> their slices of the flat gradient arena --------------
// The framework's AccumulateGrad node adds every incoming gradient into a pre-set ``.grad`` with one elementwise kernel
// per parameter (~320 launches, 0.9 ms per step in round 2's CUPTI table).  With ``.grad = None`` it simply keeps the
// incoming tensor; this kernel then moves a whole list of them into the arena at once (tensor / chunk tables in
// kernel-parameter space, multi-tensor-apply style; a null source zero-fills its slice).
constexpr int kMcTensors = 48, kMcBlocks = 448, kMcChunk = 32768;     // floats per chunk
struct MultiCopyMeta {
  const float* src[kMcTensors];
  float* dst[kMcTensors];
  int numel[kMcTensors];
  uint8_t block_tensor[kMcBlocks];
  uint16_t block_chunk[kMcBlocks];
};

__global__ void __launch_bounds__(256) multi_copy_kernel(const __grid_constant__ MultiCopyMeta m) {
  const int t = m.block_tensor[blockIdx.x];
  const int c0 = (int)m.block_chunk[blockIdx.x] * kMcChunk;
  const int n = min(kMcChunk, m.numel[t] - c0);
  const float* __restrict__ src = m.src[t] ? m.src[t] + c0 : nullptr;
  float* __restrict__ dst = m.dst[t] + c0;
  const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (!src || (reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = src ? reinterpret_cast<const float4*>(src)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(dst)[i] = v;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = src ? src[i] : 0.f;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src ? src[i] : 0.f;
  }
}

void launch_multi_copy(const float* const* srcs, float* const* dsts, const int64_t* numels, int count, cudaStream_t stream) {
  MultiCopyMeta m;
  int nt = 0, nb = 0;
  auto flush = [&]() {
    if (nb > 0) multi_copy_kernel<<<nb, 256, 0, stream>>>(m);
    nt = 0; nb = 0;
  };
  for (int i = 0; i < count; ++i) {
    int64_t done = 0;
    while (done < numels[i]) {
      if (nt == kMcTensors || nb == kMcBlocks) flush();
      // at most 65535 chunks per table entry and as many blocks as still fit into this launch
      const int64_t left = numels[i] - done;
      int64_t chunks = (left + kMcChunk - 1) / kMcChunk;
      if (chunks > kMcBlocks - nb) chunks = kMcBlocks - nb;
      const int64_t take = chunks * kMcChunk < left ? chunks * kMcChunk : left;
      m.src[nt] = srcs[i] ? srcs[i] + done : nullptr;
      m.dst[nt] = dsts[i] + done;
      m.numel[nt] = (int)take;
      for (int c = 0; c < (int)chunks; ++c) { m.block_tensor[nb] = (uint8_t)nt; m.block_chunk[nb] = (uint16_t)c; ++nb; }
      ++nt;
      done += take;
    }
  }
  flush();
}

}  // namespace mine
