// Memory-bound companions of the conv engine for the ResNet encoder (NHWC bf16, 16-byte vectors = 8 channels,
// C a power of two in [16, 2048]).  Same structure as decoder_elem.cu without the padded-buffer bookkeeping:
//
//   bn_res_act_fwd        : out = act(BN(y) [+ residual]) from the conv epilogue's batch sums; act(v) = v > 0 ? v :
//                           slope * v covers ReLU (0), LeakyReLU (0.1) and identity (1) - replaces ATen
//                           batch_norm_elemt + add + relu (three passes) with one read of y (+ residual), one write.
//   bn_res_act_bwd_reduce : g = dout * [out > 0] (also the gradient of the residual branch) and the two BatchNorm
//                           backward sums per channel; the apply step is bn_bwd_apply of decoder_elem.cu.
//   channel_stats         : per-channel sum / sum of squares for a tensor produced outside the engine (stem conv).
//
// Every thread keeps one 8-channel group for its whole grid-stride loop (the stride is a multiple of C/8), so
// per-channel coefficients and partial sums live in registers; block partials are combined in shared memory and
// published with one atomic per channel per block.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "act_types.cuh"
#include "ll_exchange.cuh"
#include "kernels.h"

namespace mine {

namespace {

// grid-stride over `total` 8-channel vectors with a stride that preserves (index mod cg)
struct Walk {
  unsigned i0, stride, total;
  int c0;
  bool active;
};
__device__ __forceinline__ Walk make_walk(unsigned total, int C) {
  const int cg = C >> 3;
  const int cg_shift = 31 - __clz(cg);
  Walk w;
  w.total = total;
  w.stride = ((gridDim.x * blockDim.x) >> cg_shift) << cg_shift;
  w.i0 = blockIdx.x * blockDim.x + threadIdx.x;
  w.active = w.i0 < w.stride;
  w.c0 = (int)(w.i0 & (unsigned)(cg - 1)) * 8;
  return w;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_res_act_fwd_kernel(
    const T* __restrict__ y, const float* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, const T* __restrict__ res, T* __restrict__ out,
    unsigned total, int C, float slope, float inv_count, float eps, const LLExchange x, float* __restrict__ red_out) {
  extern __shared__ float s_red[];       // [2C] cross-GPU reduced statistics (only when x.world > 1)
  if (x.world > 1) { ll_exchange_sum(stats, s_red, 2 * C, x, red_out); stats = s_red; }
  const Walk w = make_walk(total, C);
  if (!w.active) return;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = stats[w.c0 + j] * inv_count;
    float var = stats[C + w.c0 + j] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    a[j] = gamma[w.c0 + j] * rsqrtf(var + eps);
    b[j] = beta[w.c0 + j] - m * a[j];
  }
  for (unsigned i = w.i0; i < w.total; i += w.stride) {
    const size_t o = (size_t)i * 8;
    V8 v = ld8(y + o);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.f[j] = v.f[j] * a[j] + b[j];
    if (res) {
      const V8 r = ld8(res + o);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.f[j] += r.f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v.f[j] = v.f[j] > 0.f ? v.f[j] : slope * v.f[j];
    st8_op(out + o, v);
  }
}

// publishes acc1/acc2 (8 channels starting at c0) through a [2][C] shared buffer
__device__ __forceinline__ void publish_sums(float* s_sum, float* __restrict__ sums, int C, int c0, bool active,
                                             const float* acc1, const float* acc2) {
  // lanes with equal (lane % cg) own the same channels: fold the warp with shuffles first (narrow layers), then one
  // shared-memory atomic per channel and warp instead of one per thread
  const int cg = C >> 3, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v1 = active ? acc1[j] : 0.f, v2 = active ? acc2[j] : 0.f;
    for (int o2 = 16; o2 >= cg && o2 > 0; o2 >>= 1) {
      v1 += __shfl_xor_sync(0xffffffffu, v1, o2);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o2);
    }
    if (active && (lane < cg || cg >= 32)) { atomicAdd(&s_sum[c0 + j], v1); atomicAdd(&s_sum[C + c0 + j], v2); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const float v = s_sum[i];
    if (v != 0.f) atomicAdd(&sums[i], v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_res_act_bwd_reduce_kernel(
    const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ y,
    const float* __restrict__ stats, T* __restrict__ g_out, float* __restrict__ sums, unsigned total, int C,
    float slope, float inv_count, float eps) {
  extern __shared__ float s_sum[];       // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i] = 0.f;
  __syncthreads();
  const Walk w = make_walk(total, C);
  float mean[8], invstd[8], acc1[8], acc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = stats[w.c0 + j] * inv_count;
    float var = stats[C + w.c0 + j] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    mean[j] = m; invstd[j] = rsqrtf(var + eps);
    acc1[j] = acc2[j] = 0.f;
  }
  if (w.active) {
    for (unsigned i = w.i0; i < w.total; i += w.stride) {
      const size_t o = (size_t)i * 8;
      V8 g = ld8(dout + o);
      if (slope != 1.f) {
        const V8 a = ld8(out + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g.f[j] = a.f[j] > 0.f ? g.f[j] : slope * g.f[j];
      }
      const V8 yv = ld8(y + o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc1[j] += g.f[j];
        acc2[j] += g.f[j] * (yv.f[j] - mean[j]) * invstd[j];
      }
      st8(g_out + o, g);
    }
  }
  publish_sums(s_sum, sums, C, w.c0, w.active, acc1, acc2);
}

template <typename T>
__global__ void __launch_bounds__(256) channel_stats_kernel(const T* __restrict__ y, float* __restrict__ sums,
                                                            unsigned total, int C) {
  extern __shared__ float s_sum[];       // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i] = 0.f;
  __syncthreads();
  const Walk w = make_walk(total, C);
  float acc1[8], acc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc1[j] = acc2[j] = 0.f;
  if (w.active) {
    for (unsigned i = w.i0; i < w.total; i += w.stride) {
      const V8 v = ld8(y + (size_t)i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc1[j] += v.f[j]; acc2[j] += v.f[j] * v.f[j]; }
    }
  }
  publish_sums(s_sum, sums, C, w.c0, w.active, acc1, acc2);
}

// running_mean/var <- (1 - m) * running + m * batch statistic (unbiased variance), num_batches_tracked += 1:
// one launch instead of the ~11 framework ops per BatchNorm layer and step
__global__ void bn_update_running_kernel(const float* __restrict__ stats, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, long long* __restrict__ num_batches, int C,
                                         float inv_count, float unbias, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches) *num_batches += 1;
  if (c >= C) return;
  const float mean = stats[c] * inv_count;
  float var = stats[C + c] * inv_count - mean * mean;
  var = var < 0.f ? 0.f : var;
  running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
}

int blocks_for(size_t total, int C, int cap) {
  size_t b = (total + 255) / 256;
  if (b > (size_t)cap) b = cap;
  const size_t min_blocks = ((size_t)(C / 8) + 255) / 256;       // the stride must cover one full channel period
  if (b < min_blocks) b = min_blocks;
  return (int)(b == 0 ? 1 : b);
}

}  // namespace

void launch_bn_res_act_fwd(const void* y, const float* stats, const float* gamma, const float* beta, const void* res,
                           void* out, size_t npix, int C, float slope, float inv_count, float eps, int es,
                           const LLExchange* x, float* red_out, cudaStream_t stream) {
  const size_t total = npix * (size_t)(C / 8);
  LLExchange xx{};
  if (x) xx = *x;
  const size_t smem = xx.world > 1 ? 2 * (size_t)C * sizeof(float) : 0;
  MINE_DISPATCH_ES(es, T, (bn_res_act_fwd_kernel<T><<<blocks_for(total, C, 148 * 16), 256, smem, stream>>>(
      (const T*)y, stats, gamma, beta, (const T*)res, (T*)out, (unsigned)total, C, slope, inv_count, eps, xx, red_out)));
}

void launch_bn_res_act_bwd_reduce(const void* dout, const void* out, const void* y, const float* stats, void* g_out,
                                  float* sums, size_t npix, int C, float slope, float inv_count, float eps, int es,
                                  cudaStream_t stream) {
  const size_t total = npix * (size_t)(C / 8);
  MINE_DISPATCH_ES(es, T, (bn_res_act_bwd_reduce_kernel<T><<<blocks_for(total, C, 148 * 8), 256, 2 * C * sizeof(float), stream>>>(
      (const T*)dout, (const T*)out, (const T*)y, stats, (T*)g_out, sums, (unsigned)total, C, slope, inv_count, eps)));
}

void launch_bn_update_running(const float* stats, float* running_mean, float* running_var, long long* num_batches, int C,
                              float count, float momentum, cudaStream_t stream) {
  const float unbias = count > 1.f ? count / (count - 1.f) : 1.f;
  bn_update_running_kernel<<<(C + 255) / 256, 256, 0, stream>>>(stats, running_mean, running_var, num_batches, C,
                                                               1.f / count, unbias, momentum);
}

void launch_channel_stats(const void* y, float* sums, size_t npix, int C, int es, cudaStream_t stream) {
  const size_t total = npix * (size_t)(C / 8);
  MINE_DISPATCH_ES(es, T, (channel_stats_kernel<T><<<blocks_for(total, C, 148 * 8), 256, 2 * C * sizeof(float), stream>>>(
      (const T*)y, sums, (unsigned)total, C)));
}

}  // namespace mine
