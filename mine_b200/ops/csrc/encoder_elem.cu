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
    unsigned total, int C, float slope, float inv_count, float eps, const LLExchange x, float* __restrict__ red_out,
    int round_out) {
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
    if (round_out) st8_op(out + o, v);     // consumer is a tcgen05 kernel: fp32 storage rounded to TF32 here
    else st8(out + o, v);                  // consumer is a library convolution ("hybrid" encoder): plain fp32
  }
}

// Bitwise reproducible reduction of the per-thread partial sums acc1/acc2 (8 channels starting at c0) to sums[2][C]: no
// float atomics anywhere, every addition order is fixed by the launch geometry.
//   block: every thread parks its 16 partials in shared memory; channel c is summed over the threads that own it in thread
//          order and written to scratch[block][2C];
//   group: blocks form groups of kGroup; the block that draws a group's last ticket sums the group's partials in block
//          order into scratch[nb + group][2C];
//   grid:  the group-finisher that draws the last top-level ticket sums the group partials in group order into ``sums``
//          (no zero fill of the output).  Tickets are reset by their last drawer (graph-replay safe).
// Two levels keep both combining passes short (<= 16 x 2C resp. <= 19 x 2C values for one block) while the data pass keeps
// two blocks per SM.  fp32 atomics made the statistics differ in the last bits from run to run; through ~50 BatchNorm
// layers over a few dozen samples each (random initialisation, small images) that noise decorrelated the gradients of two
// identical runs (scripts/graph_vs_eager_probe.py: gradient cosine 0.05 -> 0.9998).  The ticket array is per device:
// the reductions of a device are stream ordered (they all run on the compute stream).
constexpr int kGroup = 16;
__device__ __forceinline__ void sum_rows(const float* __restrict__ rows, int nrows, int C2, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < C2; i += blockDim.x) {
    float v = 0.f;
    int r = 0;
    for (; r + 4 <= nrows; r += 4) {       // four loads in flight, additions in row order
      const float a0 = __ldcg(rows + (size_t)r * C2 + i), a1 = __ldcg(rows + (size_t)(r + 1) * C2 + i);
      const float a2 = __ldcg(rows + (size_t)(r + 2) * C2 + i), a3 = __ldcg(rows + (size_t)(r + 3) * C2 + i);
      v = (((v + a0) + a1) + a2) + a3;
    }
    for (; r < nrows; ++r) v += __ldcg(rows + (size_t)r * C2 + i);
    dst[i] = v;
  }
}
__device__ __forceinline__ void publish_sums(float* s_part, float* __restrict__ scratch, unsigned* __restrict__ tickets,
                                             float* __restrict__ sums, int C, int c0, bool active, const float* acc1,
                                             const float* acc2) {
  __shared__ bool s_last;
  const int cg = C >> 3;                 // <= 256 (launcher): thread t owns channel group t % cg
  const int C2 = 2 * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_part[threadIdx.x * 16 + j] = active ? acc1[j] : 0.f;
    s_part[threadIdx.x * 16 + 8 + j] = active ? acc2[j] : 0.f;
  }
  __syncthreads();
  float* mine_part = scratch + (size_t)blockIdx.x * C2;
  const int owners = blockDim.x / cg;    // threads per channel group (>= 1)
  for (int i = threadIdx.x; i < C2; i += blockDim.x) {
    const int which = i >= C ? 1 : 0, c = i - which * C;
    const int grp = c >> 3, j = c & 7;
    float v = 0.f;
    for (int k = 0; k < owners; ++k) v += s_part[(grp + k * cg) * 16 + which * 8 + j];
    mine_part[i] = v;
  }
  const int nb = gridDim.x, ngroups = (nb + kGroup - 1) / kGroup;
  const int group = blockIdx.x / kGroup, gfirst = group * kGroup;
  const int gsize = nb - gfirst < kGroup ? nb - gfirst : kGroup;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + 1 + group, 1u) == (unsigned)gsize - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* gparts = scratch + (size_t)nb * C2;
  sum_rows(scratch + (size_t)gfirst * C2, gsize, C2, ngroups == 1 ? sums : gparts + (size_t)group * C2);
  if (threadIdx.x == 0) tickets[1 + group] = 0u;
  if (ngroups == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets, 1u) == (unsigned)ngroups - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  sum_rows(gparts, ngroups, C2, sums);
  if (threadIdx.x == 0) tickets[0] = 0u;
}

// Default (fast) variant: warp fold with shuffles (lanes with equal lane % cg own the same channels), one shared-memory
// atomic per channel and warp, one global fp32 atomic per channel and block into the ZEROED ``sums``.  Results vary in the
// last bits from run to run (addition order); ``det`` selects publish_sums above instead.
__device__ __forceinline__ void publish_sums_atomic(float* s_sum, float* __restrict__ sums, int C, int c0, bool active,
                                                    const float* acc1, const float* acc2) {
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i] = 0.f;
  __syncthreads();
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
    float slope, float inv_count, float eps, float* __restrict__ scratch, unsigned* __restrict__ ticket, int det) {
  __shared__ float s_part[256 * 16];     // deterministic: [thread][16] partials; atomic: [2][C] block sums (C <= 2048)
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
  if (det) publish_sums(s_part, scratch, ticket, sums, C, w.c0, w.active, acc1, acc2);
  else publish_sums_atomic(s_part, sums, C, w.c0, w.active, acc1, acc2);
}

template <typename T>
__global__ void __launch_bounds__(256) channel_stats_kernel(const T* __restrict__ y, float* __restrict__ sums,
                                                            unsigned total, int C, float* __restrict__ scratch,
                                                            unsigned* __restrict__ ticket, int det) {
  __shared__ float s_part[256 * 16];     // deterministic: [thread][16] partials; atomic: [2][C] block sums (C <= 2048)
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
  if (det) publish_sums(s_part, scratch, ticket, sums, C, w.c0, w.active, acc1, acc2);
  else publish_sums_atomic(s_part, sums, C, w.c0, w.active, acc1, acc2);
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

// The same update for up to kRunMax layers in one launch (block = layer): the trainer defers the per-layer updates of a
// step and flushes them together (67 launches of ~3 us each otherwise).
constexpr int kRunMax = 48;
struct RunBatch {
  const float* stats[kRunMax];
  float* mean[kRunMax];
  float* var[kRunMax];
  long long* nbt[kRunMax];
  int C[kRunMax];
  float inv_count[kRunMax], unbias[kRunMax], momentum[kRunMax];
};
__global__ void bn_update_running_multi_kernel(const __grid_constant__ RunBatch b) {
  const int l = blockIdx.x;
  const int C = b.C[l];
  const float inv_count = b.inv_count[l], unbias = b.unbias[l], momentum = b.momentum[l];
  if (threadIdx.x == 0 && b.nbt[l]) *b.nbt[l] += 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = b.stats[l][c] * inv_count;
    float var = b.stats[l][C + c] * inv_count - mean * mean;
    var = var < 0.f ? 0.f : var;
    b.mean[l][c] = (1.f - momentum) * b.mean[l][c] + momentum * mean;
    b.var[l][c] = (1.f - momentum) * b.var[l][c] + momentum * var * unbias;
  }
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
                           const LLExchange* x, float* red_out, int round_out, cudaStream_t stream) {
  const size_t total = npix * (size_t)(C / 8);
  LLExchange xx{};
  if (x) xx = *x;
  const size_t smem = xx.world > 1 ? 2 * (size_t)C * sizeof(float) : 0;
  MINE_DISPATCH_ES(es, T, (bn_res_act_fwd_kernel<T><<<blocks_for(total, C, 148 * 16), 256, smem, stream>>>(
      (const T*)y, stats, gamma, beta, (const T*)res, (T*)out, (unsigned)total, C, slope, inv_count, eps, xx, red_out,
      round_out)));
}

// grid of the reproducible reductions: two blocks per SM for the data pass, fewer for the wide (hence small) tensors so
// that the combining passes stay short; scratch = (blocks + groups) x 2C floats
int reduce_blocks(size_t npix, int C) {
  const size_t total = npix * (size_t)(C / 8);
  int cap = 65536 / C;                    // C = 2048 -> 32, 1024 -> 64, 512 -> 128, <= 256 -> 256
  cap = cap > 256 ? 256 : cap;
  return blocks_for(total, C, cap);
}
size_t reduce_scratch_floats(size_t npix, int C) {
  const int nb = reduce_blocks(npix, C);
  return (size_t)(nb + (nb + kGroup - 1) / kGroup) * 2 * (size_t)C;
}

void launch_bn_res_act_bwd_reduce(const void* dout, const void* out, const void* y, const float* stats, void* g_out,
                                  float* sums, size_t npix, int C, float slope, float inv_count, float eps, int es,
                                  float* scratch, unsigned* ticket, cudaStream_t stream) {
  // scratch != null: reproducible two-level reduction (``sums`` need not be zeroed); null: fp32 atomics into zeroed sums
  const size_t total = npix * (size_t)(C / 8);
  const int blocks = scratch ? reduce_blocks(npix, C) : blocks_for(total, C, 148 * 2);
  MINE_DISPATCH_ES(es, T, (bn_res_act_bwd_reduce_kernel<T><<<blocks, 256, 0, stream>>>(
      (const T*)dout, (const T*)out, (const T*)y, stats, (T*)g_out, sums, (unsigned)total, C, slope, inv_count, eps,
      scratch, ticket, scratch ? 1 : 0)));
}

void launch_bn_update_running(const float* stats, float* running_mean, float* running_var, long long* num_batches, int C,
                              float count, float momentum, cudaStream_t stream) {
  const float unbias = count > 1.f ? count / (count - 1.f) : 1.f;
  bn_update_running_kernel<<<(C + 255) / 256, 256, 0, stream>>>(stats, running_mean, running_var, num_batches, C,
                                                               1.f / count, unbias, momentum);
}

void launch_bn_update_running_multi(int n, const float* const* stats, float* const* mean, float* const* var,
                                    long long* const* nbt, const int* C, const float* count, const float* momentum,
                                    cudaStream_t stream) {
  for (int i0 = 0; i0 < n; i0 += kRunMax) {
    RunBatch b{};
    const int m = n - i0 < kRunMax ? n - i0 : kRunMax;
    for (int i = 0; i < m; ++i) {
      b.stats[i] = stats[i0 + i]; b.mean[i] = mean[i0 + i]; b.var[i] = var[i0 + i]; b.nbt[i] = nbt[i0 + i];
      b.C[i] = C[i0 + i];
      const float cnt = count[i0 + i];
      b.inv_count[i] = 1.f / cnt; b.unbias[i] = cnt > 1.f ? cnt / (cnt - 1.f) : 1.f; b.momentum[i] = momentum[i0 + i];
    }
    bn_update_running_multi_kernel<<<m, 256, 0, stream>>>(b);
  }
}

void launch_channel_stats(const void* y, float* sums, size_t npix, int C, int es, float* scratch, unsigned* ticket,
                          cudaStream_t stream) {
  const size_t total = npix * (size_t)(C / 8);
  const int blocks = scratch ? reduce_blocks(npix, C) : blocks_for(total, C, 148 * 2);
  MINE_DISPATCH_ES(es, T, (channel_stats_kernel<T><<<blocks, 256, 0, stream>>>(
      (const T*)y, sums, (unsigned)total, C, scratch, ticket, scratch ? 1 : 0)));
}

}  // namespace mine
