// Memory-bound companions of the tcgen05 conv engine (NHWC bf16, 16-byte vectors = 8 channels).
//
//   bn_act_pad_fwd    : y (pre-BN conv output) -> ELU(BN(y)) written into a 1-pixel padded buffer with
//                       reflection or replication borders (what the next conv's TMA boxes read).  Replaces
//                       ATen batch_norm_elemt + ELU + ReflectionPad2d (+ upsample_nearest2d + cat) of the
//                       reference: one read, one write.
//   bn_act_bwd_reduce : folds the padded gradient back (adjoint of reflect/replicate pad), multiplies by
//                       ELU', accumulates sum(g) and sum(g * xhat) per channel (the two BatchNorm backward
//                       reductions, all-reduced across GPUs by the caller), stores g.
//   bn_bwd_apply      : dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); loops over the S planes of an
//                       image so the shared-skip gradient (sum over planes) stays in registers; per-plane
//                       embedding-bias gradients are reduced in shared memory.
//   head_bwd          : gradient of the MPI head activation, emitted as the 16-channel bf16 tensor the
//                       dgrad/wgrad GEMMs consume.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "act_types.cuh"
#include "conv_engine.h"
#include "ll_exchange.cuh"
#include "kernels.h"

namespace mine {

__device__ __forceinline__ int pad_src(int p, int n, int mode) {   // padded index -> source index
  int s = p - 1;
  if (mode == 0) { if (s < 0) s = -s; if (s >= n) s = 2 * n - 2 - s; }
  else { s = s < 0 ? 0 : (s >= n ? n - 1 : s); }
  return s;
}

struct BnCoef {      // per-channel affine of training-mode BN from the reduced statistics
  float a[8], b[8], mean[8], invstd[8];
};
__device__ __forceinline__ BnCoef bn_coef(const float* __restrict__ stats, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, int C, int c0, float inv_count, float eps) {
  BnCoef k;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = stats[c0 + j] * inv_count;
    float var = stats[C + c0 + j] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    const float is = rsqrtf(var + eps);
    k.mean[j] = m; k.invstd[j] = is;
    k.a[j] = gamma[c0 + j] * is;
    k.b[j] = beta[c0 + j] - m * k.a[j];
  }
  return k;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_pad_fwd_kernel(
    const T* __restrict__ y, const float* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, T* __restrict__ out, int N, int H, int W, int C, int pad_mode,
    float inv_count, float eps, const LLExchange x, float* __restrict__ red_out, const FastDiv fd_wp,
    const FastDiv fd_hp) {
  extern __shared__ float s_red[];               // [2C] cross-GPU reduced statistics (only when x.world > 1)
  if (x.world > 1) { ll_exchange_sum(stats, s_red, 2 * C, x, red_out); stats = s_red; }
  const int cg = C >> 3;                         // power of two (C in {16,...,256})
  const int cg_shift = 31 - __clz(cg);
  const int Hp = H + 2, Wp = W + 2;
  const unsigned total = (unsigned)N * Hp * Wp * cg;          // < 2^31 for every decoder tensor (checked on the host)
  const unsigned stride = ((gridDim.x * blockDim.x) >> cg_shift) << cg_shift;   // keeps the channel group per thread
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= stride) return;
  const int c0 = (int)(i0 & (cg - 1)) * 8;
  const BnCoef k = bn_coef(stats, gamma, beta, C, c0, inv_count, eps);
  for (unsigned i = i0; i < total; i += stride) {
    int px, py, n, t;                              // multiply-high decomposition (conv_engine.h::FastDiv)
    fdivmod((int)(i >> cg_shift), fd_wp, t, px);
    fdivmod(t, fd_hp, n, py);
    const int sy = pad_src(py, H, pad_mode), sx = pad_src(px, W, pad_mode);
    V8 v = ld8(y + (((size_t)n * H + sy) * W + sx) * C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = v.f[j] * k.a[j] + k.b[j];
      v.f[j] = u > 0.f ? u : (__expf(u) - 1.f);
    }
    st8_op(out + (((size_t)n * Hp + py) * Wp + px) * C + c0, v);
  }
}

// g = fold(dapad) * ELU'(bn(y)); sums[0][c] += g, sums[1][c] += g * xhat
// BN coefficients live in shared memory (4 x C floats) instead of 32 registers per thread; interior pixels
// (no pad adjoint) take a branch-free fast path.
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(
    const T* __restrict__ dapad, const T* __restrict__ y, const float* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ g_out,
    float* __restrict__ sums, int N, int H, int W, int C, int pad_mode, float inv_count, float eps) {
  extern __shared__ float s_mem[];     // [2][C] sums, then a[C], b[C], mean[C], invstd[C]
  float* s_sum = s_mem;
  float* s_a = s_mem + 2 * C;
  float* s_b = s_a + C;
  float* s_mean = s_b + C;
  float* s_is = s_mean + C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i] = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m = stats[c] * inv_count;
    float var = stats[C + c] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    const float is = rsqrtf(var + eps);
    s_mean[c] = m; s_is[c] = is; s_a[c] = gamma[c] * is; s_b[c] = beta[c] - m * gamma[c] * is;
  }
  __syncthreads();
  const int cg = C >> 3;                         // power of two
  const int cg_shift = 31 - __clz(cg);
  const int Hp = H + 2, Wp = W + 2;
  const unsigned total = (unsigned)N * H * W * cg;            // < 2^31 (checked on the host)
  // a thread keeps its channel group for the whole grid-stride loop when the stride is a multiple of cg
  const unsigned stride = ((gridDim.x * blockDim.x) >> cg_shift) << cg_shift;
  float acc1[8], acc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc1[j] = acc2[j] = 0.f;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 & (cg - 1)) * 8;
  const int lo = pad_mode == 0 ? 1 : 0;          // reflection folds the border onto row/col 1 and n-2, replication onto 0 and n-1
  if (i0 < stride) {
    for (unsigned i = i0; i < total; i += stride) {
      unsigned pix = i >> cg_shift;
      const int x = (int)(pix % (unsigned)W); pix /= (unsigned)W;
      const int yy = (int)(pix % (unsigned)H);
      const int n = (int)(pix / (unsigned)H);
      const T* base = dapad + ((size_t)n * Hp * Wp) * C + c0;
      V8 d = ld8(base + ((size_t)(yy + 1) * Wp + (x + 1)) * C);
      const bool top = (yy == lo), bot = (yy == H - 1 - lo), lef = (x == lo), rig = (x == W - 1 - lo);
      if (top | bot | lef | rig) {               // border pixels also receive the gradient of their pad copies
        int ry[3], rx[3], ny = 1, nx = 1;
        ry[0] = yy + 1; rx[0] = x + 1;
        if (top) ry[ny++] = 0;
        if (bot) ry[ny++] = H + 1;
        if (lef) rx[nx++] = 0;
        if (rig) rx[nx++] = W + 1;
        for (int a = 0; a < ny; ++a)
          for (int b = 0; b < nx; ++b) {
            if (a == 0 && b == 0) continue;
            const V8 t = ld8(base + ((size_t)ry[a] * Wp + rx[b]) * C);
#pragma unroll
            for (int j = 0; j < 8; ++j) d.f[j] += t.f[j];
          }
      }
      const size_t o = (((size_t)n * H + yy) * W + x) * C + c0;
      const V8 yv = ld8(y + o);
      V8 g;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = yv.f[j] * s_a[c0 + j] + s_b[c0 + j];
        const float de = u > 0.f ? 1.f : __expf(u);
        g.f[j] = d.f[j] * de;
        const float xhat = (yv.f[j] - s_mean[c0 + j]) * s_is[c0 + j];
        acc1[j] += g.f[j];
        acc2[j] += g.f[j] * xhat;
      }
      st8(g_out + o, g);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { atomicAdd(&s_sum[c0 + j], acc1[j]); atomicAdd(&s_sum[C + c0 + j], acc2[j]); }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], s_sum[i]);
}

// Variant 2 of the reduction above (MINE_B200_BN_REDUCE=v2, opt-in until measured): the per-channel affine (a, b)
// lives in 16 registers like in the forward kernel instead of four shared-memory arrays (32 LDS per 8 elements made
// v1 issue bound at 0.23 of HBM peak), and the second sum is accumulated as sum(g * y); every thread converts its
// partial to sum(g * xhat) = invstd * (sum(g*y) - mean * sum(g)) once, before the block reduction.
template <typename T>
__global__ void __launch_bounds__(256, 2) bn_act_bwd_reduce_v2_kernel(
    const T* __restrict__ dapad, const T* __restrict__ y, const float* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ g_out,
    float* __restrict__ sums, int N, int H, int W, int C, int pad_mode, float inv_count, float eps,
    const FastDiv fd_w, const FastDiv fd_h) {
  extern __shared__ float s_mem[];     // [2][C] block partial sums
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_mem[i] = 0.f;
  __syncthreads();
  const int cg = C >> 3;
  const int cg_shift = 31 - __clz(cg);
  const int Hp = H + 2, Wp = W + 2;
  const unsigned total = (unsigned)N * H * W * cg;
  const unsigned stride = ((gridDim.x * blockDim.x) >> cg_shift) << cg_shift;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 & (cg - 1)) * 8;
  float a[8], b[8], acc1[8], acc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {        // only (a, b) stay live in the loop; mean / invstd are recomputed at the end
    const float m = stats[c0 + j] * inv_count;
    float var = stats[C + c0 + j] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    a[j] = gamma[c0 + j] * rsqrtf(var + eps);
    b[j] = beta[c0 + j] - m * a[j];
    acc1[j] = acc2[j] = 0.f;
  }
  const int lo = pad_mode == 0 ? 1 : 0;
  // element i -> (image, row, column) by multiply-high (no integer divides), folded border gradient + activation input
  auto load = [&](unsigned i, V8& d, V8& yv, size_t& o) {
    int pix = (int)(i >> cg_shift), x, yy, n, t;
    fdivmod(pix, fd_w, t, x);
    fdivmod(t, fd_h, n, yy);
    const T* base = dapad + ((size_t)n * Hp * Wp) * C + c0;
    d = ld8(base + ((size_t)(yy + 1) * Wp + (x + 1)) * C);
    o = (((size_t)n * H + yy) * W + x) * C + c0;
    yv = ld8(y + o);
    const bool top = (yy == lo), bot = (yy == H - 1 - lo), lef = (x == lo), rig = (x == W - 1 - lo);
    if (top | bot | lef | rig) {
      int ry[3], rx[3], ny = 1, nx = 1;
      ry[0] = yy + 1; rx[0] = x + 1;
      if (top) ry[ny++] = 0;
      if (bot) ry[ny++] = H + 1;
      if (lef) rx[nx++] = 0;
      if (rig) rx[nx++] = W + 1;
      for (int p = 0; p < ny; ++p)
        for (int q = 0; q < nx; ++q) {
          if (p == 0 && q == 0) continue;
          const V8 tt = ld8(base + ((size_t)ry[p] * Wp + rx[q]) * C);
#pragma unroll
          for (int j = 0; j < 8; ++j) d.f[j] += tt.f[j];
        }
    }
  };
  auto finish = [&](const V8& d, const V8& yv, size_t o) {
    V8 g;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = fmaf(yv.f[j], a[j], b[j]);
      g.f[j] = d.f[j] * (u > 0.f ? 1.f : __expf(u));
      acc1[j] += g.f[j];
      acc2[j] = fmaf(g.f[j], yv.f[j], acc2[j]);
    }
    if (g_out) st8(g_out + o, g);       // null: the apply kernel recomputes g from dapad (one tensor less written + read)
  };
  if (i0 < stride) {
    // two elements per trip: four independent 32-byte loads in flight per thread
    for (unsigned i = i0; i < total; i += 2 * stride) {
      V8 d0, y0, d1, y1;
      size_t o0, o1 = 0;
      const bool two = i + stride < total;
      load(i, d0, y0, o0);
      if (two) load(i + stride, d1, y1, o1);
      finish(d0, y0, o0);
      if (two) finish(d1, y1, o1);
    }
  }
  // block reduction: lanes with equal (lane % cg) own the same channels (cg and the warp size are powers of two), so the
  // warp is folded with shuffles first and only cg lanes per warp touch shared memory (4096 contended shared-memory
  // atomics per block on 2C addresses cost a third of the kernel for the 16-channel layers)
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = stats[c0 + j] * inv_count;
    float var = stats[C + c0 + j] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    float v1 = acc1[j], v2 = rsqrtf(var + eps) * (acc2[j] - m * acc1[j]);
    for (int o2 = 16; o2 >= cg && o2 > 0; o2 >>= 1) {
      v1 += __shfl_xor_sync(0xffffffffu, v1, o2);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o2);
    }
    if (lane < cg || cg >= 32) {
      atomicAdd(&s_mem[c0 + j], v1);
      atomicAdd(&s_mem[C + c0 + j], v2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], s_mem[i]);
}

// One thread = 8 channels of one pixel of one IMAGE; loops over its S planes.
// FUSED: ``g`` is the PADDED upstream gradient ``dapad``; the pad adjoint and ELU' are recomputed here (same arithmetic as
// bn_act_bwd_reduce) instead of being written by the reduction kernel and read back.
template <typename T, bool FUSED>
__global__ void __launch_bounds__(256, FUSED ? 3 : 4) bn_bwd_apply_kernel(
    const T* __restrict__ g, const T* __restrict__ y, const float* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ sums, T* __restrict__ dy,
    float* __restrict__ dshared, float* __restrict__ dplane_bias, int B, int S, int H, int W, int C, float inv_count,
    float eps, const LLExchange x, const float* __restrict__ beta, int pad_mode) {
  extern __shared__ float s_pb[];      // [S][C] per-plane bias gradient partials (if wanted), then [2C] reduced sums
  const bool want_pb = dplane_bias != nullptr;
  if (x.world > 1) {                   // cross-GPU SUM of the two BatchNorm backward reductions, fused into this kernel
    float* s_red = s_pb + (want_pb ? S * C : 0);
    ll_exchange_sum(sums, s_red, 2 * C, x, nullptr);
    sums = s_red;
  }
  if (want_pb) {
    for (int i = threadIdx.x; i < S * C; i += blockDim.x) s_pb[i] = 0.f;
    __syncthreads();
  }
  const int cg = C >> 3;
  const size_t per_img = (size_t)H * W * cg;
  const int b = blockIdx.y;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const bool active = i < per_img;
  const size_t ii = active ? i : 0;
  const int c0 = (int)(ii % cg) * 8;
  const size_t pix = ii / cg;
  // dy = coef * (g - mean_g - xhat * mean_gx)  with xhat = (y - mean) * invstd, folded into  coef * g + k1 * y + k0
  float coef[8], k1[8], k0[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = stats[c0 + j] * inv_count;
    float var = stats[C + c0 + j] * inv_count - m * m;
    var = var < 0.f ? 0.f : var;
    const float invstd = rsqrtf(var + eps);
    coef[j] = gamma[c0 + j] * invstd;
    const float mg = sums[c0 + j] * inv_count, mgx = sums[C + c0 + j] * inv_count;
    k1[j] = -coef[j] * mgx * invstd;
    k0[j] = -coef[j] * mg - k1[j] * m;
    bb[j] = FUSED ? beta[c0 + j] - m * coef[j] : 0.f;     // BN(y) = coef * y + bb
  }
  float ds[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ds[j] = 0.f;
  const int lane = threadIdx.x & 31;
  // FUSED: pixel coordinates inside the padded gradient and which border rows / columns fold onto this pixel
  const int Hp = H + 2, Wp = W + 2;
  const int py = (int)(pix / (size_t)W), px = (int)(pix - (size_t)py * W);
  const int lo = pad_mode == 0 ? 1 : 0;
  const bool top = (py == lo), bot = (py == H - 1 - lo), lef = (px == lo), rig = (px == W - 1 - lo);
  for (int s = 0; s < S; ++s) {
    V8 d;
#pragma unroll
    for (int j = 0; j < 8; ++j) d.f[j] = 0.f;
    if (active) {
      const size_t o = (((size_t)(b * S + s) * H * W) + pix) * C + c0;
      const V8 yv = ld8(y + o);
      V8 gv;
      if (FUSED) {
        const T* base = g + ((size_t)(b * S + s) * Hp * Wp) * C + c0;
        gv = ld8(base + ((size_t)(py + 1) * Wp + (px + 1)) * C);
        if (top | bot | lef | rig) {
          int ry[3], rx[3], ny = 1, nx = 1;
          ry[0] = py + 1; rx[0] = px + 1;
          if (top) ry[ny++] = 0;
          if (bot) ry[ny++] = H + 1;
          if (lef) rx[nx++] = 0;
          if (rig) rx[nx++] = W + 1;
          for (int p = 0; p < ny; ++p)
            for (int q = 0; q < nx; ++q) {
              if (p == 0 && q == 0) continue;
              const V8 tt = ld8(base + ((size_t)ry[p] * Wp + rx[q]) * C);
#pragma unroll
              for (int j = 0; j < 8; ++j) gv.f[j] += tt.f[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float u = fmaf(yv.f[j], coef[j], bb[j]);
          gv.f[j] *= (u > 0.f ? 1.f : __expf(u));
        }
      } else {
        gv = ld8(g + o);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d.f[j] = fmaf(coef[j], gv.f[j], fmaf(k1[j], yv.f[j], k0[j]));
        ds[j] += d.f[j];
      }
      st8_op(dy + o, d);
    }
    if (want_pb) {
      // lanes with equal (lane % cg) own the same channels (blockDim and cg are powers of two)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = d.f[j];
        for (int o2 = 16; o2 >= cg && o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
        if (lane < cg || cg >= 32) atomicAdd(&s_pb[s * C + c0 + j], v);
      }
    }
  }
  if (dshared && active) {
    float* dsp = dshared + ((size_t)b * H * W + pix) * C + c0;
    *reinterpret_cast<float4*>(dsp) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    *reinterpret_cast<float4*>(dsp + 4) = make_float4(ds[4], ds[5], ds[6], ds[7]);
  }
  if (want_pb) {
    __syncthreads();
    for (int k = threadIdx.x; k < S * C; k += blockDim.x) {
      const float v = s_pb[k];
      if (v != 0.f) atomicAdd(&dplane_bias[(size_t)b * S * C + k], v);
    }
  }
}

// MPI head backward: g_mpi fp32 [.,4], mpi fp32 [.,4] (activated), sign int8 -> dz bf16 [.,16] (channels 4..15 zero)
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float4* __restrict__ g_mpi, const float4* __restrict__ mpi,
                                                       const int8_t* __restrict__ sign, T* __restrict__ dz,
                                                       float* __restrict__ dbias, size_t npix, int use_alpha) {
  __shared__ float s_db[4];
  if (threadIdx.x < 4) s_db[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const float4 gg = g_mpi[i], o = mpi[i];
    float d0 = gg.x * o.x * (1.f - o.x), d1 = gg.y * o.y * (1.f - o.y), d2 = gg.z * o.z * (1.f - o.z);
    float d3 = use_alpha ? gg.w * o.w * (1.f - o.w) : gg.w * (float)sign[i];
    acc[0] += d0; acc[1] += d1; acc[2] += d2; acc[3] += d3;
    V8 v0, v1;
    v0.f[0] = d0; v0.f[1] = d1; v0.f[2] = d2; v0.f[3] = d3;
#pragma unroll
    for (int j = 4; j < 8; ++j) v0.f[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v1.f[j] = 0.f;
    st8_op(dz + i * 16, v0);
    st8(dz + i * 16 + 8, v1);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_db[j], v);
  }
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(&dbias[threadIdx.x], s_db[threadIdx.x]);
}

// fp32 NCHW image-like tensor -> bf16 NHWC (encoder skip features etc. use torch channels_last instead)
static int grid_for(size_t total, int cap = 148 * 16) {
  size_t b = (total + 255) / 256;
  return (int)(b > (size_t)cap ? cap : (b == 0 ? 1 : b));
}

void launch_bn_act_pad_fwd(const void* y, const float* stats, const float* gamma, const float* beta, void* out, int N,
                           int H, int W, int C, int pad_mode, float inv_count, float eps, int es, const LLExchange* x,
                           float* red_out, cudaStream_t stream) {
  const size_t total = (size_t)N * (H + 2) * (W + 2) * (C / 8);
  LLExchange xx{};
  if (x) xx = *x;
  const size_t smem = xx.world > 1 ? 2 * (size_t)C * sizeof(float) : 0;
  MINE_DISPATCH_ES(es, T, (bn_act_pad_fwd_kernel<T><<<grid_for(total, 148 * 32), 256, smem, stream>>>(
      (const T*)y, stats, gamma, beta, (T*)out, N, H, W, C, pad_mode, inv_count, eps, xx, red_out, make_fastdiv(W + 2),
      make_fastdiv(H + 2))));
}

void launch_bn_act_bwd_reduce(const void* dapad, const void* y, const float* stats, const float* gamma,
                              const float* beta, void* g_out, float* sums, int N, int H, int W, int C, int pad_mode,
                              float inv_count, float eps, int es, cudaStream_t stream) {
  const size_t total = (size_t)N * H * W * (C / 8);
  int blocks = grid_for(total, 148 * 8);
  static const bool v1 = getenv("MINE_B200_BN_REDUCE") && getenv("MINE_B200_BN_REDUCE")[0] == 'o';   // "old"
  if (!v1) {
    MINE_DISPATCH_ES(es, T, (bn_act_bwd_reduce_v2_kernel<T><<<blocks, 256, 2 * C * sizeof(float), stream>>>(
        (const T*)dapad, (const T*)y, stats, gamma, beta, (T*)g_out, sums, N, H, W, C, pad_mode, inv_count, eps,
        make_fastdiv(W), make_fastdiv(H))));
    return;
  }
  // the grid-stride must be a multiple of the channel-group count so every thread keeps its channels
  MINE_DISPATCH_ES(es, T, (bn_act_bwd_reduce_kernel<T><<<blocks, 256, 6 * C * sizeof(float), stream>>>(
      (const T*)dapad, (const T*)y, stats, gamma, beta, (T*)g_out, sums, N, H, W, C, pad_mode, inv_count, eps)));
}

void launch_bn_bwd_apply(const void* g, const void* y, const float* stats, const float* gamma, const float* sums,
                         void* dy, float* dshared, float* dplane_bias, int B, int S, int H, int W, int C,
                         float inv_count, float eps, int es, const LLExchange* x, const float* beta, int pad_mode,
                         cudaStream_t stream) {
  // beta != null: ``g`` is the padded upstream gradient [B*S, H+2, W+2, C] (fused pad adjoint + ELU')
  const size_t per_img = (size_t)H * W * (C / 8);
  dim3 grid((unsigned)((per_img + 255) / 256), B);
  LLExchange xx{};
  if (x) xx = *x;
  const size_t smem = (dplane_bias ? (size_t)S * C * sizeof(float) : 0) + (xx.world > 1 ? 2 * (size_t)C * sizeof(float) : 0);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(bn_bwd_apply_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(bn_bwd_apply_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(bn_bwd_apply_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(bn_bwd_apply_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    attr = true;
  }
  if (beta) {
    MINE_DISPATCH_ES(es, T, (bn_bwd_apply_kernel<T, true><<<grid, 256, smem, stream>>>(
        (const T*)g, (const T*)y, stats, gamma, sums, (T*)dy, dshared, dplane_bias, B, S, H, W, C, inv_count, eps, xx,
        beta, pad_mode)));
  } else {
    MINE_DISPATCH_ES(es, T, (bn_bwd_apply_kernel<T, false><<<grid, 256, smem, stream>>>(
        (const T*)g, (const T*)y, stats, gamma, sums, (T*)dy, dshared, dplane_bias, B, S, H, W, C, inv_count, eps, xx,
        nullptr, 0)));
  }
}

void launch_head_bwd(const float* g_mpi, const float* mpi, const int8_t* sign, void* dz, float* dbias, size_t npix,
                     int use_alpha, int es, cudaStream_t stream) {
  MINE_DISPATCH_ES(es, T, (head_bwd_kernel<T><<<grid_for(npix, 148 * 8), 256, 0, stream>>>(
      (const float4*)g_mpi, (const float4*)mpi, sign, (T*)dz, dbias, npix, use_alpha)));
}

}  // namespace mine
