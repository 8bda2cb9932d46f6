// MPI head convolution for the narrow full-resolution levels (16 / 32 input channels -> 4 outputs) on CUDA cores.
//
// These two layers move ~310 MB (level 0) for 3.6 GFMA: they are bandwidth bound, and as an implicit GEMM they waste
// the tensor pipe (N = 4 padded to 16, K = 16 per tap) while paying nine TMA re-loads of 32-byte rows
// (profiles/ncu_head_0.txt: 0.67 ms, 7 % of HBM peak).  Here one CTA stages a (16+2) x (32+2) halo tile of the
// pre-padded NHWC bf16 activation in shared memory ONCE (16-byte chunks, structure-of-arrays so that a warp reads
// consecutive 16-byte words: conflict free), every thread produces two output pixels (weights are read from
// shared memory as broadcast float4 = the four output channels of one (tap, ci) and reused for both pixels), and
// the epilogue applies the head activation (sigmoid rgb, |x| + 1e-4 or sigmoid for sigma) and writes the packed
// fp32 MPI texel plus the sign byte the backward pass needs - the same outputs as the conv_taps head epilogue.
//
// Reference semantics: network/monodepth2/depth_decoder.py:137-141 (dispconv + sigmoid / abs + 1e-4).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mine {

namespace {

constexpr int kTW = 32, kTH = 16, kHW = kTW + 2, kHH = kTH + 2, kNPix = kHW * kHH;

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ void fma4(float4& acc, float x, const float4& w) {
  acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
}

__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + __expf(-v)); }

template <int C>
__global__ void __launch_bounds__(256) head_conv_direct_kernel(
    const __nv_bfloat16* __restrict__ apad, const float4* __restrict__ wpk, const float* __restrict__ bias,
    float4* __restrict__ mpi, int8_t* __restrict__ sign, int H, int W, int use_alpha) {
  constexpr int CH = C / 8;                               // 16-byte chunks per pixel
  __shared__ uint4 s_x[CH][kNPix];
  __shared__ float4 s_w[9 * C];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, n = blockIdx.z;
  const int Hp = H + 2, Wp = W + 2;
  for (int i = tid; i < 9 * C; i += 256) s_w[i] = wpk[i];
  const __nv_bfloat16* img = apad + (size_t)n * Hp * Wp * C;
  for (int i = tid; i < kNPix * CH; i += 256) {
    const int p = i / CH, c = i - p * CH;
    const int py = p / kHW, px = p - py * kHW;
    const int gy = y0 + py, gx = x0 + px;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gy < Hp && gx < Wp) v = *reinterpret_cast<const uint4*>(img + ((size_t)gy * Wp + gx) * C + c * 8);
    s_x[c][p] = v;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;                 // pixels (ty, tx) and (ty + 8, tx) of the tile
  const float4 b4 = make_float4(bias[0], bias[1], bias[2], bias[3]);
  float4 acc0 = b4, acc1 = b4;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int p0 = (ty + ky) * kHW + tx + kx, p1 = p0 + 8 * kHW;
      const float4* wt = s_w + (ky * 3 + kx) * C;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint4 a = s_x[c][p0], b = s_x[c][p1];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w0 = wt[c * 8 + 2 * j], w1 = wt[c * 8 + 2 * j + 1];
          fma4(acc0, bf_lo(aw[j]), w0); fma4(acc0, bf_hi(aw[j]), w1);
          fma4(acc1, bf_lo(bw[j]), w0); fma4(acc1, bf_hi(bw[j]), w1);
        }
      }
    }
  }
  const int ox = x0 + tx;
  if (ox >= W) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int oy = y0 + ty + half * 8;
    if (oy >= H) continue;
    const float4 z = half ? acc1 : acc0;
    float4 o;
    o.x = sigmoidf(z.x); o.y = sigmoidf(z.y); o.z = sigmoidf(z.z);
    o.w = use_alpha ? sigmoidf(z.w) : fabsf(z.w) + 1e-4f;
    const size_t pix = ((size_t)n * H + oy) * W + ox;
    mpi[pix] = o;
    sign[pix] = z.w >= 0.f ? (int8_t)1 : (int8_t)-1;
  }
}

}  // namespace

// apad: bf16 [N, H+2, W+2, C] (C = 16 or 32); wpk: fp32 [9][C][4]; bias fp32 [4]; mpi fp32 [N,H,W,4]; sign int8 [N,H,W]
const char* launch_head_conv_direct(const void* apad, const float* wpk, const float* bias, float* mpi, int8_t* sign, int N,
                                    int H, int W, int C, int use_alpha, cudaStream_t stream) {
  dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, N);
  if (grid.z > 65535 || grid.y > 65535) return "head_conv_direct: grid too large";
  if (C == 16)
    head_conv_direct_kernel<16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)apad, (const float4*)wpk, bias,
                                                          (float4*)mpi, sign, H, W, use_alpha);
  else if (C == 32)
    head_conv_direct_kernel<32><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)apad, (const float4*)wpk, bias,
                                                          (float4*)mpi, sign, H, W, use_alpha);
  else
    return "head_conv_direct: C must be 16 or 32";
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace mine
