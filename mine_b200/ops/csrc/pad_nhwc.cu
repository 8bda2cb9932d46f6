// 1-pixel reflection pad of an NHWC tensor and its adjoint (fp32 or bf16 storage, C a multiple of 8).
//
// The shared-skip convolution of every decoder level is conv3x3(reflect_pad(feature)) (reference
// network/monodepth2/layers.py Conv3x3, depth_decoder.py:124-146).  With the framework's reflection_pad2d the padded
// tensor comes back NCHW-contiguous: one layout copy in front of the channels-last library convolution and two more in
// backward (profiles: 3 copies + 2 pad kernels per level and step, ~0.27 ms in total).  These two kernels keep
// everything NHWC, so the padded tensor is consumed - and its gradient produced - by the library without a copy.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "act_types.cuh"
#include "conv_engine.h"
#include "kernels.h"

namespace mine {

namespace {

__device__ __forceinline__ int reflect1(int p, int n) {   // padded index -> source index (n >= 2)
  int s = p - 1;
  if (s < 0) s = -s;
  if (s >= n) s = 2 * n - 2 - s;
  return s;
}

// out[n, py, px, :] = x[n, reflect(py), reflect(px), :]; one thread = 8 channels of one padded pixel
template <typename T>
__global__ void __launch_bounds__(256) pad_reflect_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int N, int H,
                                                              int W, int C, const FastDiv fd_cg, const FastDiv fd_wp,
                                                              const FastDiv fd_hp) {
  const int cg = C >> 3, Hp = H + 2, Wp = W + 2;
  const unsigned total = (unsigned)N * Hp * Wp * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int pix, g, t, px, n, py;
    fdivmod((int)i, fd_cg, pix, g);
    fdivmod(pix, fd_wp, t, px);
    fdivmod(t, fd_hp, n, py);
    const V8 v = ld8(x + (((size_t)n * H + reflect1(py, H)) * W + reflect1(px, W)) * C + g * 8);
    st8(out + (size_t)i * 8, v);
  }
}

// adjoint: gx[n, y, x, :] = sum of the padded positions that read (y, x): itself plus the border rows / columns that
// reflect onto it (row 1 receives padded row 0, row H-2 receives padded row H+1, same for columns; corners fold twice)
template <typename T>
__global__ void __launch_bounds__(256) pad_reflect_bwd_kernel(const T* __restrict__ gp, T* __restrict__ gx, int N, int H,
                                                              int W, int C, const FastDiv fd_cg, const FastDiv fd_w,
                                                              const FastDiv fd_h) {
  const int cg = C >> 3, Hp = H + 2, Wp = W + 2;
  const unsigned total = (unsigned)N * H * W * cg;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int pix, g, t, x, n, y;
    fdivmod((int)i, fd_cg, pix, g);
    fdivmod(pix, fd_w, t, x);
    fdivmod(t, fd_h, n, y);
    const T* base = gp + ((size_t)n * Hp * Wp) * C + g * 8;
    int ry[3], rx[3], ny = 1, nx = 1;
    ry[0] = y + 1; rx[0] = x + 1;
    if (y == 1) ry[ny++] = 0;
    if (y == H - 2) ry[ny++] = H + 1;
    if (x == 1) rx[nx++] = 0;
    if (x == W - 2) rx[nx++] = W + 1;
    V8 acc = ld8(base + ((size_t)ry[0] * Wp + rx[0]) * C);
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) {
        if (a == 0 && b == 0) continue;
        const V8 tt = ld8(base + ((size_t)ry[a] * Wp + rx[b]) * C);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.f[j] += tt.f[j];
      }
    st8(gx + (size_t)i * 8, acc);
  }
}

int grid_of(size_t total) {
  size_t b = (total + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : (b == 0 ? 1 : b));
}

}  // namespace

void launch_pad_reflect_nhwc(const void* x, void* out, int N, int H, int W, int C, int es, cudaStream_t stream) {
  const size_t total = (size_t)N * (H + 2) * (W + 2) * (C / 8);
  MINE_DISPATCH_ES(es, T, (pad_reflect_fwd_kernel<T><<<grid_of(total), 256, 0, stream>>>(
      (const T*)x, (T*)out, N, H, W, C, make_fastdiv(C / 8), make_fastdiv(W + 2), make_fastdiv(H + 2))));
}

void launch_pad_reflect_nhwc_bwd(const void* gp, void* gx, int N, int H, int W, int C, int es, cudaStream_t stream) {
  const size_t total = (size_t)N * H * W * (C / 8);
  MINE_DISPATCH_ES(es, T, (pad_reflect_bwd_kernel<T><<<grid_of(total), 256, 0, stream>>>(
      (const T*)gp, (T*)gx, N, H, W, C, make_fastdiv(C / 8), make_fastdiv(W), make_fastdiv(H))));
}

}  // namespace mine
