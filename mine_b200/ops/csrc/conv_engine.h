// Parameter blocks of the tcgen05 convolution engine (shared by conv_tcgen05.cu and the bindings).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mine {

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (the persistent kernels decompose work-item indices in
// every role of every tile; an integer divide is ~40 dependent instructions on the epilogue warps' critical path)
struct FastDiv { uint32_t mul, shr, d; };
inline FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, (uint32_t)(d < 1 ? 1 : d)};
  if (d <= 1) return f;
  int lg = 0;
  while ((1u << lg) < (uint32_t)d) ++lg;
  const int p = 31 + lg;
  f.mul = (uint32_t)((((uint64_t)1 << p) + (uint32_t)d - 1) / (uint32_t)d);
  f.shr = (uint32_t)(p - 32);
  return f;
}

#ifdef __CUDACC__
// n / d and n % d through the launcher-made multiplier (conv_engine.h::make_fastdiv)
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return f.d == 1u ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}
__device__ __forceinline__ void fdivmod(int n, const FastDiv& f, int& q, int& r) {
  q = fdiv(n, f);
  r = n - q * (int)f.d;
}
#endif

struct ConvParams {
  // logical GEMM pixel grid (per image) and its 128-pixel tiling
  int N, Hg, Wg, TH, TW, tiles_x, tiles_y;
  // reduction: G groups (sub-pixel phases) x T taps x (Ci / KB) channel blocks
  int G, T, Ci, KB, kblocks;
  int in_stride;                         // 1, or 2 for the strided gather of the phase-conv dgrad
  int16_t tap_y[4][16], tap_x[4][16];    // input offset of every (group, tap), in input pixels
  // output tensor [N, Ho, Wo, Co]; GEMM pixel (oy, ox) of group g lands on (oy*out_sy + out_oy[g], ox*out_sx + out_ox[g])
  // Co output channels are produced in CB blocks of BN (<= 256) accumulator columns: one work item = (tile, image,
  // group, channel block); CB == 1 for the decoder (BN = Co padded to 16), > 1 for the wide encoder layers
  int Co, BN, CB, Ho, Wo, out_sy, out_sx;
  int16_t out_oy[4], out_ox[4];
  void* out;
  int out_fp32, accumulate;
  // epilogue extras
  const float* chan_bias;                // [Co] or null
  const float* plane_bias;               // [N, Co] or null (embedding term + conv bias)
  const float* shared_map;               // [N / planes_per_image, Ho, Wo, Co] fp32 or null (shared-skip conv)
  int planes_per_image;
  float* stats;                          // [2, Co] fp32 (sum, sum of squares) or null
  int act, head_alpha;                   // act = 1: MPI head (packed fp32 [.,4] output)
  void* raw_out;                         // head: sign of the sigma pre-activation, int8 [N, Ho, Wo] or null
  int es;                                // operand element size: 2 = bf16 (kind::f16), 4 = fp32 storage (kind::tf32)
  // filled by the launcher
  int stages, tmem_cols, ipb;
  // halo kernel (conv_halo.cu): every tap offset lies in a 3x3 window -> tap (g, t) = window origin + (rel_y, rel_x)
  int halo_y0, halo_x0;
  int8_t rel_y[4][16], rel_x[4][16];
  int box_stride, w_tile_bytes, w_bytes;
  // TMA-store epilogue of the sub-pixel upsample form (conv_halo.cu): per phase the 128 output rows are staged in shared
  // memory (swizzled, two buffers) and written by one bulk tensor store
  int tma_out, out_rb, out_buf_bytes, out_stage_off;
  // divisors of the work-item decomposition (filled by the launchers)
  FastDiv fd_tiles, fd_tiles_x, fd_n, fd_g, fd_planes;
};

struct ConvLaunch {
  ConvParams p;
  const void* x; int Hi, Wi;             // NHWC bf16 input [N, Hi, Wi, Ci]
  const void* w; int w_rows;             // packed bf16 weights [G*T, w_rows = CB*BN, Ci]
};

struct WgradParams {
  int N, Hg, Wg, TH, TW, KP, tiles_x, tiles_y;
  int G, T, Co, Ci;
  int dy_stride;                         // 1, or 2 when dy is addressed through sub-pixel phases
  int x_stride;                          // 1, or 2 for the weight gradient of a stride-2 convolution
  int16_t dy_oy[4], dy_ox[4];            // phase offsets into dy
  int16_t tap_y[4][16], tap_x[4][16];    // offsets into x
  float* dw;                             // fp32 [G*T, Co, Ci], accumulated with atomics
  int es;                                // operand element size: 2 = bf16, 4 = fp32 storage / TF32 math
  // filled by the launcher
  int a_cb, b_cb, a_slabs, b_slabs, co_blocks, ci_blocks, NB, taps_per_chunk, tap_chunks, stages, tmem_cols;
  // TF32 (fp32 storage): MN-major operands exist only as 128-byte rows (32 fp32) with the 32-byte-atom swizzle.  A
  // 16-channel tensor has 64-byte pixels, so one operand row holds TWO adjacent pixels ("diag" modes): the accumulator
  // then holds the (half_a, half_b) cross terms and the epilogue keeps the two diagonal blocks.
  //   pair = 2     : K rows are pairs of adjacent GEMM pixels (same-resolution layers); both diagonal blocks belong to
  //                  the same weight gradient.
  //   phase_pair   : sub-pixel (upsample) form, Co = Ci = 16: an A row holds the px = 0 / px = 1 gradients of one GEMM
  //                  pixel (adjacent in dy), a B row the low-res pixels x + b, x + b + 1; diagonal block h is the
  //                  gradient of phase (py, px = h).  Groups become the two py phases.
  int pair;                              // 1, or 2
  int phase_pair;                        // 0 / 1
  int rows;                              // K rows per pixel tile
  int a_pairdim, b_pairdim;              // operand read through the overlapping {32 = 2 pixels x 16 ch, W - 1, H, N} map
  int NBp;                               // accumulator columns per tap = NB * (diag ? 2 : 1)
};

struct WgradLaunch {
  WgradParams p;
  const void* dy; int dyH, dyW;          // NHWC bf16 [N, dyH, dyW, Co]
  const void* x; int xH, xW;             // NHWC bf16 [N, xH, xW, Ci]
};

const char* launch_conv_taps(const ConvLaunch& L, cudaStream_t stream);
// conv_halo.cu: returns true when the layer was eligible and has been launched (err set on launch failure)
bool try_launch_conv_halo(const ConvLaunch& L, cudaStream_t stream, const char** err);
void launch_pack_weights(const float* w, int64_t so, int64_t si, int64_t sy, int64_t sx, int Co, int Ci, int mode,
                         int rows_pad, void* out, int es, cudaStream_t stream);
// one entry of the device-resident job table of the batched weight packer (256 threads per block)
struct PackJob {
  const float* w; void* out;
  int64_t so, si, sy, sx;
  int Co, Ci, mode, rows_pad, total, block0;
};
void launch_pack_weights_multi(const PackJob* jobs, int njobs, int nblocks, int es, cudaStream_t stream);
const char* launch_wgrad_taps(const WgradLaunch& L, cudaStream_t stream);
// wgrad_halo.cu: returns true when the layer was eligible and has been launched (err set on launch failure)
bool try_launch_wgrad_halo(const WgradLaunch& L, cudaStream_t stream, const char** err);

}  // namespace mine
