// tcgen05 / TMEM / TMA implicit-GEMM convolution engine for sm_100a.
//
// One kernel family serves every 3x3 convolution of the per-plane MPI decoder (the >95 % FLOP share
// of MINE) in all three directions:
//
//   conv_taps  : D[128 pixels, Co] = sum_taps sum_kblk  X_tap[128 px, KB ch] * W_tap[Co, KB ch]^T
//                - fprop (9 taps on a pre-padded NHWC input),
//                - fused nearest-x2-upsample + 3x3 conv as 4 sub-pixel phases x 4 taps on the
//                  replicate-padded LOW-resolution input (4/9 of the FLOPs, no upsampled tensor),
//                - dgrad (same kernel, transposed tap table / weight pack; zero padding and the
//                  stride-2 gather of the phase form come from TMA out-of-bounds fill and TMA
//                  element strides).
//                A operand = TMA 4-D box {KB ch, TW, TH, 1} of the NHWC tensor shifted by the tap
//                offset -> a K-major 128-row smem tile (hardware swizzle 32/64/128 B);
//                B operand = TMA box of the packed weights [group*tap][Co][Ci]; accumulator in TMEM;
//                one elected thread issues tcgen05.mma; a 4..6-stage mbarrier ring decouples TMA and MMA.
//                Epilogue (4 warps, tcgen05.ld): + shared-skip map (per image, broadcast over planes)
//                + per-plane embedding bias, BatchNorm partial statistics (sum, sum^2 per channel ->
//                smem -> one atomic per channel per CTA), bf16 NHWC store (strided for the phase form),
//                or the MPI head activation (sigmoid rgb, |x|+1e-4 sigma) written straight into the
//                packed [B,S,H,W,4] fp32 MPI the render kernels consume.
//   wgrad_taps : dW[tap][Co, Ci] = sum_pixels dY[px, Co]^T X_tap[px, Ci]  (both operands MN-major smem
//                tiles straight from TMA, all taps of a chunk accumulate side by side in TMEM,
//                persistent over pixel tiles, fp32 vector atomics into the packed weight gradient).
//
// SASS evidence (cuobjdump -sass): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "conv_common.cuh"
#include "conv_engine.h"

namespace mine {


// ------------------------------------------------------------------------------------------------
// conv_taps: fprop / phase-upsample fprop / dgrad
// ------------------------------------------------------------------------------------------------
// MMA-issue loop: the whole warp walks it in uniform control flow (descriptors, TMEM addresses and the instruction
// descriptor stay in uniform registers; conv_common.cuh::elect_one), one elected lane issues.  For the narrow layers
// (N = 16..64) the issue stream - not the tensor pipe - paces the kernel, so the loop is specialised on the operand kind
// and the K steps per row and works on the low descriptor words only.
template <bool TF32, int KS>
__device__ __forceinline__ void taps_issue_loop(const ConvParams& p, uint32_t ring_base, uint32_t stage_bytes,
                                                uint32_t sub_bytes, uint32_t a_bytes, uint32_t tmem_base,
                                                uint64_t* full_bar, uint64_t* empty_bar, uint64_t* accum_full,
                                                uint64_t* accum_empty, int total_work, int iters) {
  const int row_bytes = p.KB * p.es;
  const uint32_t idesc = make_idesc(128, p.BN, 0, 0, TF32);
  const uint64_t desc0 = make_smem_desc(0, 16, 8u * row_bytes, layout_type_for(row_bytes));
  const uint32_t lo0 = (uint32_t)desc0, hi = (uint32_t)(desc0 >> 32);
  const uint32_t sub16 = sub_bytes >> 4, ab16 = a_bytes >> 4;
  const int ipb = p.ipb, stages = p.stages;
  int s = 0, j = 0;
  uint32_t par = 0;
  for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++j) {
    const int as = j & 1;
    if (j >= 2) mbar_wait(&accum_empty[as], ((j >> 1) - 1) & 1);     // epilogue drained this accumulator
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
    uint32_t first = 0;                              // 0 for the very first MMA of the tile (overwrite)
    for (int remaining = iters; remaining > 0;) {
      const int n_in = remaining < ipb ? remaining : ipb;
      mbar_wait(&full_bar[s], par);
      tc_fence_after();
      remaining -= n_in;
      if (elect_one()) {
        uint32_t a_lo = lo0 + ((ring_base + (uint32_t)s * stage_bytes) >> 4);
        for (int u = 0; u < n_in; ++u, a_lo += sub16) {
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            umma_lohi<TF32>(d_tmem, a_lo + 2u * k, a_lo + ab16 + 2u * k, hi, idesc, (u | k) ? 1u : first);
          }
        }
        umma_commit(&empty_bar[s]);           // frees the smem slot once these MMAs retire
        if (remaining == 0) umma_commit(&accum_full[as]);   // accumulator of this tile complete
      }
      __syncwarp();
      first = 1u;
      if (++s == stages) { s = 0; par ^= 1u; }
    }
  }
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_taps_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const ConvParams p) {
  // Persistent: every CTA walks work items (group, image, tile) with a grid stride.  The smem ring runs
  // across tile boundaries and the accumulator is double buffered in TMEM, so the TMA/MMA of tile i+1
  // overlap the epilogue of tile i and the per-CTA setup (barriers, TMEM allocation) is paid once.
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_full[2];
  __shared__ __align__(8) uint64_t accum_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_stats[2][256];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  const int total_work = tiles * p.N * p.G * p.CB;     // w -> (tile, image, group, channel block), tile fastest

  const int row_bytes = p.KB * p.es;                  // == swizzle span
  const uint32_t a_bytes = 128u * row_bytes, b_bytes = (uint32_t)p.BN * row_bytes;
  const uint32_t sub_bytes = a_bytes + ((b_bytes + 1023u) / 1024u) * 1024u;   // one (tap, k-block) operand pair
  const uint32_t stage_bytes = sub_bytes * p.ipb;      // ipb iterations share one barrier round trip
  uint8_t* smem_aligned = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&accum_full[s], 1); mbar_init(&accum_empty[s], 4); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_stats[0][0])[i] = 0.f;
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int iters = p.T * p.kblocks;
  if (warp == 0) {
    // ---------------- TMA producer: one elected lane (an elect.sync region is uniform code for the compiler:
    // coordinates, addresses and the tap table of the parameter block stay on the uniform datapath) ----------------
    if (elect_one()) {
      int s = 0;                                        // ring slot and its parity, carried across tiles
      uint32_t par = 0;
      bool ring_full = false;                           // becomes true once every slot has been used once
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int tile, wi, n_img, gc, g, cb, tile_y, tile_x;      // multiply-high decomposition (conv_engine.h::FastDiv)
        fdivmod(w, p.fd_tiles, wi, tile);
        fdivmod(wi, p.fd_n, gc, n_img);
        fdivmod(gc, p.fd_g, cb, g);
        fdivmod(tile, p.fd_tiles_x, tile_y, tile_x);
        const int wcol0 = cb * p.BN;
        const int oy0 = tile_y * p.TH * p.in_stride, ox0 = tile_x * p.TW * p.in_stride;
        const int wrow0 = g * p.T;
        int u = 0, n_in = 0, remaining = iters;
        uint8_t* slot = nullptr;
        for (int t = 0; t < p.T; ++t) {
          const int iy = oy0 + p.tap_y[g][t], ix = ox0 + p.tap_x[g][t];
          for (int kb = 0; kb < p.kblocks; ++kb) {
            if (u == 0) {                               // open the next pipeline stage
              if (ring_full) mbar_wait(&empty_bar[s], par ^ 1u);
              n_in = remaining < p.ipb ? remaining : p.ipb;
              mbar_expect_tx(&full_bar[s], (uint32_t)n_in * (a_bytes + b_bytes));
              slot = smem_aligned + (size_t)s * stage_bytes;
            }
            tma_load_4d(&map_x, &full_bar[s], slot, kb * p.KB, ix, iy, n_img);
            tma_load_3d(&map_w, &full_bar[s], slot + a_bytes, kb * p.KB, wcol0, wrow0 + t);
            slot += sub_bytes;
            if (++u == n_in) {
              u = 0; remaining -= n_in;
              if (++s == p.stages) { s = 0; par ^= 1u; ring_full = true; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issue: warp-uniform loop, one elected lane issues ----------------
    const int ks = row_bytes / 32;
#define TAPS_ISSUE(TF, KS_) taps_issue_loop<TF, KS_>(p, smem_u32(smem_aligned), stage_bytes, sub_bytes, a_bytes, tmem_base, full_bar, empty_bar, accum_full, accum_empty, total_work, iters)
    if (p.es == 4) { if (ks == 4) TAPS_ISSUE(true, 4); else if (ks == 2) TAPS_ISSUE(true, 2); else TAPS_ISSUE(true, 1); }
    else { if (ks == 4) TAPS_ISSUE(false, 4); else if (ks == 2) TAPS_ISSUE(false, 2); else TAPS_ISSUE(false, 1); }
#undef TAPS_ISSUE
  } else {
    // ---------------- epilogue: 4 warps, warp q owns TMEM lanes [32q, 32q+32) ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // row of the 128-pixel tile
    const int ty = r / p.TW, tx = r - ty * p.TW;
    // BN <= 32: BatchNorm partial sums stay in registers across ALL tiles of this CTA (one cross-lane
    // reduction at the end) instead of 2 x 16 x 5 shuffles per tile
    const bool reg_stats = p.stats && p.BN <= 32 && p.CB == 1;
    float ra1[16], ra2[16], rb1[16], rb2[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) ra1[jj] = ra2[jj] = rb1[jj] = rb2[jj] = 0.f;
    int j = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++j) {
      int tile, wi, n_img, gc, g, cb, tile_y, tile_x;
      fdivmod(w, p.fd_tiles, wi, tile);
      fdivmod(wi, p.fd_n, gc, n_img);
      fdivmod(gc, p.fd_g, cb, g);
      fdivmod(tile, p.fd_tiles_x, tile_y, tile_x);
      const int cbase = cb * p.BN;                           // first output channel of this work item's block
      const int oy = tile_y * p.TH + ty, ox = tile_x * p.TW + tx;
      const bool valid = (oy < p.Hg) && (ox < p.Wg);
      const int out_y = oy * p.out_sy + p.out_oy[g], out_x = ox * p.out_sx + p.out_ox[g];
      const size_t out_pix = ((size_t)n_img * p.Ho + out_y) * p.Wo + out_x;
      const int as = j & 1;
      mbar_wait(&accum_full[as], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(as * p.BN) + ((uint32_t)(q * 32) << 16);
      const float* pbias = p.plane_bias ? p.plane_bias + (size_t)n_img * p.Co + cbase : nullptr;
      const float* cbias = p.chan_bias ? p.chan_bias + cbase : nullptr;
      const float* smap = nullptr;
      if (p.shared_map && valid)
        smap = p.shared_map + (((size_t)fdiv(n_img, p.fd_planes) * p.Ho + out_y) * p.Wo + out_x) * p.Co + cbase;
      const int co_left = p.Co - cbase;                      // real channels in this block (may be < BN: padding)
      auto process_chunk = [&](const uint32_t (&v)[16], const int c0) {
        if (c0 >= co_left) return;
        float f[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) f[jj] = __uint_as_float(v[jj]);
        const bool full16 = (c0 + 16 <= co_left) && ((p.Co & 3) == 0);     // whole chunk real + 16-byte aligned rows
        if (cbias) {
          if (full16) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(cbias + c0 + jj));
              f[jj] += b.x; f[jj + 1] += b.y; f[jj + 2] += b.z; f[jj + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) if (c0 + jj < co_left) f[jj] += cbias[c0 + jj];
          }
        }
        if (pbias) {
          if (full16) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(pbias + c0 + jj));
              f[jj] += b.x; f[jj + 1] += b.y; f[jj + 2] += b.z; f[jj + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) if (c0 + jj < co_left) f[jj] += pbias[c0 + jj];
          }
        }
        if (smap) {
#pragma unroll
          for (int jj = 0; jj < 16; jj += 4) {
            const float4 m = *reinterpret_cast<const float4*>(smap + c0 + jj);
            f[jj] += m.x; f[jj + 1] += m.y; f[jj + 2] += m.z; f[jj + 3] += m.w;
          }
        }
        if (reg_stats) {
          if (c0 == 0) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) { const float x = valid ? f[jj] : 0.f; ra1[jj] += x; ra2[jj] += x * x; }
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) { const float x = valid ? f[jj] : 0.f; rb1[jj] += x; rb2[jj] += x * x; }
          }
        } else if (p.stats) {
          // wide layers: 16 channel sums and 16 sums of squares over the warp's 32 pixels with 2 x 16 shuffles
          // (round 1: 2 x 16 x 5 - the epilogue, not the MMA stream, paced the 64..256-channel layers)
          float x1[16], x2[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) { x1[jj] = valid ? f[jj] : 0.f; x2[jj] = x1[jj] * x1[jj]; }
          const float s1 = warp_reduce16(x1, lane), s2 = warp_reduce16(x2, lane);
          if ((lane & 1) == 0) {
            const int ch = c0 + ((lane >> 1) & 15);
            atomicAdd(&s_stats[0][ch], s1); atomicAdd(&s_stats[1][ch], s2);
          }
        }
        if (valid) {
          if (p.act == 1) {                     // MPI head: 4 real channels -> packed fp32 MPI (+ sign of sigma)
            float4 o;
            o.x = __fdividef(1.f, 1.f + __expf(-f[0]));
            o.y = __fdividef(1.f, 1.f + __expf(-f[1]));
            o.z = __fdividef(1.f, 1.f + __expf(-f[2]));
            o.w = p.head_alpha ? __fdividef(1.f, 1.f + __expf(-f[3])) : fabsf(f[3]) + 1e-4f;
            reinterpret_cast<float4*>(p.out)[out_pix] = o;
            if (p.raw_out) reinterpret_cast<int8_t*>(p.raw_out)[out_pix] = f[3] >= 0.f ? (int8_t)1 : (int8_t)-1;
          } else if (p.out_fp32) {
            float* dst = reinterpret_cast<float*>(p.out) + out_pix * p.Co + cbase + c0;
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              float4 o = make_float4(f[jj], f[jj + 1], f[jj + 2], f[jj + 3]);
              if (p.accumulate) { const float4 e = *reinterpret_cast<float4*>(dst + jj); o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
              *reinterpret_cast<float4*>(dst + jj) = o;
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + out_pix * p.Co + cbase + c0;
            if (p.accumulate) {
              const uint4 e0 = *reinterpret_cast<const uint4*>(dst), e1 = *reinterpret_cast<const uint4*>(dst + 8);
              const __nv_bfloat16* eb0 = reinterpret_cast<const __nv_bfloat16*>(&e0);
              const __nv_bfloat16* eb1 = reinterpret_cast<const __nv_bfloat16*>(&e1);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) { f[jj] += __bfloat162float(eb0[jj]); f[8 + jj] += __bfloat162float(eb1[jj]); }
            }
            uint4 o0, o1;
            __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
            __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              h0[jj] = __floats2bfloat162_rn(f[2 * jj], f[2 * jj + 1]);
              h1[jj] = __floats2bfloat162_rn(f[8 + 2 * jj], f[8 + 2 * jj + 1]);
            }
            *reinterpret_cast<uint4*>(dst) = o0;
            *reinterpret_cast<uint4*>(dst + 8) = o1;
          }
        }
      };
      // two accumulator chunks in flight: the tcgen05.ld of chunk c + 1 overlaps the arithmetic / stores of chunk c
      {
        uint32_t cur[16], nxt[16];
        tmem_ld16_nowait(t_acc, nxt);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          tmem_ld_wait16(nxt);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) cur[jj] = nxt[jj];
          if (c0 + 16 < p.BN) tmem_ld16_nowait(t_acc + (uint32_t)(c0 + 16), nxt);
          process_chunk(cur, c0);
          __syncwarp();
        }
      }
      // this warp has finished reading the accumulator: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accum_empty[as]);
      if (p.stats && p.CB > 1) {
        // consecutive work items of this CTA may belong to different channel blocks: publish and clear the
        // per-CTA partial sums after every tile (CB == 1 keeps them in shared memory until the end)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = threadIdx.x - 64; c < p.BN; c += 128) {
          if (c < co_left) {
            atomicAdd(p.stats + cbase + c, s_stats[0][c]);
            atomicAdd(p.stats + p.Co + cbase + c, s_stats[1][c]);
          }
          s_stats[0][c] = 0.f; s_stats[1][c] = 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    if (reg_stats) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float a1 = warp_sum32(ra1[jj]), a2 = warp_sum32(ra2[jj]);
        const float b1 = warp_sum32(rb1[jj]), b2 = warp_sum32(rb2[jj]);
        if (lane == 0) {
          atomicAdd(&s_stats[0][jj], a1); atomicAdd(&s_stats[1][jj], a2);
          if (p.BN > 16) { atomicAdd(&s_stats[0][16 + jj], b1); atomicAdd(&s_stats[1][16 + jj], b2); }
        }
      }
    }
    if (p.stats && p.CB == 1) {
      asm volatile("bar.sync 1, 128;" ::: "memory");         // epilogue warps only
      const int e = threadIdx.x - 64;                        // 0..127
      for (int c = e; c < p.Co; c += 128) {
        atomicAdd(p.stats + c, s_stats[0][c]);
        atomicAdd(p.stats + p.Co + c, s_stats[1][c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// wgrad_taps
// ------------------------------------------------------------------------------------------------
// Operand rows are K (pixels, or pixel pairs in TF32 pair mode), MN (channels) is contiguous: MN-major descriptors.
//   bf16 : rows of a_cb / b_cb channels (32/64/128 bytes) with the matching 16-byte-atom swizzle; one instruction
//          covers 16 K rows = two 8-row groups SBO apart.
//   tf32 : rows are always 128 bytes (32 fp32) in the SWIZZLE_128B_BASE32B layout (the only MN-major layout of
//          kind::tf32; TMA swizzle 128B_ATOM_32B); the swizzle atom is 4 K rows, one instruction covers 8 K rows =
//          two atoms SBO = 512 bytes apart.
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__global__ void __launch_bounds__(kConvThreads, 1)
wgrad_taps_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // blockIdx.y enumerates (group, tap chunk, co block, ci block)
  int rem = blockIdx.y;
  const int nb = rem % p.ci_blocks; rem /= p.ci_blocks;
  const int mb = rem % p.co_blocks; rem /= p.co_blocks;
  const int tc = rem % p.tap_chunks; rem /= p.tap_chunks;
  const int g = rem;                                     // group (phase); in phase-pair mode the py phase
  const int gt = p.phase_pair ? 2 * g : g;               // row of the tap / offset tables (phase-pair: the px = 0 group)
  const int t0 = tc * p.taps_per_chunk;
  const int ntaps = min(p.taps_per_chunk, p.T - t0);

  const bool tf32 = p.es == 4;
  const int a_row = tf32 ? 128 : p.a_cb * p.es, b_row = tf32 ? 128 : p.b_cb * p.es;   // bytes per K row
  const uint32_t a_slab = (uint32_t)p.rows * a_row, b_slab = (uint32_t)p.rows * b_row;
  const uint32_t a_bytes = a_slab * p.a_slabs;
  const uint32_t b_tap_bytes = b_slab * p.b_slabs;
  const uint32_t stage_bytes = ((a_bytes + b_tap_bytes * p.taps_per_chunk + 1023u) / 1024u) * 1024u;
  uint8_t* smem_aligned = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // pixel tiles handled by this CTA: tile ids blockIdx.x, blockIdx.x + gridDim.x, ...
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_img * p.N;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (elect_one()) {                                   // TMA producer: one elected lane, uniform code
      const int b_cblocks = p.b_slabs / 2;               // channel blocks per parity of the B operand (diag modes, Ci >= 32)
      for (int i = 0; i < my_tiles; ++i) {
        const int tid = blockIdx.x + i * gridDim.x;
        const int n_img = tid / tiles_per_img, tt = tid - n_img * tiles_per_img;
        const int tile_y = tt / p.tiles_x, tile_x = tt - tile_y * p.tiles_x;
        const int oy0 = tile_y * p.TH, ox0 = tile_x * p.TW;
        const int s = i % p.stages, round = i / p.stages;
        if (i >= p.stages) mbar_wait(&empty_bar[s], (round - 1) & 1);
        uint8_t* a_dst = smem_aligned + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + b_tap_bytes * ntaps);
        const int dy_y = oy0 * p.dy_stride + p.dy_oy[gt], dy_x = ox0 * p.dy_stride + p.dy_ox[gt];
        if (p.a_pairdim) {
          tma_load_4d(&map_dy, &full_bar[s], a_dst, 0, dy_x, dy_y, n_img);
        } else {
          for (int sl = 0; sl < p.a_slabs; ++sl)
            tma_load_4d(&map_dy, &full_bar[s], a_dst + sl * a_slab, (mb * p.a_slabs + sl) * p.a_cb, dy_x, dy_y, n_img);
        }
        for (int t = 0; t < ntaps; ++t) {
          uint8_t* b_dst = a_dst + a_bytes + (size_t)t * b_tap_bytes;
          const int iy = oy0 * p.x_stride + p.tap_y[gt][t0 + t], ix = ox0 * p.x_stride + p.tap_x[gt][t0 + t];
          if (p.b_pairdim) {
            tma_load_4d(&map_x, &full_bar[s], b_dst, 0, ix, iy, n_img);
          } else if (p.pair == 2 || p.phase_pair) {
            // slab order (parity q, channel block): accumulator column = q * NB + channel
            for (int sl = 0; sl < p.b_slabs; ++sl) {
              const int q = sl / b_cblocks, cbk = sl - q * b_cblocks;
              tma_load_4d(&map_x, &full_bar[s], b_dst + sl * b_slab, (nb * b_cblocks + cbk) * p.b_cb, ix + q * p.x_stride, iy, n_img);
            }
          } else {
            for (int sl = 0; sl < p.b_slabs; ++sl)
              tma_load_4d(&map_x, &full_bar[s], b_dst + sl * b_slab, (nb * p.b_slabs + sl) * p.b_cb, ix, iy, n_img);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                   // MMA issue: one elected lane, uniform code
      const uint32_t lta = tf32 ? 1u : layout_type_for(a_row), ltb = tf32 ? 1u : layout_type_for(b_row);
      const int kpi = 32 / p.es;                         // K rows per instruction: 16 (bf16) or 8 (tf32)
      const uint32_t sbo_a = tf32 ? 512u : 8u * a_row, sbo_b = tf32 ? 512u : 8u * b_row;
      // A: M = 128 rows = MN blocks LBO apart.  One slab: LBO = 0 (blocks alias block 0, rows >= Co are copies and never
      // stored).  Several slabs but fewer than the instruction needs (tf32, Co = 64): the blocks past Co read the bytes
      // that follow (finite operand data); their accumulator rows are never stored either.
      const uint32_t a_lbo = (p.a_slabs > 1) ? a_slab : 0u;
      // B: the tap tiles of a stage are consecutive slabs, b_slab apart == the MN-block stride of the MN-major
      // descriptor, so ONE instruction spans several taps: N = nblk * block width (<= 256).  A single thread issues every
      // MMA, so few wide instructions instead of many narrow ones is what keeps the tensor pipe busy for the
      // 16/32-channel layers.
      const int bw = tf32 ? 32 : p.b_cb;                 // accumulator columns per MN block of B
      const int total_blocks = ntaps * p.b_slabs;
      const int blk_per_mma = 256 / bw;
      const uint64_t da0 = make_smem_desc(0, a_lbo, sbo_a, lta);
      const uint64_t db0 = make_smem_desc(0, b_slab, sbo_b, ltb);
      // descriptors as (low word, high word): per MMA the issuing thread does two 32-bit adds (see taps_issue_loop)
      const uint32_t a_lo0 = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32), b_lo0 = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
      const uint32_t a_step = (uint32_t)(kpi * a_row) >> 4, b_step = (uint32_t)(kpi * b_row) >> 4, slab16 = b_slab >> 4;
      const int ksteps = p.rows / kpi;
      const uint32_t idesc_full = make_idesc(128, min(blk_per_mma, total_blocks) * bw, 1, 1, tf32);
      const int tail_blocks = total_blocks % blk_per_mma;
      const uint32_t idesc_tail = make_idesc(128, (tail_blocks ? tail_blocks : 1) * bw, 1, 1, tf32);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % p.stages, round = i / p.stages;
        mbar_wait(&full_bar[s], round & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_aligned + (size_t)s * stage_bytes);
        uint32_t a_lo = a_lo0 + (a_addr >> 4);
        uint32_t b_lo_k = b_lo0 + ((a_addr + a_bytes) >> 4);
        uint32_t acc = i > 0 ? 1u : 0u;
        for (int k = 0; k < ksteps; ++k, a_lo += a_step, b_lo_k += b_step) {
          uint32_t b_lo = b_lo_k, d = tmem_base;
          for (int b0 = 0; b0 < total_blocks; b0 += blk_per_mma, b_lo += blk_per_mma * slab16, d += blk_per_mma * bw) {
            const uint32_t idesc = (total_blocks - b0 >= blk_per_mma) ? idesc_full : idesc_tail;
            if (tf32) umma_lohi2<true>(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
            else umma_lohi2<false>(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
          }
          acc = 1u;
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;                           // TMEM lane == accumulator row
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (my_tiles > 0) {
      if (p.pair == 2 || p.phase_pair) {
        // rows: m = half * Co + co (Co == 16); columns of tap t: t * NBp + half' * NB + ci.  Only the diagonal
        // (half == half') blocks are wanted: pixel-pair mode adds both into the same gradient, phase-pair mode stores
        // block h as the gradient of phase (py, px = h).
        const int par = m / p.Co, co = m - par * p.Co;
        if (q == 0) {
          for (int t = 0; t < ntaps; ++t) {
            for (int pp = 0; pp < 2; ++pp) {
              for (int c0 = 0; c0 < p.NB; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + (uint32_t)(t * p.NBp + pp * p.NB + c0), v);
                const int ci = nb * p.NB + c0;
                if (par != pp || ci >= p.Ci) continue;
                const int gd = p.phase_pair ? 2 * g + pp : g;
                float* dst = p.dw + (((size_t)(gd * p.T + t0 + t) * p.Co + co) * p.Ci + ci);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                               "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                               : "memory");
                }
              }
            }
          }
        }
      } else {
        const int co = mb * 128 + m;
        for (int t = 0; t < ntaps; ++t) {
          for (int c0 = 0; c0 < p.NB; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.NB + c0), v);
            if (co >= p.Co) continue;
            const int ci = nb * p.NB + c0;
            if (ci >= p.Ci) continue;
            float* dst = p.dw + (((size_t)(g * p.T + t0 + t) * p.Co + co) * p.Ci + ci);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps (cached) + launchers
// ------------------------------------------------------------------------------------------------
// The driver API is resolved at run time (cudaGetDriverEntryPoint) so the extension links against
// libcudart only and can be imported on machines without libcuda.so.1 (the CPU build check).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

struct MapKey {
  const void* ptr; int d0, d1, d2, d3, b0, b1, b2, b3, e1, e2, rank, es;
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, d0, d1, d2, d3, b0, b1, b2, b3, e1, e2, rank, es) <
           std::tie(o.ptr, o.d0, o.d1, o.d2, o.d3, o.b0, o.b1, o.b2, o.b3, o.e1, o.e2, o.rank, o.es);
  }
};
static CUtensorMapDataType dtype_for(int es) { return es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

// NHWC activation: dims {C, W, H, N}; box {cb, bw * esx, bh * esy, 1}; element strides esx / esy on W / H (the box then
// holds bw x bh pixels).  mn32: SWIZZLE_128B with 32-byte atoms (MN-major kind::tf32 operands), else by row bytes.
const char* nhwc_map(CUtensorMap* out, const void* ptr, int C, int W, int H, int N, int cb, int bw, int bh, int esx,
                     int esy, int esz, bool mn32) {
  MapKey key{ptr, C, W, H, N, cb, bw * esx, bh * esy, mn32 ? 2 : 1, esx, esy, 4, esz};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return nullptr; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * esz, (cuuint64_t)W * C * esz, (cuuint64_t)H * W * C * esz};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)(bw * esx), (cuuint32_t)(bh * esy), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)esx, (cuuint32_t)esy, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled is not available from this driver";
  CUresult r = enc(out, dtype_for(esz), 4, const_cast<void*>(ptr), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      mn32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle_for(cb * esz),
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed (bad shape / stride / alignment)";
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = *out;
  return nullptr;
}

// 16-channel fp32 NHWC activation read as 128-byte rows of TWO ADJACENT pixels (the "diag" modes of the TF32 weight
// gradient): dims {32 (2 pixels x 16 channels, contiguous), W - 1 (row start: any pixel, stride ONE pixel - the rows of
// this view overlap), H, N}; box {32, bw * esx, bh * esy, 1} with element strides esx / esy -> bw x bh rows.  A start
// at the last pixel of a line is out of bounds (zero row), so a row never straddles two image lines.
const char* overlap32_map(CUtensorMap* out, const void* ptr, int W, int H, int N, int bw, int bh, int esx, int esy) {
  MapKey key{ptr, 32, W, H, N, 32, bw * esx, bh * esy, 3, esx, esy, 4, 4};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return nullptr; }
  if (W < 2) return "pair map: tensor too narrow";
  cuuint64_t dims[4] = {32, (cuuint64_t)(W - 1), (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
  cuuint32_t box[4] = {32, (cuuint32_t)(bw * esx), (cuuint32_t)(bh * esy), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)esx, (cuuint32_t)esy, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled is not available from this driver";
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the pair map (bad shape / stride / alignment)";
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = *out;
  return nullptr;
}

// Output of the sub-pixel upsample convolution [N, 2 Hg, 2 Wg, C] seen as {C, px (2), Wg, py (2), N * Hg}: the pixels of ONE
// phase (py, px) of a low-resolution tile are the box {C, 1, bw, 1, bh} at (0, px, x0, py, n * Hg + y0) - the TMA store
// of the conv_halo epilogue.  Row r of the box = low-res pixel (r / bw, r % bw) = the accumulator row of that pixel.
const char* phase_out_map(CUtensorMap* out, const void* ptr, int C, int Wg, int NH, int bw, int bh, int esz) {
  MapKey key{ptr, C, Wg, NH, 2, C, bw, bh, 5, 1, 1, 5, esz};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return nullptr; }
  const cuuint64_t rowb = (cuuint64_t)C * esz;
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)Wg, 2, (cuuint64_t)NH};
  cuuint64_t strides[4] = {rowb, 2 * rowb, 2 * (cuuint64_t)Wg * rowb, 4 * (cuuint64_t)Wg * rowb};
  cuuint32_t box[5] = {(cuuint32_t)C, 1, (cuuint32_t)bw, 1, (cuuint32_t)bh};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled is not available from this driver";
  CUresult r = enc(out, dtype_for(esz), 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for((int)rowb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the phase output map";
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = *out;
  return nullptr;
}

// packed weights: dims {Ci, Co_pad, GT}; box {kb, bn, 1}
const char* weight_map(CUtensorMap* out, const void* ptr, int Ci, int Cop, int GT, int kb, int bn, int esz) {
  MapKey key{ptr, Ci, Cop, GT, 0, kb, bn, 1, 0, 1, 1, 3, esz};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return nullptr; }
  cuuint64_t dims[3] = {(cuuint64_t)Ci, (cuuint64_t)Cop, (cuuint64_t)GT};
  cuuint64_t strides[2] = {(cuuint64_t)Ci * esz, (cuuint64_t)Cop * Ci * esz};
  cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)bn, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled is not available from this driver";
  CUresult r = enc(out, dtype_for(esz), 3, const_cast<void*>(ptr), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kb * esz),
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed (bad shape / stride / alignment)";
  g_maps[key] = *out;
  return nullptr;
}

int next_pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }
int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n;
}

// ---- weight packing: fp32 [Co,Ci,3,3] (any strides) -> bf16 GEMM operand packs, one launch --------------
// mode 0: fprop 3x3        out[ky*3+kx][co][ci]
// mode 1: fprop upsample   out[(py*2+px)*4 + a*2+b][co][ci] = sum over the 3x3 taps that collapse onto (a,b)
// mode 2: dgrad 3x3        out[ky*3+kx][ci][co]
// mode 3: dgrad upsample   out[(py*2+px)*4 + a*2+b][ci][co]
// rows (the GEMM N dimension) are zero-padded to `rows_pad` (multiple of 16).
__device__ __forceinline__ bool phase_has(int p, int a, int k) {      // does 3-tap index k fold onto 2-tap index a?
  return p == 0 ? (a == 0 ? k == 0 : k >= 1) : (a == 0 ? k <= 1 : k == 2);
}
__device__ __forceinline__ void store_operand(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void store_operand(float* p, float v) {   // round to nearest TF32 (the MMA would truncate)
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  *p = __uint_as_float(r);
}
// element `idx` of the [9 | 16, rows_pad, cols] operand pack of one [Co, Ci, 3, 3] weight (modes: header comment)
__device__ __forceinline__ float pack_value(const float* __restrict__ w, int64_t so, int64_t si, int64_t sy, int64_t sx,
                                            int Co, int Ci, int mode, int rows_pad, int idx) {
  const bool dgrad = mode >= 2, up = (mode & 1) != 0;
  const int cols = dgrad ? Co : Ci;
  const int col = idx % cols;
  const int row = (idx / cols) % rows_pad;
  const int gt = idx / (cols * rows_pad);
  const int co = dgrad ? col : row, ci = dgrad ? row : col;
  float v = 0.f;
  if (co < Co && ci < Ci) {
    const float* base = w + co * so + ci * si;
    if (!up) {
      v = base[(gt / 3) * sy + (gt % 3) * sx];
    } else {
      const int g = gt >> 2, t = gt & 3, py = g >> 1, px = g & 1, a = t >> 1, b = t & 1;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
          if (phase_has(py, a, ky) && phase_has(px, b, kx)) v += base[ky * sy + kx * sx];
    }
  }
  return v;
}

template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w, int64_t so, int64_t si, int64_t sy, int64_t sx, int Co,
                                    int Ci, int mode, int rows_pad, T* __restrict__ out, int total) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  store_operand(out + idx, pack_value(w, so, si, sy, sx, Co, Ci, mode, rows_pad, idx));
}

// All operand packs of a model in ONE launch: block b belongs to the job whose [block0, block0 + blocks) range holds it
// (the job table lives in device memory, built once per plan: conv_bindings.cpp::pack_plan_create).
template <typename T>
__global__ void pack_weights_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ int s_job;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;                          // last job with block0 <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_job = lo;
  }
  __syncthreads();
  const PackJob j = jobs[s_job];
  const int idx = ((int)blockIdx.x - j.block0) * blockDim.x + threadIdx.x;
  if (idx >= j.total) return;
  store_operand(reinterpret_cast<T*>(j.out) + idx, pack_value(j.w, j.so, j.si, j.sy, j.sx, j.Co, j.Ci, j.mode, j.rows_pad, idx));
}

void launch_pack_weights(const float* w, int64_t so, int64_t si, int64_t sy, int64_t sx, int Co, int Ci, int mode,
                         int rows_pad, void* out, int es, cudaStream_t stream) {
  const int gt = (mode & 1) ? 16 : 9;
  const int cols = mode >= 2 ? Co : Ci;
  const int total = gt * rows_pad * cols;
  if (es == 4)
    pack_weights_kernel<float><<<(total + 255) / 256, 256, 0, stream>>>(w, so, si, sy, sx, Co, Ci, mode, rows_pad,
                                                                        (float*)out, total);
  else
    pack_weights_kernel<__nv_bfloat16><<<(total + 255) / 256, 256, 0, stream>>>(w, so, si, sy, sx, Co, Ci, mode, rows_pad,
                                                                                (__nv_bfloat16*)out, total);
}

void launch_pack_weights_multi(const PackJob* jobs, int njobs, int nblocks, int es, cudaStream_t stream) {
  if (njobs <= 0 || nblocks <= 0) return;
  if (es == 4) pack_weights_multi_kernel<float><<<nblocks, 256, 0, stream>>>(jobs, njobs);
  else pack_weights_multi_kernel<__nv_bfloat16><<<nblocks, 256, 0, stream>>>(jobs, njobs);
}

const char* launch_conv_taps(const ConvLaunch& L, cudaStream_t stream) {
  {   // narrow high-resolution layers: halo kernel (three activation boxes per tile, resident weights)
    const char* herr = nullptr;
    if (try_launch_conv_halo(L, stream, &herr)) return herr;
  }
  ConvParams p = L.p;
  if (p.TH * p.TW != 128) return "tile must cover 128 pixels";
  if (p.es != 2 && p.es != 4) return "operand element size must be 2 (bf16) or 4 (fp32/tf32)";
  if (p.KB * p.es != 32 && p.KB * p.es != 64 && p.KB * p.es != 128) return "K block must span 32/64/128 bytes";
  if (p.Ci % p.KB) return "Ci must be a multiple of KB";
  if (p.BN % 16 || p.BN < 16 || p.BN > 256) return "BN must be a multiple of 16 in [16,256]";
  if (p.T > 16 || p.G > 4) return "too many taps/groups";
  if (p.CB < 1) p.CB = 1;
  if (p.CB > 1 && (p.Co != p.CB * p.BN || p.act != 0)) return "channel blocks must tile Co exactly (and no head epilogue)";
  p.kblocks = p.Ci / p.KB;
  p.tiles_x = (p.Wg + p.TW - 1) / p.TW;
  p.tiles_y = (p.Hg + p.TH - 1) / p.TH;
  p.fd_tiles = make_fastdiv(p.tiles_x * p.tiles_y);
  p.fd_tiles_x = make_fastdiv(p.tiles_x);
  p.fd_n = make_fastdiv(p.N);
  p.fd_g = make_fastdiv(p.G);
  p.fd_planes = make_fastdiv(p.planes_per_image > 0 ? p.planes_per_image : 1);
  p.tmem_cols = next_pow2_cols(2 * p.BN);                  // double-buffered accumulator
  const uint32_t sub_bytes = 128u * p.KB * p.es + (((uint32_t)p.BN * p.KB * p.es + 1023u) / 1024u) * 1024u;
  int ipb = (int)(24u * 1024u / sub_bytes);                // (tap, k-block) iterations per barrier round trip
  if (ipb > p.T * p.kblocks) ipb = p.T * p.kblocks;
  if (ipb > 12) ipb = 12;
  if (ipb < 1) ipb = 1;
  // tuning overrides for experiments (pipeline depth / iterations per barrier)
  static const int env_ipb = getenv("MINE_CONV_IPB") ? atoi(getenv("MINE_CONV_IPB")) : 0;
  static const int env_stages = getenv("MINE_CONV_STAGES") ? atoi(getenv("MINE_CONV_STAGES")) : 0;
  if (env_ipb > 0) { ipb = env_ipb; if (ipb > p.T * p.kblocks) ipb = p.T * p.kblocks; }
  p.ipb = ipb;
  const uint32_t stage_bytes = sub_bytes * ipb;
  int stages = (int)(96u * 1024u / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  if (env_stages > 0) stages = env_stages;
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + 1024;
  // resident CTAs per SM: limited by TMEM columns (512) and shared memory; pad smem so exactly that many fit
  int ctas_per_sm = 512 / p.tmem_cols;
  const int by_smem = (int)((220u * 1024u) / (smem + 2048));
  if (ctas_per_sm > by_smem) ctas_per_sm = by_smem;
  if (ctas_per_sm > 2) ctas_per_sm = 2;                      // ~126 registers x 192 threads: two CTAs per SM
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const size_t smem_floor = (220u * 1024u) / (ctas_per_sm + 1) + 1024;   // one more CTA must NOT fit
  if (smem < smem_floor) smem = smem_floor;
  if (smem > 200u * 1024u) smem = 200u * 1024u;
  CUtensorMap mx, mw;
  const char* e = nhwc_map(&mx, L.x, p.Ci, L.Wi, L.Hi, p.N, p.KB, p.TW, p.TH, p.in_stride, p.in_stride, p.es);
  if (e) return e;
  e = weight_map(&mw, L.w, p.Ci, L.w_rows, p.G * p.T, p.KB, p.BN, p.es);
  if (e) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
  const int total_work = p.tiles_x * p.tiles_y * p.N * p.G * p.CB;
  int grid_x = num_sms * ctas_per_sm;
  if (grid_x > total_work) grid_x = total_work;
  conv_taps_kernel<<<grid_x, kConvThreads, smem, stream>>>(mx, mw, p);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

const char* launch_wgrad_taps(const WgradLaunch& L, cudaStream_t stream) {
  {   // narrow / high-resolution layers: tap-sharing kernel (one activation load per horizontal offset, wgrad_halo.cu)
    const char* herr = nullptr;
    if (try_launch_wgrad_halo(L, stream, &herr)) return herr;
  }
  WgradParams p = L.p;
  if (p.TH * p.TW != p.KP || p.KP % 16 || p.KP < 32 || p.KP > 256) return "wgrad pixel tile must be 32..256 pixels";
  if (p.es != 2 && p.es != 4) return "operand element size must be 2 (bf16) or 4 (fp32/tf32)";
  if (p.x_stride < 1) p.x_stride = 1;
  const bool tf32 = p.es == 4;
  p.pair = 1; p.phase_pair = 0; p.a_pairdim = p.b_pairdim = 0;
  p.NB = p.Ci < 128 ? p.Ci : 128;                        // N per accumulator
  if (p.Ci % p.NB) return "Ci must be <= 128 or a multiple of 128";
  p.ci_blocks = p.Ci / p.NB;
  if (!tf32) {
    // operand slabs: channel block = min(C, 64) with the matching swizzle
    p.a_cb = p.Co < 64 ? p.Co : 64;
    p.b_cb = p.Ci < 64 ? p.Ci : 64;
    if (p.a_cb != 16 && p.a_cb != 32 && p.a_cb != 64) return "Co must be 16/32/multiple of 64";
    if (p.b_cb != 16 && p.b_cb != 32 && p.b_cb != 64) return "Ci must be 16/32/multiple of 64";
    if (p.Co % p.a_cb || p.Ci % p.b_cb) return "channels must be a multiple of the slab width";
    p.a_slabs = p.Co >= 128 ? 2 : 1;                     // M = 128 rows: 2 real slabs, or 1 slab aliased by LBO = 0
    p.co_blocks = (p.Co + 127) / 128;
    p.b_slabs = p.NB / p.b_cb;
  } else if (p.Co == 16 || p.Ci == 16) {
    // diag modes (see WgradParams): operand rows hold two adjacent pixels
    if (p.Co != 16) return "TF32 diag modes expect the 16-channel tensor on the gradient (dy) side";
    if (p.Ci != 16 && p.Ci % 32) return "Ci must be 16 or a multiple of 32";
    p.a_cb = 16; p.a_slabs = 1; p.a_pairdim = 1; p.co_blocks = 1;
    if (p.dy_stride == 2) {
      if (p.G != 4 || p.x_stride != 1) return "TF32 phase-pair mode needs the four sub-pixel groups";
      for (int g2 = 0; g2 < 2; ++g2) {          // group tables must be the canonical (py, px) ones: px only shifts x by one
        if (p.dy_ox[2 * g2] != 0 || p.dy_ox[2 * g2 + 1] != 1 || p.dy_oy[2 * g2] != p.dy_oy[2 * g2 + 1]) return "phase-pair: unexpected group offsets";
        for (int t = 0; t < p.T; ++t)
          if (p.tap_x[2 * g2 + 1][t] != p.tap_x[2 * g2][t] + 1 || p.tap_y[2 * g2 + 1][t] != p.tap_y[2 * g2][t]) return "phase-pair: unexpected tap table";
      }
      p.phase_pair = 1;
      if (p.Ci == 16) { p.b_cb = 16; p.b_slabs = 1; p.b_pairdim = 1; }
      else { p.b_cb = 32; p.b_slabs = 2 * (p.NB / 32); }
    } else {
      if (p.dy_stride != 1 || p.x_stride != 1) return "TF32 pixel-pair mode needs unit strides";
      if (p.TW % 2) return "TF32 pixel-pair mode needs an even tile width";
      p.pair = 2;
      if (p.Ci == 16) { p.b_cb = 16; p.b_slabs = 1; p.b_pairdim = 1; }
      else { p.b_cb = 32; p.b_slabs = 2 * (p.NB / 32); }
    }
  } else {
    if (p.Co % 32 || p.Ci % 32) return "TF32 weight gradient: channels must be 16 or a multiple of 32";
    p.a_cb = p.b_cb = 32;
    p.a_slabs = (p.Co < 128 ? p.Co : 128) / 32;
    p.co_blocks = (p.Co + 127) / 128;
    p.b_slabs = p.NB / 32;
  }
  p.NBp = p.NB * ((p.pair == 2 || p.phase_pair) ? 2 : 1);
  p.rows = p.KP / p.pair;
  if (p.rows % (32 / p.es)) return "pixel tile too small for the K step";
  const uint32_t a_rowb = tf32 ? 128u : (uint32_t)p.a_cb * p.es, b_rowb = tf32 ? 128u : (uint32_t)p.b_cb * p.es;
  p.taps_per_chunk = 512 / p.NBp;
  if (p.taps_per_chunk > p.T) p.taps_per_chunk = p.T;
  // keep the stage under ~96 KB (>= 2 stages in flight)
  for (;;) {
    const uint32_t sb = (uint32_t)p.rows * a_rowb * p.a_slabs + (uint32_t)p.rows * b_rowb * p.b_slabs * p.taps_per_chunk;
    if (sb <= 96 * 1024 || p.taps_per_chunk == 1) break;
    p.taps_per_chunk = (p.taps_per_chunk + 1) / 2;
  }
  p.tap_chunks = (p.T + p.taps_per_chunk - 1) / p.taps_per_chunk;
  p.tmem_cols = next_pow2_cols(p.taps_per_chunk * p.NBp);
  p.tiles_x = (p.Wg + p.TW - 1) / p.TW;
  p.tiles_y = (p.Hg + p.TH - 1) / p.TH;
  const uint32_t stage_bytes =
      (((uint32_t)p.rows * a_rowb * p.a_slabs + (uint32_t)p.rows * b_rowb * p.b_slabs * p.taps_per_chunk + 1023u) / 1024u) * 1024u;
  int stages = (int)(192u * 1024u / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  if (smem > 200u * 1024u) return "wgrad stage does not fit in shared memory";
  if (p.TW * p.dy_stride > 256 || p.TH * p.dy_stride > 256) return "strided wgrad tile exceeds the TMA box limit";
  if (p.TW * p.x_stride > 256 || p.TH * p.x_stride > 256) return "strided wgrad tile exceeds the TMA box limit";
  CUtensorMap mdy, mx;
  const char* e;
  if (p.phase_pair) e = overlap32_map(&mdy, L.dy, L.dyW, L.dyH, p.N, p.TW, p.TH, 2, 2);          // rows start at pixel 2 * ox
  else if (p.a_pairdim) e = overlap32_map(&mdy, L.dy, L.dyW, L.dyH, p.N, p.TW / 2, p.TH, 2, 1);   // rows start at pixel 2 * j
  else e = nhwc_map(&mdy, L.dy, p.Co, L.dyW, L.dyH, p.N, p.a_cb, p.TW, p.TH, p.dy_stride, p.dy_stride, p.es, tf32);
  if (e) return e;
  if (p.phase_pair && p.b_pairdim) e = overlap32_map(&mx, L.x, L.xW, L.xH, p.N, p.TW, p.TH, 1, 1);  // rows x + b, x + b + 1
  else if (p.phase_pair) e = nhwc_map(&mx, L.x, p.Ci, L.xW, L.xH, p.N, p.b_cb, p.TW, p.TH, 1, 1, p.es, true);
  else if (p.b_pairdim) e = overlap32_map(&mx, L.x, L.xW, L.xH, p.N, p.TW / 2, p.TH, 2, 1);
  else if (p.pair == 2) e = nhwc_map(&mx, L.x, p.Ci, L.xW, L.xH, p.N, p.b_cb, p.TW / 2, p.TH, 2 * p.x_stride, p.x_stride, p.es, true);
  else e = nhwc_map(&mx, L.x, p.Ci, L.xW, L.xH, p.N, p.b_cb, p.TW, p.TH, p.x_stride, p.x_stride, p.es, tf32);
  if (e) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int combos = (p.phase_pair ? p.G / 2 : p.G) * p.tap_chunks * p.co_blocks * p.ci_blocks;
  const int total_tiles = p.tiles_x * p.tiles_y * p.N;
  int splits = (148 * 2 + combos - 1) / combos;
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  dim3 grid(splits, combos, 1);
  wgrad_taps_kernel<<<grid, kConvThreads, smem, stream>>>(mdy, mx, p);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

}  // namespace mine
