// Halo variant of the tcgen05 implicit-GEMM convolution for the narrow, high-resolution layers of the MPI decoder
// (16..64 channels at full / half resolution: same-resolution 3x3 convs, the fused x2-upsample conv in its sub-pixel
// form, the MPI heads, and the 3x3 data gradients).
//
// conv_taps_kernel (conv_tcgen05.cu) issues one TMA box per (tile, tap, k-block) plus the matching weight box: nine
// activation re-loads of 32..128-byte rows and nine weight re-loads per 128-pixel tile.  For the narrow layers that
// TMA row traffic - not HBM, not the tensor pipe - sets the time (profiles/ncu_head_0.txt: DRAM 7 %, tensor 3 %).
// Here, per tile:
//   * activations: THREE boxes of (TH + 2) x TW pixels (one per horizontal tap offset).  A vertical tap offset dy is
//     a shift of dy * TW rows inside a box; with TW a multiple of 8 that shift is a whole number of 8-row swizzle
//     atoms, so the K-major UMMA descriptor of tap (dy, dx) is just  box[dx] + dy * TW * row_bytes  - nine MMAs read
//     three loads (2.4x - 2.7x fewer rows through the TMA unit);
//   * weights: ALL taps / k-blocks of the layer stay resident in shared memory for the lifetime of the CTA (loaded once);
//   * the sub-pixel (upsample) form keeps the four phase accumulators of a low-resolution tile in TMEM together, so
//     one halo feeds 16 MMAs and the epilogue writes the 2x2 output pixels of every input pixel.
// Pipeline / roles as in conv_taps_kernel: warp 0 TMA producer, warp 1 single-thread MMA issue (kind::f16 or
// kind::tf32), warps 2..5 epilogue (tcgen05.ld, bias / shared map / BatchNorm partial sums / head activation),
// persistent CTAs, TMEM accumulators double buffered across tiles.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "conv_common.cuh"
#include "conv_engine.h"

namespace mine {

// MMA-issue loop of one CTA: the whole warp walks it (uniform control flow -> descriptors in uniform registers), one
// elected lane issues.  Tap (g, t) of k-block kb reads the halo set at  rel_x * box_stride + rel_y * line_bytes  and the
// resident weight tile (g * T + t) * kblocks + kb; both offsets come straight from the parameter block.
template <bool TF32, int KS>
__device__ __forceinline__ void halo_issue_loop(const ConvParams& p, uint32_t w_base, uint32_t ring_base, uint32_t set_bytes,
                                                uint32_t tmem_base, uint64_t* full_bar, uint64_t* empty_bar,
                                                uint64_t* accum_full, uint64_t* accum_empty, int total_work) {
  const int row_bytes = p.KB * p.es;
  const uint32_t idesc = make_idesc(128, p.BN, 0, 0, TF32);
  const uint64_t desc0 = make_smem_desc(0, 16, 8u * row_bytes, layout_type_for(row_bytes));
  const uint32_t lo0 = (uint32_t)desc0, hi = (uint32_t)(desc0 >> 32);
  const int G = p.G, T = p.T, KBK = p.kblocks, BN = p.BN;
  const uint32_t line16 = (uint32_t)(p.TW * row_bytes) >> 4, box16 = (uint32_t)p.box_stride >> 4;
  const uint32_t wt16 = (uint32_t)p.w_tile_bytes >> 4;
  int s = 0, j = 0;
  uint32_t par = 0;
  for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++j) {
    const int as = j & 1;
    if (j >= 2) mbar_wait(&accum_empty[as], ((j >> 1) - 1) & 1);
    tc_fence_after();
    for (int kb = 0; kb < KBK; ++kb) {
      mbar_wait(&full_bar[s], par);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t slot_lo = lo0 + ((ring_base + (uint32_t)s * set_bytes) >> 4);
        uint32_t b_lo = lo0 + (w_base >> 4) + (uint32_t)kb * wt16;
        uint32_t d_tmem = tmem_base + (uint32_t)(as * G * BN);
        for (int g = 0; g < G; ++g, d_tmem += BN) {
          uint32_t acc = kb ? 1u : 0u;
          for (int t = 0; t < T; ++t, b_lo += (uint32_t)KBK * wt16) {
            const uint32_t a_lo = slot_lo + (uint32_t)p.rel_x[g][t] * box16 + (uint32_t)p.rel_y[g][t] * line16;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
              umma_lohi<TF32>(d_tmem, a_lo + 2u * k, b_lo + 2u * k, hi, idesc, acc);
              acc = 1u;
            }
          }
        }
        umma_commit(&empty_bar[s]);
        if (kb == KBK - 1) umma_commit(&accum_full[as]);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; par ^= 1u; }
    }
  }
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_o, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_full[2];
  __shared__ __align__(8) uint64_t accum_empty[2];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_stats[2][64];
  __shared__ __align__(16) float s_cbias[64];                        // per-channel bias (zeros without one): read once per CTA

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  const int total_work = tiles * p.N;                  // w -> (tile, image); the G groups share one halo
  const int row_bytes = p.KB * p.es;
  const uint32_t box_bytes = (uint32_t)(p.TH + 2) * p.TW * row_bytes;
  const uint32_t set_bytes = 3u * p.box_stride;
  uint8_t* smem_aligned = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint8_t* w_smem = smem_aligned;                      // [G*T][kblocks] tiles of BN x KB, w_tile_bytes apart
  uint8_t* ring = smem_aligned + p.w_bytes;            // stages x {3 boxes}

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&accum_full[s], 1); mbar_init(&accum_empty[s], 4); }
    mbar_init(&w_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) (&s_stats[0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_cbias[i] = (p.chan_bias && i < p.Co) ? p.chan_bias[i] : 0.f;
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ---------------- TMA producer: warp-uniform loop, one elected lane issues ----------------
    if (elect_one()) {
      // resident weights: every (group, tap, k-block) tile once per CTA
      const int gt_n = p.G * p.T;
      mbar_expect_tx(&w_bar, (uint32_t)gt_n * p.kblocks * (uint32_t)p.BN * row_bytes);
      for (int gt = 0; gt < gt_n; ++gt)
        for (int kb = 0; kb < p.kblocks; ++kb)
          tma_load_3d(&map_w, &w_bar, w_smem + (size_t)(gt * p.kblocks + kb) * p.w_tile_bytes, kb * p.KB, 0, gt);
    }
    __syncwarp();
    int s = 0;
    uint32_t par = 0;
    bool ring_full = false;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int tile, n_img, tile_y, tile_x;
      fdivmod(w, p.fd_tiles, n_img, tile);
      fdivmod(tile, p.fd_tiles_x, tile_y, tile_x);
      const int iy = tile_y * p.TH + p.halo_y0, ix = tile_x * p.TW + p.halo_x0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        if (ring_full) mbar_wait(&empty_bar[s], par ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], 3u * box_bytes);
          uint8_t* slot = ring + (size_t)s * set_bytes;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            tma_load_4d(&map_x, &full_bar[s], slot + dx * p.box_stride, kb * p.KB, ix + dx, iy, n_img);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; par ^= 1u; ring_full = true; }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issue: warp-uniform loop, one elected lane issues ----------------
    mbar_wait(&w_bar, 0);
    const int ks = row_bytes / 32;
#define HALO_ISSUE(TF, KS_) halo_issue_loop<TF, KS_>(p, smem_u32(w_smem), smem_u32(ring), set_bytes, tmem_base, full_bar, empty_bar, accum_full, accum_empty, total_work)
    if (p.es == 4) { if (ks == 4) HALO_ISSUE(true, 4); else if (ks == 2) HALO_ISSUE(true, 2); else HALO_ISSUE(true, 1); }
    else { if (ks == 4) HALO_ISSUE(false, 4); else if (ks == 2) HALO_ISSUE(false, 2); else HALO_ISSUE(false, 1); }
#undef HALO_ISSUE
  } else {
    // ---------------- epilogue: 4 warps, warp q owns TMEM lanes [32q, 32q+32) ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ty = r / p.TW, tx = r - ty * p.TW;
    const bool reg_stats = p.stats != nullptr;           // BN <= 32 (launcher): sums stay in registers across tiles
    const bool has_cbias = p.chan_bias != nullptr;
    float ra1[16], ra2[16], rb1[16], rb2[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) ra1[jj] = ra2[jj] = rb1[jj] = rb2[jj] = 0.f;
    int j = 0;
    // TMA-store epilogue (sub-pixel upsample form): row r of a staging buffer = this thread's pixel, 16-byte pieces
    // XOR-swizzled like the tensor map expects; `gcount` alternates the two buffers across phases and tiles
    uint8_t* out_stage = smem_aligned + p.out_stage_off;
    const uint32_t swz_mask = p.out_rb >= 128 ? 7u : (p.out_rb >= 64 ? 3u : 1u);
    const int epi_tid = (int)threadIdx.x - 64;
    uint32_t gcount = 0;
    if (p.tma_out && epi_tid == 0) tma_prefetch_desc(&map_o);
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++j) {
      int tile, n_img, tile_y, tile_x;                     // multiply-high decomposition: no integer divides per tile
      fdivmod(w, p.fd_tiles, n_img, tile);
      fdivmod(tile, p.fd_tiles_x, tile_y, tile_x);
      const int oy = tile_y * p.TH + ty, ox = tile_x * p.TW + tx;
      const bool valid = (oy < p.Hg) && (ox < p.Wg);
      const int as = j & 1;
      const float* pbias = p.plane_bias ? p.plane_bias + (size_t)n_img * p.Co : nullptr;
      const size_t img_pix = (size_t)n_img * p.Ho;
      const size_t smap_img = p.shared_map ? (size_t)fdiv(n_img, p.fd_planes) * p.Ho : 0;
      mbar_wait(&accum_full[as], (j >> 1) & 1);
      tc_fence_after();
      // the accumulators of all groups are contiguous in TMEM: walk them as one range, two 16-column loads in flight per
      // wait (a 16-column chunk never straddles two groups: BN is a multiple of 16)
      const uint32_t t_acc0 = tmem_base + (uint32_t)(as * p.G * p.BN) + ((uint32_t)(q * 32) << 16);
      const int ncols_all = p.G * p.BN;
      int g = 0, cc = 0;                                   // group / first channel of the current 16-column chunk
      for (int f0 = 0; f0 < ncols_all; f0 += 32) {
        uint32_t v[32];
        const bool two = f0 + 16 < ncols_all;
        tmem_ld16_nowait(t_acc0 + (uint32_t)f0, v);
        if (two) tmem_ld16_nowait(t_acc0 + (uint32_t)(f0 + 16), v + 16);
        tmem_ld_wait16(v);
        tmem_ld_wait16(v + 16);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !two) break;
          const int g_now = g, cc_now = cc;
          cc += 16;
          if (cc >= p.BN) { cc = 0; ++g; }
          if (cc_now >= p.Co) continue;
          const int out_y = oy * p.out_sy + p.out_oy[g_now], out_x = ox * p.out_sx + p.out_ox[g_now];
          const size_t out_pix = (img_pix + out_y) * p.Wo + out_x;
          float fv[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) fv[jj] = __uint_as_float(v[16 * h + jj]);
          if (has_cbias) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              const float4 b = *reinterpret_cast<const float4*>(&s_cbias[cc_now + jj]);
              fv[jj] += b.x; fv[jj + 1] += b.y; fv[jj + 2] += b.z; fv[jj + 3] += b.w;
            }
          }
          if (pbias) {
            if (cc_now + 16 <= p.Co) {            // Co is a multiple of 4 whenever a plane bias exists (decoder widths)
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(pbias + cc_now + jj));
                fv[jj] += b.x; fv[jj + 1] += b.y; fv[jj + 2] += b.z; fv[jj + 3] += b.w;
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) if (cc_now + jj < p.Co) fv[jj] += pbias[cc_now + jj];
            }
          }
          if (p.shared_map && valid) {
            const float* smap = p.shared_map + ((smap_img + out_y) * p.Wo + out_x) * p.Co + cc_now;
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
              const float4 m = *reinterpret_cast<const float4*>(smap + jj);
              fv[jj] += m.x; fv[jj + 1] += m.y; fv[jj + 2] += m.z; fv[jj + 3] += m.w;
            }
          }
          if (reg_stats) {
            if (cc_now == 0) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) { const float x = valid ? fv[jj] : 0.f; ra1[jj] += x; ra2[jj] += x * x; }
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) { const float x = valid ? fv[jj] : 0.f; rb1[jj] += x; rb2[jj] += x * x; }
            }
          }
          if (p.tma_out) {
            // stage this chunk in shared memory; when the phase is complete one bulk tensor store writes its 128 rows
            // (full 16-byte pieces from every lane instead of 32 scattered pieces per STG; the store queue was the limit)
            uint8_t* buf = out_stage + (gcount & 1u) * (uint32_t)p.out_buf_bytes;
            const uint32_t row_off = (uint32_t)r * (uint32_t)p.out_rb;
            if (p.out_fp32) {
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4) {
                uint32_t o = row_off + (uint32_t)(cc_now + jj) * 4u;
                o ^= ((o >> 7) & swz_mask) << 4;
                *reinterpret_cast<float4*>(buf + o) = make_float4(fv[jj], fv[jj + 1], fv[jj + 2], fv[jj + 3]);
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; jj += 8) {
                uint4 pk;
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) hh[k2] = __floats2bfloat162_rn(fv[jj + 2 * k2], fv[jj + 2 * k2 + 1]);
                uint32_t o = row_off + (uint32_t)(cc_now + jj) * 2u;
                o ^= ((o >> 7) & swz_mask) << 4;
                *reinterpret_cast<uint4*>(buf + o) = pk;
              }
            }
            if (cc_now + 16 >= p.BN) {             // last chunk of phase g_now
              fence_proxy_async();
              if (epi_tid == 0) bulk_wait_read0();   // the store issued one phase ago has read its buffer: free for the next phase
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (epi_tid == 0) {
                tma_store_5d(&map_o, buf, 0, (int)p.out_ox[g_now], tile_x * p.TW, (int)p.out_oy[g_now],
                             n_img * p.Hg + tile_y * p.TH);
                bulk_commit();
              }
              ++gcount;
            }
          } else if (valid) {
            if (p.act == 1) {                     // MPI head: 4 real channels -> packed fp32 MPI (+ sign of sigma)
              float4 o;
              o.x = __fdividef(1.f, 1.f + __expf(-fv[0]));
              o.y = __fdividef(1.f, 1.f + __expf(-fv[1]));
              o.z = __fdividef(1.f, 1.f + __expf(-fv[2]));
              o.w = p.head_alpha ? __fdividef(1.f, 1.f + __expf(-fv[3])) : fabsf(fv[3]) + 1e-4f;
              reinterpret_cast<float4*>(p.out)[out_pix] = o;
              if (p.raw_out) reinterpret_cast<int8_t*>(p.raw_out)[out_pix] = fv[3] >= 0.f ? (int8_t)1 : (int8_t)-1;
            } else if (p.out_fp32) {
              float* dst = reinterpret_cast<float*>(p.out) + out_pix * p.Co + cc_now;
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4) {
                float4 o = make_float4(fv[jj], fv[jj + 1], fv[jj + 2], fv[jj + 3]);
                if (p.accumulate) {               // second gradient contribution lands on top of the first (no add kernel)
                  const float4 e = *reinterpret_cast<const float4*>(dst + jj);
                  o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                }
                *reinterpret_cast<float4*>(dst + jj) = o;
              }
            } else {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + out_pix * p.Co + cc_now;
              if (p.accumulate) {
                const uint4 e0 = *reinterpret_cast<const uint4*>(dst), e1 = *reinterpret_cast<const uint4*>(dst + 8);
                const __nv_bfloat16* eb0 = reinterpret_cast<const __nv_bfloat16*>(&e0);
                const __nv_bfloat16* eb1 = reinterpret_cast<const __nv_bfloat16*>(&e1);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) { fv[jj] += __bfloat162float(eb0[jj]); fv[8 + jj] += __bfloat162float(eb1[jj]); }
              }
              uint4 o0, o1;
              __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
              __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                h0[jj] = __floats2bfloat162_rn(fv[2 * jj], fv[2 * jj + 1]);
                h1[jj] = __floats2bfloat162_rn(fv[8 + 2 * jj], fv[8 + 2 * jj + 1]);
              }
              *reinterpret_cast<uint4*>(dst) = o0;
              *reinterpret_cast<uint4*>(dst + 8) = o1;
            }
          }
        }
        __syncwarp();
      }
      // this warp has finished reading the accumulators of the tile: hand them back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accum_empty[as]);
    }
    if (p.tma_out && epi_tid == 0) bulk_wait_all();        // every store of this CTA complete before it exits
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
      asm volatile("bar.sync 1, 128;" ::: "memory");         // epilogue warps only
      const int e = threadIdx.x - 64;                        // 0..127
      if (e < p.Co) {
        atomicAdd(p.stats + e, s_stats[0][e]);
        atomicAdd(p.stats + p.Co + e, s_stats[1][e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool try_launch_conv_halo(const ConvLaunch& L, cudaStream_t stream, const char** err) {
  *err = nullptr;
  static const bool enabled = !(getenv("MINE_B200_HALO") && getenv("MINE_B200_HALO")[0] == '0');
  if (!enabled) return false;
  ConvParams p = L.p;
  if (p.CB < 1) p.CB = 1;
  if (p.in_stride != 1 || p.CB != 1) return false;
  if (p.accumulate && (p.act != 0 || p.stats)) return false;
  if (p.es != 2 && p.es != 4) return false;
  if (p.BN % 16 || p.BN < 16 || p.BN > 64) return false;
  if (p.stats && p.BN > 32) return false;                              // BatchNorm sums live in registers: <= 32 channels
  if (p.KB * p.es != 32 && p.KB * p.es != 64 && p.KB * p.es != 128) return false;
  if (p.Ci % p.KB) return false;
  p.kblocks = p.Ci / p.KB;
  if (p.kblocks > 4) return false;
  if (2 * p.G * p.BN > 512) return false;                              // double-buffered accumulators of all groups
  // all tap offsets inside one 3x3 window
  int y0 = 1 << 20, x0 = 1 << 20, y1 = -(1 << 20), x1 = -(1 << 20);
  for (int g = 0; g < p.G; ++g)
    for (int t = 0; t < p.T; ++t) {
      y0 = min(y0, (int)p.tap_y[g][t]); y1 = max(y1, (int)p.tap_y[g][t]);
      x0 = min(x0, (int)p.tap_x[g][t]); x1 = max(x1, (int)p.tap_x[g][t]);
    }
  if (y1 - y0 > 2 || x1 - x0 > 2) return false;
  p.halo_y0 = y0; p.halo_x0 = x0;
  for (int g = 0; g < p.G; ++g)
    for (int t = 0; t < p.T; ++t) { p.rel_y[g][t] = (int8_t)(p.tap_y[g][t] - y0); p.rel_x[g][t] = (int8_t)(p.tap_x[g][t] - x0); }
  // tile: TW a multiple of 8 so that a vertical tap shift is a whole number of swizzle atoms
  const int row_bytes = p.KB * p.es;
  const double u8 = (double)p.Hg * p.Wg / ((double)((p.Hg + 15) / 16 * 16) * ((p.Wg + 7) / 8 * 8));
  const double u16 = (double)p.Hg * p.Wg / ((double)((p.Hg + 7) / 8 * 8) * ((p.Wg + 15) / 16 * 16));
  if (u8 >= u16) { p.TW = 8; p.TH = 16; } else { p.TW = 16; p.TH = 8; }
  p.tiles_x = (p.Wg + p.TW - 1) / p.TW;
  p.tiles_y = (p.Hg + p.TH - 1) / p.TH;
  p.fd_tiles = make_fastdiv(p.tiles_x * p.tiles_y);
  p.fd_tiles_x = make_fastdiv(p.tiles_x);
  p.fd_planes = make_fastdiv(p.planes_per_image > 0 ? p.planes_per_image : 1);
  const uint32_t box_bytes = (uint32_t)(p.TH + 2) * p.TW * row_bytes;
  p.box_stride = (int)((box_bytes + 1023u) / 1024u * 1024u);
  p.w_tile_bytes = (int)(((uint32_t)p.BN * row_bytes + 1023u) / 1024u * 1024u);
  p.w_bytes = p.G * p.T * p.kblocks * p.w_tile_bytes;
  if (p.w_bytes > 112 * 1024) return false;
  const uint32_t set_bytes = 3u * p.box_stride;
  p.tmem_cols = next_pow2_cols(2 * p.G * p.BN);
  // prefer two resident CTAs per SM (the epilogue of one overlaps the loads / MMAs of the other) when >= 2 ring stages
  // still fit in half of the shared memory; otherwise one CTA with a deeper ring
  // TMA-store epilogue: sub-pixel upsample form writing an NHWC tensor whose pixel row is one swizzle span
  static const bool tma_out_on = !(getenv("MINE_B200_TMA_STORE") && getenv("MINE_B200_TMA_STORE")[0] == '0');
  const int es_out = p.out_fp32 ? 4 : 2;
  const int out_rb = p.Co * es_out;
  p.tma_out = 0; p.out_rb = out_rb; p.out_buf_bytes = 0; p.out_stage_off = 0;
  if (tma_out_on && p.G == 4 && p.act == 0 && !p.accumulate && p.out_sy == 2 && p.out_sx == 2 && p.Co == p.BN &&
      (out_rb == 64 || out_rb == 128) && p.Hg % p.TH == 0 && p.Ho == 2 * p.Hg && p.Wo == 2 * p.Wg) {
    // (32-byte rows - 16 bf16 channels - measured slower through the bulk store: 0.109 vs 0.100 ms at level 0)
    bool std_phases = true;
    for (int g = 0; g < 4; ++g) std_phases = std_phases && p.out_oy[g] == (g >> 1) && p.out_ox[g] == (g & 1);
    if (std_phases) { p.tma_out = 1; p.out_buf_bytes = (int)((128u * out_rb + 1023u) / 1024u * 1024u); }
  }
  // room for the ring next to the resident weights (and the staging buffers): two CTAs per SM when >= 2 stages fit in half
  // of the shared memory, else one CTA with a deeper ring.  The staging buffers are dropped (direct stores from registers)
  // when they would cost the second CTA or the halo form itself.
  auto plan = [&](uint32_t extra, int& ctas, int& st) {
    ctas = 1;
    const int64_t half = (int64_t)107 * 1024 - 1024 - p.w_bytes - (int64_t)extra;
    const int64_t full = (int64_t)200 * 1024 - 2048 - p.w_bytes - (int64_t)extra;
    st = half > 0 ? (int)(half / set_bytes) : 0;
    if (st >= 2 && 2 * p.tmem_cols <= 512) ctas = 2;
    else st = full > 0 ? (int)(full / set_bytes) : 0;
  };
  uint32_t out_bytes = 2u * (uint32_t)p.out_buf_bytes;
  int ctas_per_sm = 1, stages = 0;
  plan(out_bytes, ctas_per_sm, stages);
  if (p.tma_out) {
    int c0 = 1, s0 = 0;
    plan(0u, c0, s0);
    if (stages < 2 || ctas_per_sm < c0) {
      p.tma_out = 0; p.out_buf_bytes = 0; out_bytes = 0;
      ctas_per_sm = c0; stages = s0;
    }
  }
  if (stages > 4) stages = 4;
  if (stages < 2) return false;
  p.stages = stages;
  p.out_stage_off = p.w_bytes + stages * (int)set_bytes;          // multiples of 1024: swizzle-aligned staging buffers
  size_t smem = (size_t)p.w_bytes + (size_t)stages * set_bytes + out_bytes + 1024;
  const size_t smem_floor = (220u * 1024u) / (ctas_per_sm + 1) + 1024;   // one more CTA must NOT fit
  if (smem < smem_floor) smem = smem_floor;
  if (smem > 200u * 1024u) smem = 200u * 1024u;
  CUtensorMap mx, mw, mo;
  const char* e = nhwc_map(&mx, L.x, p.Ci, L.Wi, L.Hi, p.N, p.KB, p.TW, p.TH + 2, 1, 1, p.es);
  if (e) { *err = e; return true; }
  e = weight_map(&mw, L.w, p.Ci, L.w_rows, p.G * p.T, p.KB, p.BN, p.es);
  if (e) { *err = e; return true; }
  if (p.tma_out) {
    e = phase_out_map(&mo, p.out, p.Co, p.Wg, p.N * p.Hg, p.TW, p.TH, es_out);
    if (e) { *err = e; return true; }
  } else {
    mo = mw;                                               // unused by the kernel
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int total_work = p.tiles_x * p.tiles_y * p.N;
  int grid_x = sm_count() * ctas_per_sm;
  if (grid_x > total_work) grid_x = total_work;
  conv_halo_kernel<<<grid_x, kConvThreads, smem, stream>>>(mx, mw, mo, p);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) *err = cudaGetErrorString(ce);
  return true;
}

}  // namespace mine
