// Tap-sharing weight gradient for the narrow, high-resolution layers (the weight-gradient counterpart of conv_halo.cu).
//
// wgrad_taps_kernel (conv_tcgen05.cu) loads one activation tile PER TAP: nine (3x3) or sixteen (sub-pixel form) boxes of
// the same pixels per 128-pixel tile, which makes the TMA row traffic - not HBM, not the tensor pipe - the bound for the
// 16..64-channel layers.  Here the B operand (activations, MN-major: one K row per pixel or pixel pair) is loaded once
// per HORIZONTAL tap offset as a box of TH + 2 image lines; a vertical tap offset is a shift by whole lines inside that
// box (a whole number of swizzle atoms), so the taps (ty = 0..2, tx) are consecutive "MN blocks" one line apart and ONE
// tcgen05.mma with N = 3 x block covers them.  All sub-pixel groups of the upsample form share the activation boxes.
// The A operand (output gradient) is loaded once per group.  Everything else - operand layouts (bf16 swizzles; TF32:
// 128-byte rows, SWIZZLE_128B_BASE32B, pixel-pair / phase-pair rows for 16-channel tensors), the single-thread issue
// loop, fp32 atomics into the packed gradient - is as in wgrad_taps_kernel.
//
// Operand roles are SWAPPED with respect to wgrad_taps_kernel: the activation tap-blocks form the M dimension (3 taps x
// 32 channels = 96 of the 128 accumulator rows are useful) and the output gradient the N dimension (16..64 columns).
// tcgen05.mma costs max(M,128) * N / 256 cycles, so the fixed-cost M side should be the wide one: with the gradient on
// M (Co = 16..32 useful rows of 128) and 96 tap columns on N the narrow layers were tensor-pipe bound at ~50 % active
// (profiles/ncu_r2_wgrad_*), 3x the math of this arrangement.
//
// The work of a layer is described by small tables built on the host (per chunk of <= 512 accumulator columns):
//   loads[] : activation boxes of a pixel tile {horizontal offset, channel block}
//   ops[]   : MMAs per K step {A slab, load, first line, number of line-blocks, accumulator column, diag half,
//             destination (group, tap) of every block}
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"
#include "conv_engine.h"

namespace mine {

struct WHLoad { int16_t x_off, cb; };
struct WHOp {
  uint8_t a_idx, load, ty0, nblk;
  uint16_t col;
  int8_t half;            // diag mode 1: B parity of this op (only accumulator rows of the same half are kept); else -1
  uint8_t dst[3];         // destination tap index (g * T + t) of every line-block (phase-pair: of half 0)
};
struct WHChunk {
  int nl, nops, cols;
  WHLoad loads[8];
  WHOp ops[12];
};
struct WHParams {
  int N, TH, TW, tiles_x, tiles_y;
  int rpl, rows, lines;                 // K rows per image line, per tile, lines per activation box (TH + 2)
  int Co, Ci, T, es;
  int NA;                               // A slabs loaded per tile: groups x channel slabs
  int a_groups, a_slabs;                // NA = a_groups * a_slabs
  int a_cb, b_cb, bw;                   // channels per A / B slab row, accumulator columns per B block
  int a_row, b_row;                     // bytes per K row
  int a_kind, b_kind;                   // 0: NHWC map {c, x, y, n}; 1: overlapping 2-pixel rows {0, x, y, n}
  int dy_stride, dy_xmul;               // dy pixel of GEMM pixel ox: ox * dy_xmul + offset (dy_xmul = dy_stride)
  int16_t a_oy[4], a_ox[4];             // per A group
  int y0, x0;                           // window origin of the taps in x
  int diag;                             // 0 none, 1 parity-slab ops, 2 halves inside a block (16-channel pair rows)
  int half_dst_stride;                  // phase-pair: destination tap index of half h = dst + h * stride; else 0
  int ncols;                            // accumulator columns per op = gradient channels (diag modes: 32 = 2 halves x 16)
  int NB;                               // input channels handled per CTA (<= 128), ci_blocks over blockIdx.z
  int stages, tmem_cols, nchunks, ci_blocks;
  float* dw;
  WHChunk chunk[4];
};

// MMA issue loop of ONE issuer (an elected lane; the loop itself is uniform code, so the per-op constants are read
// from the parameter block on the uniform datapath and descriptors never leave uniform registers).  The ops of a tile are
// dealt round-robin to `n_issuers` warps (the MMA warp and the otherwise idle epilogue warps): a single issue stream is
// latency bound, several of them keep the tensor pipe fed.  Every issuer commits its own MMAs, so the slot / accumulator
// barriers expect `n_issuers` arrivals.  Per op the K steps run back to back (same accumulator).
template <bool TF32>
__device__ __forceinline__ void wh_issue_loop(const WHParams& p, const WHChunk& ck, int issuer, int n_issuers,
                                              uint32_t ring_base, uint32_t stage_bytes, uint32_t a_lo0, uint32_t a_hi,
                                              uint32_t b_lo0, uint32_t b_hi, uint32_t a_step, uint32_t b_step, int ksteps,
                                              int my_tiles, uint32_t tmem_base, uint32_t a_slab16, uint32_t a_bytes,
                                              uint32_t b_region, uint32_t line_bytes, uint64_t* full_bar,
                                              uint64_t* empty_bar, uint64_t* accum_bar) {
  const uint32_t idesc = make_idesc(128, p.ncols, 1, 1, TF32);
  const int nops = ck.nops;
  int s = 0;
  uint32_t par = 0;
  for (int i = 0; i < my_tiles; ++i) {
    mbar_wait(&full_bar[s], par);
    tc_fence_after();
    const uint32_t base16 = (ring_base + (uint32_t)s * stage_bytes) >> 4;
    const uint32_t acc0 = i > 0 ? 1u : 0u;
    for (int o = issuer; o < nops; o += n_issuers) {
      const WHOp& op = ck.ops[o];
      uint32_t ka = a_lo0 + base16 + (uint32_t)(op.a_idx * p.a_slabs) * a_slab16;
      uint32_t kb = b_lo0 + base16 + ((a_bytes + (uint32_t)op.load * b_region + (uint32_t)op.ty0 * line_bytes) >> 4);
      const uint32_t d_tmem = tmem_base + op.col;
      uint32_t acc = acc0;
      for (int k = 0; k < ksteps; ++k, ka += a_step, kb += b_step) {
        umma_lohi2<TF32>(d_tmem, kb, b_hi, ka, a_hi, idesc, acc);            // A = activation blocks, B = gradient
        acc = 1u;
      }
    }
    umma_commit(&empty_bar[s]);
    if (++s == p.stages) { s = 0; par ^= 1u; }
  }
  umma_commit(accum_bar);
}

__global__ void __launch_bounds__(kConvThreads, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ WHParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ WHChunk ck;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&p.chunk[blockIdx.y]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&ck);
    for (int i = threadIdx.x; i < (int)(sizeof(WHChunk) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const int nb = blockIdx.z;                              // input-channel block (Ci > 128)
  const bool tf32 = p.es == 4;
  const uint32_t a_slab = (uint32_t)p.rows * p.a_row;
  const uint32_t a_bytes = a_slab * p.NA;
  const uint32_t b_region = (uint32_t)p.lines * p.rpl * p.b_row;          // one activation box
  const int n_ops_all = p.chunk[blockIdx.y].nops;
  const int n_issuers = n_ops_all < 4 ? n_ops_all : 4;      // lane 0 of warps 1..4
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], n_issuers); }
    mbar_init(&accum_bar, n_issuers);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  __syncthreads();
  // issue-loop constants (every issuer thread needs them)
  const uint32_t lta = tf32 ? 1u : layout_type_for(p.a_row), ltb = tf32 ? 1u : layout_type_for(p.b_row);
  const int kpi = 32 / p.es;                             // K rows per instruction: 16 (bf16) or 8 (tf32)
  const uint32_t sbo_a = tf32 ? 512u : 8u * p.a_row, sbo_b = tf32 ? 512u : 8u * p.b_row;
  const uint32_t a_lbo = a_slab;                          // gradient = N operand: column blocks one channel slab apart
  const uint64_t da0 = make_smem_desc(0, a_lbo, sbo_a, lta);
  const uint64_t db0 = make_smem_desc(0, (uint32_t)p.rpl * p.b_row, sbo_b, ltb);   // MN-block stride of B = one image line
  const uint32_t a_lo0 = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32), b_lo0 = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
  const uint32_t a_step = (uint32_t)(kpi * p.a_row) >> 4, b_step = (uint32_t)(kpi * p.b_row) >> 4;
  const int ksteps = p.rows / kpi;
#define WH_ISSUE(ID)                                                                                                        \
  do {                                                                                                                      \
    const WHChunk& pck = p.chunk[blockIdx.y];          /* parameter space: uniform loads */                                \
    if (tf32) wh_issue_loop<true>(p, pck, ID, n_issuers, smem_u32(smem_aligned), stage_bytes, a_lo0, a_hi, b_lo0, b_hi,     \
                                  a_step, b_step, ksteps, my_tiles, tmem_base, a_slab >> 4, a_bytes, b_region,              \
                                  (uint32_t)p.rpl * p.b_row, full_bar, empty_bar, &accum_bar);                              \
    else wh_issue_loop<false>(p, pck, ID, n_issuers, smem_u32(smem_aligned), stage_bytes, a_lo0, a_hi, b_lo0, b_hi,         \
                              a_step, b_step, ksteps, my_tiles, tmem_base, a_slab >> 4, a_bytes, b_region,                  \
                              (uint32_t)p.rpl * p.b_row, full_bar, empty_bar, &accum_bar);                                  \
  } while (0)
  const uint32_t stage_bytes = ((a_bytes + b_region * ck.nl + 1023u) / 1024u) * 1024u;
  uint8_t* smem_aligned = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_img * p.N;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (elect_one()) {                                     // TMA producer: one elected lane, uniform code
      const WHChunk& pck = p.chunk[blockIdx.y];
      int s = 0;
      uint32_t par = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tid = blockIdx.x + i * gridDim.x;
        const int n_img = tid / tiles_per_img, tt = tid - n_img * tiles_per_img;
        const int tile_y = tt / p.tiles_x, tile_x = tt - tile_y * p.tiles_x;
        const int oy0 = tile_y * p.TH, ox0 = tile_x * p.TW;
        if (i >= p.stages) mbar_wait(&empty_bar[s], par ^ 1u);
        uint8_t* a_dst = smem_aligned + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + b_region * pck.nl);
        for (int g = 0; g < p.a_groups; ++g) {
          const int ay = oy0 * p.dy_stride + p.a_oy[g], ax = ox0 * p.dy_xmul + p.a_ox[g];
          for (int sl = 0; sl < p.a_slabs; ++sl)
            tma_load_4d(&map_dy, &full_bar[s], a_dst + (size_t)(g * p.a_slabs + sl) * a_slab, p.a_kind ? 0 : sl * p.a_cb, ax, ay, n_img);
        }
        uint8_t* b_dst = a_dst + a_bytes;
        for (int l = 0; l < pck.nl; ++l, b_dst += b_region)
          tma_load_4d(&map_x, &full_bar[s], b_dst, p.b_kind ? 0 : (nb * (p.NB / p.b_cb) + pck.loads[l].cb) * p.b_cb,
                      ox0 + p.x0 + pck.loads[l].x_off, oy0 + p.y0, n_img);
        if (++s == p.stages) { s = 0; par ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) WH_ISSUE(0);
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;                           // TMEM lane == accumulator row
    if (warp - 1 < n_issuers) {                            // warps 2..4: extra issuers during the main loop
      if (elect_one()) WH_ISSUE(warp - 1);
    }
    __syncwarp();
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (my_tiles > 0) {
      // accumulator row m = (line-block j, channel c' of the block), columns = gradient channels (diag modes: 2 halves x
      // 16).  Only the diagonal (activation half == gradient half) is kept in the diag modes.
      const int j = m / p.bw, cp = m - j * p.bw;
      for (int o = 0; o < ck.nops; ++o) {
        const WHOp op = ck.ops[o];
        int ci = nb * p.NB + (p.b_kind ? 0 : ck.loads[op.load].cb * p.b_cb) + cp, half = 0;
        bool keep = j < op.nblk;
        if (p.diag == 1) half = op.half;                                  // parity-slab op: activation half of the whole op
        else if (p.diag == 2) { half = cp >> 4; ci = nb * p.NB + (cp & 15); }   // half inside the 2-pixel row
        keep = keep && ci < p.Ci;
        const int gt = keep ? op.dst[j] + half * p.half_dst_stride : 0;
        // tcgen05.ld takes ONE column address per warp: in the diag modes both gradient halves are visited with uniform
        // addresses and every lane keeps the one that matches its activation half
        const int nhalf = p.diag == 2 ? 2 : 1, ncol = p.diag ? 16 : p.Co;
        for (int hh = 0; hh < nhalf; ++hh) {
          const int col0 = p.diag == 1 ? op.half * 16 : hh * 16 * (p.diag == 2);
          const bool mine_half = p.diag != 2 || half == hh;
          for (int c0 = 0; c0 < ncol; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(op.col + col0 + c0), v);
            if (!keep || !mine_half) continue;
            float* dst = p.dw + (((size_t)gt * p.Co + c0) * p.Ci + ci);
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c0 + e < p.Co) atomicAdd(dst + (size_t)e * p.Ci, __uint_as_float(v[e]));
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
// host side
// ------------------------------------------------------------------------------------------------
bool try_launch_wgrad_halo(const WgradLaunch& L, cudaStream_t stream, const char** err) {
  *err = nullptr;
  static const bool enabled = !(getenv("MINE_B200_WGRAD_HALO") && getenv("MINE_B200_WGRAD_HALO")[0] == '0');
  if (!enabled) return false;
  const WgradParams& w = L.p;
  const int es = w.es;
  if (es != 2 && es != 4) return false;
  if ((w.x_stride > 1) || w.Co > 128) return false;
  const bool tf32 = es == 4;
  const bool up = w.G == 4 && w.T == 4 && w.dy_stride == 2;      // sub-pixel (upsample) form
  const bool same = w.G == 1 && w.T == 9 && w.dy_stride == 1;    // 3x3, same resolution
  if (!up && !same) return false;
  // measured (profiles/kernel_bench_r2_*): with the gradient on the N side this kernel wins for narrow gradients; for
  // Co >= 128 (and the sub-pixel form at Co = 64) wgrad_taps_kernel's arrangement (gradient on M) is faster
  if (w.Co > 64 || (up && w.Co > 32)) return false;
  // tap window of x: all offsets inside 3x3, canonical tables
  int y0 = 1 << 20, x0 = 1 << 20, y1 = -(1 << 20), x1 = -(1 << 20);
  for (int g = 0; g < w.G; ++g)
    for (int t = 0; t < w.T; ++t) {
      y0 = min(y0, (int)w.tap_y[g][t]); y1 = max(y1, (int)w.tap_y[g][t]);
      x0 = min(x0, (int)w.tap_x[g][t]); x1 = max(x1, (int)w.tap_x[g][t]);
    }
  if (y1 - y0 != 2 || x1 - x0 != 2) return false;
  if (same) {
    for (int t = 0; t < 9; ++t) if (w.tap_y[0][t] - y0 != t / 3 || w.tap_x[0][t] - x0 != t % 3) return false;
    if (w.dy_oy[0] != 0 || w.dy_ox[0] != 0) return false;
  } else {
    for (int g = 0; g < 4; ++g) {
      const int py = g >> 1, px = g & 1;
      if (w.dy_oy[g] != py || w.dy_ox[g] != px) return false;
      for (int t = 0; t < 4; ++t) if (w.tap_y[g][t] - y0 != py + (t >> 1) || w.tap_x[g][t] - x0 != px + (t & 1)) return false;
    }
  }
  WHParams p;
  memset(&p, 0, sizeof(p));
  p.N = w.N; p.Co = w.Co; p.Ci = w.Ci; p.T = w.T; p.es = es; p.dw = w.dw;
  p.y0 = y0; p.x0 = x0; p.dy_stride = w.dy_stride; p.dy_xmul = w.dy_stride;
  p.NB = w.Ci < 128 ? w.Ci : 128;
  if (w.Ci % p.NB) return false;
  p.ci_blocks = w.Ci / p.NB;
  int pair = 1;                    // GEMM pixels per K row
  bool phase_pair = false;
  if (!tf32) {
    p.a_cb = w.Co < 64 ? w.Co : 64; p.b_cb = p.NB < 64 ? p.NB : 64;
    if ((p.a_cb != 16 && p.a_cb != 32 && p.a_cb != 64) || (p.b_cb != 16 && p.b_cb != 32 && p.b_cb != 64)) return false;
    if (w.Co % p.a_cb || p.NB % p.b_cb) return false;
    p.a_row = p.a_cb * 2; p.b_row = p.b_cb * 2; p.bw = p.b_cb;
    p.a_slabs = w.Co >= 128 ? 2 : 1;
  } else if (w.Co == 16 || w.Ci == 16) {
    if (w.Co != 16 || (w.Ci != 16 && w.Ci % 32)) return false;
    p.a_kind = 1; p.a_cb = 16; p.a_slabs = 1; p.a_row = 128; p.b_row = 128; p.bw = 32;
    if (up) {
      if (w.Ci != 16) return false;                 // (no such layer: the 16 -> 16 upsample conv is the only one)
      phase_pair = true; p.b_kind = 1; p.b_cb = 16; p.diag = 2; p.half_dst_stride = w.T;
    } else {
      pair = 2;
      if (w.Ci == 16) { p.b_kind = 1; p.b_cb = 16; p.diag = 2; }
      else { p.b_cb = 32; p.diag = 1; }
    }
  } else {
    if (w.Co % 32 || w.Ci % 32) return false;
    p.a_cb = p.b_cb = 32; p.a_row = p.b_row = 128; p.bw = 32;
    p.a_slabs = (w.Co < 128 ? w.Co : 128) / 32;
  }
  // ---- tile: rows-per-line must be a whole number of swizzle atoms (tf32: 4 K rows, bf16: 8), stage <= ~96 KB ----
  const int atom = tf32 ? 4 : 8, kpi = 32 / es;
  const int cbn = p.b_kind ? 1 : p.NB / p.b_cb;          // channel slabs of the B operand
  const int nx = (p.diag == 1) ? 4 : (phase_pair ? 2 : 3);
  p.a_groups = up ? (phase_pair ? 2 : 4) : 1;
  p.NA = p.a_groups * p.a_slabs;
  int best_th = 0, best_tw = 0; double best_cost = 1e30;
  for (int tw = 8; tw <= 64; tw *= 2) {
    const int rpl = tw / pair;
    if (rpl % atom) continue;
    for (int th = 2; th <= 32; th *= 2) {
      const int rows = th * rpl;
      if (rows % kpi || rows < 32 || rows > 256) continue;
      if (tw * w.dy_stride * (p.a_kind && !phase_pair ? 1 : 1) > 256 || th * w.dy_stride > 256) continue;
      const size_t stage = (size_t)rows * p.a_row * p.NA + (size_t)(th + 2) * rpl * p.b_row * nx * cbn;
      if (stage > 100 * 1024) continue;
      const int tx = (w.Wg + tw - 1) / tw, ty = (w.Hg + th - 1) / th;
      // TMA rows per useful pixel (what bounds these layers) x tile quantisation
      const double rows_per_tile = (double)rows * p.NA + (double)(th + 2) * rpl * nx * cbn;
      const double cost = rows_per_tile * tx * ty / ((double)w.Hg * w.Wg);
      if (cost < best_cost) { best_cost = cost; best_th = th; best_tw = tw; }
    }
  }
  if (!best_th) return false;
  p.TH = best_th; p.TW = best_tw;
  p.rpl = p.TW / pair; p.rows = p.TH * p.rpl; p.lines = p.TH + 2;
  p.tiles_x = (w.Wg + p.TW - 1) / p.TW; p.tiles_y = (w.Hg + p.TH - 1) / p.TH;
  for (int g = 0; g < p.a_groups; ++g) {
    p.a_oy[g] = phase_pair ? g : (up ? (g >> 1) : 0);
    p.a_ox[g] = phase_pair ? 0 : (up ? (g & 1) : 0);
  }
  p.ncols = p.diag ? 32 : w.Co;
  if (p.ncols % 16 || p.ncols > 256) return false;
  const int max_blk = 128 / p.bw;                 // line-blocks per MMA (M = 128 rows)
  if (max_blk < 1) return false;
  // ---- ops, then chunks of <= 512 accumulator columns ----
  struct RawOp { int a_idx, xoff, cb, ty0, nblk, half, dst[3]; };
  RawOp raw[48]; int nraw = 0;
  if (same) {
    for (int tx = 0; tx < 3; ++tx)
      for (int hq = 0; hq < (p.diag == 1 ? 2 : 1); ++hq)
        for (int cb = 0; cb < cbn; ++cb) {
          RawOp& o = raw[nraw++];
          o.a_idx = 0; o.xoff = tx + hq; o.cb = cb; o.ty0 = 0; o.nblk = 3; o.half = p.diag == 1 ? hq : -1;
          for (int ty = 0; ty < 3; ++ty) o.dst[ty] = ty * 3 + tx;
        }
  } else if (phase_pair) {
    for (int py = 0; py < 2; ++py)
      for (int b = 0; b < 2; ++b) {
        RawOp& o = raw[nraw++];
        o.a_idx = py; o.xoff = b; o.cb = 0; o.ty0 = py; o.nblk = 2; o.half = -1;
        for (int a = 0; a < 2; ++a) o.dst[a] = (2 * py) * 4 + a * 2 + b;      // half h (= px) adds h * T
        o.dst[2] = 0;
      }
  } else {
    for (int g = 0; g < 4; ++g)
      for (int b = 0; b < 2; ++b)
        for (int cb = 0; cb < cbn; ++cb) {
          RawOp& o = raw[nraw++];
          const int py = g >> 1, px = g & 1;
          o.a_idx = g; o.xoff = px + b; o.cb = cb; o.ty0 = py; o.nblk = 2; o.half = -1;
          for (int a = 0; a < 2; ++a) o.dst[a] = g * 4 + a * 2 + b;
          o.dst[2] = 0;
        }
  }
  // an MMA covers at most max_blk line-blocks (M = 128 rows): split longer ops
  for (int i = 0; i < nraw; ++i) {
    if (raw[i].nblk <= max_blk) continue;
    if (nraw == 48) return false;
    RawOp tail = raw[i];
    tail.ty0 = raw[i].ty0 + max_blk; tail.nblk = raw[i].nblk - max_blk;
    for (int j2 = 0; j2 < 3; ++j2) tail.dst[j2] = j2 + max_blk < 3 ? raw[i].dst[j2 + max_blk] : 0;
    raw[i].nblk = max_blk;
    for (int k2 = nraw; k2 > i + 1; --k2) raw[k2] = raw[k2 - 1];
    raw[i + 1] = tail;
    ++nraw;
  }
  p.nchunks = 0;
  {
    // balanced chunks: as few as the 512 accumulator columns allow, the ops spread evenly over them
    int total_cols = nraw * p.ncols;
    int want = (total_cols + 511) / 512;
    if ((nraw + want - 1) / want > 12) want = (nraw + 11) / 12;
    const int per_chunk = (nraw + want - 1) / want;
    int i = 0;
    while (i < nraw) {
      if (p.nchunks == 4) return false;
      WHChunk& c = p.chunk[p.nchunks++];
      c.nl = 0; c.nops = 0; c.cols = 0;
      while (i < nraw && c.nops < per_chunk) {
        const RawOp& r = raw[i];
        const int cols = p.ncols;
        if (c.cols + cols > 512) break;
        int li = -1;
        for (int l = 0; l < c.nl; ++l) if (c.loads[l].x_off == r.xoff && c.loads[l].cb == r.cb) li = l;
        if (li < 0) { if (c.nl == 8) break; li = c.nl++; c.loads[li].x_off = (int16_t)r.xoff; c.loads[li].cb = (int16_t)r.cb; }
        WHOp& o = c.ops[c.nops++];
        o.a_idx = (uint8_t)r.a_idx; o.load = (uint8_t)li; o.ty0 = (uint8_t)r.ty0; o.nblk = (uint8_t)r.nblk;
        o.col = (uint16_t)c.cols; o.half = (int8_t)r.half;
        for (int j = 0; j < 3; ++j) o.dst[j] = (uint8_t)r.dst[j];
        c.cols += cols;
        ++i;
      }
      if (c.nops == 0) return false;
    }
  }
  int max_cols = 0, max_nl = 0;
  for (int c = 0; c < p.nchunks; ++c) { max_cols = max(max_cols, p.chunk[c].cols); max_nl = max(max_nl, p.chunk[c].nl); }
  p.tmem_cols = next_pow2_cols(max_cols);
  const size_t stage_bytes = (((size_t)p.rows * p.a_row * p.NA + (size_t)p.lines * p.rpl * p.b_row * max_nl + 1023) / 1024) * 1024;
  // two resident CTAs per SM (two sets of issuers / epilogue warps) when the accumulators and >= 2 stages allow it
  const size_t slack = 8 * (size_t)p.rpl * p.b_row;
  int ctas = (2 * p.tmem_cols <= 512 && 2 * stage_bytes + 1024 + slack <= 107 * 1024) ? 2 : 1;
  int stages = (int)(((ctas == 2 ? 107u : 198u) * 1024u - 1024u - slack) / stage_bytes);
  if (stages > 4) stages = 4;
  if (stages < 2) return false;
  p.stages = stages;
  // + slack: the activation operand always spans 128 accumulator rows = 128 / bw line-blocks, the unused ones read
  // (and discard) whatever follows the box
  size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (size_t)p.rpl * p.b_row;
  const size_t smem_floor = (220u * 1024u) / (ctas + 1) + 1024;          // one more CTA must NOT fit
  if (smem < smem_floor) smem = smem_floor;
  // ---- tensor maps ----
  CUtensorMap mdy, mx;
  const char* e;
  if (phase_pair) e = overlap32_map(&mdy, L.dy, L.dyW, L.dyH, p.N, p.TW, p.TH, 2, 2);            // rows start at dy pixel 2 * ox
  else if (p.a_kind) e = overlap32_map(&mdy, L.dy, L.dyW, L.dyH, p.N, p.TW / 2, p.TH, 2, 1);      // pixel pairs
  else e = nhwc_map(&mdy, L.dy, p.Co, L.dyW, L.dyH, p.N, p.a_cb, p.TW, p.TH, w.dy_stride, w.dy_stride, es, tf32);
  if (e) { *err = e; return true; }
  if (phase_pair) { p.dy_xmul = 2; }
  if (p.b_kind) e = overlap32_map(&mx, L.x, L.xW, L.xH, p.N, p.rpl, p.lines, pair == 2 ? 2 : 1, 1);
  else e = nhwc_map(&mx, L.x, p.Ci, L.xW, L.xH, p.N, p.b_cb, p.rpl, p.lines, pair == 2 ? 2 : 1, 1, es, tf32);
  if (e) { *err = e; return true; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int total_tiles = p.tiles_x * p.tiles_y * p.N;
  const int combos = p.nchunks * p.ci_blocks;
  int splits = (sm_count() * ctas + combos - 1) / combos;
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  dim3 grid(splits, p.nchunks, p.ci_blocks);
  wgrad_halo_kernel<<<grid, kConvThreads, smem, stream>>>(mdy, mx, p);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) *err = cudaGetErrorString(ce);
  return true;
}

}  // namespace mine
