// Split-K variant of the tcgen05 implicit-GEMM convolution for layers with fewer output tiles than SMs.
//
// The deep encoder layers (8x12 .. 32x48 maps, reductions of up to 9 x 2048 channels) give conv_taps_kernel only
// 2..96 work items: a handful of CTAs walk 32..288 TMA->MMA iterations serially while most of the 148 SMs idle
// (profiles/README.md: 30 us per launch).  Here the (tap, channel-block) iteration space of every output tile is
// divided among `ksplit` CTAs (blockIdx.y); each accumulates its share in TMEM exactly like conv_taps_kernel (same
// TMA boxes, same K-major descriptors, one elected MMA thread, mbarrier ring) and adds its partial tile into a
// zero-initialised fp32 output with red.global.add.v4.f32.  splitk_finalize_kernel then converts to bf16 and
// produces the BatchNorm sums (per-channel sum / sum of squares) in one pass over the small tensor.
//
// Restrictions (the callers' layers satisfy them): one tap group (G = 1), Co a multiple of 16, dense output.
// Opt-in (MINE_B200_SPLITK=1) until it has been run on hardware; semantics = conv_taps with fp32 accumulation
// (mine_b200/ops/emu.py::conv_taps_splitk).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_engine.h"
#include "kernels.h"

namespace mine {

namespace sk {

// one elected lane / provably uniform warp index: see conv_common.cuh (role code on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "SK_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SK_WAIT_DONE;\n\t"
      "bra SK_WAIT_LOOP;\n\t"
      "SK_WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// same bit layouts as conv_tcgen05.cu (K-major operands)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
__device__ __forceinline__ uint32_t layout_type_for(int swizzle_bytes) {
  return swizzle_bytes == 128 ? 2u : (swizzle_bytes == 64 ? 4u : 6u);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, bool tf32) {
  const uint32_t fmt = tf32 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kThreads = 192;
constexpr int kStages = 4;

struct Params {
  int N, Hg, Wg, TH, TW, tiles_x, tiles_y;
  int T, Ci, KB, kblocks, in_stride;
  int16_t tap_y[16], tap_x[16];
  int Co, BN, CB, Ho, Wo;
  float* out;                 // fp32 [N, Ho, Wo, Co], zero-initialised by the caller
  int ksplit, tmem_cols, es;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_splitk_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t accum_full;
  __shared__ uint32_t tmem_base_smem;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  const int w = blockIdx.x;                                 // (tile, image, channel block), tile fastest
  const int tile = w % tiles, wi = w / tiles;
  const int n_img = wi % p.N, cb = wi / p.N;
  const int tile_y = tile / p.tiles_x, tile_x = tile - tile_y * p.tiles_x;
  // this CTA's share of the flattened (tap, channel block) iteration space
  const int iters = p.T * p.kblocks;
  const int i0 = (int)(((long long)blockIdx.y * iters) / p.ksplit);
  const int i1 = (int)(((long long)(blockIdx.y + 1) * iters) / p.ksplit);
  const int my_iters = i1 - i0;

  const int row_bytes = p.KB * p.es;
  const uint32_t a_bytes = 128u * row_bytes, b_bytes = (uint32_t)p.BN * row_bytes;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) / 1024u) * 1024u;
  uint8_t* smem_aligned = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {                                       // one elected lane: uniform code (conv_common.cuh)
      const int oy0 = tile_y * p.TH * p.in_stride, ox0 = tile_x * p.TW * p.in_stride;
      int t = i0 / p.kblocks, kb = i0 - t * p.kblocks;
      for (int i = 0; i < my_iters; ++i) {
        const int s = i % kStages, round = i / kStages;
        if (i >= kStages) mbar_wait(&empty_bar[s], (round - 1) & 1);
        uint8_t* slot = smem_aligned + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
        tma_load_4d(&map_x, &full_bar[s], slot, kb * p.KB, ox0 + p.tap_x[t], oy0 + p.tap_y[t], n_img);
        tma_load_3d(&map_w, &full_bar[s], slot + a_bytes, kb * p.KB, cb * p.BN, t);
        if (++kb == p.kblocks) { kb = 0; ++t; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const bool tf32 = p.es == 4;
      const uint32_t idesc = make_idesc(128, p.BN, tf32);
      const uint64_t desc0 = make_smem_desc(0, 16, 8u * row_bytes, layout_type_for(row_bytes));
      const int ksteps = row_bytes / 32;
      uint32_t first = 0;
      for (int i = 0; i < my_iters; ++i) {
        const int s = i % kStages, round = i / kStages;
        mbar_wait(&full_bar[s], round & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_aligned + (size_t)s * stage_bytes);
        const uint64_t da = desc0 + (uint64_t)(a_addr >> 4), db = desc0 + (uint64_t)((a_addr + a_bytes) >> 4);
        for (int k = 0; k < ksteps; ++k) {
          if (tf32) umma_tf32(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, first);
          else umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, first);
          first = 1u;
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_full);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ty = r / p.TW, tx = r - ty * p.TW;
    const int oy = tile_y * p.TH + ty, ox = tile_x * p.TW + tx;
    const bool valid = (oy < p.Hg) && (ox < p.Wg) && my_iters > 0;
    mbar_wait(&accum_full, 0);
    tc_fence_after();
    const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16);
    float* dst_row = p.out + (((size_t)n_img * p.Ho + oy) * p.Wo + ox) * p.Co + cb * p.BN;
    const int co_left = p.Co - cb * p.BN;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(t_acc + (uint32_t)c0, v);
      if (valid && c0 < co_left) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c0 + j), "f"(__uint_as_float(v[j])),
                       "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                       : "memory");
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// fp32 [npix, C] -> bf16 [npix, C] (+ per-channel sum / sum of squares of the fp32 values)
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ y,
                                                              float* __restrict__ stats, unsigned total, int C) {
  extern __shared__ float s_sum[];       // [2][C]
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i] = 0.f;
    __syncthreads();
  }
  const int cg = C >> 3;
  const int cg_shift = 31 - __clz(cg);
  const unsigned stride = ((gridDim.x * blockDim.x) >> cg_shift) << cg_shift;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i0 < stride;
  const int c0 = (int)(i0 & (unsigned)(cg - 1)) * 8;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (active) {
    for (unsigned i = i0; i < total; i += stride) {
      const float4 lo = *reinterpret_cast<const float4*>(acc + (size_t)i * 8);
      const float4 hi = *reinterpret_cast<const float4*>(acc + (size_t)i * 8 + 4);
      const float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      *reinterpret_cast<uint4*>(y + (size_t)i * 8) = u;
#pragma unroll
      for (int j = 0; j < 8; ++j) { a1[j] += f[j]; a2[j] += f[j] * f[j]; }
    }
  }
  if (stats) {
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&s_sum[c0 + j], a1[j]); atomicAdd(&s_sum[C + c0 + j], a2[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      const float v = s_sum[i];
      if (v != 0.f) atomicAdd(&stats[i], v);
    }
  }
}

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

}  // namespace sk

// x: bf16 [N, Hi, Wi, Ci]; wpack: bf16 [T, rows (= Co padded / CB*BN), Ci]; out32: fp32 [N, Ho, Wo, Co] zeroed
const char* launch_conv_splitk(const void* x, int N, int Hi, int Wi, int Ci, const void* wpack, int w_rows, int T,
                               const int* tap_y, const int* tap_x, int in_stride, float* out32, int Hg, int Wg, int Co,
                               int TH, int TW, int ksplit, int es, cudaStream_t stream) {
  using namespace sk;
  Params p{};
  if (TH * TW != 128) return "tile must cover 128 pixels";
  if (T < 1 || T > 16) return "taps";
  p.N = N; p.Hg = Hg; p.Wg = Wg; p.TH = TH; p.TW = TW; p.T = T; p.Ci = Ci; p.in_stride = in_stride;
  if (es != 2 && es != 4) return "operand element size must be 2 or 4";
  p.es = es;
  p.KB = Ci >= 128 / es ? 128 / es : Ci;
  if (p.KB * es != 32 && p.KB * es != 64 && p.KB * es != 128) return "the K block must span 32, 64 or 128 bytes";
  if (Ci % p.KB) return "Ci must be a multiple of the K block";
  p.kblocks = Ci / p.KB;
  for (int t = 0; t < T; ++t) { p.tap_y[t] = (int16_t)tap_y[t]; p.tap_x[t] = (int16_t)tap_x[t]; }
  p.Co = Co; p.Ho = Hg; p.Wo = Wg; p.out = out32;
  if (w_rows <= 256) { p.BN = w_rows; p.CB = 1; }
  else { if (w_rows % 128 || w_rows != Co) return "wide weight packs need Co == rows, a multiple of 128"; p.BN = 128; p.CB = w_rows / 128; }
  if (p.BN % 16) return "BN must be a multiple of 16";
  if (Co % 16) return "Co must be a multiple of 16";
  p.tiles_x = (Wg + TW - 1) / TW; p.tiles_y = (Hg + TH - 1) / TH;
  const int iters = T * p.kblocks;
  p.ksplit = ksplit < 1 ? 1 : (ksplit > iters ? iters : ksplit);
  int cols = 32; while (cols < p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t row_bytes = p.KB * es;
  const uint32_t stage_bytes = 128u * row_bytes + (((uint32_t)p.BN * row_bytes + 1023u) / 1024u) * 1024u;
  const size_t smem = (size_t)kStages * stage_bytes + 1024;
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return "cuTensorMapEncodeTiled is not available from this driver";
  CUtensorMap mx, mw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Ci, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Ci * es, (cuuint64_t)Wi * Ci * es, (cuuint64_t)Hi * Wi * Ci * es};
    cuuint32_t box[4] = {(cuuint32_t)p.KB, (cuuint32_t)(TW * in_stride), (cuuint32_t)(TH * in_stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)in_stride, (cuuint32_t)in_stride, 1};
    if (enc(&mx, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.KB * es), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the activation";
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Ci, (cuuint64_t)w_rows, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)Ci * es, (cuuint64_t)w_rows * Ci * es};
    cuuint32_t box[3] = {(cuuint32_t)p.KB, (cuuint32_t)p.BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&mw, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wpack), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(p.KB * es), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed for the weights";
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  dim3 grid(p.tiles_x * p.tiles_y * N * p.CB, p.ksplit, 1);
  conv_splitk_kernel<<<grid, kThreads, smem, stream>>>(mx, mw, p);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

void launch_splitk_finalize(const float* acc, void* y, float* stats, size_t npix, int C, cudaStream_t stream) {
  const size_t total = npix * (size_t)(C / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const size_t min_blocks = ((size_t)(C / 8) + 255) / 256;
  if (blocks < min_blocks) blocks = min_blocks;
  if (blocks == 0) blocks = 1;
  sk::splitk_finalize_kernel<<<(unsigned)blocks, 256, stats ? 2 * C * sizeof(float) : 0, stream>>>(
      acc, (__nv_bfloat16*)y, stats, (unsigned)total, C);
}

}  // namespace mine
