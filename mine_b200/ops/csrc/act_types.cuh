// Activation element types of the conv engine: bf16 (fast mode) or fp32 storage (TF32 tensor-core math - the
// numerics class of the reference's cuDNN TF32 convolutions).  Every memory-bound companion kernel is templated
// on the storage type and works on groups of 8 channels: one 16-byte vector (bf16) or two (fp32).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mine {

struct V8 {
  float f[8];
};

__device__ __forceinline__ V8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  V8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    r.f[2 * i] = t.x; r.f[2 * i + 1] = t.y;
  }
  return r;
}
__device__ __forceinline__ V8 ld8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  V8 r;
  r.f[0] = a.x; r.f[1] = a.y; r.f[2] = a.z; r.f[3] = a.w;
  r.f[4] = b.x; r.f[5] = b.y; r.f[6] = b.z; r.f[7] = b.w;
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const V8& v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v.f[2 * i], v.f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8(float* p, const V8& v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.f[0], v.f[1], v.f[2], v.f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v.f[4], v.f[5], v.f[6], v.f[7]);
}

// Store of a tensor-core OPERAND (activation / gradient that a convolution will read): fp32 storage is rounded to the
// nearest TF32 value (cvt.rna, what cuDNN / cuBLAS do when they load fp32 data for TF32 math), because kind::tf32
// ignores the 13 low mantissa bits - truncation would bias every product towards zero.  bf16: plain store.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void st8_op(__nv_bfloat16* p, const V8& v) { st8(p, v); }
__device__ __forceinline__ void st8_op(float* p, const V8& v) {
  *reinterpret_cast<float4*>(p) = make_float4(round_tf32(v.f[0]), round_tf32(v.f[1]), round_tf32(v.f[2]), round_tf32(v.f[3]));
  *reinterpret_cast<float4*>(p + 4) = make_float4(round_tf32(v.f[4]), round_tf32(v.f[5]), round_tf32(v.f[6]), round_tf32(v.f[7]));
}

// host-side dispatch on the element size (2: bf16, 4: fp32)
#define MINE_DISPATCH_ES(es, T, ...)                 \
  do {                                               \
    if ((es) == 4) { using T = float; __VA_ARGS__; } \
    else { using T = __nv_bfloat16; __VA_ARGS__; }   \
  } while (0)

}  // namespace mine
