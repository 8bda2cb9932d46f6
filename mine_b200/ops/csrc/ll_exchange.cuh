// Cross-GPU SUM of a small per-channel statistic vector INSIDE the kernel that consumes it (SyncBatchNorm forward
// sums / backward sums: 134 exchanges per training step).  Same wire protocol as allreduce_small_ll_kernel
// (comm.cu): every value travels as one 8-byte word {fp32 bits, epoch} pushed into each peer's receive buffer over
// NVLink; receivers poll their OWN memory.  Here the exchange is the prologue of a multi-CTA elementwise kernel
// (normalise / BatchNorm-backward apply) instead of a separate single-CTA launch on the critical path:
//   * CTA (0,0) pushes this rank's vector to every peer;
//   * EVERY CTA polls the local receive buffer and sums the contributions in rank order into its shared memory
//     (identical bits on every rank and in every CTA), then runs its normal body with the reduced vector;
//   * CTA (0,0) also stores the reduced vector for later consumers (backward pass, running statistics);
//   * the device-resident epoch is advanced by the last CTA that has read it (ticket counter), so captured CUDA graphs
//     replay correctly and late-scheduled CTAs never see the next epoch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mine {

__device__ __forceinline__ uint64_t llx_global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All threads of the CTA must call this (contains __syncthreads).  local: this rank's n values (global memory);
// s_out: shared memory [n]; red_out: optional global copy of the reduced vector (written by CTA (0,0)).
__device__ __forceinline__ void ll_exchange_sum(const float* __restrict__ local, float* s_out, int n, const LLExchange& x,
                                                float* __restrict__ red_out) {
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(x.epoch) + 1;
  const size_t par_off = (size_t)(epoch & 1) * x.world * x.cap;
  const bool leader = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (leader) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t bits = __float_as_uint(local[i]);
      for (int p = 0; p < x.world; ++p) {
        if (p == x.rank) continue;
        uint2* dst = reinterpret_cast<uint2*>(x.ptr[p]) + par_off + (size_t)x.rank * x.cap + i;
        asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(dst), "r"(bits), "r"(epoch) : "memory");
      }
    }
  }
  const uint2* mine = reinterpret_cast<const uint2*>(x.ptr[x.rank]) + par_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float own = local[i];
    float acc = 0.f;
    for (int r = 0; r < x.world; ++r) {                     // rank order: identical bits everywhere
      float v = own;
      if (r != x.rank) {
        uint32_t lo, hi, spins = 0;
        uint64_t t0 = 0;
        const uint2* src = mine + (size_t)r * x.cap + i;
        do {
          asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(src) : "memory");
          if (hi != epoch && (++spins & 0xFFFF) == 0) {
            if (t0 == 0) t0 = llx_global_ns();
            else if (llx_global_ns() - t0 > 120ull * 1000ull * 1000ull * 1000ull) __trap();
          }
        } while (hi != epoch);
        v = __uint_as_float(lo);
      }
      acc += v;
    }
    s_out[i] = acc;
    if (leader && red_out) red_out[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // every thread of this CTA has read the epoch (they all passed the barrier): take a ticket; the last CTA advances it
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    __threadfence();
    if (atomicAdd(x.ticket, 1u) == total - 1) {
      *x.ticket = 0;
      *x.epoch = epoch;
      __threadfence();
    }
  }
}

}  // namespace mine
