// Hand-written all-reduce kernels over NVLink peer memory (symmetric heap), sm_100a.
//
// Replaces the two NCCL paths of the reference's data parallelism:
//   N6/N7  134 tiny SyncBatchNorm collectives per step  -> allreduce_small_oneshot: every rank publishes
//          its <=64 Ki-float vector in its own symmetric slot, one flag round over NVLink, every rank then
//          pulls all peers' vectors with P2P loads and sums them in RANK ORDER (bit-identical on all
//          ranks, so BN running statistics never diverge).  One kernel, latency ~ a flag round trip.
//   N8     DDP bucketed gradient all-reduce (+ divide by world) -> allreduce_mean_twoshot, IN PLACE on the
//          symmetric gradient arena: rank r owns slice r of the bucket, reads that slice from every peer
//          (16-byte P2P loads), fuses the 1/world scaling (and optionally a bf16 mirror for the next
//          forward), and stores the result into EVERY peer's arena (P2P stores) - reduce-scatter and
//          all-gather in one kernel with device-side flag barriers before and after.  With an NVSwitch
//          multicast mapping the same kernel uses multimem.ld_reduce / multimem.st (in-switch reduction).
// No NCCL call is made on either path; launched on a side stream so it overlaps the rest of backward.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mine {

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_v4(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Spin guard: a peer that never arrives (crashed rank, mismatched call sequence) must not hang the GPU forever:
// after ~120 s of polling the kernel traps, which surfaces as a CUDA error on the host.
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kSpinLimitNs = 120ull * 1000ull * 1000ull * 1000ull;   // ranks can be skewed by seconds during warm-up (autotuning, graph capture)

// Flag barrier between the same-numbered CTA of every rank.  flags: [channels][world] uint32 per rank
// (symmetric).  Thread p < world signals peer p and waits for peer p's signal; values only grow.
__device__ __forceinline__ void peer_barrier(const PeerTable& flags, int rank, int world, int channel, uint32_t epoch) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int p = threadIdx.x;
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(flags.ptr[p]) + channel * world + rank, epoch);
    const uint32_t* mine_flag = reinterpret_cast<const uint32_t*>(flags.ptr[rank]) + channel * world + p;
    const uint64_t t0 = global_ns();
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine_flag) - epoch) < 0) {
      if ((++spins & 0xFFFF) == 0 && global_ns() - t0 > kSpinLimitNs) __trap();
    }
  }
  __syncthreads();
}

// Epochs live in DEVICE memory and are advanced by the kernels themselves, so a captured CUDA graph can be
// replayed: every replay sees the next epoch without any host-side argument changing.
//
// ---- small one-shot SUM (BN statistics) -------------------------------------------------------------
// data: [2 slots][cap] floats per rank (symmetric); slot = epoch & 1.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) allreduce_small_oneshot_kernel(float* __restrict__ inout, int n, PeerTable data,
                                                                       PeerTable flags, int rank, int world, int cap,
                                                                       uint32_t* __restrict__ epoch_ptr) {
  const uint32_t epoch = *epoch_ptr + 1;
  float* my_slot = reinterpret_cast<float*>(data.ptr[rank]) + (size_t)(epoch & 1) * cap;
  const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(inout) & 15) == 0);
  if (vec) {
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x)
      reinterpret_cast<float4*>(my_slot)[i] = reinterpret_cast<const float4*>(inout)[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) my_slot[i] = inout[i];
  }
  peer_barrier(flags, rank, world, 0, epoch);
  if (vec) {
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < world; ++p) {                 // rank order: identical bits everywhere
        const float4 v = ld_peer_v4(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(data.ptr[p]) + (size_t)(epoch & 1) * cap) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(inout)[i] = acc;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float acc = 0.f;
      for (int p = 0; p < world; ++p) {
        const float* src = reinterpret_cast<const float*>(data.ptr[p]) + (size_t)(epoch & 1) * cap;
        float v;
        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(src + i) : "memory");
        acc += v;
      }
      inout[i] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ptr = epoch;
}

// ---- small SUM, low-latency variant ------------------------------------------------------------------
// Every value travels as one 8-byte word {fp32 bits, epoch}: the sender PUSHES its vector into every peer's
// receive buffer, the receiver polls its own memory until the word carries the current epoch.  Data and flag
// arrive together (64-bit stores are single-copy atomic), so one NVLink traversal replaces the
// publish / flag / pull round trips of the kernel above.  ll: [2 parities][world sources][cap] uint2 per rank.
__global__ void __launch_bounds__(1024) allreduce_small_ll_kernel(float* __restrict__ inout, int n, PeerTable ll, int rank,
                                                                  int world, int cap, uint32_t* __restrict__ epoch_ptr) {
  __shared__ uint2* s_dst[16];
  const uint32_t epoch = *epoch_ptr + 1;
  const size_t par_off = (size_t)(epoch & 1) * world * cap;
  if ((int)threadIdx.x < world)
    s_dst[threadIdx.x] = reinterpret_cast<uint2*>(ll.ptr[threadIdx.x]) + par_off + (size_t)rank * cap;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t bits = __float_as_uint(inout[i]);
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;
      asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(s_dst[p] + i), "r"(bits), "r"(epoch) : "memory");
    }
  }
  const uint2* mine = reinterpret_cast<const uint2*>(ll.ptr[rank]) + par_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float own = inout[i];
    float acc = 0.f;
    for (int r = 0; r < world; ++r) {                     // rank order: identical bits on every rank
      float x = own;
      if (r != rank) {
        uint32_t lo, hi;
        const uint2* src = mine + (size_t)r * cap + i;
        uint32_t spins = 0;
        uint64_t t0 = 0;
        do {
          asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(src) : "memory");
          if (hi != epoch && (++spins & 0xFFFF) == 0) {
            if (t0 == 0) t0 = global_ns();
            else if (global_ns() - t0 > kSpinLimitNs) __trap();
          }
        } while (hi != epoch);
        x = __uint_as_float(lo);
      }
      acc += x;
    }
    inout[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ptr = epoch;
}

// ---- two-shot mean, in place on the symmetric arena ---------------------------------------------------
// Elements [lo, hi) of every rank's arena (float offsets, multiples of 4).  Rank r reduces its 1/world slice.
// Four independent 16-byte peer loads per thread are in flight before the first add (NVLink latency ~2 us).
constexpr int kUnroll = 4;
__global__ void __launch_bounds__(512) allreduce_mean_twoshot_kernel(PeerTable arena, PeerTable flags, float* mc_arena,
                                                                     int64_t lo, int64_t hi, int rank, int world,
                                                                     float scale, uint32_t* __restrict__ epochs,
                                                                     int use_multimem) {
  const int channel = blockIdx.x + 1;                       // channel 0 belongs to the small one-shot kernel
  const uint32_t epoch = epochs[blockIdx.x] + 1;
  peer_barrier(flags, rank, world, channel, epoch);         // every rank's gradients for [lo, hi) are final
  const int64_t n4 = (hi - lo) / 4;
  const int64_t per = (n4 + world - 1) / world;
  const int64_t s4 = rank * per, e4 = (s4 + per < n4) ? s4 + per : n4;
  const int64_t base4 = lo / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = s4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < e4; i0 += stride * kUnroll) {
    float4 acc[kUnroll];
    if (use_multimem) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < e4) acc[u] = multimem_ld_reduce_v4(reinterpret_cast<const float4*>(mc_arena) + base4 + i);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < world; ++p) {                   // fixed order: identical bits on every rank
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int64_t i = i0 + u * stride;
          v[u] = (i < e4) ? ld_peer_v4(reinterpret_cast<const float4*>(arena.ptr[p]) + base4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= e4) continue;
      float4 r = acc[u];
      r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
      if (use_multimem) {
        multimem_st_v4(reinterpret_cast<float4*>(mc_arena) + base4 + i, r);
      } else {
        for (int p = 0; p < world; ++p) st_peer_v4(reinterpret_cast<float4*>(arena.ptr[p]) + base4 + i, r);
      }
    }
  }
  peer_barrier(flags, rank, world, channel, epoch + 1);     // all slices have landed everywhere
  if (threadIdx.x == 0) epochs[blockIdx.x] = epoch + 1;
}

void launch_allreduce_small(float* inout, int n, const PeerTable& data, const PeerTable& flags, int rank, int world,
                            int cap, uint32_t* epoch_ptr, cudaStream_t stream) {
  int threads = n >= 4096 ? 1024 : (n >= 1024 ? 512 : 256);
  allreduce_small_oneshot_kernel<<<1, threads, 0, stream>>>(inout, n, data, flags, rank, world, cap, epoch_ptr);
}

void launch_allreduce_small_ll(float* inout, int n, const PeerTable& ll, int rank, int world, int cap,
                               uint32_t* epoch_ptr, cudaStream_t stream) {
  int threads = n >= 2048 ? 1024 : (n >= 512 ? 512 : 256);
  allreduce_small_ll_kernel<<<1, threads, 0, stream>>>(inout, n, ll, rank, world, cap, epoch_ptr);
}

void launch_allreduce_mean(const PeerTable& arena, const PeerTable& flags, float* mc_arena, int64_t lo, int64_t hi,
                           int rank, int world, uint32_t* epochs, int blocks, cudaStream_t stream) {
  allreduce_mean_twoshot_kernel<<<blocks, 512, 0, stream>>>(arena, flags, mc_arena, lo, hi, rank, world, 1.0f / world,
                                                            epochs, mc_arena != nullptr ? 1 : 0);
}

}  // namespace mine
