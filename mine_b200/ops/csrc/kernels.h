// Launcher declarations shared by the .cu translation units (raw pointers + stream; no torch
// headers here so kernels compile in seconds) and bindings.cpp (tensor checks, streams).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mine {

// ---- render.cu -----------------------------------------------------------------------------
// depth_mode: 0 = sum(w z)/(sum w + 1e-5), 1 = background at infinity, 2 = raw sum(w z)
void launch_render_src_fwd(const float* mpi, const float* disparity, const float* kinv, const float* src_img,
                           float* mpi_out, float* rgb, float* depth, float* wsum, int B, int S, int H, int W,
                           bool use_alpha, bool blend, int depth_mode, cudaStream_t stream);
void launch_render_src_bwd(const float* mpi, const float* disparity, const float* kinv, const float* src_img,
                           const float* depth_fwd, const float* wsum_fwd, const float* g_rgb, const float* g_depth,
                           const float* g_blend, float* g_mpi, int B, int S, int H, int W, bool use_alpha, bool blend,
                           int depth_mode, cudaStream_t stream);
void launch_render_tgt_fwd(const float* mpi, const float* disparity, const float* g_tgt_src, const float* kinv,
                           const float* ktgt, float* rgb, float* depth, float* mask, float* wsum, int B, int S, int H,
                           int W, bool use_alpha, int depth_mode, cudaStream_t stream);
void launch_render_tgt_bwd(const float* mpi, const float* disparity, const float* g_tgt_src, const float* kinv,
                           const float* ktgt, const float* rgb_fwd, const float* depth_fwd, const float* wsum_fwd,
                           const float* g_rgb, const float* g_depth, float* g_mpi, int B, int S, int H, int W,
                           bool use_alpha, int depth_mode, cudaStream_t stream);

// ---- losses.cu -----------------------------------------------------------------------------
// SSIM (11x11 Gaussian, sigma 1.5, zero padding).  sum_out += sum of the SSIM map; when `partials`
// is non-null the three adjoint maps X1,X2,X3 [planes,3,H,W] needed by the backward are stored.
void launch_ssim_fwd(const float* a, const float* b, float* sum_out, float* partials, int planes, int H, int W,
                     cudaStream_t stream);
// grad_a = scale * (blur(X1) + 2 a blur(X2) + b blur(X3)); scale read from device memory (*scale_ptr * scale_mul)
void launch_ssim_bwd(const float* a, const float* b, const float* partials, const float* scale_ptr, float scale_mul,
                     float* grad_a, int planes, int H, int W, cudaStream_t stream);
// masked L1: sum_out += sum |a-b| * [mask >= thr]; grad_sign (optional) = sign(a-b) * [mask >= thr]
void launch_masked_l1_fwd(const float* a, const float* b, const float* mask, float thr, float* sum_out,
                          float* grad_sign, int B, int C, int HW, cudaStream_t stream);

// ---- adam.cu -------------------------------------------------------------------------------
// hyper: device array {lr, step}
void launch_fused_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1,
                       float beta2, float eps, float weight_decay, cudaStream_t stream);
// srcs[i] (null: zero fill) -> dsts[i], numels[i] floats each, a handful of multi-tensor launches
void launch_multi_copy(const float* const* srcs, float* const* dsts, const int64_t* numels, int count, cudaStream_t stream);

}  // namespace mine

namespace mine {
// cross-GPU statistic exchange fused into the consuming kernel (device side: ll_exchange.cuh); null pointer: none
struct LLExchange {
  void* ptr[16];          // per-rank receive buffers: [2 parities][world sources][cap] uint2 (symmetric allocation)
  int rank, world, cap;   // world <= 1: exchange disabled
  uint32_t* epoch;        // device counter shared with allreduce_small_ll_kernel
  uint32_t* ticket;       // device counter (zero between kernels)
};
// ---- decoder_elem.cu (NHWC bf16) ---------------------------------------------------------------
// pad_mode: 0 = reflection, 1 = replication (1 pixel); stats = [2, C] global (sum, sum of squares)
void launch_bn_act_pad_fwd(const void* y, const float* stats, const float* gamma, const float* beta, void* out, int N,
                           int H, int W, int C, int pad_mode, float inv_count, float eps, int es, const LLExchange* x,
                           float* red_out, cudaStream_t stream);
void launch_bn_act_bwd_reduce(const void* dapad, const void* y, const float* stats, const float* gamma,
                              const float* beta, void* g_out, float* sums, int N, int H, int W, int C, int pad_mode,
                              float inv_count, float eps, int es, cudaStream_t stream);
void launch_bn_bwd_apply(const void* g, const void* y, const float* stats, const float* gamma, const float* sums,
                         void* dy, float* dshared, float* dplane_bias, int B, int S, int H, int W, int C,
                         float inv_count, float eps, int es, const LLExchange* x, const float* beta, int pad_mode,
                         cudaStream_t stream);
// pad_nhwc.cu: 1-pixel reflection pad of an NHWC tensor [N,H,W,C] -> [N,H+2,W+2,C] and its adjoint (es: 2 bf16, 4 fp32)
void launch_pad_reflect_nhwc(const void* x, void* out, int N, int H, int W, int C, int es, cudaStream_t stream);
void launch_pad_reflect_nhwc_bwd(const void* gp, void* gx, int N, int H, int W, int C, int es, cudaStream_t stream);
void launch_head_bwd(const float* g_mpi, const float* mpi, const int8_t* sign, void* dz, float* dbias, size_t npix,
                     int use_alpha, int es, cudaStream_t stream);
}  // namespace mine

namespace mine {
// ---- conv_splitk.cu (split-K implicit GEMM for layers with few output tiles; fp32 partial sums + finalize) -------
const char* launch_conv_splitk(const void* x, int N, int Hi, int Wi, int Ci, const void* wpack, int w_rows, int T,
                               const int* tap_y, const int* tap_x, int in_stride, float* out32, int Hg, int Wg, int Co,
                               int TH, int TW, int ksplit, int es, cudaStream_t stream);
void launch_splitk_finalize(const float* acc, void* y, float* stats, size_t npix, int C, cudaStream_t stream);
}  // namespace mine

namespace mine {
// ---- sparse.cu (sparse-point disparity supervision: projection, gather, scale calibration, log-L1) -------
void launch_sparse_point_fwd(const float* disp, const float* K, const float* xyz, const float* scale_in, int* idx,
                             float* d_syn, float* sgn, float* scale_out, float* loss, int B, int H, int W, int N,
                             cudaStream_t stream);
void launch_sparse_point_bwd(const float* g_loss, const float* g_scale, const int* idx, const float* d_syn,
                             const float* sgn, const float* scale, float* grad_disp, float* grad_scale_in,
                             int computed_scale, int B, int H, int W, int N, cudaStream_t stream);
}  // namespace mine

namespace mine {
// ---- head_direct.cu (CUDA-core MPI head for the 16 / 32-channel levels) ----------------------------------
const char* launch_head_conv_direct(const void* apad, const float* wpk, const float* bias, float* mpi, int8_t* sign, int N,
                                    int H, int W, int C, int use_alpha, cudaStream_t stream);
}  // namespace mine

namespace mine {
// ---- encoder_elem.cu (unpadded NHWC bf16, C a power of two in [16, 2048]) -------------------------------
void launch_bn_res_act_fwd(const void* y, const float* stats, const float* gamma, const float* beta, const void* res,
                           void* out, size_t npix, int C, float slope, float inv_count, float eps, int es,
                           const LLExchange* x, float* red_out, int round_out, cudaStream_t stream);
int reduce_blocks(size_t npix, int C);   // grid of the two reproducible reductions below
size_t reduce_scratch_floats(size_t npix, int C);   // their per-call scratch; ticket: persistent zeroed unsigned[64] per device
void launch_bn_res_act_bwd_reduce(const void* dout, const void* out, const void* y, const float* stats, void* g_out,
                                  float* sums, size_t npix, int C, float slope, float inv_count, float eps, int es,
                                  float* scratch, unsigned* ticket, cudaStream_t stream);
void launch_channel_stats(const void* y, float* sums, size_t npix, int C, int es, float* scratch, unsigned* ticket,
                          cudaStream_t stream);
void launch_bn_update_running(const float* stats, float* running_mean, float* running_var, long long* num_batches, int C,
                              float count, float momentum, cudaStream_t stream);
void launch_bn_update_running_multi(int n, const float* const* stats, float* const* mean, float* const* var,
                                    long long* const* nbt, const int* C, const float* count, const float* momentum,
                                    cudaStream_t stream);
}  // namespace mine

namespace mine {
// ---- comm.cu (NVLink peer memory) -----------------------------------------------------------------
struct PeerTable { void* ptr[16]; };
// epoch counters are device-resident (advanced by the kernels): CUDA-graph replay safe
void launch_allreduce_small(float* inout, int n, const PeerTable& data, const PeerTable& flags, int rank, int world,
                            int cap, uint32_t* epoch_ptr, cudaStream_t stream);
void launch_allreduce_small_ll(float* inout, int n, const PeerTable& ll, int rank, int world, int cap,
                               uint32_t* epoch_ptr, cudaStream_t stream);
void launch_allreduce_mean(const PeerTable& arena, const PeerTable& flags, float* mc_arena, int64_t lo, int64_t hi,
                           int rank, int world, uint32_t* epochs, int blocks, cudaStream_t stream);
}  // namespace mine

namespace mine {
// ---- smooth.cu (edge-aware smoothness v1 / v2; buffers sums/stats/out/gd/hs/grad must be zeroed by the caller) ----
void launch_smooth_v2_fwd(const float* img, const float* disp, float* sums, float* out, float* g, float* gd, int B, int H,
                          int W, cudaStream_t stream);
void launch_smooth_v2_bwd(const float* g, const float* sums, const float* gd, const float* gout, float* grad, int B, int HW,
                          cudaStream_t stream);
void launch_smooth_v1_fwd(const float* img, const float* disp, float* stats, float* sob, float* out, float* hmap, float* hs,
                          float gmin, float ratio, int B, int H, int W, cudaStream_t stream);
void launch_smooth_v1_bwd(const float* sob, const float* stats, const float* hmap, const float* hs, const float* gout,
                          float* grad, int B, int H, int W, cudaStream_t stream);
}  // namespace mine
