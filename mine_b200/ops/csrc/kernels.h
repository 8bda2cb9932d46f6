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
void launch_fused_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float bias_corr1, float bias_corr2, cudaStream_t stream);

}  // namespace mine
