// Python bindings: tensor validation + raw-pointer launchers on the current CUDA stream.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "kernels.h"

namespace {

inline void check_f32(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
inline const float* opt_ptr(const c10::optional<at::Tensor>& t, const char* name) {
  if (!t.has_value() || !t->defined()) return nullptr;
  check_f32(*t, name);
  return t->data_ptr<float>();
}
inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

// mpi: [B,S,H,W,4]
std::vector<at::Tensor> render_src_fwd(const at::Tensor& mpi, const at::Tensor& disparity, const at::Tensor& kinv,
                                       const c10::optional<at::Tensor>& src_img, bool use_alpha, bool blend,
                                       int64_t depth_mode, bool write_mpi) {
  check_f32(mpi, "mpi"); check_f32(disparity, "disparity"); check_f32(kinv, "k_src_inv");
  TORCH_CHECK(mpi.dim() == 5 && mpi.size(4) == 4, "mpi must be [B,S,H,W,4]");
  const int B = mpi.size(0), S = mpi.size(1), H = mpi.size(2), W = mpi.size(3);
  TORCH_CHECK(S <= 256, "at most 256 planes");
  c10::cuda::CUDAGuard guard(mpi.device());
  auto o = mpi.options();
  at::Tensor rgb = at::empty({B, 3, H, W}, o), depth = at::empty({B, 1, H, W}, o), wsum = at::empty({B, 1, H, W}, o);
  at::Tensor out_mpi = write_mpi ? at::empty_like(mpi) : at::Tensor();
  mine::launch_render_src_fwd(mpi.data_ptr<float>(), disparity.data_ptr<float>(), kinv.data_ptr<float>(),
                              opt_ptr(src_img, "src_img"), write_mpi ? out_mpi.data_ptr<float>() : nullptr,
                              rgb.data_ptr<float>(), depth.data_ptr<float>(), wsum.data_ptr<float>(), B, S, H, W,
                              use_alpha, blend, (int)depth_mode, cur_stream());
  return {rgb, depth, wsum, out_mpi};
}

at::Tensor render_src_bwd(const at::Tensor& mpi, const at::Tensor& disparity, const at::Tensor& kinv,
                          const c10::optional<at::Tensor>& src_img, const at::Tensor& depth_fwd,
                          const at::Tensor& wsum_fwd, const c10::optional<at::Tensor>& g_rgb,
                          const c10::optional<at::Tensor>& g_depth, const c10::optional<at::Tensor>& g_blend,
                          bool use_alpha, bool blend, int64_t depth_mode) {
  check_f32(mpi, "mpi");
  const int B = mpi.size(0), S = mpi.size(1), H = mpi.size(2), W = mpi.size(3);
  c10::cuda::CUDAGuard guard(mpi.device());
  at::Tensor g_mpi = at::empty_like(mpi);
  mine::launch_render_src_bwd(mpi.data_ptr<float>(), disparity.data_ptr<float>(), kinv.data_ptr<float>(),
                              opt_ptr(src_img, "src_img"), depth_fwd.data_ptr<float>(), wsum_fwd.data_ptr<float>(),
                              opt_ptr(g_rgb, "g_rgb"), opt_ptr(g_depth, "g_depth"), opt_ptr(g_blend, "g_blend"),
                              g_mpi.data_ptr<float>(), B, S, H, W, use_alpha, blend, (int)depth_mode, cur_stream());
  return g_mpi;
}

std::vector<at::Tensor> render_tgt_fwd(const at::Tensor& mpi, const at::Tensor& disparity, const at::Tensor& g_tgt_src,
                                       const at::Tensor& kinv, const at::Tensor& ktgt, bool use_alpha,
                                       int64_t depth_mode) {
  check_f32(mpi, "mpi"); check_f32(disparity, "disparity"); check_f32(g_tgt_src, "G_tgt_src");
  check_f32(kinv, "k_src_inv"); check_f32(ktgt, "k_tgt");
  TORCH_CHECK(mpi.dim() == 5 && mpi.size(4) == 4, "mpi must be [B,S,H,W,4]");
  const int B = mpi.size(0), S = mpi.size(1), H = mpi.size(2), W = mpi.size(3);
  c10::cuda::CUDAGuard guard(mpi.device());
  auto o = mpi.options();
  at::Tensor rgb = at::empty({B, 3, H, W}, o), depth = at::empty({B, 1, H, W}, o), mask = at::empty({B, 1, H, W}, o),
             wsum = at::empty({B, 1, H, W}, o);
  mine::launch_render_tgt_fwd(mpi.data_ptr<float>(), disparity.data_ptr<float>(), g_tgt_src.data_ptr<float>(),
                              kinv.data_ptr<float>(), ktgt.data_ptr<float>(), rgb.data_ptr<float>(),
                              depth.data_ptr<float>(), mask.data_ptr<float>(), wsum.data_ptr<float>(), B, S, H, W,
                              use_alpha, (int)depth_mode, cur_stream());
  return {rgb, depth, mask, wsum};
}

at::Tensor render_tgt_bwd(const at::Tensor& mpi, const at::Tensor& disparity, const at::Tensor& g_tgt_src,
                          const at::Tensor& kinv, const at::Tensor& ktgt, const at::Tensor& rgb_fwd,
                          const at::Tensor& depth_fwd, const at::Tensor& wsum_fwd,
                          const c10::optional<at::Tensor>& g_rgb, const c10::optional<at::Tensor>& g_depth,
                          bool use_alpha, int64_t depth_mode) {
  check_f32(mpi, "mpi");
  const int B = mpi.size(0), S = mpi.size(1), H = mpi.size(2), W = mpi.size(3);
  c10::cuda::CUDAGuard guard(mpi.device());
  at::Tensor g_mpi = at::zeros_like(mpi);
  mine::launch_render_tgt_bwd(mpi.data_ptr<float>(), disparity.data_ptr<float>(), g_tgt_src.data_ptr<float>(),
                              kinv.data_ptr<float>(), ktgt.data_ptr<float>(), rgb_fwd.data_ptr<float>(),
                              depth_fwd.data_ptr<float>(), wsum_fwd.data_ptr<float>(), opt_ptr(g_rgb, "g_rgb"),
                              opt_ptr(g_depth, "g_depth"), g_mpi.data_ptr<float>(), B, S, H, W, use_alpha,
                              (int)depth_mode, cur_stream());
  return g_mpi;
}

std::vector<at::Tensor> ssim_fwd(const at::Tensor& a, const at::Tensor& b, bool need_grad) {
  check_f32(a, "a"); check_f32(b, "b");
  TORCH_CHECK(a.dim() == 4 && a.sizes() == b.sizes(), "ssim expects two [N,C,H,W] tensors");
  const int planes = a.size(0) * a.size(1), H = a.size(2), W = a.size(3);
  c10::cuda::CUDAGuard guard(a.device());
  at::Tensor sum = at::zeros({}, a.options());
  at::Tensor partials = need_grad ? at::empty({planes, 3, H, W}, a.options()) : at::Tensor();
  mine::launch_ssim_fwd(a.data_ptr<float>(), b.data_ptr<float>(), sum.data_ptr<float>(),
                        need_grad ? partials.data_ptr<float>() : nullptr, planes, H, W, cur_stream());
  return {sum, partials};
}

at::Tensor ssim_bwd(const at::Tensor& a, const at::Tensor& b, const at::Tensor& partials, const at::Tensor& scale,
                    double scale_mul) {
  check_f32(a, "a"); check_f32(b, "b"); check_f32(partials, "partials"); check_f32(scale, "scale");
  const int planes = a.size(0) * a.size(1), H = a.size(2), W = a.size(3);
  c10::cuda::CUDAGuard guard(a.device());
  at::Tensor grad = at::empty_like(a);
  mine::launch_ssim_bwd(a.data_ptr<float>(), b.data_ptr<float>(), partials.data_ptr<float>(), scale.data_ptr<float>(),
                        (float)scale_mul, grad.data_ptr<float>(), planes, H, W, cur_stream());
  return grad;
}

std::vector<at::Tensor> masked_l1_fwd(const at::Tensor& a, const at::Tensor& b, const at::Tensor& mask, double thr,
                                      bool need_grad) {
  check_f32(a, "a"); check_f32(b, "b"); check_f32(mask, "mask");
  const int B = a.size(0), C = a.size(1), HW = a.size(2) * a.size(3);
  c10::cuda::CUDAGuard guard(a.device());
  at::Tensor sum = at::zeros({}, a.options());
  at::Tensor sign = need_grad ? at::empty_like(a) : at::Tensor();
  mine::launch_masked_l1_fwd(a.data_ptr<float>(), b.data_ptr<float>(), mask.data_ptr<float>(), (float)thr,
                             sum.data_ptr<float>(), need_grad ? sign.data_ptr<float>() : nullptr, B, C, HW,
                             cur_stream());
  return {sum, sign};
}

// disp [B,1,H,W] (or [B,H,W]); k [B,3,3]; xyz [B,3,N]; scale_in [B] or none
// -> {loss (0-dim), scale [B], idx int32 [B,N], d_syn [B,N], sgn [B,N]}
std::vector<at::Tensor> sparse_point_fwd(const at::Tensor& disp, const at::Tensor& k, const at::Tensor& xyz,
                                         const c10::optional<at::Tensor>& scale_in) {
  check_f32(disp, "disp"); check_f32(k, "k"); check_f32(xyz, "xyz");
  const int B = disp.size(0), H = disp.size(-2), W = disp.size(-1), N = xyz.size(2);
  TORCH_CHECK(disp.numel() == (int64_t)B * H * W && k.numel() == B * 9 && xyz.size(0) == B && xyz.size(1) == 3, "shapes");
  const float* sc = nullptr;
  if (scale_in.has_value() && scale_in->defined()) { check_f32(*scale_in, "scale"); TORCH_CHECK(scale_in->numel() == B, "scale"); sc = scale_in->data_ptr<float>(); }
  c10::cuda::CUDAGuard guard(disp.device());
  at::Tensor loss = at::zeros({}, disp.options());
  at::Tensor scale = at::empty({B}, disp.options());
  at::Tensor idx = at::empty({B, N}, disp.options().dtype(at::kInt));
  at::Tensor d_syn = at::empty({B, N}, disp.options());
  at::Tensor sgn = at::empty({B, N}, disp.options());
  mine::launch_sparse_point_fwd(disp.data_ptr<float>(), k.data_ptr<float>(), xyz.data_ptr<float>(), sc, idx.data_ptr<int>(),
                                d_syn.data_ptr<float>(), sgn.data_ptr<float>(), scale.data_ptr<float>(),
                                loss.data_ptr<float>(), B, H, W, N, cur_stream());
  return {loss, scale, idx, d_syn, sgn};
}

// -> {grad_disp (shape of disp), grad_scale_in [B] (zeros when the scale was computed in the forward)}
std::vector<at::Tensor> sparse_point_bwd(const at::Tensor& g_loss, const c10::optional<at::Tensor>& g_scale,
                                         const at::Tensor& idx, const at::Tensor& d_syn, const at::Tensor& sgn,
                                         const at::Tensor& scale, std::vector<int64_t> disp_shape, bool computed_scale) {
  check_f32(g_loss, "g_loss"); check_f32(d_syn, "d_syn"); check_f32(sgn, "sgn"); check_f32(scale, "scale");
  const int B = d_syn.size(0), N = d_syn.size(1);
  const int H = disp_shape[disp_shape.size() - 2], W = disp_shape[disp_shape.size() - 1];
  const float* gs = nullptr;
  if (g_scale.has_value() && g_scale->defined()) { check_f32(*g_scale, "g_scale"); gs = g_scale->data_ptr<float>(); }
  c10::cuda::CUDAGuard guard(d_syn.device());
  at::Tensor grad_disp = at::zeros(disp_shape, d_syn.options());
  at::Tensor grad_scale = at::zeros({B}, d_syn.options());
  mine::launch_sparse_point_bwd(g_loss.data_ptr<float>(), gs, idx.data_ptr<int>(), d_syn.data_ptr<float>(),
                                sgn.data_ptr<float>(), scale.data_ptr<float>(), grad_disp.data_ptr<float>(),
                                grad_scale.data_ptr<float>(), computed_scale ? 1 : 0, B, H, W, N, cur_stream());
  return {grad_disp, grad_scale};
}

void fused_adam(at::Tensor p, const at::Tensor& g, at::Tensor m, at::Tensor v, const at::Tensor& hyper, double beta1,
                double beta2, double eps, double wd) {
  check_f32(p, "p"); check_f32(g, "g"); check_f32(m, "m"); check_f32(v, "v"); check_f32(hyper, "hyper");
  TORCH_CHECK(hyper.numel() == 2, "hyper = {lr, step}");
  c10::cuda::CUDAGuard guard(p.device());
  mine::launch_fused_adam(p.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(),
                          p.numel(), hyper.data_ptr<float>(), (float)beta1, (float)beta2, (float)eps, (float)wd,
                          cur_stream());
}

// gradient gather: srcs[i] (dense fp32 tensors in their parameter's memory order; undefined -> zero fill) into
// flat[offsets[i] : offsets[i] + numels[i]]
void gather_into_arena(at::Tensor flat, std::vector<c10::optional<at::Tensor>> srcs, std::vector<int64_t> offsets,
                       std::vector<int64_t> numels) {
  check_f32(flat, "flat");
  TORCH_CHECK(srcs.size() == offsets.size() && srcs.size() == numels.size(), "list lengths differ");
  c10::cuda::CUDAGuard guard(flat.device());
  std::vector<const float*> sp(srcs.size());
  std::vector<float*> dp(srcs.size());
  for (size_t i = 0; i < srcs.size(); ++i) {
    TORCH_CHECK(offsets[i] >= 0 && offsets[i] + numels[i] <= flat.numel(), "slice outside the arena");
    dp[i] = flat.data_ptr<float>() + offsets[i];
    if (srcs[i].has_value() && srcs[i]->defined()) {
      const at::Tensor& t = *srcs[i];
      TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.numel() == numels[i] && t.is_non_overlapping_and_dense(),
                  "gradient ", i, " must be a dense CUDA fp32 tensor of its parameter's size");
      sp[i] = t.data_ptr<float>();
    } else {
      sp[i] = nullptr;
    }
  }
  mine::launch_multi_copy(sp.data(), dp.data(), numels.data(), (int)srcs.size(), cur_stream());
}

std::vector<at::Tensor> smooth_v2_fwd(const at::Tensor& img, const at::Tensor& disp, bool need_grad) {
  check_f32(img, "img"); check_f32(disp, "disp");
  TORCH_CHECK(img.dim() == 4 && img.size(1) == 3 && disp.dim() == 4 && disp.size(1) == 1, "img Bx3xHxW, disp Bx1xHxW");
  const int B = disp.size(0), H = disp.size(2), W = disp.size(3);
  c10::cuda::CUDAGuard guard(disp.device());
  auto o = disp.options();
  at::Tensor out = at::zeros({}, o), sums = at::zeros({B}, o), gd = at::zeros({B}, o);
  at::Tensor g = need_grad ? at::empty({B, 1, H, W}, o) : at::Tensor();
  mine::launch_smooth_v2_fwd(img.data_ptr<float>(), disp.data_ptr<float>(), sums.data_ptr<float>(), out.data_ptr<float>(),
                             need_grad ? g.data_ptr<float>() : nullptr, need_grad ? gd.data_ptr<float>() : nullptr, B, H, W,
                             cur_stream());
  return {out, sums, g, gd};
}

at::Tensor smooth_v2_bwd(const at::Tensor& g, const at::Tensor& sums, const at::Tensor& gd, const at::Tensor& gout) {
  check_f32(g, "g"); check_f32(gout, "gout");
  c10::cuda::CUDAGuard guard(g.device());
  at::Tensor grad = at::empty_like(g);
  const int B = g.size(0), HW = g.size(2) * g.size(3);
  mine::launch_smooth_v2_bwd(g.data_ptr<float>(), sums.data_ptr<float>(), gd.data_ptr<float>(), gout.data_ptr<float>(),
                             grad.data_ptr<float>(), B, HW, cur_stream());
  return grad;
}

std::vector<at::Tensor> smooth_v1_fwd(const at::Tensor& img, const at::Tensor& disp, double gmin, double ratio,
                                      bool need_grad) {
  check_f32(img, "img"); check_f32(disp, "disp");
  TORCH_CHECK(img.dim() == 4 && img.size(1) == 3 && disp.dim() == 4 && disp.size(1) == 1, "img Bx3xHxW, disp Bx1xHxW");
  const int B = disp.size(0), H = disp.size(2), W = disp.size(3);
  c10::cuda::CUDAGuard guard(disp.device());
  auto o = disp.options();
  at::Tensor out = at::zeros({}, o), stats = at::zeros({B, 6}, o), hs = at::zeros({B, 4}, o);
  at::Tensor sob = at::empty({B, 2, H, W}, o);
  at::Tensor hmap = need_grad ? at::empty({B, 2, H, W}, o) : at::Tensor();
  mine::launch_smooth_v1_fwd(img.data_ptr<float>(), disp.data_ptr<float>(), stats.data_ptr<float>(), sob.data_ptr<float>(),
                             out.data_ptr<float>(), need_grad ? hmap.data_ptr<float>() : nullptr, hs.data_ptr<float>(),
                             (float)gmin, (float)ratio, B, H, W, cur_stream());
  return {out, stats, sob, hmap, hs};
}

at::Tensor smooth_v1_bwd(const at::Tensor& sob, const at::Tensor& stats, const at::Tensor& hmap, const at::Tensor& hs,
                         const at::Tensor& gout) {
  check_f32(sob, "sob"); check_f32(hmap, "hmap"); check_f32(gout, "gout");
  c10::cuda::CUDAGuard guard(sob.device());
  const int B = sob.size(0), H = sob.size(2), W = sob.size(3);
  at::Tensor grad = at::zeros({B, 1, H, W}, sob.options());
  mine::launch_smooth_v1_bwd(sob.data_ptr<float>(), stats.data_ptr<float>(), hmap.data_ptr<float>(), hs.data_ptr<float>(),
                             gout.data_ptr<float>(), grad.data_ptr<float>(), B, H, W, cur_stream());
  return grad;
}

}  // namespace

void register_conv(pybind11::module_& m);     // conv_bindings.cpp
void register_comm(pybind11::module_& m);     // comm_bindings.cpp

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("render_src_fwd", &render_src_fwd);
  m.def("render_src_bwd", &render_src_bwd);
  m.def("render_tgt_fwd", &render_tgt_fwd);
  m.def("render_tgt_bwd", &render_tgt_bwd);
  m.def("ssim_fwd", &ssim_fwd);
  m.def("ssim_bwd", &ssim_bwd);
  m.def("masked_l1_fwd", &masked_l1_fwd);
  m.def("sparse_point_fwd", &sparse_point_fwd);
  m.def("sparse_point_bwd", &sparse_point_bwd);
  m.def("fused_adam", &fused_adam);
  m.def("gather_into_arena", &gather_into_arena);
  m.def("smooth_v2_fwd", &smooth_v2_fwd);
  m.def("smooth_v2_bwd", &smooth_v2_bwd);
  m.def("smooth_v1_fwd", &smooth_v1_fwd);
  m.def("smooth_v1_bwd", &smooth_v1_bwd);
  register_conv(m);
  register_comm(m);
}
