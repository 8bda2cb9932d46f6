// Bindings of the tcgen05 conv engine and its elementwise companions.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstring>
#include <tuple>

#include "conv_engine.h"
#include "kernels.h"

namespace {

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

// Operand storage type of the engine (set once per process by conv_engine.set_precision): 2 = bf16, 4 = fp32 (TF32 math).
// Only the entry points whose OUTPUT type cannot be inferred from an input (weight packs, head gradient) read it.
int g_operand_es = 2;
inline at::ScalarType operand_dtype() { return g_operand_es == 4 ? at::kFloat : at::kBFloat16; }

inline void check_act_nhwc(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && (t.scalar_type() == at::kBFloat16 || t.scalar_type() == at::kFloat) && t.is_contiguous() &&
                  t.dim() == 4, name, " must be a contiguous CUDA bf16 or fp32 [N,H,W,C] tensor");
  TORCH_CHECK(t.size(3) % 8 == 0, name, ": channels must be a multiple of 8");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0, name, " must be 16-byte aligned");
}
inline int esize(const at::Tensor& t) { return t.scalar_type() == at::kFloat ? 4 : 2; }
inline void check_same_type(const at::Tensor& a, const at::Tensor& b, const char* what) {
  TORCH_CHECK(a.scalar_type() == b.scalar_type(), what, ": operand dtypes differ");
}
inline const float* opt_f32(const c10::optional<at::Tensor>& t, const char* name) {
  if (!t.has_value() || !t->defined()) return nullptr;
  TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous(), name, " must be contiguous CUDA fp32");
  return t->data_ptr<float>();
}

// Fused statistic exchange (ll_exchange.cuh): peer receive buffers of the P2P communicator + its device counters.
struct FusedX {
  mine::LLExchange x{};
  bool on = false;
};
FusedX make_fused(const std::vector<int64_t>& ll_ptrs, int64_t rank, int64_t cap, const c10::optional<at::Tensor>& epoch,
                  const c10::optional<at::Tensor>& ticket, int64_t n) {
  FusedX f;
  if (ll_ptrs.size() < 2) return f;
  TORCH_CHECK(ll_ptrs.size() <= 16, "at most 16 peers");
  TORCH_CHECK(epoch.has_value() && ticket.has_value() && epoch->is_cuda() && ticket->is_cuda() &&
                  epoch->scalar_type() == at::kInt && ticket->scalar_type() == at::kInt, "epoch / ticket: CUDA int32 tensors");
  TORCH_CHECK(n <= cap, "statistic vector larger than the exchange slot");
  for (size_t i = 0; i < ll_ptrs.size(); ++i) f.x.ptr[i] = reinterpret_cast<void*>(ll_ptrs[i]);
  f.x.rank = (int)rank; f.x.world = (int)ll_ptrs.size(); f.x.cap = (int)cap;
  f.x.epoch = reinterpret_cast<uint32_t*>(epoch->data_ptr<int>());
  f.x.ticket = reinterpret_cast<uint32_t*>(ticket->data_ptr<int>());
  f.on = true;
  return f;
}

void conv_taps(const at::Tensor& x, const at::Tensor& wpack, at::Tensor out, int64_t Hg, int64_t Wg, int64_t G, int64_t T,
               std::vector<int64_t> tap_y, std::vector<int64_t> tap_x, int64_t in_stride, int64_t Co, int64_t out_sy,
               int64_t out_sx, std::vector<int64_t> out_oy, std::vector<int64_t> out_ox, bool accumulate,
               const c10::optional<at::Tensor>& chan_bias, const c10::optional<at::Tensor>& plane_bias,
               const c10::optional<at::Tensor>& shared_map, int64_t planes_per_image,
               const c10::optional<at::Tensor>& stats, int64_t act, bool head_alpha,
               const c10::optional<at::Tensor>& raw_out, int64_t TH, int64_t TW) {
  check_act_nhwc(x, "x");
  TORCH_CHECK(wpack.is_cuda() && wpack.scalar_type() == x.scalar_type() && wpack.is_contiguous() && wpack.dim() == 3,
              "wpack must be [G*T, BN, Ci] in the dtype of x");
  TORCH_CHECK(wpack.size(0) == G * T && wpack.size(2) == x.size(3), "wpack shape mismatch");
  TORCH_CHECK((int64_t)tap_y.size() == G * T && (int64_t)tap_x.size() == G * T, "tap table size");
  TORCH_CHECK((int64_t)out_oy.size() == G && (int64_t)out_ox.size() == G, "group offset size");
  TORCH_CHECK(out.is_cuda() && out.is_contiguous(), "out must be contiguous CUDA");
  c10::cuda::CUDAGuard guard(x.device());
  mine::ConvLaunch L{};
  mine::ConvParams& p = L.p;
  p.N = x.size(0); p.Hg = Hg; p.Wg = Wg; p.TH = TH; p.TW = TW;
  p.G = G; p.T = T; p.Ci = x.size(3);
  p.es = esize(x);
  p.KB = p.Ci >= 128 / p.es ? 128 / p.es : p.Ci;
  p.in_stride = in_stride;
  for (int g = 0; g < G; ++g) {
    for (int t = 0; t < T; ++t) { p.tap_y[g][t] = (int16_t)tap_y[g * T + t]; p.tap_x[g][t] = (int16_t)tap_x[g * T + t]; }
    p.out_oy[g] = (int16_t)out_oy[g]; p.out_ox[g] = (int16_t)out_ox[g];
  }
  // wide layers (encoder): Co is produced in blocks of 128 accumulator columns; <= 256 rows: one block (decoder)
  const int64_t rows = wpack.size(1);
  p.Co = Co;
  if (rows <= 256) { p.BN = rows; p.CB = 1; }
  else {
    TORCH_CHECK(rows % 128 == 0 && rows == Co, "wide weight packs need Co == rows, a multiple of 128");
    p.BN = 128; p.CB = rows / 128;
  }
  L.w_rows = rows;
  p.out_sy = out_sy; p.out_sx = out_sx;
  p.out = out.data_ptr();
  p.act = act; p.head_alpha = head_alpha;
  if (act == 1) {
    TORCH_CHECK(out.scalar_type() == at::kFloat && out.size(-1) == 4, "head output must be fp32 [...,4]");
    p.Ho = out.size(-3); p.Wo = out.size(-2);
    TORCH_CHECK(out.numel() == (int64_t)p.N * p.Ho * p.Wo * 4, "head output shape");
  } else {
    TORCH_CHECK(out.dim() == 4 && out.size(0) == p.N && out.size(3) == Co, "out must be [N,Ho,Wo,Co]");
    p.Ho = out.size(1); p.Wo = out.size(2);
    TORCH_CHECK(out.scalar_type() == at::kBFloat16 || out.scalar_type() == at::kFloat, "out dtype");
    p.out_fp32 = out.scalar_type() == at::kFloat;
    TORCH_CHECK(Co % 16 == 0, "Co must be a multiple of 16 for tensor outputs");
  }
  p.accumulate = accumulate;
  p.chan_bias = opt_f32(chan_bias, "chan_bias");
  p.plane_bias = opt_f32(plane_bias, "plane_bias");
  p.shared_map = opt_f32(shared_map, "shared_map");
  p.planes_per_image = planes_per_image > 0 ? planes_per_image : 1;
  if (p.shared_map) TORCH_CHECK(shared_map->numel() == (int64_t)(p.N / p.planes_per_image) * p.Ho * p.Wo * Co, "shared_map shape");
  if (p.plane_bias) TORCH_CHECK(plane_bias->numel() == (int64_t)p.N * Co, "plane_bias shape");
  p.stats = const_cast<float*>(opt_f32(stats, "stats"));
  if (p.stats) TORCH_CHECK(stats->numel() == 2 * Co, "stats must be [2, Co]");
  p.raw_out = (raw_out.has_value() && raw_out->defined()) ? raw_out->data_ptr() : nullptr;
  L.x = x.data_ptr(); L.Hi = x.size(1); L.Wi = x.size(2);
  L.w = wpack.data_ptr();
  const char* err = mine::launch_conv_taps(L, cur_stream());
  TORCH_CHECK(err == nullptr, "conv_taps: ", err ? err : "");
}

void set_operand_size(int64_t es) {
  TORCH_CHECK(es == 2 || es == 4, "operand element size must be 2 (bf16) or 4 (fp32 storage, TF32 math)");
  g_operand_es = (int)es;
}
int64_t get_operand_size() { return g_operand_es; }

// encoder activations: rounded to TF32 when the next convolution is a tcgen05 kernel (default), plain fp32 when it is a
// library convolution ("hybrid" encoder)
static bool g_round_encoder_out = true;
void set_output_rounding(bool on) { g_round_encoder_out = on; }

at::Tensor pack_weights(const at::Tensor& w, int64_t mode) {
  TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.dim() == 4 && w.size(2) == 3 && w.size(3) == 3,
              "pack_weights expects a CUDA fp32 [Co,Ci,3,3] tensor (any strides)");
  c10::cuda::CUDAGuard guard(w.device());
  const int Co = w.size(0), Ci = w.size(1);
  const int rows = mode >= 2 ? Ci : Co, cols = mode >= 2 ? Co : Ci;
  const int rows_pad = (rows + 15) / 16 * 16;
  at::Tensor out = at::empty({(mode & 1) ? 16 : 9, rows_pad, cols}, w.options().dtype(operand_dtype()));
  mine::launch_pack_weights(w.data_ptr<float>(), w.stride(0), w.stride(1), w.stride(2), w.stride(3), Co, Ci, (int)mode,
                            rows_pad, out.data_ptr(), g_operand_es, cur_stream());
  return out;
}

// Batched packer: the operand packs of many weights in one launch per step.  ``pack_plan_create`` allocates the
// (persistent) outputs and uploads the job table once; ``pack_plan_run`` re-packs all of them from the current weights.
std::tuple<std::vector<at::Tensor>, at::Tensor, int64_t> pack_plan_create(const std::vector<at::Tensor>& ws,
                                                                         const std::vector<int64_t>& modes) {
  TORCH_CHECK(ws.size() == modes.size() && !ws.empty(), "pack_plan_create: one mode per weight");
  c10::cuda::CUDAGuard guard(ws[0].device());
  std::vector<at::Tensor> outs;
  std::vector<mine::PackJob> jobs(ws.size());
  int block0 = 0;
  for (size_t i = 0; i < ws.size(); ++i) {
    const at::Tensor& w = ws[i];
    const int mode = (int)modes[i];
    TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.dim() == 4 && w.size(2) == 3 && w.size(3) == 3,
                "pack_plan_create expects CUDA fp32 [Co,Ci,3,3] tensors (any strides)");
    const int Co = w.size(0), Ci = w.size(1);
    const int rows = mode >= 2 ? Ci : Co, cols = mode >= 2 ? Co : Ci;
    const int rows_pad = (rows + 15) / 16 * 16;
    at::Tensor out = at::empty({(mode & 1) ? 16 : 9, rows_pad, cols}, w.options().dtype(operand_dtype()));
    mine::PackJob& j = jobs[i];
    j.w = w.data_ptr<float>(); j.out = out.data_ptr();
    j.so = w.stride(0); j.si = w.stride(1); j.sy = w.stride(2); j.sx = w.stride(3);
    j.Co = Co; j.Ci = Ci; j.mode = mode; j.rows_pad = rows_pad; j.total = (int)out.numel(); j.block0 = block0;
    block0 += (j.total + 255) / 256;
    outs.push_back(out);
  }
  const int64_t bytes = (int64_t)(jobs.size() * sizeof(mine::PackJob));
  at::Tensor host = at::empty({bytes}, at::TensorOptions().dtype(at::kByte));
  memcpy(host.data_ptr(), jobs.data(), (size_t)bytes);
  at::Tensor table = host.to(ws[0].device());
  return std::make_tuple(outs, table, (int64_t)block0);
}

void pack_plan_run(const at::Tensor& table, int64_t nblocks) {
  TORCH_CHECK(table.is_cuda() && table.scalar_type() == at::kByte, "pack_plan_run: job table");
  c10::cuda::CUDAGuard guard(table.device());
  mine::launch_pack_weights_multi(reinterpret_cast<const mine::PackJob*>(table.data_ptr()),
                                  (int)(table.numel() / (int64_t)sizeof(mine::PackJob)), (int)nblocks, g_operand_es,
                                  cur_stream());
}

void wgrad_taps(const at::Tensor& dy, const at::Tensor& x, at::Tensor dw, int64_t Hg, int64_t Wg, int64_t G, int64_t T,
                std::vector<int64_t> tap_y, std::vector<int64_t> tap_x, int64_t dy_stride, std::vector<int64_t> dy_oy,
                std::vector<int64_t> dy_ox, int64_t TH, int64_t TW, int64_t x_stride) {
  check_act_nhwc(dy, "dy"); check_act_nhwc(x, "x"); check_same_type(dy, x, "wgrad_taps");
  TORCH_CHECK(dw.is_cuda() && dw.scalar_type() == at::kFloat && dw.is_contiguous() && dw.dim() == 3, "dw must be fp32 [G*T,Co,Ci]");
  TORCH_CHECK(dw.size(0) == G * T && dw.size(1) == dy.size(3) && dw.size(2) == x.size(3), "dw shape");
  TORCH_CHECK(dy.size(0) == x.size(0), "batch mismatch");
  c10::cuda::CUDAGuard guard(x.device());
  mine::WgradLaunch L{};
  mine::WgradParams& p = L.p;
  p.N = x.size(0); p.Hg = Hg; p.Wg = Wg; p.TH = TH; p.TW = TW; p.KP = TH * TW;
  p.G = G; p.T = T; p.Co = dy.size(3); p.Ci = x.size(3);
  p.es = esize(x);
  p.dy_stride = dy_stride;
  p.x_stride = x_stride;
  for (int g = 0; g < G; ++g) {
    for (int t = 0; t < T; ++t) { p.tap_y[g][t] = (int16_t)tap_y[g * T + t]; p.tap_x[g][t] = (int16_t)tap_x[g * T + t]; }
    p.dy_oy[g] = (int16_t)dy_oy[g]; p.dy_ox[g] = (int16_t)dy_ox[g];
  }
  p.dw = dw.data_ptr<float>();
  L.dy = dy.data_ptr(); L.dyH = dy.size(1); L.dyW = dy.size(2);
  L.x = x.data_ptr(); L.xH = x.size(1); L.xW = x.size(2);
  const char* err = mine::launch_wgrad_taps(L, cur_stream());
  TORCH_CHECK(err == nullptr, "wgrad_taps: ", err ? err : "");
}

// `stats` are this rank's sums; with a fused exchange (ll_ptrs of >= 2 ranks) the kernel reduces them across GPUs in its
// prologue and the reduced vector comes back as the second result (`count` is then the GLOBAL element count).
std::vector<at::Tensor> bn_act_pad_fwd_x(const at::Tensor& y, const at::Tensor& stats, const at::Tensor& gamma,
                                         const at::Tensor& beta, int64_t pad_mode, double count, double eps,
                                         std::vector<int64_t> ll_ptrs, int64_t rank, int64_t cap,
                                         const c10::optional<at::Tensor>& epoch, const c10::optional<at::Tensor>& ticket) {
  check_act_nhwc(y, "y");
  c10::cuda::CUDAGuard guard(y.device());
  const int N = y.size(0), H = y.size(1), W = y.size(2), C = y.size(3);
  TORCH_CHECK((C & (C - 1)) == 0 && C >= 16, "channels must be a power of two >= 16");
  TORCH_CHECK((int64_t)N * (H + 2) * (W + 2) * (C / 8) < (1ll << 31), "tensor too large for 32-bit indexing");
  TORCH_CHECK(opt_f32(stats, "stats") && stats.numel() == 2 * C, "stats must be fp32 [2, C]");
  at::Tensor out = at::empty({N, H + 2, W + 2, C}, y.options());
  FusedX f = make_fused(ll_ptrs, rank, cap, epoch, ticket, 2 * C);
  at::Tensor red = f.on ? at::empty_like(stats) : stats;
  mine::launch_bn_act_pad_fwd(y.data_ptr(), stats.data_ptr<float>(), gamma.data_ptr<float>(), beta.data_ptr<float>(),
                              out.data_ptr(), N, H, W, C, (int)pad_mode, (float)(1.0 / count), (float)eps, esize(y),
                              f.on ? &f.x : nullptr, f.on ? red.data_ptr<float>() : nullptr, cur_stream());
  return {out, red};
}

at::Tensor bn_act_pad_fwd(const at::Tensor& y, const at::Tensor& stats, const at::Tensor& gamma, const at::Tensor& beta,
                          int64_t pad_mode, double count, double eps) {
  return bn_act_pad_fwd_x(y, stats, gamma, beta, pad_mode, count, eps, {}, 0, 0, c10::nullopt, c10::nullopt)[0];
}

static std::vector<at::Tensor> bn_act_bwd_reduce_impl(const at::Tensor& dapad, const at::Tensor& y, const at::Tensor& stats,
                                                      const at::Tensor& gamma, const at::Tensor& beta, int64_t pad_mode,
                                                      double count, double eps, bool want_g) {
  check_act_nhwc(dapad, "dapad"); check_act_nhwc(y, "y"); check_same_type(dapad, y, "bn_act_bwd_reduce");
  c10::cuda::CUDAGuard guard(y.device());
  const int N = y.size(0), H = y.size(1), W = y.size(2), C = y.size(3);
  TORCH_CHECK(dapad.size(0) == N && dapad.size(1) == H + 2 && dapad.size(2) == W + 2 && dapad.size(3) == C, "dapad shape");
  TORCH_CHECK((C & (C - 1)) == 0 && C >= 16, "channels must be a power of two >= 16");
  TORCH_CHECK((int64_t)N * (H + 2) * (W + 2) * (C / 8) < (1ll << 31), "tensor too large for 32-bit indexing");
  at::Tensor g = want_g ? at::empty_like(y) : at::Tensor();
  at::Tensor sums = at::zeros({2, C}, stats.options());
  mine::launch_bn_act_bwd_reduce(dapad.data_ptr(), y.data_ptr(), stats.data_ptr<float>(), gamma.data_ptr<float>(),
                                 beta.data_ptr<float>(), want_g ? g.data_ptr() : nullptr, sums.data_ptr<float>(), N, H, W,
                                 C, (int)pad_mode, (float)(1.0 / count), (float)eps, esize(y), cur_stream());
  return {g, sums};
}

std::vector<at::Tensor> bn_act_bwd_reduce(const at::Tensor& dapad, const at::Tensor& y, const at::Tensor& stats,
                                          const at::Tensor& gamma, const at::Tensor& beta, int64_t pad_mode, double count,
                                          double eps) {
  return bn_act_bwd_reduce_impl(dapad, y, stats, gamma, beta, pad_mode, count, eps, true);
}

// the two BatchNorm backward sums only (the activation-gradient tensor is not materialised: bn_bwd_apply_fused_x)
at::Tensor bn_act_bwd_sums(const at::Tensor& dapad, const at::Tensor& y, const at::Tensor& stats, const at::Tensor& gamma,
                           const at::Tensor& beta, int64_t pad_mode, double count, double eps) {
  return bn_act_bwd_reduce_impl(dapad, y, stats, gamma, beta, pad_mode, count, eps, false)[1];
}

static std::vector<at::Tensor> bn_bwd_apply_impl(const at::Tensor& g, const at::Tensor& y, const at::Tensor& stats,
                                                 const at::Tensor& gamma, const at::Tensor& sums, int64_t planes_per_image,
                                                 bool want_shared, bool want_plane_bias, double count, double eps,
                                                 std::vector<int64_t> ll_ptrs, int64_t rank, int64_t cap,
                                                 const c10::optional<at::Tensor>& epoch,
                                                 const c10::optional<at::Tensor>& ticket, const at::Tensor* beta,
                                                 int64_t pad_mode) {
  check_act_nhwc(g, "g"); check_act_nhwc(y, "y"); check_same_type(g, y, "bn_bwd_apply");
  c10::cuda::CUDAGuard guard(y.device());
  const int N = y.size(0), H = y.size(1), W = y.size(2), C = y.size(3);
  if (beta) {
    TORCH_CHECK(g.size(0) == N && g.size(1) == H + 2 && g.size(2) == W + 2 && g.size(3) == C, "dapad shape");
    TORCH_CHECK(beta->is_cuda() && beta->scalar_type() == at::kFloat && beta->numel() == C, "beta");
  } else {
    TORCH_CHECK(g.sizes() == y.sizes(), "g shape");
  }
  const int S = planes_per_image, B = N / S;
  FusedX f = make_fused(ll_ptrs, rank, cap, epoch, ticket, 2 * C);
  at::Tensor dy = at::empty_like(y);
  at::Tensor dshared = want_shared ? at::empty({B, H, W, C}, stats.options()) : at::Tensor();
  at::Tensor dpb = want_plane_bias ? at::zeros({N, C}, stats.options()) : at::Tensor();
  mine::launch_bn_bwd_apply(g.data_ptr(), y.data_ptr(), stats.data_ptr<float>(), gamma.data_ptr<float>(),
                            sums.data_ptr<float>(), dy.data_ptr(), want_shared ? dshared.data_ptr<float>() : nullptr,
                            want_plane_bias ? dpb.data_ptr<float>() : nullptr, B, S, H, W, C, (float)(1.0 / count),
                            (float)eps, esize(y), f.on ? &f.x : nullptr, beta ? beta->data_ptr<float>() : nullptr,
                            (int)pad_mode, cur_stream());
  return {dy, dshared, dpb};
}

std::vector<at::Tensor> bn_bwd_apply_x(const at::Tensor& g, const at::Tensor& y, const at::Tensor& stats,
                                       const at::Tensor& gamma, const at::Tensor& sums, int64_t planes_per_image,
                                       bool want_shared, bool want_plane_bias, double count, double eps,
                                       std::vector<int64_t> ll_ptrs, int64_t rank, int64_t cap,
                                       const c10::optional<at::Tensor>& epoch, const c10::optional<at::Tensor>& ticket) {
  return bn_bwd_apply_impl(g, y, stats, gamma, sums, planes_per_image, want_shared, want_plane_bias, count, eps, ll_ptrs,
                           rank, cap, epoch, ticket, nullptr, 0);
}

std::vector<at::Tensor> bn_bwd_apply(const at::Tensor& g, const at::Tensor& y, const at::Tensor& stats,
                                     const at::Tensor& gamma, const at::Tensor& sums, int64_t planes_per_image,
                                     bool want_shared, bool want_plane_bias, double count, double eps) {
  return bn_bwd_apply_x(g, y, stats, gamma, sums, planes_per_image, want_shared, want_plane_bias, count, eps, {}, 0, 0,
                        c10::nullopt, c10::nullopt);
}

// BatchNorm + ELU + pad backward straight from the PADDED upstream gradient (pad adjoint and ELU' recomputed in-kernel)
std::vector<at::Tensor> bn_bwd_apply_fused_x(const at::Tensor& dapad, const at::Tensor& y, const at::Tensor& stats,
                                             const at::Tensor& gamma, const at::Tensor& beta, const at::Tensor& sums,
                                             int64_t planes_per_image, bool want_shared, bool want_plane_bias, double count,
                                             double eps, int64_t pad_mode, std::vector<int64_t> ll_ptrs, int64_t rank,
                                             int64_t cap, const c10::optional<at::Tensor>& epoch,
                                             const c10::optional<at::Tensor>& ticket) {
  return bn_bwd_apply_impl(dapad, y, stats, gamma, sums, planes_per_image, want_shared, want_plane_bias, count, eps, ll_ptrs,
                           rank, cap, epoch, ticket, &beta, pad_mode);
}

std::vector<at::Tensor> bn_bwd_apply_fused(const at::Tensor& dapad, const at::Tensor& y, const at::Tensor& stats,
                                           const at::Tensor& gamma, const at::Tensor& beta, const at::Tensor& sums,
                                           int64_t planes_per_image, bool want_shared, bool want_plane_bias, double count,
                                           double eps, int64_t pad_mode) {
  return bn_bwd_apply_fused_x(dapad, y, stats, gamma, beta, sums, planes_per_image, want_shared, want_plane_bias, count, eps,
                              pad_mode, {}, 0, 0, c10::nullopt, c10::nullopt);
}

std::vector<at::Tensor> head_bwd(const at::Tensor& g_mpi, const at::Tensor& mpi, const at::Tensor& sign, bool use_alpha) {
  TORCH_CHECK(g_mpi.is_cuda() && g_mpi.scalar_type() == at::kFloat && g_mpi.is_contiguous(), "g_mpi");
  TORCH_CHECK(mpi.is_contiguous() && mpi.scalar_type() == at::kFloat && sign.scalar_type() == at::kChar, "mpi/sign");
  c10::cuda::CUDAGuard guard(mpi.device());
  const int64_t npix = mpi.numel() / 4;
  auto sizes = sign.sizes().vec();           // [N, H, W]
  at::Tensor dz = at::empty({sizes[0], sizes[1], sizes[2], 16}, mpi.options().dtype(operand_dtype()));
  at::Tensor dbias = at::zeros({4}, mpi.options());
  mine::launch_head_bwd(g_mpi.data_ptr<float>(), mpi.data_ptr<float>(), sign.data_ptr<int8_t>(), dz.data_ptr(),
                        dbias.data_ptr<float>(), (size_t)npix, use_alpha ? 1 : 0, g_operand_es, cur_stream());
  return {dz, dbias};
}

std::vector<at::Tensor> head_conv_direct(const at::Tensor& apad, const at::Tensor& wpk, const at::Tensor& bias,
                                         bool use_alpha) {
  check_act_nhwc(apad, "apad");
  TORCH_CHECK(apad.scalar_type() == at::kBFloat16, "head_conv_direct: bf16 operands only");
  const int64_t N = apad.size(0), H = apad.size(1) - 2, W = apad.size(2) - 2, C = apad.size(3);
  TORCH_CHECK(C == 16 || C == 32, "head_conv_direct handles 16 or 32 input channels");
  TORCH_CHECK(wpk.is_cuda() && wpk.scalar_type() == at::kFloat && wpk.is_contiguous() && wpk.numel() == 9 * C * 4,
              "wpk must be contiguous fp32 [9, C, 4]");
  TORCH_CHECK(bias.is_cuda() && bias.scalar_type() == at::kFloat && bias.is_contiguous() && bias.numel() == 4, "bias");
  c10::cuda::CUDAGuard guard(apad.device());
  at::Tensor mpi = at::empty({N, H, W, 4}, apad.options().dtype(at::kFloat));
  at::Tensor sign = at::empty({N, H, W}, apad.options().dtype(at::kChar));
  const char* err = mine::launch_head_conv_direct(apad.data_ptr(), wpk.data_ptr<float>(), bias.data_ptr<float>(),
                                                  mpi.data_ptr<float>(), sign.data_ptr<int8_t>(), (int)N, (int)H, (int)W,
                                                  (int)C, use_alpha ? 1 : 0, cur_stream());
  TORCH_CHECK(err == nullptr, err ? err : "");
  return {mpi, sign};
}

inline void check_channels(const at::Tensor& y) {
  const int64_t C = y.size(3);
  TORCH_CHECK((C & (C - 1)) == 0 && C >= 16 && C <= 2048, "channels must be a power of two in [16, 2048]");
  TORCH_CHECK(y.numel() / 8 < (1ll << 31), "tensor too large for 32-bit indexing");
}

std::vector<at::Tensor> bn_res_act_fwd_x(const at::Tensor& y, const at::Tensor& stats, const at::Tensor& gamma,
                                         const at::Tensor& beta, const c10::optional<at::Tensor>& residual, double slope,
                                         double count, double eps, std::vector<int64_t> ll_ptrs, int64_t rank, int64_t cap,
                                         const c10::optional<at::Tensor>& epoch, const c10::optional<at::Tensor>& ticket) {
  check_act_nhwc(y, "y"); check_channels(y);
  TORCH_CHECK(stats.numel() == 2 * y.size(3) && opt_f32(stats, "stats") && opt_f32(gamma, "gamma") && opt_f32(beta, "beta"),
              "stats/gamma/beta");
  const void* res = nullptr;
  if (residual.has_value() && residual->defined()) {
    check_act_nhwc(*residual, "residual"); check_same_type(*residual, y, "bn_res_act_fwd");
    TORCH_CHECK(residual->sizes() == y.sizes(), "residual shape");
    res = residual->data_ptr();
  }
  c10::cuda::CUDAGuard guard(y.device());
  at::Tensor out = at::empty_like(y);
  FusedX f = make_fused(ll_ptrs, rank, cap, epoch, ticket, 2 * y.size(3));
  at::Tensor red = f.on ? at::empty_like(stats) : stats;
  mine::launch_bn_res_act_fwd(y.data_ptr(), stats.data_ptr<float>(), gamma.data_ptr<float>(), beta.data_ptr<float>(), res,
                              out.data_ptr(), (size_t)(y.numel() / y.size(3)), (int)y.size(3), (float)slope,
                              (float)(1.0 / count), (float)eps, esize(y), f.on ? &f.x : nullptr,
                              f.on ? red.data_ptr<float>() : nullptr, g_round_encoder_out ? 1 : 0, cur_stream());
  return {out, red};
}

at::Tensor bn_res_act_fwd(const at::Tensor& y, const at::Tensor& stats, const at::Tensor& gamma, const at::Tensor& beta,
                          const c10::optional<at::Tensor>& residual, double slope, double count, double eps) {
  return bn_res_act_fwd_x(y, stats, gamma, beta, residual, slope, count, eps, {}, 0, 0, c10::nullopt, c10::nullopt)[0];
}

// engine.deterministic: bitwise reproducible BatchNorm reductions in the encoder kernels (two-level fixed-order sums
// instead of fp32 atomics; ~10 us more per layer and direction)
static bool g_deterministic = false;
void set_deterministic(bool on) { g_deterministic = on; }
bool get_deterministic() { return g_deterministic; }

// ticket counters of the reproducible reductions (encoder_elem.cu::publish_sums), one zeroed array per device; every
// launch leaves them at zero
static unsigned* reduce_tickets(const at::Device& dev) {
  static std::vector<at::Tensor> tickets(64);
  at::Tensor& t = tickets[dev.index() < 0 ? 0 : dev.index()];
  if (!t.defined()) t = at::zeros({64}, at::TensorOptions().device(dev).dtype(at::kInt));
  return reinterpret_cast<unsigned*>(t.data_ptr<int>());
}

std::vector<at::Tensor> bn_res_act_bwd_reduce(const at::Tensor& dout, const at::Tensor& out, const at::Tensor& y,
                                              const at::Tensor& stats, const at::Tensor& gamma, const at::Tensor& beta,
                                              double slope, double count, double eps) {
  check_act_nhwc(dout, "dout"); check_act_nhwc(out, "out"); check_act_nhwc(y, "y"); check_channels(y);
  TORCH_CHECK(dout.sizes() == y.sizes() && out.sizes() == y.sizes(), "shape mismatch");
  check_same_type(dout, y, "bn_res_act_bwd_reduce"); check_same_type(out, y, "bn_res_act_bwd_reduce");
  TORCH_CHECK(stats.numel() == 2 * y.size(3) && opt_f32(stats, "stats"), "stats");
  (void)gamma; (void)beta;                     // the reduction needs mean / invstd only; kept for a uniform signature
  c10::cuda::CUDAGuard guard(y.device());
  at::Tensor g = at::empty_like(y);
  const int64_t C = y.size(3), npix = y.numel() / C;
  TORCH_CHECK(C <= 2048, "at most 2048 channels");
  const bool det = g_deterministic;
  at::Tensor sums = det ? at::empty({2, C}, stats.options()) : at::zeros({2, C}, stats.options());
  at::Tensor scratch = det ? at::empty({(int64_t)mine::reduce_scratch_floats((size_t)npix, (int)C)}, stats.options())
                           : at::Tensor();
  mine::launch_bn_res_act_bwd_reduce(dout.data_ptr(), out.data_ptr(), y.data_ptr(), stats.data_ptr<float>(), g.data_ptr(),
                                     sums.data_ptr<float>(), (size_t)npix, (int)C, (float)slope, (float)(1.0 / count),
                                     (float)eps, esize(y), det ? scratch.data_ptr<float>() : nullptr,
                                     det ? reduce_tickets(y.device()) : nullptr, cur_stream());
  return {g, sums};
}

at::Tensor channel_stats(const at::Tensor& y) {
  check_act_nhwc(y, "y"); check_channels(y);
  c10::cuda::CUDAGuard guard(y.device());
  const int64_t C = y.size(3), npix = y.numel() / C;
  TORCH_CHECK(C <= 2048, "at most 2048 channels");
  const bool det = g_deterministic;
  at::Tensor sums = det ? at::empty({2, C}, y.options().dtype(at::kFloat)) : at::zeros({2, C}, y.options().dtype(at::kFloat));
  at::Tensor scratch = det ? at::empty({(int64_t)mine::reduce_scratch_floats((size_t)npix, (int)C)}, sums.options())
                           : at::Tensor();
  mine::launch_channel_stats(y.data_ptr(), sums.data_ptr<float>(), (size_t)npix, (int)C, esize(y),
                             det ? scratch.data_ptr<float>() : nullptr, det ? reduce_tickets(y.device()) : nullptr,
                             cur_stream());
  return sums;
}

void conv_taps_splitk(const at::Tensor& x, const at::Tensor& wpack, at::Tensor out32, int64_t Hg, int64_t Wg, int64_t T,
                      std::vector<int64_t> tap_y, std::vector<int64_t> tap_x, int64_t in_stride, int64_t Co, int64_t TH,
                      int64_t TW, int64_t ksplit) {
  check_act_nhwc(x, "x");
  TORCH_CHECK(wpack.is_cuda() && wpack.scalar_type() == x.scalar_type() && wpack.is_contiguous() && wpack.dim() == 3 &&
                  wpack.size(0) == T && wpack.size(2) == x.size(3), "wpack must be [T, rows, Ci] in the dtype of x");
  TORCH_CHECK((int64_t)tap_y.size() == T && (int64_t)tap_x.size() == T, "tap table size");
  TORCH_CHECK(out32.is_cuda() && out32.scalar_type() == at::kFloat && out32.is_contiguous() && out32.dim() == 4 &&
                  out32.size(0) == x.size(0) && out32.size(1) == Hg && out32.size(2) == Wg && out32.size(3) == Co,
              "out32 must be a zero-initialised fp32 [N, Hg, Wg, Co] tensor");
  c10::cuda::CUDAGuard guard(x.device());
  std::vector<int> ty(tap_y.begin(), tap_y.end()), tx(tap_x.begin(), tap_x.end());
  const char* err = mine::launch_conv_splitk(x.data_ptr(), (int)x.size(0), (int)x.size(1), (int)x.size(2), (int)x.size(3),
                                             wpack.data_ptr(), (int)wpack.size(1), (int)T, ty.data(), tx.data(),
                                             (int)in_stride, out32.data_ptr<float>(), (int)Hg, (int)Wg, (int)Co, (int)TH,
                                             (int)TW, (int)ksplit, esize(x), cur_stream());
  TORCH_CHECK(err == nullptr, "conv_taps_splitk: ", err ? err : "");
}

// fp32 partial-sum tensor -> {bf16 tensor, stats [2, C] (empty when not wanted)}
std::vector<at::Tensor> splitk_finalize(const at::Tensor& acc, bool want_stats) {
  TORCH_CHECK(acc.is_cuda() && acc.scalar_type() == at::kFloat && acc.is_contiguous() && acc.dim() == 4, "acc");
  const int64_t C = acc.size(3);
  TORCH_CHECK((C & (C - 1)) == 0 && C >= 16 && C <= 2048, "channels must be a power of two in [16, 2048]");
  TORCH_CHECK(acc.numel() / 8 < (1ll << 31), "tensor too large for 32-bit indexing");
  c10::cuda::CUDAGuard guard(acc.device());
  at::Tensor stats = want_stats ? at::zeros({2, C}, acc.options()) : at::empty({0}, acc.options());
  if (g_operand_es == 4) {       // fp32 operands: the partial-sum tensor IS the activation; only the BN sums are left
    if (want_stats) {
      const bool det = g_deterministic;       // ``stats`` is zero filled above: fine for both variants
      at::Tensor scratch = det ? at::empty({(int64_t)mine::reduce_scratch_floats((size_t)(acc.numel() / C), (int)C)}, acc.options())
                               : at::Tensor();
      mine::launch_channel_stats(acc.data_ptr(), stats.data_ptr<float>(), (size_t)(acc.numel() / C), (int)C, 4,
                                 det ? scratch.data_ptr<float>() : nullptr, det ? reduce_tickets(acc.device()) : nullptr,
                                 cur_stream());
    }
    return {acc, stats};
  }
  at::Tensor y = at::empty(acc.sizes(), acc.options().dtype(at::kBFloat16));
  mine::launch_splitk_finalize(acc.data_ptr<float>(), y.data_ptr(), want_stats ? stats.data_ptr<float>() : nullptr,
                               (size_t)(acc.numel() / C), (int)C, cur_stream());
  return {y, stats};
}

void bn_update_running(const at::Tensor& stats, at::Tensor running_mean, at::Tensor running_var,
                       at::Tensor num_batches_tracked, double count, double momentum) {
  const int64_t C = running_mean.numel();
  TORCH_CHECK(opt_f32(stats, "stats") && stats.numel() == 2 * C, "stats must be fp32 [2, C]");
  TORCH_CHECK(opt_f32(running_mean, "running_mean") && opt_f32(running_var, "running_var") && running_var.numel() == C,
              "running statistics must be contiguous CUDA fp32");
  TORCH_CHECK(num_batches_tracked.is_cuda() && num_batches_tracked.scalar_type() == at::kLong, "num_batches_tracked");
  c10::cuda::CUDAGuard guard(stats.device());
  mine::launch_bn_update_running(stats.data_ptr<float>(), running_mean.data_ptr<float>(), running_var.data_ptr<float>(),
                                 reinterpret_cast<long long*>(num_batches_tracked.data_ptr<int64_t>()), (int)C,
                                 (float)count, (float)momentum, cur_stream());
}

at::Tensor pad_reflect_nhwc(const at::Tensor& x) {
  check_act_nhwc(x, "x");
  TORCH_CHECK(x.size(1) >= 2 && x.size(2) >= 2, "reflection pad needs at least 2 x 2 pixels");
  TORCH_CHECK(x.numel() / 8 + (int64_t)x.size(0) * (2 * (x.size(1) + x.size(2)) + 4) * (x.size(3) / 8) < (1ll << 31),
              "tensor too large for 32-bit indexing");
  c10::cuda::CUDAGuard guard(x.device());
  at::Tensor out = at::empty({x.size(0), x.size(1) + 2, x.size(2) + 2, x.size(3)}, x.options());
  mine::launch_pad_reflect_nhwc(x.data_ptr(), out.data_ptr(), (int)x.size(0), (int)x.size(1), (int)x.size(2),
                                (int)x.size(3), esize(x), cur_stream());
  return out;
}

at::Tensor pad_reflect_nhwc_bwd(const at::Tensor& gp) {
  check_act_nhwc(gp, "gp");
  TORCH_CHECK(gp.size(1) >= 4 && gp.size(2) >= 4, "padded gradient too small");
  TORCH_CHECK(gp.numel() / 8 < (1ll << 31), "tensor too large for 32-bit indexing");
  c10::cuda::CUDAGuard guard(gp.device());
  at::Tensor gx = at::empty({gp.size(0), gp.size(1) - 2, gp.size(2) - 2, gp.size(3)}, gp.options());
  mine::launch_pad_reflect_nhwc_bwd(gp.data_ptr(), gx.data_ptr(), (int)gp.size(0), (int)gp.size(1) - 2,
                                    (int)gp.size(2) - 2, (int)gp.size(3), esize(gp), cur_stream());
  return gx;
}

// deferred form of bn_update_running: all layers of a step in one launch per 48 layers
void bn_update_running_multi(const std::vector<at::Tensor>& stats, std::vector<at::Tensor> running_mean,
                             std::vector<at::Tensor> running_var, std::vector<at::Tensor> num_batches_tracked,
                             const std::vector<double>& count, const std::vector<double>& momentum) {
  const size_t n = stats.size();
  TORCH_CHECK(running_mean.size() == n && running_var.size() == n && num_batches_tracked.size() == n && count.size() == n &&
              momentum.size() == n, "bn_update_running_multi: list lengths differ");
  if (n == 0) return;
  c10::cuda::CUDAGuard guard(stats[0].device());
  std::vector<const float*> ps(n);
  std::vector<float*> pm(n), pv(n);
  std::vector<long long*> pn(n);
  std::vector<int> cs(n);
  std::vector<float> cnt(n), mom(n);
  for (size_t i = 0; i < n; ++i) {
    const int64_t C = running_mean[i].numel();
    TORCH_CHECK(opt_f32(stats[i], "stats") && stats[i].numel() == 2 * C, "stats must be fp32 [2, C]");
    TORCH_CHECK(opt_f32(running_mean[i], "running_mean") && opt_f32(running_var[i], "running_var") &&
                running_var[i].numel() == C, "running statistics must be contiguous CUDA fp32");
    TORCH_CHECK(num_batches_tracked[i].is_cuda() && num_batches_tracked[i].scalar_type() == at::kLong, "num_batches_tracked");
    ps[i] = stats[i].data_ptr<float>(); pm[i] = running_mean[i].data_ptr<float>(); pv[i] = running_var[i].data_ptr<float>();
    pn[i] = reinterpret_cast<long long*>(num_batches_tracked[i].data_ptr<int64_t>());
    cs[i] = (int)C; cnt[i] = (float)count[i]; mom[i] = (float)momentum[i];
  }
  mine::launch_bn_update_running_multi((int)n, ps.data(), pm.data(), pv.data(), pn.data(), cs.data(), cnt.data(),
                                       mom.data(), cur_stream());
}

// host-side evaluation of the launcher-made multiply-high division (conv_engine.h::make_fastdiv; the device code is
// ``__umulhi(n, mul) >> shr``): lets the CPU test tier check the constants for every divisor the launchers can produce
int64_t fastdiv_host(int64_t n, int64_t d) {
  TORCH_CHECK(n >= 0 && n < (1ll << 31) && d >= 1 && d < (1ll << 31), "fastdiv_host: 0 <= n < 2^31, 1 <= d < 2^31");
  const mine::FastDiv f = mine::make_fastdiv((int)d);
  if (f.d == 1u) return n;
  return (int64_t)((uint32_t)(((uint64_t)(uint32_t)n * (uint64_t)f.mul) >> 32) >> f.shr);
}

}  // namespace

void register_conv(pybind11::module_& m) {
  m.def("set_operand_size", &set_operand_size);
  m.def("get_operand_size", &get_operand_size);
  m.def("set_output_rounding", &set_output_rounding);
  m.def("fastdiv_host", &fastdiv_host);
  m.def("set_deterministic", &set_deterministic);
  m.def("get_deterministic", &get_deterministic);
  m.def("conv_taps", &conv_taps);
  m.def("wgrad_taps", &wgrad_taps);
  m.def("pack_weights", &pack_weights);
  m.def("pack_plan_create", &pack_plan_create);
  m.def("pack_plan_run", &pack_plan_run);
  m.def("bn_act_pad_fwd", &bn_act_pad_fwd);
  m.def("bn_act_pad_fwd_x", &bn_act_pad_fwd_x);
  m.def("bn_bwd_apply_x", &bn_bwd_apply_x);
  m.def("bn_res_act_fwd_x", &bn_res_act_fwd_x);
  m.def("bn_act_bwd_reduce", &bn_act_bwd_reduce);
  m.def("bn_bwd_apply", &bn_bwd_apply);
  m.def("bn_act_bwd_sums", &bn_act_bwd_sums);
  m.def("bn_bwd_apply_fused", &bn_bwd_apply_fused);
  m.def("bn_bwd_apply_fused_x", &bn_bwd_apply_fused_x);
  m.def("head_bwd", &head_bwd);
  m.def("bn_res_act_fwd", &bn_res_act_fwd);
  m.def("bn_res_act_bwd_reduce", &bn_res_act_bwd_reduce);
  m.def("channel_stats", &channel_stats);
  m.def("head_conv_direct", &head_conv_direct);
  m.def("bn_update_running", &bn_update_running);
  m.def("bn_update_running_multi", &bn_update_running_multi);
  m.def("pad_reflect_nhwc", &pad_reflect_nhwc);
  m.def("pad_reflect_nhwc_bwd", &pad_reflect_nhwc_bwd);
  m.def("conv_taps_splitk", &conv_taps_splitk);
  m.def("splitk_finalize", &splitk_finalize);
}
