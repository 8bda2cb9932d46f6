// Bindings of the NVLink all-reduce kernels.  Peer pointers come from torch's symmetric-memory
// rendezvous (handle exchange only); the reductions themselves are ours (comm.cu).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "kernels.h"

namespace {

mine::PeerTable table_from(const std::vector<int64_t>& ptrs) {
  TORCH_CHECK(ptrs.size() <= 16, "at most 16 peers");
  mine::PeerTable t{};
  for (size_t i = 0; i < ptrs.size(); ++i) t.ptr[i] = reinterpret_cast<void*>(ptrs[i]);
  return t;
}

void allreduce_small(at::Tensor inout, std::vector<int64_t> data_ptrs, std::vector<int64_t> flag_ptrs, int64_t rank,
                     int64_t cap, at::Tensor epoch) {
  TORCH_CHECK(epoch.is_cuda() && epoch.scalar_type() == at::kInt && epoch.numel() >= 1, "epoch must be a CUDA int32 tensor");
  TORCH_CHECK(inout.is_cuda() && inout.scalar_type() == at::kFloat && inout.is_contiguous(), "inout must be contiguous fp32");
  TORCH_CHECK(inout.numel() <= cap, "vector larger than the one-shot slot");
  c10::cuda::CUDAGuard guard(inout.device());
  mine::launch_allreduce_small(inout.data_ptr<float>(), (int)inout.numel(), table_from(data_ptrs), table_from(flag_ptrs),
                               (int)rank, (int)data_ptrs.size(), (int)cap, reinterpret_cast<uint32_t*>(epoch.data_ptr<int>()),
                               at::cuda::getCurrentCUDAStream().stream());
}

void allreduce_small_ll(at::Tensor inout, std::vector<int64_t> ll_ptrs, int64_t rank, int64_t cap, at::Tensor epoch) {
  TORCH_CHECK(inout.is_cuda() && inout.scalar_type() == at::kFloat && inout.is_contiguous(), "inout must be contiguous fp32");
  TORCH_CHECK(inout.numel() <= cap, "vector larger than the LL slot");
  TORCH_CHECK(epoch.is_cuda() && epoch.scalar_type() == at::kInt, "epoch must be a CUDA int32 tensor");
  c10::cuda::CUDAGuard guard(inout.device());
  mine::launch_allreduce_small_ll(inout.data_ptr<float>(), (int)inout.numel(), table_from(ll_ptrs), (int)rank,
                                  (int)ll_ptrs.size(), (int)cap, reinterpret_cast<uint32_t*>(epoch.data_ptr<int>()),
                                  at::cuda::getCurrentCUDAStream().stream());
}

void allreduce_mean(std::vector<int64_t> arena_ptrs, std::vector<int64_t> flag_ptrs, int64_t mc_ptr, int64_t lo, int64_t hi,
                    int64_t rank, at::Tensor epochs, int64_t blocks) {
  TORCH_CHECK(lo % 4 == 0 && hi % 4 == 0, "bucket bounds must be multiples of 4 floats");
  TORCH_CHECK(epochs.is_cuda() && epochs.scalar_type() == at::kInt && epochs.numel() >= blocks, "epochs: CUDA int32 [blocks]");
  c10::cuda::CUDAGuard guard(epochs.device());
  mine::launch_allreduce_mean(table_from(arena_ptrs), table_from(flag_ptrs), reinterpret_cast<float*>(mc_ptr), lo, hi,
                              (int)rank, (int)arena_ptrs.size(), reinterpret_cast<uint32_t*>(epochs.data_ptr<int>()), (int)blocks,
                              at::cuda::getCurrentCUDAStream().stream());
}

}  // namespace

void register_comm(pybind11::module_& m) {
  m.def("allreduce_small", &allreduce_small);
  m.def("allreduce_small_ll", &allreduce_small_ll);
  m.def("allreduce_mean", &allreduce_mean);
}
