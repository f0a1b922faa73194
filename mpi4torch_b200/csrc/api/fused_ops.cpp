// Tensor-core ops exposed to Python: the tcgen05 GEMM and the fused
// Allreduce->GEMM linear forward (BASELINE.json config "4096x4096 linear layer:
// Allreduce(params)->GEMM fused").  The reference has no fused op and no GPU
// kernel at all; its equivalent is Allreduce + `/ size` + matmul
// as three library calls (reference examples/simple_linear_regression.py:29).
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>
#include <torch/torch.h>

#include "../runtime/cuda_backend.h"
#include "../runtime/world.h"

namespace m4t {

namespace {

using torch::Tensor;

CudaBackend& backend() {
  World& w = World::instance();
  TORCH_CHECK(w.cuda_ready(), "mpi4torch_b200: the CUDA backend is not initialised");
  return *w.cuda();
}

void check_2d_bf16(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.dim() == 2 && t.scalar_type() == at::kBFloat16 && t.stride(1) == 1,
              "mpi4torch_b200: ", name, " must be a 2-d bf16 CUDA tensor with unit inner stride");
}

// y = x @ w^T on the tcgen05 path.
Tensor gemm_bf16_tn(const Tensor& x, const Tensor& w) {
  check_2d_bf16(x, "x");
  check_2d_bf16(w, "w");
  TORCH_CHECK(x.size(1) == w.size(1), "mpi4torch_b200: inner dimensions differ");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({x.size(0), w.size(0)}, x.options());
  if (y.numel() == 0) return y;
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  backend().gemm_bf16_tn(x.data_ptr(), w.data_ptr(), y.data_ptr(), x.size(0), w.size(0), x.size(1), x.stride(0),
                         w.stride(0), y.stride(0), c10::cuda::getCurrentCUDAStream(x.device().index()).stream());
  return y;
}

bool gemm_bf16_tn_ok(const Tensor& x, const Tensor& w) {
  if (!(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.scalar_type() == at::kBFloat16 &&
        w.scalar_type() == at::kBFloat16 && x.stride(1) == 1 && w.stride(1) == 1 && x.size(1) == w.size(1)))
    return false;
  return World::instance().cuda_ready() &&
         gemm_bf16_tn_supported(x.size(0), w.size(0), x.size(1), x.data_ptr(), w.data_ptr(), x.data_ptr(), x.stride(0),
                                w.stride(0), w.size(0));
}

bool allreduce_linear_supported(const Tensor& x, const Tensor& w) {
  if (!World::instance().cuda_ready()) return false;
  if (!(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.scalar_type() == at::kBFloat16 &&
        w.scalar_type() == at::kBFloat16 && x.is_contiguous() && w.is_contiguous() && x.size(1) == w.size(1)))
    return false;
  return backend().fused_linear_available(w.size(0), w.size(1));
}

// (y, w_avg) with y = x @ (scale * sum_ranks w)^T, one kernel.  w_avg aliases
// symmetric-heap memory that stays valid until the call after next.
std::tuple<Tensor, Tensor> allreduce_linear_fused(const Tensor& x, const Tensor& w, double scale) {
  TORCH_CHECK(allreduce_linear_supported(x, w), "mpi4torch_b200: fused Allreduce->GEMM does not support these tensors");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({x.size(0), w.size(0)}, x.options());
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  const void* wavg = backend().fused_allreduce_linear(
      x.data_ptr(), w.data_ptr(), y.data_ptr(), x.size(0), w.size(0), x.size(1), x.stride(0), y.stride(0),
      static_cast<float>(scale), c10::cuda::getCurrentCUDAStream(x.device().index()).stream());
  Tensor w_avg = torch::from_blob(const_cast<void*>(wavg), {w.size(0), w.size(1)}, w.options());
  return {y, w_avg};
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(mpi4torch_b200, m) {
  m.def("gemm_bf16_tn(Tensor x, Tensor w) -> Tensor", &gemm_bf16_tn);
  m.def("gemm_bf16_tn_supported(Tensor x, Tensor w) -> bool", &gemm_bf16_tn_ok);
  m.def("allreduce_linear_supported(Tensor x, Tensor w) -> bool", &allreduce_linear_supported);
  m.def("allreduce_linear_fused(Tensor x, Tensor w, float scale) -> (Tensor, Tensor)", &allreduce_linear_fused);
}

}  // namespace m4t
