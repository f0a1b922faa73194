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
#include "communicator.h"

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

// Same GEMM on CTA pairs (cta_group::2); exposed separately for A/B testing.
Tensor gemm_bf16_tn_2cta(const Tensor& x, const Tensor& w) {
  check_2d_bf16(x, "x");
  check_2d_bf16(w, "w");
  TORCH_CHECK(x.size(1) == w.size(1), "mpi4torch_b200: inner dimensions differ");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor y = at::empty({x.size(0), w.size(0)}, x.options());
  if (y.numel() == 0) return y;
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  launch_gemm_bf16_tn_2cta(x.data_ptr(), w.data_ptr(), y.data_ptr(), x.size(0), w.size(0), x.size(1), x.stride(0),
                           w.stride(0), y.stride(0), backend().device_comm().sm_count,
                           c10::cuda::getCurrentCUDAStream(x.device().index()).stream());
  return y;
}

// y = a @ b with both operands row-major (dgrad: gy[M,N] @ W[N,K]); the weight is read in place through an
// MN-major tcgen05 operand, no transposed copy.
bool gemm_bf16_nn_ok(const Tensor& a, const Tensor& b) {
  if (!(a.is_cuda() && b.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.scalar_type() == at::kBFloat16 &&
        b.scalar_type() == at::kBFloat16 && a.stride(1) == 1 && b.stride(1) == 1 && a.size(1) == b.size(0)))
    return false;
  return World::instance().cuda_ready() &&
         gemm_bf16_nn_supported(a.size(0), b.size(1), a.size(1), a.data_ptr(), b.data_ptr(), a.data_ptr(), a.stride(0),
                                b.stride(0), b.size(1));
}

Tensor gemm_bf16_nn(const Tensor& a, const Tensor& b) {
  check_2d_bf16(a, "a");
  check_2d_bf16(b, "b");
  TORCH_CHECK(a.size(1) == b.size(0), "mpi4torch_b200: inner dimensions differ");
  c10::cuda::CUDAGuard guard(a.device());
  Tensor y = at::empty({a.size(0), b.size(1)}, a.options());
  if (y.numel() == 0) return y;
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  launch_gemm_bf16_nn_2cta(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.size(0), b.size(1), a.size(1), a.stride(0),
                           b.stride(0), y.stride(0), backend().device_comm().sm_count,
                           c10::cuda::getCurrentCUDAStream(a.device().index()).stream());
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

// Training forward of the data-parallel linear layer with everything fused:
//   W_avg = scale * sum_ranks W          (inside the GEMM kernel when NVLS is up)
//   y     = x @ W_avg^T                  (tcgen05, never stored)
//   dy    = grad_scale * (y - target)    (GEMM epilogue)
//   loss  = loss_scale * sum((y-target)^2)  (GEMM epilogue, one atomic per warp per tile)
// Returns (dy, loss[1] fp32, W_avg).
std::tuple<Tensor, Tensor, Tensor> linear_mse_forward(const Tensor& x, const Tensor& w, const Tensor& target, double scale,
                                                      double loss_scale, double grad_scale, bool allow_fused) {
  check_2d_bf16(x, "x");
  check_2d_bf16(w, "w");
  check_2d_bf16(target, "target");
  TORCH_CHECK(x.size(1) == w.size(1) && target.size(0) == x.size(0) && target.size(1) == w.size(0),
              "mpi4torch_b200: linear_mse_forward shape mismatch");
  c10::cuda::CUDAGuard guard(x.device());
  cudaStream_t stream = c10::cuda::getCurrentCUDAStream(x.device().index()).stream();
  Tensor dy = at::empty({x.size(0), w.size(0)}, x.options());
  Tensor loss = at::zeros({1}, x.options().dtype(at::kFloat));
  MseEpilogue mse;
  mse.target = target.data_ptr();
  mse.ldt = target.stride(0);
  mse.loss_acc = loss.data_ptr<float>();
  mse.loss_scale = static_cast<float>(loss_scale);
  mse.grad_scale = static_cast<float>(grad_scale);
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  CudaBackend& be = backend();
  Tensor w_avg;
  if (allow_fused && be.size() > 1 && x.is_contiguous() && w.is_contiguous() && be.fused_linear_available(w.size(0), w.size(1))) {
    const void* wavg = be.fused_allreduce_linear(x.data_ptr(), w.data_ptr(), dy.data_ptr(), x.size(0), w.size(0), x.size(1),
                                                 x.stride(0), dy.stride(0), static_cast<float>(scale), stream, &mse);
    w_avg = torch::from_blob(const_cast<void*>(wavg), {w.size(0), w.size(1)}, w.options());
  } else {
    if (be.size() > 1) {
      w_avg = at::empty_like(w, at::MemoryFormat::Contiguous);
      Tensor wc = w.contiguous();
      Epilogue epi;
      epi.scale = scale;
      epi.has_scale = true;
      be.allreduce(wc.data_ptr(), w_avg.data_ptr(), wc.numel(), DType::BF16, ReduceOp::SUM, epi, stream);
    } else {
      w_avg = w;
    }
    be.gemm_bf16_tn(x.data_ptr(), w_avg.data_ptr(), dy.data_ptr(), x.size(0), w_avg.size(0), x.size(1), x.stride(0),
                    w_avg.stride(0), dy.stride(0), stream, &mse);
  }
  return {dy, loss, w_avg};
}

// Tensor backed by the symmetric heap's persistent arena (same offset on every
// rank when all ranks allocate in the same order).  Kernels that need peer- or
// switch-visible inputs use such tensors in place instead of staging them.
Tensor symmetric_empty(c10::IntArrayRef shape, c10::ScalarType dtype) {
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  CudaBackend& be = backend();
  int64_t numel = 1;
  for (int64_t d : shape) numel *= d;
  const int64_t bytes = numel * static_cast<int64_t>(c10::elementSize(dtype));
  auto opts = at::TensorOptions().dtype(dtype).device(c10::Device(c10::kCUDA, static_cast<c10::DeviceIndex>(be.device())));
  if (be.size() <= 1) return at::empty(shape, opts);  // no peers: ordinary memory is as good
  const int64_t off = be.symm_alloc(bytes);
  return torch::from_blob(be.symm_ptr(off), shape, opts);
}

// param <- param + scale * Allreduce(grad): the gradient all-reduce with the
// optimizer update as its epilogue, in place.
void allreduce_axpy_(Tensor param, const Tensor& grad, double scale, int64_t max_blocks) {
  auto comm = c10::make_intrusive<Communicator>();
  comm->raw_allreduce_axpy_(param, grad, scale, max_blocks);
}

const float* grad_scale_ptr(const c10::optional<Tensor>& gs, const Tensor& like) {
  if (!gs.has_value() || !gs->defined()) return nullptr;
  TORCH_CHECK(gs->is_cuda() && gs->device() == like.device() && gs->scalar_type() == at::kFloat && gs->numel() == 1,
              "mpi4torch_b200: grad_scale must be a one-element fp32 tensor on the same device");
  return gs->data_ptr<float>();
}

// G = grad_scale * dy^T @ x on the tcgen05 path with MN-major operands (no transposed copies).
// grad_scale is a device scalar (the upstream gradient of the loss as autograd delivers it).
Tensor wgrad_bf16(const Tensor& dy, const Tensor& x, const c10::optional<Tensor>& grad_scale) {
  check_2d_bf16(dy, "dy");
  check_2d_bf16(x, "x");
  TORCH_CHECK(dy.size(0) == x.size(0), "mpi4torch_b200: batch dimensions differ");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor gw = at::empty({dy.size(1), x.size(1)}, x.options());
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  launch_wgrad_bf16(dy.data_ptr(), x.data_ptr(), gw.data_ptr(), dy.size(0), dy.size(1), x.size(1), dy.stride(0),
                    x.stride(0), gw.stride(0), backend().device_comm().sm_count,
                    c10::cuda::getCurrentCUDAStream(x.device().index()).stream(), grad_scale_ptr(grad_scale, x));
  return gw;
}

// w += scale * grad_scale * dy^T @ x: the single-rank SGD step as the wgrad GEMM's own epilogue
// (no gradient tensor, one kernel).  Not a collective.
void wgrad_sgd_(Tensor w, const Tensor& dy, const Tensor& x, double scale, const c10::optional<Tensor>& grad_scale) {
  check_2d_bf16(dy, "dy");
  check_2d_bf16(x, "x");
  check_2d_bf16(w, "w");
  TORCH_CHECK(dy.size(0) == x.size(0) && w.size(0) == dy.size(1) && w.size(1) == x.size(1),
              "mpi4torch_b200: wgrad_sgd_ shape mismatch");
  TORCH_CHECK(scale != 0.0, "mpi4torch_b200: wgrad_sgd_ needs a non-zero scale");
  c10::cuda::CUDAGuard guard(x.device());
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  launch_wgrad_bf16(dy.data_ptr(), x.data_ptr(), w.data_ptr(), dy.size(0), dy.size(1), x.size(1), dy.stride(0),
                    x.stride(0), w.stride(0), backend().device_comm().sm_count,
                    c10::cuda::getCurrentCUDAStream(x.device().index()).stream(), grad_scale_ptr(grad_scale, x),
                    static_cast<float>(scale));
}

bool wgrad_bf16_ok(const Tensor& dy, const Tensor& x) {
  if (!(dy.is_cuda() && x.is_cuda() && dy.dim() == 2 && x.dim() == 2 && dy.scalar_type() == at::kBFloat16 &&
        x.scalar_type() == at::kBFloat16 && dy.stride(1) == 1 && x.stride(1) == 1 && dy.size(0) == x.size(0)))
    return false;
  return World::instance().cuda_ready() &&
         wgrad_bf16_supported(dy.size(0), dy.size(1), x.size(1), dy.data_ptr(), x.data_ptr(), x.data_ptr(), dy.stride(0),
                              x.stride(0), x.size(1));
}

bool wgrad_allreduce_sgd_supported(const Tensor& w, const Tensor& dy, const Tensor& x) {
  if (!World::instance().cuda_ready() || !wgrad_bf16_ok(dy, x)) return false;
  if (!(w.is_cuda() && w.dim() == 2 && w.scalar_type() == at::kBFloat16 && w.is_contiguous() &&
        w.size(0) == dy.size(1) && w.size(1) == x.size(1)))
    return false;
  return backend().fused_wgrad_available(w.data_ptr(), dy.size(0), w.size(0), w.size(1));
}

// w += scale * sum_ranks(dy^T @ x): backward GEMM, gradient all-reduce and SGD step in one kernel.
// Collective; `w` must be a replicated symmetric_empty() tensor (M4T_FUSED_WGRAD=0 disables the kernel).
void wgrad_allreduce_sgd_(Tensor w, const Tensor& dy, const Tensor& x, double scale,
                          const c10::optional<Tensor>& grad_scale) {
  TORCH_CHECK(wgrad_allreduce_sgd_supported(w, dy, x), "mpi4torch_b200: fused wgrad->Allreduce->SGD does not support these tensors");
  c10::cuda::CUDAGuard guard(x.device());
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  backend().fused_wgrad_update(w.data_ptr(), dy.data_ptr(), x.data_ptr(), dy.size(0), w.size(0), w.size(1), dy.stride(0),
                               x.stride(0), static_cast<float>(scale),
                               c10::cuda::getCurrentCUDAStream(x.device().index()).stream(), false,
                               grad_scale_ptr(grad_scale, x));
}

// Same kernel, and additionally the parameter all-reduce of the NEXT forward: returns
// W_avg = (1/size) * sum_ranks w_new (symmetric-heap memory, valid until the next call).
Tensor wgrad_allreduce_sgd_prefetch_(Tensor w, const Tensor& dy, const Tensor& x, double scale,
                                     const c10::optional<Tensor>& grad_scale) {
  TORCH_CHECK(wgrad_allreduce_sgd_supported(w, dy, x), "mpi4torch_b200: fused wgrad->Allreduce->SGD does not support these tensors");
  c10::cuda::CUDAGuard guard(x.device());
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  const void* wavg = backend().fused_wgrad_update(w.data_ptr(), dy.data_ptr(), x.data_ptr(), dy.size(0), w.size(0), w.size(1),
                                                  dy.stride(0), x.stride(0), static_cast<float>(scale),
                                                  c10::cuda::getCurrentCUDAStream(x.device().index()).stream(), true,
                                                  grad_scale_ptr(grad_scale, x));
  return torch::from_blob(const_cast<void*>(wavg), {w.size(0), w.size(1)}, w.options().requires_grad(false));
}

// Forward + fused MSE epilogue against weights that are ALREADY averaged (no collective):
// (dL/dy, loss[1]).  Pairs with wgrad_allreduce_sgd_prefetch_.
std::tuple<Tensor, Tensor> linear_mse_forward_local(const Tensor& x, const Tensor& w_avg, const Tensor& target,
                                                    double loss_scale, double grad_scale) {
  check_2d_bf16(x, "x");
  check_2d_bf16(w_avg, "w_avg");
  check_2d_bf16(target, "target");
  TORCH_CHECK(x.size(1) == w_avg.size(1) && target.size(0) == x.size(0) && target.size(1) == w_avg.size(0),
              "mpi4torch_b200: linear_mse_forward_local shape mismatch");
  c10::cuda::CUDAGuard guard(x.device());
  cudaStream_t stream = c10::cuda::getCurrentCUDAStream(x.device().index()).stream();
  Tensor dy = at::empty({x.size(0), w_avg.size(0)}, x.options());
  Tensor loss = at::zeros({1}, x.options().dtype(at::kFloat));
  MseEpilogue mse;
  mse.target = target.data_ptr();
  mse.ldt = target.stride(0);
  mse.loss_acc = loss.data_ptr<float>();
  mse.loss_scale = static_cast<float>(loss_scale);
  mse.grad_scale = static_cast<float>(grad_scale);
  std::lock_guard<std::recursive_mutex> g(World::instance().mutex());
  backend().gemm_bf16_tn(x.data_ptr(), w_avg.data_ptr(), dy.data_ptr(), x.size(0), w_avg.size(0), x.size(1), x.stride(0),
                         w_avg.stride(0), dy.stride(0), stream, &mse);
  return {dy, loss};
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(mpi4torch_b200, m) {
  m.def("wgrad_allreduce_sgd_prefetch_(Tensor(a!) w, Tensor dy, Tensor x, float scale, Tensor? grad_scale=None) -> Tensor",
        &wgrad_allreduce_sgd_prefetch_);
  m.def("wgrad_sgd_(Tensor(a!) w, Tensor dy, Tensor x, float scale, Tensor? grad_scale=None) -> ()", &wgrad_sgd_);
  m.def("linear_mse_forward_local(Tensor x, Tensor w_avg, Tensor target, float loss_scale, float grad_scale) -> (Tensor, Tensor)",
        &linear_mse_forward_local);
  m.def("wgrad_bf16(Tensor dy, Tensor x, Tensor? grad_scale=None) -> Tensor", &wgrad_bf16);
  m.def("wgrad_bf16_supported(Tensor dy, Tensor x) -> bool", &wgrad_bf16_ok);
  m.def("wgrad_allreduce_sgd_supported(Tensor w, Tensor dy, Tensor x) -> bool", &wgrad_allreduce_sgd_supported);
  m.def("wgrad_allreduce_sgd_(Tensor(a!) w, Tensor dy, Tensor x, float scale, Tensor? grad_scale=None) -> ()",
        &wgrad_allreduce_sgd_);
  m.def("linear_mse_forward(Tensor x, Tensor w, Tensor target, float scale, float loss_scale, float grad_scale, "
        "bool allow_fused) -> (Tensor, Tensor, Tensor)",
        &linear_mse_forward);
  m.def("allreduce_axpy_(Tensor(a!) param, Tensor grad, float scale, int max_blocks=0) -> ()", &allreduce_axpy_);
  m.def("symmetric_empty(int[] shape, ScalarType dtype) -> Tensor", &symmetric_empty);
  m.def("gemm_bf16_tn(Tensor x, Tensor w) -> Tensor", &gemm_bf16_tn);
  m.def("gemm_bf16_nn(Tensor a, Tensor b) -> Tensor", &gemm_bf16_nn);
  m.def("gemm_bf16_nn_supported(Tensor a, Tensor b) -> bool", &gemm_bf16_nn_ok);
  m.def("gemm_bf16_tn_2cta(Tensor x, Tensor w) -> Tensor", &gemm_bf16_tn_2cta);
  m.def("gemm_bf16_tn_supported(Tensor x, Tensor w) -> bool", &gemm_bf16_tn_ok);
  m.def("allreduce_linear_supported(Tensor x, Tensor w) -> bool", &allreduce_linear_supported);
  m.def("allreduce_linear_fused(Tensor x, Tensor w, float scale) -> (Tensor, Tensor)", &allreduce_linear_fused);
}

}  // namespace m4t
