// Rooted collectives: Bcast_ and Reduce_ (each other's adjoints, reference
// csrc/extension.cpp:324-331 and :381-393), in place on the caller's buffer.
//
//   Bcast_  : root pushes its buffer into every rank's staging half with ONE
//             multimem.st per 16 bytes (the NVSwitch replicates), per-block
//             barrier, non-roots copy out of their own HBM.  Without a
//             multicast mapping the root stages locally and peers pull.
//   Reduce_ : every rank stages, per-block barrier, the root reduces
//             (multimem.ld_reduce issued by the root only, or P peer loads in
//             rank order), non-roots zero-fill in the same launch (reference
//             :443-447 zero-fills in a separate ATen op).
#include <algorithm>

#include "kernels.h"
#include "vec_ops.cuh"

namespace m4t {

namespace {

constexpr int kThreads = 512;

struct RootedArgs {
  SyncCtx sync;
  char* heap[kMaxGpuPeers];
  char* mc_heap;
  void* buf;
  int64_t stage_off;
  int64_t half_bytes;
  int64_t n;
  int64_t nvec;
  int64_t chunk_vecs;
  int root;
  int aligned;
};

template <DType DT>
__global__ void __launch_bounds__(kThreads) bcast_kernel(const RootedArgs a) {
  const SyncCtx& c = a.sync;
  const unsigned long long fb = read_flag_base(c);
  const int par = static_cast<int>(read_op_count(c) & 1ull);
  const int64_t half = a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const bool al = a.aligned != 0;
  const bool is_root = c.rank == a.root;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  int bar = 0;
  for (int64_t base = 0; base < a.nvec; base += a.chunk_vecs) {
    const int64_t end = min(base + a.chunk_vecs, a.nvec);
    if (is_root) {
      for (int64_t i = base + first; i < end; i += stride) {
        const Vec16 v = load_private<DT>(a.buf, i, a.n, al);
        if (a.mc_heap) multimem_st_vec(a.mc_heap + half + i * 16, v);
        else st_vec(a.heap[c.rank] + half + i * 16, v);
      }
    }
    block_barrier_all(c, fb, bar++);
    if (!is_root) {
      const char* src = (a.mc_heap ? a.heap[c.rank] : a.heap[a.root]) + half;
      for (int64_t i = base + first; i < end; i += stride)
        store_private<DT>(a.buf, i, a.n, al, ld_vec_sys(src + i * 16));
    }
  }
  finish_op(c, static_cast<unsigned int>(bar));
}

template <DType DT, ReduceOp OP, NvlsKind NK>
__global__ void __launch_bounds__(kThreads) reduce_kernel(const RootedArgs a) {
  using V = VecOf<DT>;
  const SyncCtx& c = a.sync;
  const unsigned long long fb = read_flag_base(c);
  const int par = static_cast<int>(read_op_count(c) & 1ull);
  const int64_t half = a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const bool al = a.aligned != 0;
  const bool is_root = c.rank == a.root;
  const int P = c.size;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  Vec16 zero;
  zero.w[0] = zero.w[1] = zero.w[2] = zero.w[3] = 0;
  int bar = 0;
  for (int64_t base = 0; base < a.nvec; base += a.chunk_vecs) {
    const int64_t end = min(base + a.chunk_vecs, a.nvec);
    for (int64_t i = base + first; i < end; i += stride) {
      st_vec(a.heap[c.rank] + half + i * 16, load_private<DT>(a.buf, i, a.n, al));
      if (!is_root) store_private<DT>(a.buf, i, a.n, al, zero);  // non-root result (reference :443-447)
    }
    block_barrier_all(c, fb, bar++);
    if (is_root) {
      if constexpr (NK != NvlsKind::NONE) {
        // four in-switch reductions in flight per thread (one NVSwitch round trip each)
        for (int64_t i0 = base + first; i0 < end; i0 += 4 * stride) {
          Vec16 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u * stride < end) x[u] = multimem_ld_reduce_vec<NK>(a.mc_heap + half + (i0 + u * stride) * 16);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u * stride < end) store_private<DT>(a.buf, i0 + u * stride, a.n, al, x[u]);
        }
      } else {
        for (int64_t i = base + first; i < end; i += stride) {
          typename V::A acc[V::N];
          init_from<DT, OP>(acc, ld_vec_sys(a.heap[0] + half + i * 16));
#pragma unroll 1
          for (int p = 1; p < P; ++p) combine_into<DT, OP>(acc, ld_vec_sys(a.heap[p] + half + i * 16));
          store_private<DT>(a.buf, i, a.n, al, V::pack(acc));
        }
      }
    }
  }
  finish_op(c, static_cast<unsigned int>(bar));
}

template <DType DT, ReduceOp OP> struct LaunchReduce {
  static void run(const RootedArgs& a, int blocks, cudaStream_t s) {
    reduce_kernel<DT, OP, NvlsKind::NONE><<<blocks, kThreads, 0, s>>>(a);
  }
};

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, what << " launch failed: " << cudaGetErrorString(e));
  note_kernel_launch();
}

RootedArgs make_args(const DeviceComm& dc, void* buf, int64_t n, DType dt, int root, int& blocks) {
  RootedArgs a;
  a.sync = dc.sync;
  for (int p = 0; p < kMaxGpuPeers; ++p) a.heap[p] = dc.heap[p];
  a.mc_heap = dc.mc_heap;
  a.buf = buf;
  a.stage_off = dc.stage_off;
  a.half_bytes = dc.half_bytes;
  a.n = n;
  a.nvec = (n * dtype_size(dt) + 15) / 16;
  a.root = root;
  a.aligned = (reinterpret_cast<uintptr_t>(buf) & 15u) == 0;
  M4T_CHECK(a.nvec * 16 <= dc.half_bytes, "rooted collective of " << n << " elements exceeds the staging half");
  blocks = std::max(1, std::min(blocks, kMaxChannels));
  // ~8 pipeline steps for large buffers, one for small ones
  const int64_t per_step = static_cast<int64_t>(blocks) * kThreads * 8;
  a.chunk_vecs = std::max<int64_t>(per_step, (a.nvec + 7) / 8);
  a.chunk_vecs = std::max<int64_t>(1, a.chunk_vecs);
  blocks = static_cast<int>(std::min<int64_t>(blocks, std::max<int64_t>(1, (std::min(a.chunk_vecs, std::max<int64_t>(a.nvec, 1)) + kThreads - 1) / kThreads)));
  return a;
}

}  // namespace

void launch_bcast(const DeviceComm& dc, void* buf, int64_t n, DType dt, int root, int blocks, cudaStream_t stream) {
  RootedArgs a = make_args(dc, buf, n, dt, root, blocks);
  // data movement only: dispatch on element size
  switch (dtype_size(dt)) {
    case 1: bcast_kernel<DType::U8><<<blocks, kThreads, 0, stream>>>(a); break;
    case 2: bcast_kernel<DType::I16><<<blocks, kThreads, 0, stream>>>(a); break;
    case 4: bcast_kernel<DType::I32><<<blocks, kThreads, 0, stream>>>(a); break;
    default: bcast_kernel<DType::I64><<<blocks, kThreads, 0, stream>>>(a); break;
  }
  check_launch("bcast");
}

void launch_reduce(const DeviceComm& dc, void* buf, int64_t n, DType dt, ReduceOp op, int root, int blocks,
                   cudaStream_t stream) {
  check_op_dtype(op, dt);
  RootedArgs a = make_args(dc, buf, n, dt, root, blocks);
  const bool nvls = dc.mc_heap != nullptr && nvls_supported(dt, op);
  if (nvls && op == ReduceOp::SUM && dt == DType::BF16) {
    reduce_kernel<DType::BF16, ReduceOp::SUM, NvlsKind::ADD_BF16><<<blocks, kThreads, 0, stream>>>(a);
  } else if (nvls && op == ReduceOp::SUM && dt == DType::F16) {
    reduce_kernel<DType::F16, ReduceOp::SUM, NvlsKind::ADD_F16><<<blocks, kThreads, 0, stream>>>(a);
  } else if (nvls && op == ReduceOp::SUM && dt == DType::F32) {
    reduce_kernel<DType::F32, ReduceOp::SUM, NvlsKind::ADD_F32><<<blocks, kThreads, 0, stream>>>(a);
  } else {
    M4T_DISPATCH_DTYPE_OP(dt, op, LaunchReduce, a, blocks, stream);
  }
  check_launch("reduce");
}

}  // namespace m4t
