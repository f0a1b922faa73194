// Allgather by NVSwitch multicast push (experimental, M4T_AG_PUSH=1):
//   K1  every rank multimem.st's its slab straight into the GATHERED layout of
//       the staging half of ALL ranks (egress S per GPU instead of (P-1)*S for
//       the pull kernel; the switch replicates),
//   K2  per-block barrier with all peers, then a flat local copy staging -> out.
// Requires 16-byte aligned slabs; anything else stays on slab_pull_kernel.
#include <algorithm>

#include "kernels.h"
#include "vec_ops.cuh"

namespace m4t {

namespace {

constexpr int kThreads = 512;
constexpr int kUnroll = 4;

struct PushArgs {
  SyncCtx sync;
  char* mc_heap;
  char* my_heap;
  const char* in;
  char* out;
  int64_t stage_off, half_bytes;
  // my slab: rows x run_vecs 16-byte vectors
  int64_t src_off, dst_off;  // bytes
  int64_t n1, n2;
  int64_t ss[3], ds[3];      // bytes
  int64_t run_vecs, total_items;
  int64_t out_bytes;
};

__global__ void __launch_bounds__(kThreads) slab_push_mc_kernel(const PushArgs a) {
  const int par = static_cast<int>(read_op_count(a.sync) & 1ull);
  char* dst = a.mc_heap + a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t it0 = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; it0 < a.total_items;
       it0 += kUnroll * stride) {
    Vec16 v[kUnroll];
    int64_t dof[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t it = it0 + u * stride;
      dof[u] = -1;
      if (it < a.total_items) {
        const int64_t row = it / a.run_vecs;
        const int64_t vec = it - row * a.run_vecs;
        const int64_t i2 = row % a.n2;
        const int64_t t = row / a.n2;
        const int64_t i1 = t % a.n1;
        const int64_t i0 = t / a.n1;
        const int64_t so = a.src_off + i0 * a.ss[0] + i1 * a.ss[1] + i2 * a.ss[2] + vec * 16;
        dof[u] = a.dst_off + i0 * a.ds[0] + i1 * a.ds[1] + i2 * a.ds[2] + vec * 16;
        v[u] = ld_vec_stream(a.in + so);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dof[u] >= 0) multimem_st_vec(dst + dof[u], v[u]);
  }
  __threadfence_system();  // my multicast stores are performed before this kernel retires
}

__global__ void __launch_bounds__(kThreads) stage_out_kernel(const PushArgs a) {
  const SyncCtx& c = a.sync;
  const unsigned long long fb = read_flag_base(c);
  const int par = static_cast<int>(read_op_count(c) & 1ull);
  block_barrier_all(c, fb, 0);  // every peer has entered K2, hence finished its K1
  const char* src = a.my_heap + a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const int64_t nvec = a.out_bytes / 16;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
    Vec16 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = ld_vec_sys(src + (i + u * stride) * 16);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) st_vec(a.out + (i + u * stride) * 16, v[u]);
  }
  for (; i < nvec; i += stride) st_vec(a.out + i * 16, ld_vec_sys(src + i * 16));
  finish_op(c, 1);
}

}  // namespace

// Returns false (and launches nothing) when the plan is not eligible.
bool launch_allgather_push(const DeviceComm& dc, const PullPlan& plan, const void* in, void* out, DType dt, int blocks,
                           cudaStream_t stream) {
  if (!plan.replicated_output || dc.mc_heap == nullptr || plan.out_elems == 0) return false;
  const int64_t es = dtype_size(dt);
  const SlabJob* mine = nullptr;
  for (const auto& j : plan.jobs)
    if (j.peer == dc.sync.rank) mine = &j;
  auto al16 = [](int64_t v) { return (v & 15) == 0; };
  const int64_t out_bytes = plan.out_elems * es;
  if (!al16(out_bytes) || out_bytes > dc.half_bytes) return false;  // rank-independent (replicated output)
  // eligibility must be decided identically on every rank: all slabs share the
  // same `after` stride pattern, so checking the global quantities suffices
  for (const auto& j : plan.jobs) {
    if (!al16(j.run * es) || !al16(j.dst_off * es)) return false;
    for (int k = 0; k < 3; ++k)
      if (j.n[k] > 1 && (!al16(j.ss[k] * es) || !al16(j.ds[k] * es))) return false;
  }
  // Everything above is identical on every rank, so all ranks take the same path.  Pointer
  // alignment is rank-local: it cannot select a different algorithm without desynchronising
  // the ranks, so it is a hard requirement of the opt-in path.
  M4T_CHECK(al16(reinterpret_cast<intptr_t>(in)) && al16(reinterpret_cast<intptr_t>(out)),
            "M4T_AG_PUSH needs 16-byte aligned input and output tensors");
  PushArgs a{};
  a.sync = dc.sync;
  a.mc_heap = dc.mc_heap;
  a.my_heap = dc.heap[dc.sync.rank];
  a.in = static_cast<const char*>(in);
  a.out = static_cast<char*>(out);
  a.stage_off = dc.stage_off;
  a.half_bytes = dc.half_bytes;
  a.out_bytes = out_bytes;
  if (mine != nullptr) {
    a.src_off = mine->src_off * es;
    a.dst_off = mine->dst_off * es;
    a.n1 = mine->n[1];
    a.n2 = mine->n[2];
    for (int k = 0; k < 3; ++k) {
      a.ss[k] = mine->ss[k] * es;
      a.ds[k] = mine->ds[k] * es;
    }
    a.run_vecs = mine->run * es / 16;
    a.total_items = mine->rows() * a.run_vecs;
  } else {
    a.n1 = a.n2 = 1;
    a.run_vecs = 1;
    a.total_items = 0;  // this rank contributes an empty slab
  }
  blocks = std::max(1, std::min(blocks, kMaxChannels));
  if (a.total_items > 0) {
    const int pb = static_cast<int>(std::min<int64_t>((a.total_items + kThreads * kUnroll - 1) / (kThreads * kUnroll), 4LL * dc.sm_count));
    slab_push_mc_kernel<<<std::max(1, pb), kThreads, 0, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    M4T_CHECK(e == cudaSuccess, "slab_push_mc launch failed: " << cudaGetErrorString(e));
    note_kernel_launch();
  }
  stage_out_kernel<<<blocks, kThreads, 0, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, "stage_out launch failed: " << cudaGetErrorString(e));
  note_kernel_launch();
  return true;
}

}  // namespace m4t
