// Axis-aware data movement: Gather / Allgather / Scatter / Alltoall (and their
// adjoints, which are the same kernels with swapped plans) plus the
// reduce-scatter box that is Allgather's true adjoint.
//
// The reference builds four MPI derived datatypes per call and issues 2..4*P
// blocking collectives (csrc/extension.cpp:516-591, :942-946); here the host
// turns the call into a PullPlan (runtime/plan.h) and ONE kernel pulls every
// box straight out of the peers' HBM over NVLink with 16-byte loads:
//   stage_in kernel : private input -> my staging half (so peers can read it)
//   slab kernel     : per-block barrier with all peers, then strided pulls.
// A rank that pulls from itself reads its private input directly.
#include <algorithm>

#include "kernels.h"
#include "vec_ops.cuh"

namespace m4t {

namespace {

constexpr int kThreads = 512;

struct DevJob {
  int64_t src_off;  // bytes
  int64_t dst_off;  // bytes
  int64_t n1, n2;   // inner loop extents (n0 implied by rows)
  int64_t ss[3];    // bytes
  int64_t ds[3];    // bytes
  int64_t run_vecs;
  int64_t item_begin;  // prefix sum of rows*run_vecs
  int peer;
};

struct SlabArgs {
  SyncCtx sync;
  char* heap[kMaxGpuPeers];
  char* mc_heap;
  const char* in;
  char* out;
  int64_t stage_off;
  int64_t half_bytes;
  int64_t total_items;
  int njobs;
  int do_barrier;
  DevJob jobs[kMaxGpuPeers];
};

template <int VB> struct Mover;
template <> struct Mover<16> {
  static __device__ __forceinline__ void copy(char* d, const char* s) { st_vec(d, ld_vec_sys(s)); }
};
template <> struct Mover<8> {
  static __device__ __forceinline__ void copy(char* d, const char* s) {
    uint32_t a, b;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(s) : "memory");
    asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(d), "r"(a), "r"(b) : "memory");
  }
};
template <> struct Mover<4> {
  static __device__ __forceinline__ void copy(char* d, const char* s) {
    *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<const volatile uint32_t*>(s);
  }
};
template <> struct Mover<2> {
  static __device__ __forceinline__ void copy(char* d, const char* s) {
    *reinterpret_cast<uint16_t*>(d) = *reinterpret_cast<const volatile uint16_t*>(s);
  }
};
template <> struct Mover<1> {
  static __device__ __forceinline__ void copy(char* d, const char* s) {
    *reinterpret_cast<uint8_t*>(d) = *reinterpret_cast<const volatile uint8_t*>(s);
  }
};

__device__ __forceinline__ void item_to_offsets(const DevJob& j, int64_t local, int64_t vb, int64_t& so,
                                                int64_t& d_o) {
  const int64_t row = local / j.run_vecs;
  const int64_t v = local - row * j.run_vecs;
  const int64_t i2 = row % j.n2;
  const int64_t t = row / j.n2;
  const int64_t i1 = t % j.n1;
  const int64_t i0 = t / j.n1;
  so = j.src_off + i0 * j.ss[0] + i1 * j.ss[1] + i2 * j.ss[2] + v * vb;
  d_o = j.dst_off + i0 * j.ds[0] + i1 * j.ds[1] + i2 * j.ds[2] + v * vb;
}

// Loads kPullUnroll items before storing any: peer loads take ~2 us, so the
// achievable NVLink bandwidth is proportional to the requests in flight.
constexpr int kPullUnroll = 4;
constexpr int kReduceUnroll = 4;  // slab_reduce_vec: switch / peer reductions in flight per thread

template <int VB> struct MoverReg;
template <> struct MoverReg<16> {
  using T = Vec16;
  static __device__ __forceinline__ T ld(const char* s) { return ld_vec_sys(s); }
  static __device__ __forceinline__ void st(char* d, const T& v) { st_vec(d, v); }
};
template <> struct MoverReg<8> {
  using T = uint2;
  static __device__ __forceinline__ T ld(const char* s) {
    T v;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(s) : "memory");
    return v;
  }
  static __device__ __forceinline__ void st(char* d, const T& v) {
    asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(d), "r"(v.x), "r"(v.y) : "memory");
  }
};
template <> struct MoverReg<4> {
  using T = uint32_t;
  static __device__ __forceinline__ T ld(const char* s) { return *reinterpret_cast<const volatile uint32_t*>(s); }
  static __device__ __forceinline__ void st(char* d, const T& v) { *reinterpret_cast<uint32_t*>(d) = v; }
};
template <> struct MoverReg<2> {
  using T = uint16_t;
  static __device__ __forceinline__ T ld(const char* s) { return *reinterpret_cast<const volatile uint16_t*>(s); }
  static __device__ __forceinline__ void st(char* d, const T& v) { *reinterpret_cast<uint16_t*>(d) = v; }
};
template <> struct MoverReg<1> {
  using T = uint8_t;
  static __device__ __forceinline__ T ld(const char* s) { return *reinterpret_cast<const volatile uint8_t*>(s); }
  static __device__ __forceinline__ void st(char* d, const T& v) { *reinterpret_cast<uint8_t*>(d) = v; }
};

template <int VB>
__global__ void __launch_bounds__(kThreads) slab_pull_kernel(const SlabArgs a) {
  using MR = MoverReg<VB>;
  const SyncCtx& c = a.sync;
  unsigned long long fb = 0;
  int par = 0;
  if (a.do_barrier) {
    fb = read_flag_base(c);
    par = static_cast<int>(read_op_count(c) & 1ull);
    block_barrier_all(c, fb, 0);
  }
  const int64_t half = a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t it0 = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; it0 < a.total_items;
       it0 += kPullUnroll * stride) {
    typename MR::T v[kPullUnroll];
    int64_t dof[kPullUnroll];
#pragma unroll
    for (int u = 0; u < kPullUnroll; ++u) {
      const int64_t it = it0 + u * stride;
      dof[u] = -1;
      if (it < a.total_items) {
        int jx = 0;
#pragma unroll 1
        while (jx + 1 < a.njobs && it >= a.jobs[jx + 1].item_begin) ++jx;
        const DevJob& j = a.jobs[jx];
        int64_t so;
        item_to_offsets(j, it - j.item_begin, VB, so, dof[u]);
        const char* src = (j.peer == c.rank || !a.do_barrier) ? a.in : (a.heap[j.peer] + half);
        v[u] = MR::ld(src + so);
      }
    }
#pragma unroll
    for (int u = 0; u < kPullUnroll; ++u)
      if (dof[u] >= 0) MR::st(a.out + dof[u], v[u]);
  }
  if (a.do_barrier) finish_op(c, 1);
}

// Private -> staging half copy.  Byte-granular tail, 16-byte body when aligned.
__global__ void __launch_bounds__(kThreads) stage_in_kernel(SyncCtx c, char* my_heap, int64_t stage_off,
                                                            int64_t half_bytes, const char* in, int64_t bytes,
                                                            int aligned) {
  const int par = static_cast<int>(read_op_count(c) & 1ull);
  char* dst = my_heap + stage_off + static_cast<int64_t>(par) * half_bytes;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (aligned) {
    const int64_t nvec = bytes / 16;
    for (int64_t i = tid; i < nvec; i += stride) st_vec(dst + i * 16, ld_vec_stream(in + i * 16));
    for (int64_t i = nvec * 16 + tid; i < bytes; i += stride) dst[i] = in[i];
  } else {
    for (int64_t i = tid; i < bytes; i += stride) dst[i] = in[i];
  }
}

// ---------------------------------------------------------------------------
// reduce-scatter box: out[box] = epilogue(reduce_p staged_p[box])
// ---------------------------------------------------------------------------
struct ReduceArgs {
  SyncCtx sync;
  char* heap[kMaxGpuPeers];
  char* mc_heap;
  const char* in;
  char* out;
  DevEpilogue epi;
  int64_t stage_off;
  int64_t half_bytes;
  int64_t total_items;
  int64_t out_elems;
  DevJob job;
  int do_barrier;
};

template <DType DT, ReduceOp OP, NvlsKind NK>
__global__ void __launch_bounds__(kThreads) slab_reduce_vec_kernel(const ReduceArgs a) {
  using V = VecOf<DT>;
  const SyncCtx& c = a.sync;
  unsigned long long fb = 0;
  int par = 0;
  if (a.do_barrier) {
    fb = read_flag_base(c);
    par = static_cast<int>(read_op_count(c) & 1ull);
    block_barrier_all(c, fb, 0);
  }
  const int P = a.do_barrier ? c.size : 1;
  const int64_t half = a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  // kReduceUnroll items per thread and trip: the in-switch reductions (one ~3 us NVSwitch round
  // trip each) are all requested before the first result is consumed
  for (int64_t it0 = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; it0 < a.total_items;
       it0 += kReduceUnroll * stride) {
    int64_t so[kReduceUnroll], d_o[kReduceUnroll];
    Vec16 x[kReduceUnroll];
#pragma unroll
    for (int u = 0; u < kReduceUnroll; ++u) {
      const int64_t it = it0 + u * stride;
      d_o[u] = -1;
      if (it < a.total_items) {
        item_to_offsets(a.job, it, 16, so[u], d_o[u]);
        if constexpr (NK != NvlsKind::NONE) x[u] = multimem_ld_reduce_vec<NK>(a.mc_heap + half + so[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kReduceUnroll; ++u) {
      if (d_o[u] < 0) continue;
      typename V::A acc[V::N];
      if constexpr (NK != NvlsKind::NONE) {
        V::unpack(x[u], acc);
      } else {
        // peers in batches of four: four NVLink loads in flight, combined in rank order
        const char* s0 = (c.rank == 0 || !a.do_barrier) ? a.in : (a.heap[0] + half);
        init_from<DT, OP>(acc, ld_vec_sys(s0 + so[u]));
#pragma unroll 1
        for (int p0 = 1; p0 < P; p0 += 4) {
          Vec16 v[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int p = p0 + w;
            if (p < P) {
              const char* sp = (p == c.rank) ? a.in : (a.heap[p] + half);
              v[w] = ld_vec_sys(sp + so[u]);
            }
          }
#pragma unroll
          for (int w = 0; w < 4; ++w)
            if (p0 + w < P) combine_into<DT, OP>(acc, v[w]);
        }
      }
      apply_scale<DT>(acc, a.epi);
      // the box is 16-byte aligned in the output, so vector index = byte offset / 16
      apply_accumulate<DT>(acc, a.epi, d_o[u] / 16, a.out_elems, true);
      st_vec(a.out + d_o[u], V::pack(acc));
    }
  }
  if (a.do_barrier) finish_op(c, 1);
}

// Element-granular fallback (unaligned boxes).
template <DType DT, ReduceOp OP>
__global__ void __launch_bounds__(kThreads) slab_reduce_elem_kernel(const ReduceArgs a) {
  using E = Elem<DT>;
  using S = typename E::storage;
  using A = typename E::acc;
  using C = Combine<OP, A, E::is_float>;
  const SyncCtx& c = a.sync;
  unsigned long long fb = 0;
  int par = 0;
  if (a.do_barrier) {
    fb = read_flag_base(c);
    par = static_cast<int>(read_op_count(c) & 1ull);
    block_barrier_all(c, fb, 0);
  }
  const int P = a.do_barrier ? c.size : 1;
  const int64_t half = a.stage_off + static_cast<int64_t>(par) * a.half_bytes;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t it = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; it < a.total_items; it += stride) {
    int64_t so, d_o;
    item_to_offsets(a.job, it, sizeof(S), so, d_o);
    const char* s0 = (c.rank == 0 || !a.do_barrier) ? a.in : (a.heap[0] + half);
    A acc = normalise_single<OP, A>(E::load(*reinterpret_cast<const volatile S*>(s0 + so)));
#pragma unroll 1
    for (int p = 1; p < P; ++p) {
      const char* sp = (p == c.rank) ? a.in : (a.heap[p] + half);
      acc = C::apply(acc, E::load(*reinterpret_cast<const volatile S*>(sp + so)));
    }
    if (a.epi.has_scale) acc = ScaleAcc<A>::apply(acc, a.epi);
    if (a.epi.acc) acc = acc + E::load(*reinterpret_cast<const S*>(static_cast<const char*>(a.epi.acc) + d_o));
    *reinterpret_cast<S*>(a.out + d_o) = E::store(acc);
  }
  if (a.do_barrier) finish_op(c, 1);
}

__global__ void __launch_bounds__(kThreads) zero_kernel(char* p, int64_t bytes, int aligned) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (aligned) {
    const int64_t nvec = bytes / 16;
    Vec16 z;
    z.w[0] = z.w[1] = z.w[2] = z.w[3] = 0;
    for (int64_t i = tid; i < nvec; i += stride) st_vec(p + i * 16, z);
    for (int64_t i = nvec * 16 + tid; i < bytes; i += stride) p[i] = 0;
  } else {
    for (int64_t i = tid; i < bytes; i += stride) p[i] = 0;
  }
}

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, what << " launch failed: " << cudaGetErrorString(e));
  note_kernel_launch();
}

int64_t gcd64(int64_t a, int64_t b) {
  a = a < 0 ? -a : a;
  b = b < 0 ? -b : b;
  while (b) {
    int64_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// Largest power-of-two vector width (<=16 bytes) dividing every offset/stride/run.
int pick_vb(const std::vector<SlabJob>& jobs, int64_t es, const void* in, const void* out, int64_t stage_off) {
  int64_t g = 16;
  g = gcd64(g, static_cast<int64_t>(reinterpret_cast<uintptr_t>(in) & 15));
  g = gcd64(g, static_cast<int64_t>(reinterpret_cast<uintptr_t>(out) & 15));
  g = gcd64(g, stage_off & 15);
  for (const auto& j : jobs) {
    g = gcd64(g, j.run * es);
    g = gcd64(g, j.src_off * es);
    g = gcd64(g, j.dst_off * es);
    for (int k = 0; k < 3; ++k) {
      if (j.n[k] > 1) {
        g = gcd64(g, j.ss[k] * es);
        g = gcd64(g, j.ds[k] * es);
      }
    }
  }
  if (g == 0) g = 16;
  int vb = 1;
  while (vb * 2 <= 16 && g % (vb * 2) == 0) vb *= 2;
  return vb;
}

void fill_dev_job(DevJob& d, const SlabJob& j, int64_t es, int vb, int64_t item_begin) {
  d.peer = j.peer;
  d.src_off = j.src_off * es;
  d.dst_off = j.dst_off * es;
  d.n1 = j.n[1];
  d.n2 = j.n[2];
  for (int k = 0; k < 3; ++k) {
    d.ss[k] = j.ss[k] * es;
    d.ds[k] = j.ds[k] * es;
  }
  d.run_vecs = j.run * es / vb;
  d.item_begin = item_begin;
}

template <DType DT, ReduceOp OP> struct LaunchReduceVec {
  static void run(const ReduceArgs& a, int blocks, cudaStream_t s) {
    slab_reduce_vec_kernel<DT, OP, NvlsKind::NONE><<<blocks, kThreads, 0, s>>>(a);
  }
};
template <DType DT, ReduceOp OP> struct LaunchReduceElem {
  static void run(const ReduceArgs& a, int blocks, cudaStream_t s) {
    slab_reduce_elem_kernel<DT, OP><<<blocks, kThreads, 0, s>>>(a);
  }
};

void launch_pull_impl(SlabArgs& a, int vb, int blocks, cudaStream_t stream) {
  switch (vb) {
    case 16: slab_pull_kernel<16><<<blocks, kThreads, 0, stream>>>(a); break;
    case 8: slab_pull_kernel<8><<<blocks, kThreads, 0, stream>>>(a); break;
    case 4: slab_pull_kernel<4><<<blocks, kThreads, 0, stream>>>(a); break;
    case 2: slab_pull_kernel<2><<<blocks, kThreads, 0, stream>>>(a); break;
    default: slab_pull_kernel<1><<<blocks, kThreads, 0, stream>>>(a); break;
  }
  check_launch("slab_pull");
}

void build_pull_args(SlabArgs& a, const PullPlan& plan, const void* in, void* out, DType dt, int64_t stage_off,
                     int& vb) {
  const int64_t es = dtype_size(dt);
  M4T_CHECK(plan.jobs.size() <= static_cast<size_t>(kMaxGpuPeers), "too many slab jobs");
  vb = pick_vb(plan.jobs, es, in, out, stage_off);
  a.in = static_cast<const char*>(in);
  a.out = static_cast<char*>(out);
  a.njobs = static_cast<int>(plan.jobs.size());
  int64_t items = 0;
  for (int k = 0; k < a.njobs; ++k) {
    fill_dev_job(a.jobs[k], plan.jobs[k], es, vb, items);
    items += plan.jobs[k].rows() * a.jobs[k].run_vecs;
  }
  a.total_items = items;
}

}  // namespace

namespace {
__global__ void __launch_bounds__(kThreads) copy_bytes_kernel(char* dst, const char* src, int64_t bytes, int aligned) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (aligned) {
    const int64_t nvec = bytes / 16;
    int64_t i = tid;
    // 4 independent 16-byte loads in flight per thread
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      const Vec16 a = ld_vec_stream(src + i * 16), b = ld_vec_stream(src + (i + stride) * 16),
                  c = ld_vec_stream(src + (i + 2 * stride) * 16), d = ld_vec_stream(src + (i + 3 * stride) * 16);
      st_vec(dst + i * 16, a);
      st_vec(dst + (i + stride) * 16, b);
      st_vec(dst + (i + 2 * stride) * 16, c);
      st_vec(dst + (i + 3 * stride) * 16, d);
    }
    for (; i < nvec; i += stride) st_vec(dst + i * 16, ld_vec_stream(src + i * 16));
    for (int64_t j = nvec * 16 + tid; j < bytes; j += stride) dst[j] = src[j];
  } else {
    for (int64_t j = tid; j < bytes; j += stride) dst[j] = src[j];
  }
}
}  // namespace

void launch_copy_bytes(void* dst, const void* src, int64_t bytes, int sm_count, cudaStream_t stream) {
  if (bytes <= 0) return;
  const int aligned = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0;
  const int blocks = static_cast<int>(std::min<int64_t>((bytes / 64 + kThreads) / kThreads, 4LL * sm_count));
  copy_bytes_kernel<<<blocks, kThreads, 0, stream>>>(static_cast<char*>(dst), static_cast<const char*>(src), bytes, aligned);
  check_launch("copy_bytes");
}

void launch_zero(void* p, int64_t bytes, int sm_count, cudaStream_t stream) {
  if (bytes <= 0) return;
  const int aligned = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
  const int blocks = static_cast<int>(std::min<int64_t>((bytes / 16 + kThreads) / kThreads, 4LL * sm_count));
  zero_kernel<<<blocks, kThreads, 0, stream>>>(static_cast<char*>(p), bytes, aligned);
  check_launch("zero");
}

void launch_stage_in(const DeviceComm& dc, const void* in, int64_t bytes, int sm_count, cudaStream_t stream) {
  if (bytes <= 0) return;
  M4T_CHECK(bytes <= dc.half_bytes, "staging " << bytes << " B exceeds the staging half (" << dc.half_bytes
                                               << " B); raise M4T_STAGE_MB (M4T_SUB_STAGE_MB for communicators created by Split)");
  const int aligned = (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (dc.stage_off & 15) == 0;
  const int blocks = static_cast<int>(std::min<int64_t>((bytes / 16 + kThreads) / kThreads, 4LL * sm_count));
  stage_in_kernel<<<blocks, kThreads, 0, stream>>>(dc.sync, dc.heap[dc.sync.rank], dc.stage_off, dc.half_bytes,
                                                   static_cast<const char*>(in), bytes, aligned);
  check_launch("stage_in");
}

void launch_slab_pull(const DeviceComm& dc, const PullPlan& plan, const void* in, void* out, DType dt, int blocks,
                      cudaStream_t stream) {
  SlabArgs a;
  a.sync = dc.sync;
  for (int p = 0; p < kMaxGpuPeers; ++p) a.heap[p] = dc.heap[p];
  a.mc_heap = dc.mc_heap;
  a.stage_off = dc.stage_off;
  a.half_bytes = dc.half_bytes;
  a.do_barrier = 1;
  int vb = 16;
  build_pull_args(a, plan, in, out, dt, dc.stage_off, vb);
  // grid must be identical on every rank (per-block barriers): derive it from
  // globally known quantities only.
  blocks = std::max(1, std::min(blocks, kMaxChannels));
  launch_pull_impl(a, vb, blocks, stream);
}

void launch_slab_local(const PullPlan& plan, const void* in, void* out, DType dt, int sm_count,
                       cudaStream_t stream) {
  if (plan.jobs.empty()) return;
  SlabArgs a;
  a.sync = SyncCtx{};
  for (int p = 0; p < kMaxGpuPeers; ++p) a.heap[p] = nullptr;
  a.mc_heap = nullptr;
  a.stage_off = 0;
  a.half_bytes = 0;
  a.do_barrier = 0;
  int vb = 16;
  build_pull_args(a, plan, in, out, dt, 0, vb);
  if (a.total_items == 0) return;
  const int blocks = static_cast<int>(std::min<int64_t>((a.total_items + kThreads - 1) / kThreads, 4LL * sm_count));
  launch_pull_impl(a, vb, blocks, stream);
}

void launch_slab_reduce(const DeviceComm& dc, const ReducePlan& plan, const void* in, void* out, DType dt,
                        ReduceOp op, const Epilogue& epi, bool use_nvls, int blocks, cudaStream_t stream) {
  check_op_dtype(op, dt);
  const int64_t es = dtype_size(dt);
  ReduceArgs a;
  a.sync = dc.sync;
  for (int p = 0; p < kMaxGpuPeers; ++p) a.heap[p] = dc.heap[p];
  a.mc_heap = dc.mc_heap;
  a.in = static_cast<const char*>(in);
  a.out = static_cast<char*>(out);
  a.epi = make_dev_epilogue(epi);
  a.stage_off = dc.stage_off;
  a.half_bytes = dc.half_bytes;
  a.out_elems = plan.out_elems;
  a.do_barrier = dc.sync.size > 1 ? 1 : 0;
  std::vector<SlabJob> one{plan.box};
  int vb = plan.out_elems > 0 ? pick_vb(one, es, in, out, dc.stage_off) : 16;
  if (epi.accumulate && (reinterpret_cast<uintptr_t>(epi.accumulate) & 15u)) vb = std::min<int>(vb, static_cast<int>(es));
  const bool vec = (vb == 16);
  const int item_bytes = vec ? 16 : static_cast<int>(es);
  fill_dev_job(a.job, plan.box, es, item_bytes, 0);
  a.total_items = plan.out_elems > 0 ? plan.box.rows() * a.job.run_vecs : 0;
  blocks = std::max(1, std::min(blocks, kMaxChannels));
  if (!a.do_barrier) {
    if (a.total_items == 0) return;
    blocks = static_cast<int>(std::min<int64_t>((a.total_items + kThreads - 1) / kThreads, 4LL * dc.sm_count));
  }
  if (vec) {
    const bool nvls = use_nvls && a.do_barrier && dc.mc_heap && nvls_supported(dt, op);
    if (nvls && dt == DType::BF16 && op == ReduceOp::SUM) {
      slab_reduce_vec_kernel<DType::BF16, ReduceOp::SUM, NvlsKind::ADD_BF16><<<blocks, kThreads, 0, stream>>>(a);
    } else if (nvls && dt == DType::F16 && op == ReduceOp::SUM) {
      slab_reduce_vec_kernel<DType::F16, ReduceOp::SUM, NvlsKind::ADD_F16><<<blocks, kThreads, 0, stream>>>(a);
    } else if (nvls && dt == DType::F32 && op == ReduceOp::SUM) {
      slab_reduce_vec_kernel<DType::F32, ReduceOp::SUM, NvlsKind::ADD_F32><<<blocks, kThreads, 0, stream>>>(a);
    } else {
      M4T_DISPATCH_DTYPE_OP(dt, op, LaunchReduceVec, a, blocks, stream);
    }
  } else {
    M4T_DISPATCH_DTYPE_OP(dt, op, LaunchReduceElem, a, blocks, stream);
  }
  check_launch("slab_reduce");
}

}  // namespace m4t
