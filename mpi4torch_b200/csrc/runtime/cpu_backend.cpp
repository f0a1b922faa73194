#include "cpu_backend.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>

#include <omp.h>

#include "reduce_ops.h"

namespace m4t {

namespace {

// Large copies / reductions are memory-bound; a rank may use the OpenMP threads the launcher
// left it (OMP_NUM_THREADS = cores / ranks).  Below the threshold one thread is faster.
constexpr size_t kParallelBytes = 1u << 20;

inline void par_memcpy(void* dst, const void* src, size_t bytes) {
  if (bytes < kParallelBytes || omp_get_max_threads() <= 1 || omp_in_parallel()) {
    std::memcpy(dst, src, bytes);
    return;
  }
  constexpr size_t kChunk = 256u << 10;
  const int64_t chunks = static_cast<int64_t>((bytes + kChunk - 1) / kChunk);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < chunks; ++c) {
    const size_t off = static_cast<size_t>(c) * kChunk;
    std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, std::min(kChunk, bytes - off));
  }
}

constexpr size_t kMinArena = 1u << 20;
constexpr int64_t kTwoPhaseBytes = 256 * 1024;

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename A> inline A scale_acc(A v, double s) {
  return static_cast<A>(static_cast<double>(v) * s);
}
template <> inline float scale_acc<float>(float v, double s) { return v * static_cast<float>(s); }

// out[i] = epi(combine_k srcs[k][i]) for i in [lo, hi).  Sources are combined in index order
// k = 0..nsrc-1 per element (identical result on every rank); the loops are blocked so that
// the inner ones run over contiguous elements of ONE source and vectorise.
template <DType DT, ReduceOp OP> struct CpuReduceRange {
  using E = Elem<DT>;
  using S = typename E::storage;
  using A = typename E::acc;
  using C = Combine<OP, A, E::is_float>;

  static void run_serial(const void* const* srcs, int nsrc, S* o, const S* accp, bool scaled, double s, int64_t lo,
                         int64_t hi) {
    constexpr int kBlock = 512;
    A buf[kBlock];
    for (int64_t b0 = lo; b0 < hi; b0 += kBlock) {
      const int n = static_cast<int>(std::min<int64_t>(kBlock, hi - b0));
      const S* s0 = static_cast<const S*>(srcs[0]) + b0;
      for (int j = 0; j < n; ++j) buf[j] = normalise_single<OP, A>(E::load(s0[j]));
      for (int k = 1; k < nsrc; ++k) {
        const S* sk = static_cast<const S*>(srcs[k]) + b0;
        for (int j = 0; j < n; ++j) buf[j] = C::apply(buf[j], E::load(sk[j]));
      }
      if (scaled)
        for (int j = 0; j < n; ++j) buf[j] = scale_acc<A>(buf[j], s);
      if (accp)
        for (int j = 0; j < n; ++j) buf[j] = buf[j] + E::load(accp[b0 + j]);
      for (int j = 0; j < n; ++j) o[b0 + j] = E::store(buf[j]);
    }
  }

  static void run(const void* const* srcs, int nsrc, void* out, int64_t lo, int64_t hi,
                  const Epilogue* epi) {
    S* o = static_cast<S*>(out);
    const S* accp = (epi && epi->accumulate) ? static_cast<const S*>(epi->accumulate) : nullptr;
    const bool scaled = epi && epi->has_scale;
    const double s = epi ? epi->scale : 1.0;
    const int64_t n = hi - lo;
    const int threads = omp_get_max_threads();
    if (n <= 0) return;
    if (static_cast<size_t>(n) * sizeof(S) < kParallelBytes || threads <= 1 || omp_in_parallel()) {
      run_serial(srcs, nsrc, o, accp, scaled, s, lo, hi);
      return;
    }
    const int64_t per = ((n + threads - 1) / threads + 511) / 512 * 512;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < threads; ++t) {
      const int64_t a = lo + static_cast<int64_t>(t) * per;
      const int64_t b = std::min(hi, a + per);
      if (a < b) run_serial(srcs, nsrc, o, accp, scaled, s, a, b);
    }
  }
};

// dst[i] = acc[i] + src[i]  (phase-2 epilogue of the two-phase allreduce)
template <DType DT, ReduceOp OP> struct CpuAccumulateCopy {
  static void run(const void* src, const void* acc, void* dst, int64_t lo, int64_t hi) {
    using E = Elem<DT>;
    using S = typename E::storage;
    const S* s = static_cast<const S*>(src);
    const S* a = static_cast<const S*>(acc);
    S* d = static_cast<S*>(dst);
    for (int64_t i = lo; i < hi; ++i) d[i] = E::store(E::load(a[i]) + E::load(s[i]));
  }
};

inline void copy_rows(const SlabJob& j, const char* src, char* dst, int64_t es) {
  const size_t run_bytes = static_cast<size_t>(j.run * es);
  for (int64_t i0 = 0; i0 < j.n[0]; ++i0)
    for (int64_t i1 = 0; i1 < j.n[1]; ++i1)
      for (int64_t i2 = 0; i2 < j.n[2]; ++i2) {
        const int64_t so = j.src_off + i0 * j.ss[0] + i1 * j.ss[1] + i2 * j.ss[2];
        const int64_t d_o = j.dst_off + i0 * j.ds[0] + i1 * j.ds[1] + i2 * j.ds[2];
        std::memcpy(dst + d_o * es, src + so * es, run_bytes);
      }
}

}  // namespace

CpuBackend::CpuBackend(Control& ctl) : ctl_(ctl) {}

CpuBackend::~CpuBackend() {
  for (int par = 0; par < 2; ++par) {
    if (own_[par].ptr) {
      munmap(own_[par].ptr, own_[par].cap);
      shm_unlink(arena_name(ctl_.rank(), par, own_[par].gen).c_str());
    }
    for (int p = 0; p < kMaxRanks; ++p)
      if (peers_[p][par].ptr) munmap(peers_[p][par].ptr, peers_[p][par].cap);
  }
  // drop payload segments of messages that were never received
  for (int s = 0; s < kMaxRanks; ++s)
    for (auto& d : unexpected_[s])
      if (d.kind == 1) shm_unlink(msg_name(s, ctl_.rank(), d.seq).c_str());
}

std::string CpuBackend::arena_name(int rank, int par, uint64_t gen) const {
  return ctl_.seg_name("a" + std::to_string(rank) + "_" + std::to_string(par) + "_" + std::to_string(gen));
}

std::string CpuBackend::msg_name(int src, int dst, uint64_t seq) const {
  return ctl_.seg_name("m" + std::to_string(src) + "_" + std::to_string(dst) + "_" + std::to_string(seq));
}

char* CpuBackend::stage(int par, size_t bytes) {
  Mapping& m = own_[par];
  if (bytes <= m.cap) return m.ptr;
  const size_t newcap = round_up(std::max({bytes, 2 * m.cap, kMinArena}), 4096);
  const uint64_t newgen = m.gen + 1;
  const std::string name = arena_name(ctl_.rank(), par, newgen);
  shm_unlink(name.c_str());
  int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  M4T_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << std::strerror(errno));
  M4T_CHECK(ftruncate(fd, static_cast<off_t>(newcap)) == 0,
            "ftruncate(" << newcap << ") failed: " << std::strerror(errno) << " (is /dev/shm large enough?)");
  void* p = mmap(nullptr, newcap, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  M4T_CHECK(p != MAP_FAILED, "mmap of arena failed: " << std::strerror(errno));
  if (m.ptr) {
    // Safe: by the parity argument no peer can still be reading the arena that
    // served the op two steps ago.
    munmap(m.ptr, m.cap);
    shm_unlink(arena_name(ctl_.rank(), par, m.gen).c_str());
  }
  m.ptr = static_cast<char*>(p);
  m.cap = newcap;
  m.gen = newgen;
  auto& slot = ctl_.block()->slots[ctl_.rank()];
  slot.arena_cap[par].store(newcap, std::memory_order_relaxed);
  slot.arena_gen[par].store(newgen, std::memory_order_release);
  return m.ptr;
}

const char* CpuBackend::peer_arena(int peer, int par) {
  if (peer == ctl_.rank()) return own_[par].ptr;
  auto& slot = ctl_.block()->slots[peer];
  const uint64_t gen = slot.arena_gen[par].load(std::memory_order_acquire);
  Mapping& m = peers_[peer][par];
  if (gen != m.gen) {
    if (m.ptr) munmap(m.ptr, m.cap);
    m = Mapping{};
    const size_t cap = slot.arena_cap[par].load(std::memory_order_relaxed);
    const std::string name = arena_name(peer, par, gen);
    int fd = shm_open(name.c_str(), O_RDWR, 0600);
    M4T_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << std::strerror(errno));
    void* p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    M4T_CHECK(p != MAP_FAILED, "mmap of peer arena failed: " << std::strerror(errno));
    m.ptr = static_cast<char*>(p);
    m.cap = cap;
    m.gen = gen;
  }
  return m.ptr;
}

void CpuBackend::allreduce(const void* in, void* out, int64_t n, DType dt, ReduceOp op,
                           const Epilogue& epi, void*) {
  check_op_dtype(op, dt);
  const int P = ctl_.size(), r = ctl_.rank();
  const int64_t es = dtype_size(dt);
  const size_t bytes = static_cast<size_t>(n * es);
  if (P == 1) {
    const void* srcs[1] = {in};
    M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceRange, srcs, 1, out, 0, n, &epi);
    return;
  }
  const int par = static_cast<int>(op_seq_++ & 1);
  const bool two_phase = static_cast<int64_t>(bytes) >= kTwoPhaseBytes;
  char* mine = stage(par, two_phase ? 2 * bytes : bytes);
  if (bytes) par_memcpy(mine, in, bytes);
  ctl_.barrier();
  const void* srcs[kMaxRanks];
  for (int p = 0; p < P; ++p) srcs[p] = peer_arena(p, par);
  if (!two_phase) {
    M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceRange, srcs, P, out, 0, n, &epi);
    return;
  }
  // phase 1: reduce my slice into the result half of my arena (scale fused)
  Epilogue e1;
  e1.scale = epi.scale;
  e1.has_scale = epi.has_scale;
  const int64_t lo = n * r / P, hi = n * (r + 1) / P;
  M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceRange, srcs, P, mine + bytes, lo, hi, &e1);
  ctl_.barrier();
  // phase 2: collect every owner's slice (accumulate fused)
  for (int p = 0; p < P; ++p) {
    const int64_t plo = n * p / P, phi = n * (p + 1) / P;
    const char* res = peer_arena(p, par) + bytes;
    if (epi.accumulate) {
      M4T_DISPATCH_DTYPE_OP(dt, ReduceOp::SUM, CpuAccumulateCopy, res, epi.accumulate, out, plo, phi);
    } else {
      par_memcpy(static_cast<char*>(out) + plo * es, res + plo * es, static_cast<size_t>((phi - plo) * es));
    }
  }
}

void CpuBackend::bcast(void* buf, int64_t n, DType dt, int root, void*) {
  const int P = ctl_.size(), r = ctl_.rank();
  M4T_CHECK(root >= 0 && root < P, "Bcast_: root " << root << " out of range");
  if (P == 1) return;
  const size_t bytes = static_cast<size_t>(n * dtype_size(dt));
  const int par = static_cast<int>(op_seq_++ & 1);
  if (r == root) {
    char* mine = stage(par, bytes);
    if (bytes) par_memcpy(mine, buf, bytes);
  }
  ctl_.barrier();
  if (r != root && bytes) par_memcpy(buf, peer_arena(root, par), bytes);
}

void CpuBackend::reduce(void* buf, int64_t n, DType dt, ReduceOp op, int root, void*) {
  check_op_dtype(op, dt);
  const int P = ctl_.size(), r = ctl_.rank();
  M4T_CHECK(root >= 0 && root < P, "Reduce_: root " << root << " out of range");
  const size_t bytes = static_cast<size_t>(n * dtype_size(dt));
  if (P == 1) {
    const void* srcs[1] = {buf};
    M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceRange, srcs, 1, buf, 0, n, nullptr);
    return;
  }
  const int par = static_cast<int>(op_seq_++ & 1);
  char* mine = stage(par, bytes);
  if (bytes) par_memcpy(mine, buf, bytes);
  ctl_.barrier();
  if (r == root) {
    const void* srcs[kMaxRanks];
    for (int p = 0; p < P; ++p) srcs[p] = peer_arena(p, par);
    M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceRange, srcs, P, buf, 0, n, nullptr);
  } else if (bytes) {
    std::memset(buf, 0, bytes);
  }
}

void CpuBackend::pull(const PullPlan& plan, const void* in, void* out, DType dt, void*) {
  const int P = ctl_.size(), r = ctl_.rank();
  const int64_t es = dtype_size(dt);
  if (P == 1) {
    for (const auto& j : plan.jobs) copy_rows(j, static_cast<const char*>(in), static_cast<char*>(out), es);
    return;
  }
  const int par = static_cast<int>(op_seq_++ & 1);
  if (plan.stage_elems > 0) {
    const size_t bytes = static_cast<size_t>(plan.stage_elems * es);
    char* mine = stage(par, bytes);
    par_memcpy(mine, in, bytes);
  }
  ctl_.barrier();
  for (const auto& j : plan.jobs) {
    const char* src = (j.peer == r) ? static_cast<const char*>(in) : peer_arena(j.peer, par);
    copy_rows(j, src, static_cast<char*>(out), es);
  }
}

namespace {
template <DType DT, ReduceOp OP> struct CpuReduceBox {
  static void run(const SlabJob& j, const char* const* srcs, int nsrc, char* out, const Epilogue* epi) {
    const int64_t es = dtype_size(DT);
    for (int64_t i0 = 0; i0 < j.n[0]; ++i0)
      for (int64_t i1 = 0; i1 < j.n[1]; ++i1)
        for (int64_t i2 = 0; i2 < j.n[2]; ++i2) {
          const int64_t so = j.src_off + i0 * j.ss[0] + i1 * j.ss[1] + i2 * j.ss[2];
          const int64_t d_o = j.dst_off + i0 * j.ds[0] + i1 * j.ds[1] + i2 * j.ds[2];
          const void* row_srcs[kMaxRanks];
          for (int k = 0; k < nsrc; ++k) row_srcs[k] = srcs[k] + so * es;
          Epilogue e;
          if (epi) {
            e = *epi;
            if (e.accumulate) e.accumulate = static_cast<const char*>(e.accumulate) + d_o * es;
          }
          CpuReduceRange<DT, OP>::run(row_srcs, nsrc, out + d_o * es, 0, j.run, epi ? &e : nullptr);
        }
  }
};
}  // namespace

void CpuBackend::reduce_pull(const ReducePlan& plan, const void* in, void* out, DType dt, ReduceOp op,
                             const Epilogue& epi, void*) {
  check_op_dtype(op, dt);
  const int P = ctl_.size(), r = ctl_.rank();
  const int64_t es = dtype_size(dt);
  const char* srcs[kMaxRanks];
  if (P == 1) {
    srcs[0] = static_cast<const char*>(in);
    if (plan.out_elems > 0) {
      M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceBox, plan.box, srcs, 1, static_cast<char*>(out), &epi);
    }
    return;
  }
  const int par = static_cast<int>(op_seq_++ & 1);
  const size_t bytes = static_cast<size_t>(plan.stage_elems * es);
  char* mine = stage(par, bytes);
  if (bytes) par_memcpy(mine, in, bytes);
  ctl_.barrier();
  if (plan.out_elems == 0) return;
  for (int p = 0; p < P; ++p) srcs[p] = (p == r) ? static_cast<const char*>(in) : peer_arena(p, par);
  M4T_DISPATCH_DTYPE_OP(dt, op, CpuReduceBox, plan.box, srcs, P, static_cast<char*>(out), &epi);
}

// ---------------------------------------------------------------------------
// point-to-point: eager, buffered sends through per-pair descriptor rings
// ---------------------------------------------------------------------------
int64_t CpuBackend::isend(const void* buf, int64_t bytes, int dest, int64_t tag, void*) {
  const int P = ctl_.size(), r = ctl_.rank();
  M4T_CHECK(dest >= 0 && dest < P, "Isend: destination rank " << dest << " out of range");
  PairRing& ring = ctl_.block()->rings[r][dest];
  MsgDesc d{};
  d.seq = ++send_seq_[dest];
  d.tag = tag;
  d.bytes = static_cast<uint64_t>(bytes);
  if (bytes <= kInlineBytes) {
    d.kind = 0;
    if (bytes) std::memcpy(d.inline_data, buf, static_cast<size_t>(bytes));
  } else {
    d.kind = 1;
    const std::string name = msg_name(r, dest, d.seq);
    shm_unlink(name.c_str());
    int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    M4T_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << std::strerror(errno));
    M4T_CHECK(ftruncate(fd, static_cast<off_t>(bytes)) == 0,
              "ftruncate(" << bytes << ") failed: " << std::strerror(errno));
    void* p = mmap(nullptr, static_cast<size_t>(bytes), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    M4T_CHECK(p != MAP_FAILED, "mmap of message segment failed: " << std::strerror(errno));
    par_memcpy(p, buf, static_cast<size_t>(bytes));
    munmap(p, static_cast<size_t>(bytes));
  }
  const uint64_t head = ring.head.load(std::memory_order_relaxed);
  ctl_.wait_until([&] { return head - ring.tail.load(std::memory_order_acquire) < kMailboxDepth; },
                  "space in the send ring");
  ring.entries[head % kMailboxDepth] = d;
  ring.head.store(head + 1, std::memory_order_release);
  const int64_t id = next_request_++;
  Request rq;
  rq.is_recv = false;
  rq.peer = dest;
  rq.tag = tag;
  rq.bytes = bytes;
  requests_[id] = rq;
  return id;
}

int64_t CpuBackend::irecv(void* buf, int64_t bytes, int source, int64_t tag, void*) {
  const int P = ctl_.size();
  M4T_CHECK(source >= 0 && source < P, "Irecv: source rank " << source << " out of range");
  const int64_t id = next_request_++;
  Request rq;
  rq.is_recv = true;
  rq.buf = buf;
  rq.bytes = bytes;
  rq.peer = source;
  rq.tag = tag;
  requests_[id] = rq;
  return id;
}

void CpuBackend::deliver(const MsgDesc& d, int source, Request& rq) {
  M4T_CHECK(static_cast<int64_t>(d.bytes) <= rq.bytes,
            "message truncated: " << d.bytes << " bytes sent by rank " << source << " (tag " << d.tag
                                  << ") into a " << rq.bytes << "-byte receive buffer");
  if (d.kind == 0) {
    if (d.bytes) std::memcpy(rq.buf, d.inline_data, d.bytes);
    return;
  }
  const std::string name = msg_name(source, ctl_.rank(), d.seq);
  int fd = shm_open(name.c_str(), O_RDWR, 0600);
  M4T_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << std::strerror(errno));
  void* p = mmap(nullptr, d.bytes, PROT_READ, MAP_SHARED, fd, 0);
  close(fd);
  M4T_CHECK(p != MAP_FAILED, "mmap of message segment failed: " << std::strerror(errno));
  par_memcpy(rq.buf, p, d.bytes);
  munmap(p, d.bytes);
  shm_unlink(name.c_str());
}

void CpuBackend::wait(int64_t request, void*) {
  auto it = requests_.find(request);
  M4T_CHECK(it != requests_.end(), "Wait: unknown or already completed request " << request
                                       << " (a WaitHandle may only be waited on once)");
  Request rq = it->second;
  requests_.erase(it);
  if (!rq.is_recv) return;  // sends are buffered: complete at post time
  const int src = rq.peer;
  auto& pending = unexpected_[src];
  for (auto u = pending.begin(); u != pending.end(); ++u) {
    if (u->tag == rq.tag) {
      deliver(*u, src, rq);
      pending.erase(u);
      return;
    }
  }
  PairRing& ring = ctl_.block()->rings[src][ctl_.rank()];
  for (;;) {
    const uint64_t tail = ring.tail.load(std::memory_order_relaxed);
    ctl_.wait_until([&] { return ring.head.load(std::memory_order_acquire) > tail; }, "a matching message");
    MsgDesc d = ring.entries[tail % kMailboxDepth];
    ring.tail.store(tail + 1, std::memory_order_release);
    if (d.tag == rq.tag) {
      deliver(d, src, rq);
      return;
    }
    pending.push_back(d);
  }
}

}  // namespace m4t
