// mpi_shim.cpp — a small single-node MPI over POSIX shared memory.
//
// Purpose: the image has no MPI, and the reference (helmholtz-analytics/mpi4torch) builds only
// through mpicc/mpicxx against <mpi.h>.  This library provides the ~35 entry points the
// reference calls so the UNMODIFIED reference can be built and run as the comparison arm of
// bench.py.  It is written to be a fair stand-in for a stock (non CUDA-aware) MPI's shared-memory
// path: chunked, double-buffered collectives where every rank reduces 1/P of each chunk,
// derived datatypes (vector + resized) packed/unpacked exactly as the standard prescribes, and a
// chunk-ring point-to-point engine with tag matching and an unexpected-message queue.
//
// Process discovery (first match wins):
//   MPISHIM_RANK / MPISHIM_SIZE / MPISHIM_JOB     (set by ../bin/mpirun)
//   OMPI_COMM_WORLD_RANK / _SIZE, PMI_RANK / PMI_SIZE
//   RANK / WORLD_SIZE / MASTER_PORT               (torchrun)
//   otherwise a singleton world of size 1.
#include "mpi.h"

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------ util
double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

[[noreturn]] void die(const char* what) {
  std::fprintf(stderr, "[mpishim] fatal: %s\n", what);
  std::fflush(stderr);
  _exit(86);
}

int64_t env_i64(const char* name, int64_t dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoll(v) : dflt;
}

inline void cpu_relax() {
#if defined(__x86_64__)
  __builtin_ia32_pause();
#else
  std::atomic_signal_fence(std::memory_order_seq_cst);
#endif
}

// ------------------------------------------------------------------------------- shared segment
constexpr uint64_t kMagic = 0x4d50495348494d31ull;  // "MPISHIM1"
constexpr int kMaxRanks = 64;
constexpr int kRingChunks = 4;

struct alignas(64) PaddedU64 {
  std::atomic<uint64_t> v;
  char pad[56];
};

struct alignas(64) ChanHdr {  // one per ordered (src, dst) pair
  std::atomic<uint64_t> head;  // chunks produced (sender)
  char pad0[56];
  std::atomic<uint64_t> tail;  // chunks consumed (receiver)
  char pad1[56];
};

struct ChunkHdr {
  int32_t tag;
  int32_t first;  // 1 on the first chunk of a message
  int64_t total;  // message bytes
  int64_t nbytes;  // payload bytes in this chunk
  char pad[40];
};
static_assert(sizeof(ChunkHdr) == 64, "chunk header is one cache line");

struct ShmHeader {
  std::atomic<uint64_t> magic;
  int32_t size;
  int32_t pad0;
  int64_t slot_bytes;
  int64_t chunk_bytes;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> abort_flag;
  alignas(64) std::atomic<uint32_t> bar_count;
  alignas(64) std::atomic<uint32_t> bar_gen;
  alignas(64) PaddedU64 meta[kMaxRanks];  // per-rank scalar published before a barrier
};

struct World {
  bool initialized = false, finalized = false;
  int rank = 0, size = 1;
  std::string shm_name;
  char* base = nullptr;
  size_t bytes = 0;
  ShmHeader* hdr = nullptr;
  int64_t slot = 0;   // bytes per collective slot
  int64_t chunk = 0;  // payload bytes per p2p chunk
  char* coll = nullptr;  // [2][size + 1][slot]   (slot index `size` = result area)
  ChanHdr* chans = nullptr;  // [size][size]
  char* rings = nullptr;     // [size][size][kRingChunks][64 + chunk]
  uint64_t coll_seq = 0;     // chunk counter -> parity
  double timeout_s = 600;
  std::recursive_mutex mu;
};
World W;

char* coll_slot(int parity, int idx) { return W.coll + (static_cast<int64_t>(parity) * (W.size + 1) + idx) * W.slot; }
ChanHdr& chan(int src, int dst) { return W.chans[src * W.size + dst]; }
char* ring_chunk(int src, int dst, int i) {
  return W.rings + ((static_cast<int64_t>(src) * W.size + dst) * kRingChunks + i) * (64 + W.chunk);
}

void check_abort() {
  if (W.hdr && W.hdr->abort_flag.load(std::memory_order_relaxed)) die("another rank aborted");
}

struct Backoff {
  int spins = 0;
  double t0 = 0;
  void pause() {
    if (++spins < 2000) {
      cpu_relax();
      return;
    }
    if ((spins & 1023) == 0) {
      check_abort();
      if (t0 == 0) t0 = now_s();
      else if (now_s() - t0 > W.timeout_s) die("timeout waiting for a peer (MPISHIM_TIMEOUT_S)");
    }
    sched_yield();
  }
};

void barrier() {
  if (W.size == 1) return;
  ShmHeader* h = W.hdr;
  const uint32_t gen = h->bar_gen.load(std::memory_order_acquire);
  if (h->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == static_cast<uint32_t>(W.size)) {
    h->bar_count.store(0, std::memory_order_relaxed);
    h->bar_gen.store(gen + 1, std::memory_order_release);
  } else {
    Backoff b;
    while (h->bar_gen.load(std::memory_order_acquire) == gen) b.pause();
  }
}

void world_init() {
  if (W.initialized) return;
  const char* r = std::getenv("MPISHIM_RANK");
  const char* s = std::getenv("MPISHIM_SIZE");
  std::string job;
  if (r && s) {
    job = std::getenv("MPISHIM_JOB") ? std::getenv("MPISHIM_JOB") : "job";
  } else if ((r = std::getenv("OMPI_COMM_WORLD_RANK")) && (s = std::getenv("OMPI_COMM_WORLD_SIZE"))) {
    job = "ompi" + std::to_string(getppid());
  } else if ((r = std::getenv("PMI_RANK")) && (s = std::getenv("PMI_SIZE"))) {
    job = "pmi" + std::to_string(getppid());
  } else if ((r = std::getenv("RANK")) && (s = std::getenv("WORLD_SIZE"))) {
    const char* port = std::getenv("MASTER_PORT");
    job = std::string("tr") + (port ? port : "0") + "_" + std::to_string(getppid());
  } else {
    r = "0";
    s = "1";
  }
  W.rank = std::atoi(r);
  W.size = std::atoi(s);
  if (W.size < 1 || W.size > kMaxRanks || W.rank < 0 || W.rank >= W.size) die("bad rank/size in the environment");
  W.timeout_s = static_cast<double>(env_i64("MPISHIM_TIMEOUT_S", 600));
  W.initialized = true;
  if (W.size == 1) return;

  W.slot = env_i64("MPISHIM_SLOT_KB", 4096) * 1024;
  W.chunk = env_i64("MPISHIM_CHUNK_KB", 512) * 1024;
  const size_t hdr_bytes = (sizeof(ShmHeader) + 4095) & ~size_t(4095);
  const size_t chan_bytes = (sizeof(ChanHdr) * W.size * W.size + 4095) & ~size_t(4095);
  const size_t coll_bytes = static_cast<size_t>(2) * (W.size + 1) * W.slot;
  const size_t ring_bytes = static_cast<size_t>(W.size) * W.size * kRingChunks * (64 + W.chunk);
  W.bytes = hdr_bytes + chan_bytes + coll_bytes + ring_bytes;
  W.shm_name = "/mpishim_" + job;

  int fd = -1;
  if (W.rank == 0) {
    shm_unlink(W.shm_name.c_str());
    fd = shm_open(W.shm_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) die("shm_open(create) failed");
    if (ftruncate(fd, static_cast<off_t>(W.bytes)) != 0) die("ftruncate failed");
  } else {
    const double t0 = now_s();
    while (true) {
      fd = shm_open(W.shm_name.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && static_cast<size_t>(st.st_size) == W.bytes) break;
        close(fd);
        fd = -1;
      }
      if (now_s() - t0 > W.timeout_s) die("timeout opening the shared segment");
      usleep(1000);
    }
  }
  void* p = mmap(nullptr, W.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) die("mmap failed");
  W.base = static_cast<char*>(p);
  W.hdr = reinterpret_cast<ShmHeader*>(W.base);
  W.chans = reinterpret_cast<ChanHdr*>(W.base + hdr_bytes);
  W.coll = W.base + hdr_bytes + chan_bytes;
  W.rings = W.coll + coll_bytes;
  if (W.rank == 0) {
    W.hdr->size = W.size;
    W.hdr->slot_bytes = W.slot;
    W.hdr->chunk_bytes = W.chunk;
    W.hdr->magic.store(kMagic, std::memory_order_release);
  } else {
    Backoff b;
    while (W.hdr->magic.load(std::memory_order_acquire) != kMagic) b.pause();
    if (W.hdr->size != W.size || W.hdr->slot_bytes != W.slot || W.hdr->chunk_bytes != W.chunk)
      die("shared segment was created with a different geometry");
  }
  W.hdr->attached.fetch_add(1, std::memory_order_acq_rel);
  Backoff b;
  while (W.hdr->attached.load(std::memory_order_acquire) < static_cast<uint32_t>(W.size)) b.pause();
  // everybody is mapped: the name is no longer needed (no stale segments after a crash)
  if (W.rank == 0) shm_unlink(W.shm_name.c_str());
}

// ------------------------------------------------------------------------------------ datatypes
enum Prim { P_NONE = 0, P_U8, P_I8, P_I16, P_I32, P_I64, P_F32, P_F64, P_U32, P_U64 };

struct Dtype {
  bool used = false;
  Prim prim = P_NONE;
  int64_t size = 0;  // bytes of data in one item
  int64_t lb = 0, extent = 0;
  std::vector<std::pair<int64_t, int64_t>> blocks;  // (byte offset, byte length), in signature order
  std::vector<int64_t> prefix;  // prefix[i] = packed offset of block i
  bool contiguous() const { return blocks.size() == 1 && blocks[0].first == 0 && extent == size; }
  void finish() {
    // merge adjacent blocks, build prefix sums
    std::vector<std::pair<int64_t, int64_t>> m;
    for (auto& b : blocks) {
      if (b.second == 0) continue;
      if (!m.empty() && m.back().first + m.back().second == b.first) m.back().second += b.second;
      else m.push_back(b);
    }
    blocks.swap(m);
    prefix.resize(blocks.size());
    int64_t acc = 0;
    for (size_t i = 0; i < blocks.size(); ++i) {
      prefix[i] = acc;
      acc += blocks[i].second;
    }
    size = acc;
  }
};

std::vector<Dtype> g_types;
constexpr int kFirstDerived = 32;

void types_init() {
  if (!g_types.empty()) return;
  g_types.resize(kFirstDerived);
  auto prim = [&](int h, Prim p, int64_t sz) {
    Dtype& t = g_types[h];
    t.used = true;
    t.prim = p;
    t.extent = sz;
    t.blocks = {{0, sz}};
    t.finish();
  };
  prim(MPI_BYTE, P_U8, 1);
  prim(MPI_CHAR, P_I8, 1);
  prim(MPI_SHORT, P_I16, 2);
  prim(MPI_INT, P_I32, 4);
  prim(MPI_LONG, P_I64, 8);
  prim(MPI_FLOAT, P_F32, 4);
  prim(MPI_DOUBLE, P_F64, 8);
  prim(MPI_UNSIGNED_CHAR, P_U8, 1);
  prim(MPI_UNSIGNED, P_U32, 4);
  prim(MPI_UNSIGNED_LONG, P_U64, 8);
}

const Dtype* get_type(MPI_Datatype h) {
  types_init();
  if (h <= 0 || h >= static_cast<int>(g_types.size()) || !g_types[h].used) return nullptr;
  return &g_types[h];
}

int new_type(Dtype&& t) {
  types_init();
  for (size_t i = kFirstDerived; i < g_types.size(); ++i)
    if (!g_types[i].used) {
      g_types[i] = std::move(t);
      g_types[i].used = true;
      return static_cast<int>(i);
    }
  g_types.push_back(std::move(t));
  g_types.back().used = true;
  return static_cast<int>(g_types.size() - 1);
}

// Copies bytes [off, off+len) of the packed stream of `count` items of type t laid out at `base`
// to (PACK) or from (!PACK) the linear buffer `lin`.
template <bool PACK>
void stream_copy(char* base, const Dtype& t, int64_t off, int64_t len, char* lin) {
  if (len <= 0) return;
  if (t.contiguous()) {
    if (PACK) std::memcpy(lin, base + off, static_cast<size_t>(len));
    else std::memcpy(base + off, lin, static_cast<size_t>(len));
    return;
  }
  int64_t item = off / t.size;
  int64_t rem = off % t.size;
  size_t bi = static_cast<size_t>(std::upper_bound(t.prefix.begin(), t.prefix.end(), rem) - t.prefix.begin()) - 1;
  int64_t in_block = rem - t.prefix[bi];
  while (len > 0) {
    const auto& b = t.blocks[bi];
    const int64_t n = std::min(len, b.second - in_block);
    char* p = base + item * t.extent + b.first + in_block;
    if (PACK) std::memcpy(lin, p, static_cast<size_t>(n));
    else std::memcpy(p, lin, static_cast<size_t>(n));
    lin += n;
    len -= n;
    in_block = 0;
    if (++bi == t.blocks.size()) {
      bi = 0;
      ++item;
    }
  }
}

// ------------------------------------------------------------------------------------ reductions
template <typename T, int OP>
void reduce_into(T* __restrict acc, const T* __restrict in, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    T a = acc[i], b = in[i];
    if constexpr (OP == MPI_MAX) acc[i] = a < b ? b : a;
    else if constexpr (OP == MPI_MIN) acc[i] = b < a ? b : a;
    else if constexpr (OP == MPI_SUM) acc[i] = a + b;
    else if constexpr (OP == MPI_PROD) acc[i] = a * b;
    else if constexpr (OP == MPI_LAND) acc[i] = static_cast<T>((a != T(0)) && (b != T(0)));
    else if constexpr (OP == MPI_LOR) acc[i] = static_cast<T>((a != T(0)) || (b != T(0)));
    else if constexpr (OP == MPI_LXOR) acc[i] = static_cast<T>((a != T(0)) != (b != T(0)));
    else if constexpr (std::is_integral<T>::value) {
      if constexpr (OP == MPI_BAND) acc[i] = a & b;
      else if constexpr (OP == MPI_BOR) acc[i] = a | b;
      else if constexpr (OP == MPI_BXOR) acc[i] = a ^ b;
    }
  }
}

template <typename T>
bool reduce_typed(int op, void* acc, const void* in, int64_t n) {
  T* a = static_cast<T*>(acc);
  const T* b = static_cast<const T*>(in);
  switch (op) {
    case MPI_MAX: reduce_into<T, MPI_MAX>(a, b, n); return true;
    case MPI_MIN: reduce_into<T, MPI_MIN>(a, b, n); return true;
    case MPI_SUM: reduce_into<T, MPI_SUM>(a, b, n); return true;
    case MPI_PROD: reduce_into<T, MPI_PROD>(a, b, n); return true;
    case MPI_LAND: reduce_into<T, MPI_LAND>(a, b, n); return true;
    case MPI_LOR: reduce_into<T, MPI_LOR>(a, b, n); return true;
    case MPI_LXOR: reduce_into<T, MPI_LXOR>(a, b, n); return true;
    case MPI_BAND:
      if (!std::is_integral<T>::value) return false;
      reduce_into<T, MPI_BAND>(a, b, n);
      return true;
    case MPI_BOR:
      if (!std::is_integral<T>::value) return false;
      reduce_into<T, MPI_BOR>(a, b, n);
      return true;
    case MPI_BXOR:
      if (!std::is_integral<T>::value) return false;
      reduce_into<T, MPI_BXOR>(a, b, n);
      return true;
    default: return false;
  }
}

bool reduce_bytes(Prim p, int op, void* acc, const void* in, int64_t nelem) {
  switch (p) {
    case P_U8: return reduce_typed<uint8_t>(op, acc, in, nelem);
    case P_I8: return reduce_typed<int8_t>(op, acc, in, nelem);
    case P_I16: return reduce_typed<int16_t>(op, acc, in, nelem);
    case P_I32: return reduce_typed<int32_t>(op, acc, in, nelem);
    case P_I64: return reduce_typed<int64_t>(op, acc, in, nelem);
    case P_U32: return reduce_typed<uint32_t>(op, acc, in, nelem);
    case P_U64: return reduce_typed<uint64_t>(op, acc, in, nelem);
    case P_F32: return reduce_typed<float>(op, acc, in, nelem);
    case P_F64: return reduce_typed<double>(op, acc, in, nelem);
    default: return false;
  }
}

bool valid_comm(MPI_Comm c) { return c == MPI_COMM_WORLD || c == MPI_COMM_SELF; }
int comm_size(MPI_Comm c) { return c == MPI_COMM_SELF ? 1 : W.size; }
int comm_rank(MPI_Comm c) { return c == MPI_COMM_SELF ? 0 : W.rank; }

// Reduction collective shared by Allreduce and Reduce: chunked, double-buffered; in every chunk
// each rank copies in, reduces 1/P of the chunk over all P slots into the result area, and the
// receivers copy the result out.  Two barriers per chunk.
int reduce_collective(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, bool all,
                      MPI_Comm comm) {
  const Dtype* t = get_type(datatype);
  if (!t || t->prim == P_NONE) return MPI_ERR_TYPE;
  if (op < MPI_MAX || op > MPI_BXOR) return MPI_ERR_OP;
  if (count < 0) return MPI_ERR_ARG;
  const int P = comm_size(comm), me = comm_rank(comm);
  const int64_t es = t->size;
  const bool i_recv = all || me == root;
  const char* src = static_cast<const char*>(sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf);
  if (P == 1) {
    if (sendbuf != MPI_IN_PLACE && count > 0) std::memmove(recvbuf, sendbuf, static_cast<size_t>(count) * es);
    return MPI_SUCCESS;
  }
  // probe op/type compatibility once so that every rank fails the same way
  {
    double a = 0, b = 0;
    if (!reduce_bytes(t->prim, op, &a, &b, 0)) return MPI_ERR_OP;
  }
  const int64_t total = static_cast<int64_t>(count) * es;
  const int64_t chunk_elems = W.slot / es;
  for (int64_t off = 0; off < total || (total == 0 && off == 0); off += chunk_elems * es) {
    const int64_t nb = std::min(total - off, chunk_elems * es);
    const int64_t ne = nb / es;
    const int par = static_cast<int>(W.coll_seq++ & 1);
    if (nb > 0) std::memcpy(coll_slot(par, me), src + off, static_cast<size_t>(nb));
    barrier();
    // my slice of the chunk, in elements, aligned to 64 bytes
    const int64_t align = std::max<int64_t>(1, 64 / es);
    int64_t per = (ne + P - 1) / P;
    per = (per + align - 1) / align * align;
    const int64_t lo = std::min(ne, per * me), hi = std::min(ne, per * (me + 1));
    if (hi > lo) {
      char* res = coll_slot(par, P) + lo * es;
      std::memcpy(res, coll_slot(par, 0) + lo * es, static_cast<size_t>(hi - lo) * es);
      for (int r = 1; r < P; ++r) reduce_bytes(t->prim, op, res, coll_slot(par, r) + lo * es, hi - lo);
    }
    barrier();
    if (i_recv && nb > 0) std::memcpy(static_cast<char*>(recvbuf) + off, coll_slot(par, P), static_cast<size_t>(nb));
    if (total == 0) break;
  }
  return MPI_SUCCESS;
}

// One party's packed stream: `count` items of `type` at `base + displ * extent(type)`.
struct Stream {
  char* base = nullptr;
  const Dtype* type = nullptr;
  int64_t bytes = 0;
};

Stream make_stream(const void* buf, int64_t count, const Dtype* t, int64_t displ) {
  Stream s;
  s.type = t;
  s.base = const_cast<char*>(static_cast<const char*>(buf)) + displ * t->extent;
  s.bytes = count * t->size;
  return s;
}

// Generic "slot per non-root party" exchange used by gather(v), allgather(v) and scatter(v).
//   mode 0: gather to root     — party r's send stream lands in root's recv stream r
//   mode 1: allgather          — party r's send stream lands in everyone's recv stream r
//   mode 2: scatter from root  — root's send stream r lands in party r's recv stream
// Streams longer than a slot are moved in rounds; the number of rounds is agreed through the
// per-rank meta words.
int exchange(int mode, int root, const Stream* send /*1 or P*/, const Stream* recv /*1 or P*/, MPI_Comm comm) {
  const int P = comm_size(comm), me = comm_rank(comm);
  if (P == 1) {
    const Stream& s = send[0];
    const Stream& r = recv[0];
    if (s.bytes != r.bytes) return MPI_ERR_TRUNCATE;
    std::vector<char> tmp(static_cast<size_t>(s.bytes));
    stream_copy<true>(s.base, *s.type, 0, s.bytes, tmp.data());
    stream_copy<false>(r.base, *r.type, 0, r.bytes, tmp.data());
    return MPI_SUCCESS;
  }
  // publish the longest stream this rank sends
  int64_t mymax = 0;
  if (mode == 2) {
    if (me == root)
      for (int r = 0; r < P; ++r) mymax = std::max(mymax, send[r].bytes);
  } else {
    mymax = send[0].bytes;
  }
  W.hdr->meta[me].v.store(static_cast<uint64_t>(mymax), std::memory_order_relaxed);
  barrier();
  int64_t longest = 0;
  for (int r = 0; r < P; ++r) longest = std::max<int64_t>(longest, static_cast<int64_t>(W.hdr->meta[r].v.load(std::memory_order_relaxed)));
  int rc = MPI_SUCCESS;
  const int64_t rounds = std::max<int64_t>(1, (longest + W.slot - 1) / W.slot);
  for (int64_t k = 0; k < rounds; ++k) {
    const int64_t off = k * W.slot;
    const int par = static_cast<int>(W.coll_seq++ & 1);
    if (mode == 2) {
      if (me == root)
        for (int r = 0; r < P; ++r) {
          const int64_t n = std::min(W.slot, send[r].bytes - off);
          if (n > 0) stream_copy<true>(send[r].base, *send[r].type, off, n, coll_slot(par, r));
        }
    } else {
      const int64_t n = std::min(W.slot, send[0].bytes - off);
      if (n > 0) stream_copy<true>(send[0].base, *send[0].type, off, n, coll_slot(par, me));
    }
    barrier();
    if (mode == 2) {
      const int64_t n = std::min(W.slot, recv[0].bytes - off);
      if (n > 0) stream_copy<false>(recv[0].base, *recv[0].type, off, n, coll_slot(par, me));
    } else if (mode == 1 || me == root) {
      for (int r = 0; r < P; ++r) {
        const int64_t n = std::min(W.slot, recv[r].bytes - off);
        if (n > 0) stream_copy<false>(recv[r].base, *recv[r].type, off, n, coll_slot(par, r));
      }
    }
    // the next round uses the other parity; the round after that is separated from this round's
    // readers by the next round's barrier
  }
  // Buffer reuse: a slot of parity p is rewritten two rounds later at the earliest, and at least
  // two barriers (this op's next round or the next collective's first barriers) lie in between,
  // so no rank can still be reading it.
  return rc;
}

// ------------------------------------------------------------------------------- point to point
struct Request {
  bool active = false, done = false, is_send = false;
  int peer = 0, tag = 0;
  char* buf = nullptr;
  int64_t bytes = 0;     // send: message size, recv: capacity
  int64_t moved = 0;     // bytes pushed (send) / received (recv)
  int64_t msg_bytes = 0;  // recv: size of the matched message
  bool matched = false;  // recv: bound to an incoming message
  int err = MPI_SUCCESS;
  MPI_Status st{};
};

struct Unexpected {  // message that arrived before a matching receive was posted
  int src = 0, tag = 0;
  int64_t total = 0, have = 0;
  std::vector<char> data;
  int claimed_by = 0;  // request id that will take the data when complete (0 = none)
};

std::vector<Request> g_reqs(1);  // id 0 = MPI_REQUEST_NULL
std::vector<std::deque<int>> g_sendq;  // per destination: pending send request ids (FIFO)
std::vector<std::deque<int>> g_recvq;  // per source: posted, unmatched receives (FIFO)
std::deque<Unexpected*> g_unexp;
// per source: where the chunks of the message currently arriving go
struct InFlight {
  int req = 0;
  Unexpected* ux = nullptr;
  int64_t left = 0;
};
std::vector<InFlight> g_inflight;

int new_request() {
  for (size_t i = 1; i < g_reqs.size(); ++i)
    if (!g_reqs[i].active) {
      g_reqs[i] = Request();
      g_reqs[i].active = true;
      return static_cast<int>(i);
    }
  g_reqs.emplace_back();
  g_reqs.back().active = true;
  return static_cast<int>(g_reqs.size() - 1);
}

void p2p_init() {
  if (g_sendq.empty()) {
    g_sendq.resize(W.size);
    g_recvq.resize(W.size);
    g_inflight.resize(W.size);
  }
}

void finish_recv(Request& rq, int src, int tag, int64_t total) {
  rq.done = true;
  rq.st.MPI_SOURCE = src;
  rq.st.MPI_TAG = tag;
  rq.st.shim_bytes = total;
  rq.st.MPI_ERROR = rq.err;
}

// Advances every pending transfer as far as possible without blocking; true if anything moved.
bool progress() {
  bool any = false;
  const int P = W.size, me = W.rank;
  if (P == 1) return false;
  // sends: only the oldest pending message per destination may use the channel (FIFO order)
  for (int d = 0; d < P; ++d) {
    while (!g_sendq[d].empty()) {
      Request& rq = g_reqs[g_sendq[d].front()];
      ChanHdr& c = chan(me, d);
      bool blocked = false;
      while (rq.moved < rq.bytes || (rq.bytes == 0 && !rq.done)) {
        const uint64_t head = c.head.load(std::memory_order_relaxed);
        if (head - c.tail.load(std::memory_order_acquire) >= kRingChunks) {
          blocked = true;
          break;
        }
        char* slot = ring_chunk(me, d, static_cast<int>(head % kRingChunks));
        ChunkHdr* h = reinterpret_cast<ChunkHdr*>(slot);
        const int64_t n = std::min(W.chunk, rq.bytes - rq.moved);
        h->tag = rq.tag;
        h->first = rq.moved == 0;
        h->total = rq.bytes;
        h->nbytes = n;
        if (n > 0) std::memcpy(slot + 64, rq.buf + rq.moved, static_cast<size_t>(n));
        rq.moved += n;
        c.head.store(head + 1, std::memory_order_release);
        any = true;
        if (rq.bytes == 0) break;
      }
      if (blocked) break;
      rq.done = true;  // buffered-send semantics: the user buffer is reusable
      g_sendq[d].pop_front();
    }
  }
  // receives
  for (int s = 0; s < P; ++s) {
    ChanHdr& c = chan(s, me);
    while (true) {
      const uint64_t tail = c.tail.load(std::memory_order_relaxed);
      if (c.head.load(std::memory_order_acquire) == tail) break;
      char* slot = ring_chunk(s, me, static_cast<int>(tail % kRingChunks));
      const ChunkHdr* h = reinterpret_cast<const ChunkHdr*>(slot);
      InFlight& fl = g_inflight[s];
      if (h->first) {
        fl = InFlight();
        fl.left = h->total;
        // first posted receive from s whose tag matches
        for (auto it = g_recvq[s].begin(); it != g_recvq[s].end(); ++it) {
          Request& rq = g_reqs[*it];
          if (rq.tag == MPI_ANY_TAG || rq.tag == h->tag) {
            fl.req = *it;
            rq.matched = true;
            rq.msg_bytes = h->total;
            if (h->total > rq.bytes) rq.err = MPI_ERR_TRUNCATE;
            g_recvq[s].erase(it);
            break;
          }
        }
        if (!fl.req) {
          fl.ux = new Unexpected();
          fl.ux->src = s;
          fl.ux->tag = h->tag;
          fl.ux->total = h->total;
          fl.ux->data.resize(static_cast<size_t>(h->total));
          g_unexp.push_back(fl.ux);
        }
      }
      if (fl.req) {
        Request& rq = g_reqs[fl.req];
        const int64_t room = std::max<int64_t>(0, rq.bytes - rq.moved);
        const int64_t n = std::min(room, h->nbytes);
        if (n > 0) std::memcpy(rq.buf + rq.moved, slot + 64, static_cast<size_t>(n));
        rq.moved += n;
      } else {
        std::memcpy(fl.ux->data.data() + fl.ux->have, slot + 64, static_cast<size_t>(h->nbytes));
        fl.ux->have += h->nbytes;
      }
      fl.left -= h->nbytes;
      const int tag = h->tag;
      const int64_t total = h->total;
      c.tail.store(tail + 1, std::memory_order_release);
      any = true;
      if (fl.left <= 0) {
        if (fl.req) {
          finish_recv(g_reqs[fl.req], s, tag, total);
        } else if (fl.ux->claimed_by) {
          Request& rq = g_reqs[fl.ux->claimed_by];
          const int64_t n = std::min(rq.bytes, fl.ux->total);
          if (n > 0) std::memcpy(rq.buf, fl.ux->data.data(), static_cast<size_t>(n));
          if (fl.ux->total > rq.bytes) rq.err = MPI_ERR_TRUNCATE;
          rq.moved = n;
          finish_recv(rq, s, tag, total);
          g_unexp.erase(std::find(g_unexp.begin(), g_unexp.end(), fl.ux));
          delete fl.ux;
        }
        fl = InFlight();
      }
    }
  }
  return any;
}

void wait_request(int id) {
  Request& rq = g_reqs[id];
  Backoff b;
  while (!rq.done) {
    if (progress()) b = Backoff();
    else b.pause();
  }
}

int start_send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request) {
  const Dtype* t = get_type(datatype);
  if (!t) return MPI_ERR_TYPE;
  if (!t->contiguous()) return MPI_ERR_TYPE;  // point-to-point: contiguous types only
  if (!valid_comm(comm) || count < 0) return MPI_ERR_ARG;
  const int id = new_request();
  Request& rq = g_reqs[id];
  rq.is_send = true;
  rq.tag = tag;
  rq.buf = const_cast<char*>(static_cast<const char*>(buf));
  rq.bytes = static_cast<int64_t>(count) * t->size;
  *request = id;
  if (dest == MPI_PROC_NULL) {
    rq.done = true;
    return MPI_SUCCESS;
  }
  if (comm == MPI_COMM_SELF) dest = W.rank;
  if (dest < 0 || dest >= W.size) return MPI_ERR_ARG;
  rq.peer = dest;
  if (W.size == 1 || dest == W.rank) {
    // self send: buffer it as an unexpected message (or hand it to a posted receive)
    p2p_init();
    for (auto it = g_recvq[W.rank].begin(); it != g_recvq[W.rank].end(); ++it) {
      Request& r = g_reqs[*it];
      if (r.tag == MPI_ANY_TAG || r.tag == tag) {
        const int64_t n = std::min(r.bytes, rq.bytes);
        if (n > 0) std::memcpy(r.buf, rq.buf, static_cast<size_t>(n));
        if (rq.bytes > r.bytes) r.err = MPI_ERR_TRUNCATE;
        r.moved = n;
        finish_recv(r, W.rank, tag, rq.bytes);
        g_recvq[W.rank].erase(it);
        rq.done = true;
        return MPI_SUCCESS;
      }
    }
    auto* ux = new Unexpected();
    ux->src = W.rank;
    ux->tag = tag;
    ux->total = ux->have = rq.bytes;
    ux->data.assign(rq.buf, rq.buf + rq.bytes);
    g_unexp.push_back(ux);
    rq.done = true;
    return MPI_SUCCESS;
  }
  p2p_init();
  g_sendq[dest].push_back(id);
  progress();
  return MPI_SUCCESS;
}

int start_recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Request* request) {
  const Dtype* t = get_type(datatype);
  if (!t) return MPI_ERR_TYPE;
  if (!t->contiguous()) return MPI_ERR_TYPE;
  if (!valid_comm(comm) || count < 0) return MPI_ERR_ARG;
  if (source == MPI_ANY_SOURCE) return MPI_ERR_ARG;  // not needed by the reference
  const int id = new_request();
  Request& rq = g_reqs[id];
  rq.tag = tag;
  rq.buf = static_cast<char*>(buf);
  rq.bytes = static_cast<int64_t>(count) * t->size;
  *request = id;
  if (source == MPI_PROC_NULL) {
    rq.done = true;
    return MPI_SUCCESS;
  }
  if (comm == MPI_COMM_SELF) source = W.rank;
  if (source < 0 || source >= W.size) return MPI_ERR_ARG;
  rq.peer = source;
  p2p_init();
  // an earlier unexpected message from this source with a matching tag?
  for (auto it = g_unexp.begin(); it != g_unexp.end(); ++it) {
    Unexpected* ux = *it;
    if (ux->src != source || ux->claimed_by || !(tag == MPI_ANY_TAG || tag == ux->tag)) continue;
    if (ux->have == ux->total) {
      const int64_t n = std::min(rq.bytes, ux->total);
      if (n > 0) std::memcpy(rq.buf, ux->data.data(), static_cast<size_t>(n));
      if (ux->total > rq.bytes) rq.err = MPI_ERR_TRUNCATE;
      rq.moved = n;
      finish_recv(rq, source, ux->tag, ux->total);
      g_unexp.erase(it);
      delete ux;
    } else {
      ux->claimed_by = id;  // still arriving: take it over when complete
      rq.matched = true;
    }
    return MPI_SUCCESS;
  }
  g_recvq[source].push_back(id);
  if (W.size > 1) progress();
  return MPI_SUCCESS;
}

struct Lock {
  std::lock_guard<std::recursive_mutex> g;
  Lock() : g(W.mu) {}
};


}  // namespace

// ================================================================================== C interface
extern "C" {

int MPI_Init(int*, char***) {
  Lock l;
  world_init();
  types_init();
  return MPI_SUCCESS;
}

int MPI_Init_thread(int* argc, char*** argv, int /*required*/, int* provided) {
  if (provided) *provided = MPI_THREAD_MULTIPLE;  // every entry point takes the global lock
  return MPI_Init(argc, argv);
}

int MPI_Initialized(int* flag) {
  *flag = W.initialized ? 1 : 0;
  return MPI_SUCCESS;
}

int MPI_Finalized(int* flag) {
  *flag = W.finalized ? 1 : 0;
  return MPI_SUCCESS;
}

int MPI_Query_thread(int* provided) {
  *provided = MPI_THREAD_MULTIPLE;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  Lock l;
  if (!W.initialized || W.finalized) return MPI_SUCCESS;
  if (W.size > 1) {
    // drain our pending sends so peers that still wait can finish
    for (int d = 0; d < W.size; ++d)
      while (!g_sendq.empty() && !g_sendq[d].empty()) wait_request(g_sendq[d].front());
    barrier();
    munmap(W.base, W.bytes);
    W.base = nullptr;
    W.hdr = nullptr;
  }
  W.finalized = true;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int errorcode) {
  if (W.hdr) W.hdr->abort_flag.store(1, std::memory_order_relaxed);
  std::fprintf(stderr, "[mpishim] MPI_Abort(%d) on rank %d\n", errorcode, W.rank);
  _exit(errorcode ? errorcode : 1);
}

double MPI_Wtime(void) { return now_s(); }

int MPI_Error_string(int errorcode, char* string, int* resultlen) {
  int n = std::snprintf(string, MPI_MAX_ERROR_STRING, "mpishim error %d", errorcode);
  if (resultlen) *resultlen = n;
  return MPI_SUCCESS;
}

int MPI_Get_processor_name(char* name, int* resultlen) {
  if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) std::strcpy(name, "localhost");
  if (resultlen) *resultlen = static_cast<int>(std::strlen(name));
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int* rank) {
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  *rank = comm_rank(comm);
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int* size) {
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  *size = comm_size(comm);
  return MPI_SUCCESS;
}

MPI_Comm MPI_Comm_f2c(MPI_Fint comm) { return comm; }
MPI_Fint MPI_Comm_c2f(MPI_Comm comm) { return comm; }
MPI_Request MPI_Request_f2c(MPI_Fint request) { return request; }
MPI_Fint MPI_Request_c2f(MPI_Request request) { return request; }

int MPI_Barrier(MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  if (comm_size(comm) > 1) barrier();
  return MPI_SUCCESS;
}

int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
  Lock l;
  const Dtype* t = get_type(datatype);
  if (!t || !valid_comm(comm) || count < 0) return MPI_ERR_ARG;
  const int P = comm_size(comm), me = comm_rank(comm);
  if (P == 1) return MPI_SUCCESS;
  if (root < 0 || root >= P) return MPI_ERR_ARG;
  const int64_t total = static_cast<int64_t>(count) * t->size;
  // the root's stream goes through the result slot; one barrier per chunk plus a trailing one
  // every second chunk is implied by the parity rule (see reduce_collective)
  for (int64_t off = 0; off < total; off += W.slot) {
    const int64_t n = std::min(W.slot, total - off);
    const int par = static_cast<int>(W.coll_seq++ & 1);
    if (me == root) stream_copy<true>(static_cast<char*>(buffer), *t, off, n, coll_slot(par, P));
    barrier();
    if (me != root) stream_copy<false>(static_cast<char*>(buffer), *t, off, n, coll_slot(par, P));
    barrier();
  }
  return MPI_SUCCESS;
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  if (root < 0 || root >= comm_size(comm)) return MPI_ERR_ARG;
  return reduce_collective(sendbuf, recvbuf, count, datatype, op, root, false, comm);
}

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  return reduce_collective(sendbuf, recvbuf, count, datatype, op, 0, true, comm);
}

int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                const int* displs, MPI_Datatype recvtype, int root, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  const int P = comm_size(comm), me = comm_rank(comm);
  if (root < 0 || root >= P) return MPI_ERR_ARG;
  const Dtype* st = get_type(sendtype);
  if (!st) return MPI_ERR_TYPE;
  Stream s = make_stream(sendbuf, sendcount, st, 0);
  std::vector<Stream> r(P);
  if (me == root) {
    const Dtype* rt = get_type(recvtype);
    if (!rt) return MPI_ERR_TYPE;
    for (int i = 0; i < P; ++i) r[i] = make_stream(recvbuf, recvcounts[i], rt, displs[i]);
    if (sendbuf == MPI_IN_PLACE) s = r[me];
  }
  return exchange(0, root, &s, r.data(), comm);
}

int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
               int root, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  const int P = comm_size(comm);
  std::vector<int> counts(P, recvcount), displs(P);
  for (int i = 0; i < P; ++i) displs[i] = i * recvcount;
  return MPI_Gatherv(sendbuf, sendcount, sendtype, recvbuf, counts.data(), displs.data(), recvtype, root, comm);
}

int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                   const int* displs, MPI_Datatype recvtype, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  const int P = comm_size(comm), me = comm_rank(comm);
  const Dtype* rt = get_type(recvtype);
  if (!rt) return MPI_ERR_TYPE;
  std::vector<Stream> r(P);
  for (int i = 0; i < P; ++i) r[i] = make_stream(recvbuf, recvcounts[i], rt, displs[i]);
  Stream s;
  if (sendbuf == MPI_IN_PLACE) {
    s = r[me];
  } else {
    const Dtype* st = get_type(sendtype);
    if (!st) return MPI_ERR_TYPE;
    s = make_stream(sendbuf, sendcount, st, 0);
  }
  return exchange(1, 0, &s, r.data(), comm);
}

int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  const int P = comm_size(comm);
  std::vector<int> counts(P, recvcount), displs(P);
  for (int i = 0; i < P; ++i) displs[i] = i * recvcount;
  return MPI_Allgatherv(sendbuf, sendcount, sendtype, recvbuf, counts.data(), displs.data(), recvtype, comm);
}

int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  const int P = comm_size(comm), me = comm_rank(comm);
  if (root < 0 || root >= P) return MPI_ERR_ARG;
  std::vector<Stream> s(P);
  if (me == root) {
    const Dtype* st = get_type(sendtype);
    if (!st) return MPI_ERR_TYPE;
    for (int i = 0; i < P; ++i) s[i] = make_stream(sendbuf, sendcounts[i], st, displs[i]);
  }
  Stream r;
  if (recvbuf == MPI_IN_PLACE) {
    r = s[me];  // root keeps its part in place: copying a stream onto itself is harmless
  } else {
    const Dtype* rt = get_type(recvtype);
    if (!rt) return MPI_ERR_TYPE;
    r = make_stream(recvbuf, recvcount, rt, 0);
  }
  return exchange(2, root, s.data(), &r, comm);
}

int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                int root, MPI_Comm comm) {
  Lock l;
  if (!valid_comm(comm)) return MPI_ERR_ARG;
  const int P = comm_size(comm);
  std::vector<int> counts(P, sendcount), displs(P);
  for (int i = 0; i < P; ++i) displs[i] = i * sendcount;
  return MPI_Scatterv(sendbuf, counts.data(), displs.data(), sendtype, recvbuf, recvcount, recvtype, root, comm);
}

int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request) {
  Lock l;
  return start_send(buf, count, datatype, dest, tag, comm, request);
}

int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Request* request) {
  Lock l;
  return start_recv(buf, count, datatype, source, tag, comm, request);
}

int MPI_Wait(MPI_Request* request, MPI_Status* status) {
  Lock l;
  const int id = *request;
  if (id == MPI_REQUEST_NULL) return MPI_SUCCESS;
  if (id < 0 || id >= static_cast<int>(g_reqs.size()) || !g_reqs[id].active) return MPI_ERR_REQUEST;
  wait_request(id);
  Request& rq = g_reqs[id];
  if (status) *status = rq.st;
  const int err = rq.err;
  rq.active = false;
  *request = MPI_REQUEST_NULL;
  return err;
}

int MPI_Test(MPI_Request* request, int* flag, MPI_Status* status) {
  Lock l;
  const int id = *request;
  if (id == MPI_REQUEST_NULL) {
    *flag = 1;
    return MPI_SUCCESS;
  }
  if (id < 0 || id >= static_cast<int>(g_reqs.size()) || !g_reqs[id].active) return MPI_ERR_REQUEST;
  if (W.size > 1) progress();
  *flag = g_reqs[id].done ? 1 : 0;
  if (*flag) return MPI_Wait(request, status);
  return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[]) {
  int rc = MPI_SUCCESS;
  for (int i = 0; i < count; ++i) {
    int e = MPI_Wait(&requests[i], statuses ? &statuses[i] : nullptr);
    if (e != MPI_SUCCESS) rc = e;
  }
  return rc;
}

int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm) {
  MPI_Request r;
  int e = MPI_Isend(buf, count, datatype, dest, tag, comm, &r);
  return e != MPI_SUCCESS ? e : MPI_Wait(&r, MPI_STATUS_IGNORE);
}

int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status* status) {
  MPI_Request r;
  int e = MPI_Irecv(buf, count, datatype, source, tag, comm, &r);
  return e != MPI_SUCCESS ? e : MPI_Wait(&r, status);
}

int MPI_Get_count(const MPI_Status* status, MPI_Datatype datatype, int* count) {
  const Dtype* t = get_type(datatype);
  if (!t || t->size == 0) return MPI_ERR_TYPE;
  *count = static_cast<int>(status->shim_bytes / t->size);
  return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  return MPI_Type_vector(1, count, count, oldtype, newtype);
}

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  Lock l;
  const Dtype* o = get_type(oldtype);
  if (!o || count < 0 || blocklength < 0) return MPI_ERR_ARG;
  Dtype t;
  const Dtype old = *o;  // new_type may reallocate the table
  t.blocks.reserve(static_cast<size_t>(count) * (old.contiguous() ? 1 : static_cast<size_t>(blocklength) * old.blocks.size()));
  for (int64_t i = 0; i < count; ++i) {
    const int64_t start = i * static_cast<int64_t>(stride) * old.extent;
    if (old.contiguous()) {
      t.blocks.emplace_back(start, static_cast<int64_t>(blocklength) * old.size);
    } else {
      for (int64_t j = 0; j < blocklength; ++j)
        for (auto& b : old.blocks) t.blocks.emplace_back(start + j * old.extent + b.first, b.second);
    }
  }
  t.lb = 0;
  t.extent = count > 0 ? ((static_cast<int64_t>(count) - 1) * stride + blocklength) * old.extent : 0;
  t.finish();
  *newtype = new_type(std::move(t));
  return MPI_SUCCESS;
}

int MPI_Type_create_resized(MPI_Datatype oldtype, MPI_Aint lb, MPI_Aint extent, MPI_Datatype* newtype) {
  Lock l;
  const Dtype* o = get_type(oldtype);
  if (!o) return MPI_ERR_TYPE;
  Dtype t = *o;
  t.used = false;
  t.prim = P_NONE;
  t.lb = lb;
  t.extent = extent;
  *newtype = new_type(std::move(t));
  return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype* datatype) { return get_type(*datatype) ? MPI_SUCCESS : MPI_ERR_TYPE; }

int MPI_Type_free(MPI_Datatype* datatype) {
  Lock l;
  if (*datatype < kFirstDerived || !get_type(*datatype)) return MPI_ERR_TYPE;
  g_types[*datatype] = Dtype();
  *datatype = MPI_DATATYPE_NULL;
  return MPI_SUCCESS;
}

int MPI_Type_get_extent(MPI_Datatype datatype, MPI_Aint* lb, MPI_Aint* extent) {
  const Dtype* t = get_type(datatype);
  if (!t) return MPI_ERR_TYPE;
  *lb = t->lb;
  *extent = t->extent;
  return MPI_SUCCESS;
}

int MPI_Type_size(MPI_Datatype datatype, int* size) {
  const Dtype* t = get_type(datatype);
  if (!t) return MPI_ERR_TYPE;
  *size = static_cast<int>(t->size);
  return MPI_SUCCESS;
}

}  // extern "C"
