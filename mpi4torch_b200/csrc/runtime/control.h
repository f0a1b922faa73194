// Host control plane: one POSIX shared-memory segment per job that every rank
// of the (single-node, one NVLink domain) world attaches to.
//
// This replaces MPI_Init / mpirun wire-up (reference csrc/extension.cpp:1306-1394):
//   * rank / size come from the launcher environment (RANK, WORLD_SIZE),
//   * the job id (M4T_JOB_ID, or MASTER_PORT + parent pid under torchrun)
//     names the segment,
//   * flag barriers, a small metadata all-gather, per-pair descriptor rings for
//     point-to-point matching and SCM_RIGHTS fd passing (for CUDA VMM handles)
//     are all the control plane ever does - no collective payload crosses it
//     on the GPU path.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "common.h"

namespace m4t {

constexpr int kMaxRanks = 64;
constexpr int kMetaWords = 192;    // int64 words per rank per parity for metadata exchange
constexpr int kMailboxDepth = 64;  // outstanding un-received messages per directed pair
constexpr int kInlineBytes = 32;

struct MsgDesc {
  uint64_t seq;    // per (src,dst) message number, 1-based
  int64_t tag;
  uint64_t bytes;
  uint32_t kind;   // 0 = payload inline, 1 = payload in a shm segment, 2 = device FIFO
  uint32_t dtype;
  uint8_t inline_data[kInlineBytes];
};

struct alignas(64) PairRing {  // SPSC: producer = src rank, consumer = dst rank
  std::atomic<uint64_t> head;  // messages published
  char pad0[56];
  std::atomic<uint64_t> tail;  // messages consumed
  char pad1[56];
  MsgDesc entries[kMailboxDepth];
};

struct alignas(64) RankSlot {
  std::atomic<uint64_t> arrive;        // barrier arrival counter
  std::atomic<uint64_t> arena_gen[2];  // generation of each data arena (CPU backend)
  std::atomic<uint64_t> arena_cap[2];
  std::atomic<int32_t> pid;
  std::atomic<int32_t> state;          // 0 = absent, 1 = attached, 2 = detached
  char pad[64 - 8 - 16 - 16 - 8];
  int64_t meta[2][kMetaWords];
};

struct ControlBlock {
  std::atomic<uint64_t> magic;
  std::atomic<uint32_t> world;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> abort_flag;  // set by any rank on a fatal error
  std::atomic<uint32_t> abort_rank;
  std::atomic<uint32_t> departing;   // ranks that reached quiesce() (teardown handshake)
  char pad[64 - 28];
  RankSlot slots[kMaxRanks];
  PairRing rings[kMaxRanks][kMaxRanks];  // [src][dst]
};

// A rank's attachment to the job-wide control block.
class Control {
 public:
  // Attaches (rank 0 creates).  Blocks until all `size` ranks are attached.
  Control(int rank, int size, const std::string& job_id);
  ~Control();
  Control(const Control&) = delete;
  Control& operator=(const Control&) = delete;

  int rank() const { return rank_; }
  int size() const { return size_; }
  const std::string& job_id() const { return job_; }
  ControlBlock* block() { return cb_; }

  // Host barrier over all ranks.  Throws on peer abort / timeout.
  void barrier();
  // Teardown handshake (call at most once): waits until every rank has also reached its
  // quiesce(), so nobody unlinks a shared segment a slower peer has not opened yet.  Never
  // throws; gives up (false) on abort, on a dead peer or after `timeout_s`.
  bool quiesce(double timeout_s) noexcept;

  // All-gather `k` int64 words per rank; `all` is [size][k] row-major.
  // One barrier; double-buffered so consecutive calls need no trailing barrier.
  void allgather_i64(const int64_t* mine, int k, int64_t* all);

  // Broadcast `k` int64 words from `root`.
  void bcast_i64(int64_t* data, int k, int root);

  // Spin helper used by all host-side waits: returns when pred() is true,
  // throws if a peer aborted or the timeout (M4T_TIMEOUT_S, default 300) hit.
  template <typename Pred>
  void wait_until(Pred pred, const char* what);

  void signal_abort() noexcept;
  bool aborted() const;

  // fd passing between ranks over abstract unix sockets (CUDA VMM handles).
  // All ranks call; `mine` are this rank's fds; returns fds[peer][i] (own row
  // duplicates `mine`).  Collective.
  std::vector<std::vector<int>> exchange_fds(const std::vector<int>& mine);

  // name helpers for auxiliary shm segments
  std::string seg_name(const std::string& suffix) const;

  uint64_t next_host_seq() { return ++host_seq_; }

  // Called now and then while a wait on this segment spins.  A job that spans nodes installs the TCP engine's progress
  // here: a rank that waits for a neighbour on its node must keep moving the bytes it still owes to other nodes.
  void set_idle_hook(std::function<void()> hook) { idle_hook_ = std::move(hook); }

 private:
  void backoff(uint64_t& spins);
  void check_abort_timeout(uint64_t start_ns, const char* what);

  int rank_, size_;
  std::string job_;
  std::string seg_;
  ControlBlock* cb_ = nullptr;
  uint64_t barrier_count_ = 0;
  uint64_t meta_seq_ = 0;
  uint64_t host_seq_ = 0;
  uint64_t fd_round_ = 0;
  double timeout_s_;
  std::function<void()> idle_hook_;
  bool in_quiesce_ = false;
};

uint64_t now_ns();

template <typename Pred>
void Control::wait_until(Pred pred, const char* what) {
  uint64_t spins = 0;
  uint64_t start = 0;
  while (!pred()) {
    backoff(spins);
    if ((spins & 0x3ff) == 0) {
      if (start == 0) start = now_ns();
      check_abort_timeout(start, what);
    }
  }
}

// Resolves (rank, size, job id) from the launcher environment.  World size 1
// with no environment yields a private single-rank job.
struct WorldEnv {
  int rank = 0;
  int size = 1;
  int local_rank = 0;
  std::string job_id;
};
WorldEnv world_env_from_environment();

}  // namespace m4t
