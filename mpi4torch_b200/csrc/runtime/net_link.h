// TCP transport for jobs that span more than one node (the reference gets this from MPI itself; on one node the
// POSIX shared-memory control plane and the NVLink symmetric heap are used instead and none of this is active).
//
//   NetEngine - one per process: a full mesh of non-blocking TCP connections to every other rank of the job, framed
//               messages (communicator id, kind, tag, byte count), posted receives matched in FIFO order per
//               (source, communicator, kind, tag), an unexpected-message queue, and ONE progress loop (poll) that
//               drains incoming frames while it pushes queued sends - so two ranks that send each other large messages
//               at the same time cannot block each other.  There is no progress thread: like an MPI without
//               asynchronous progress, data moves while some rank is inside a library call.
//   NetLink   - a communicator's view of the engine: member list (communicator rank -> world rank), its own id so that
//               frames of different communicators never match each other, and the host control operations the API
//               layer needs (barrier, small all-gather / broadcast of int64 words).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <poll.h>
#include <list>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace m4t {

constexpr uint32_t kNetCtrl = 1;  // barrier / metadata words
constexpr uint32_t kNetColl = 2;  // collective payload
constexpr uint32_t kNetP2p = 3;   // user Isend / Irecv

// Byte buffer without value-initialisation (a std::vector<char> would zero-fill - and page-fault - every message).
struct NetBuffer {
  std::unique_ptr<char[]> ptr;
  size_t bytes = 0;
  void allocate(size_t n) {
    ptr.reset(n ? new char[n] : nullptr);
    bytes = n;
  }
  void assign(const void* src, size_t n) {
    allocate(n);
    if (n) std::memcpy(ptr.get(), src, n);
  }
  char* data() { return ptr.get(); }
  const char* data() const { return ptr.get(); }
  size_t size() const { return bytes; }
  bool empty() const { return bytes == 0; }
};

class NetEngine {
 public:
  // Opens the listening socket of this rank (ephemeral port on all interfaces); returns the fd.
  static int listen_any(int* port_out);

  // Builds the mesh: addrs[p] = "host:port" of rank p's listening socket.  Ranks connect to every lower rank and
  // accept from every higher one.  Takes ownership of listen_fd (closed once the mesh stands).
  // `job_key`: greetings carrying another key (a different job, a port scanner) are turned away.
  NetEngine(int rank, int size, int listen_fd, const std::vector<std::string>& addrs, double timeout_s, uint32_t job_key = 0);
  ~NetEngine();
  NetEngine(const NetEngine&) = delete;
  NetEngine& operator=(const NetEngine&) = delete;

  int rank() const { return rank_; }
  int size() const { return size_; }
  double timeout_s() const { return timeout_s_; }
  void set_timeout_s(double t) { timeout_s_ = t; }

  // `data` must stay valid until wait() returned for the operation, unless `copy` (the engine then keeps its own
  // copy until the bytes have left).  peer == rank() is a local hand-over.
  uint64_t post_send(int peer, uint32_t comm, uint32_t kind, int64_t tag, const void* data, size_t bytes,
                     bool copy = false);
  // data == nullptr: the engine allocates the buffer (any size up to cap), fetch it through wait()'s `owned`.
  uint64_t post_recv(int peer, uint32_t comm, uint32_t kind, int64_t tag, void* data, size_t cap);
  // Progresses until the operation is complete; returns the message size; forgets the operation.
  size_t wait(uint64_t op, NetBuffer* owned = nullptr);
  void wait_all(const std::vector<uint64_t>& ops);
  // One non-blocking progress pass (pending sends leave, arriving frames are taken in); for waits outside the engine.
  void poke() { progress(0); }

 private:
  struct Header {
    uint32_t magic, comm, kind, reserved;
    int64_t tag;
    uint64_t bytes;
  };
  struct Op {
    bool done = false, is_recv = false, engine_buffer = false;
    int peer = 0;
    uint32_t comm = 0, kind = 0;
    int64_t tag = 0;
    char* data = nullptr;
    size_t cap = 0, bytes = 0;
    NetBuffer owned;
  };
  struct SendItem {
    Header h;
    const char* data;
    size_t off;  // bytes of header + payload already written
    uint64_t op;
  };
  struct Unexpected {
    Header h;
    NetBuffer data;
    bool complete = false;
    uint64_t claimed = 0;  // posted receive waiting for the rest of this frame
  };
  struct Peer {
    int fd = -1;
    bool closed = false;
    std::deque<SendItem> sendq;
    std::list<uint64_t> posted;  // receive ops in post order
    std::list<Unexpected> ux;    // arrival order
    // frame being received
    Header in_h{};
    size_t in_hdr = 0, in_got = 0;
    char* in_dst = nullptr;
    uint64_t in_op = 0;
    Unexpected* in_ux = nullptr;
    bool in_active = false;
  };

  static bool matches(const Header& h, const Op& o) { return h.comm == o.comm && h.kind == o.kind && h.tag == o.tag; }
  void progress(int timeout_ms);
  void pump_in(int p);
  void pump_out(int p);
  void frame_started(int p);
  void frame_finished(int p);
  void deliver_unexpected(Peer& pr, std::list<Unexpected>::iterator it, uint64_t opid);
  void check_peer_alive(const Op& o) const;
  [[noreturn]] void fail(const std::string& what) const;

  int rank_, size_;
  double timeout_s_;
  std::vector<Peer> peers_;
  std::unordered_map<uint64_t, Op> ops_;
  uint64_t next_op_ = 1;
  std::vector<struct pollfd> poll_fds_;
  std::vector<int> poll_who_;
};

class NetLink {
 public:
  NetLink(std::shared_ptr<NetEngine> eng, uint32_t comm_id, std::vector<int> members, int my_index);

  int rank() const { return rank_; }
  int size() const { return static_cast<int>(members_.size()); }
  NetEngine& engine() { return *eng_; }
  const std::shared_ptr<NetEngine>& engine_ptr() const { return eng_; }
  const std::vector<int>& members() const { return members_; }
  uint32_t comm_id() const { return comm_; }

  // tag of the next collective: identical on all members because collectives are called in the same order
  int64_t next_seq() { return ++coll_seq_; }
  uint64_t send(int peer, uint32_t kind, int64_t tag, const void* data, size_t bytes) {
    return eng_->post_send(members_[peer], comm_, kind, tag, data, bytes);
  }
  uint64_t recv(int peer, uint32_t kind, int64_t tag, void* data, size_t cap) {
    return eng_->post_recv(members_[peer], comm_, kind, tag, data, cap);
  }

  void barrier();
  void allgather_i64(const int64_t* mine, int k, int64_t* all);
  void bcast_i64(int64_t* data, int k, int root);
  // teardown handshake: a barrier that never throws and gives up after `timeout_s` (a rank that leaves early - an
  // exception, sys.exit - must not keep its peers waiting for the full operation timeout)
  bool quiesce(double timeout_s) noexcept;

 private:
  std::shared_ptr<NetEngine> eng_;
  uint32_t comm_;
  std::vector<int> members_;
  int rank_;
  int64_t coll_seq_ = 0, ctrl_seq_ = 0;
};

// 32-bit id of a communicator from its job string (identical on all members, distinct between communicators).
uint32_t net_comm_id(const std::string& job);

}  // namespace m4t
