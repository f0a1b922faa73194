#include "net_link.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <stdexcept>

#include "control.h"  // now_ns

namespace m4t {

namespace {
constexpr uint32_t kNetMagic = 0x6d34746eu;  // "m4tn"
constexpr size_t kPumpBudget = 8u << 20;     // bytes one peer may move per progress pass (fairness)

void set_nonblocking(int fd) {
  const int fl = fcntl(fd, F_GETFL, 0);
  fcntl(fd, F_SETFL, fl | O_NONBLOCK);
}

void tune_socket(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  int buf = 4 << 20;
  setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
  setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
}

// blocking helpers used only while the mesh is being built
void write_full(int fd, const void* p, size_t n, uint64_t deadline_ns) {
  const char* c = static_cast<const char*>(p);
  while (n) {
    const ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
    if (w > 0) {
      c += w;
      n -= static_cast<size_t>(w);
    } else if (w < 0 && (errno == EINTR || errno == EAGAIN)) {
      M4T_CHECK(now_ns() < deadline_ns, "timed out while greeting a peer");
    } else {
      M4T_CHECK(false, "send() failed while building the TCP mesh: " << std::strerror(errno));
    }
  }
}

void read_full(int fd, void* p, size_t n, uint64_t deadline_ns) {
  char* c = static_cast<char*>(p);
  while (n) {
    pollfd pf{fd, POLLIN, 0};
    const int pr = ::poll(&pf, 1, 200);
    M4T_CHECK(now_ns() < deadline_ns, "timed out waiting for a peer's greeting");
    if (pr <= 0) continue;
    const ssize_t r = ::recv(fd, c, n, 0);
    if (r > 0) {
      c += r;
      n -= static_cast<size_t>(r);
    } else if (r == 0) {
      M4T_CHECK(false, "a peer closed the connection while the TCP mesh was being built");
    } else if (errno != EINTR && errno != EAGAIN) {
      M4T_CHECK(false, "recv() failed while building the TCP mesh: " << std::strerror(errno));
    }
  }
}

int connect_to(const std::string& addr, uint64_t deadline_ns) {
  const size_t colon = addr.rfind(':');
  M4T_CHECK(colon != std::string::npos, "malformed peer address '" << addr << "'");
  const std::string host = addr.substr(0, colon), port = addr.substr(colon + 1);
  for (;;) {
    addrinfo hints{};
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    addrinfo* res = nullptr;
    const int rc = getaddrinfo(host.c_str(), port.c_str(), &hints, &res);
    M4T_CHECK(rc == 0, "cannot resolve peer address '" << addr << "': " << gai_strerror(rc));
    int fd = -1;
    for (addrinfo* ai = res; ai; ai = ai->ai_next) {
      fd = ::socket(ai->ai_family, ai->ai_socktype, ai->ai_protocol);
      if (fd < 0) continue;
      if (::connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) break;
      ::close(fd);
      fd = -1;
    }
    freeaddrinfo(res);
    if (fd >= 0) return fd;
    M4T_CHECK(now_ns() < deadline_ns, "could not connect to peer at " << addr << ": " << std::strerror(errno));
    usleep(50 * 1000);  // the peer may not be listening yet
  }
}
}  // namespace

uint32_t net_comm_id(const std::string& job) {
  uint32_t h = 2166136261u;  // FNV-1a
  for (unsigned char c : job) h = (h ^ c) * 16777619u;
  return h ? h : 1u;
}

int NetEngine::listen_any(int* port_out) {
  int fd = ::socket(AF_INET6, SOCK_STREAM, 0);
  bool v6 = fd >= 0;
  if (v6) {
    int off = 0;
    setsockopt(fd, IPPROTO_IPV6, IPV6_V6ONLY, &off, sizeof(off));  // dual stack
    sockaddr_in6 a{};
    a.sin6_family = AF_INET6;
    a.sin6_addr = in6addr_any;
    a.sin6_port = 0;
    if (::bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0) {
      ::close(fd);
      fd = -1;
      v6 = false;
    }
  }
  if (!v6) {
    fd = ::socket(AF_INET, SOCK_STREAM, 0);
    M4T_CHECK(fd >= 0, "socket() failed: " << std::strerror(errno));
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_ANY);
    a.sin_port = 0;
    M4T_CHECK(::bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) == 0, "bind() failed: " << std::strerror(errno));
  }
  M4T_CHECK(::listen(fd, 128) == 0, "listen() failed: " << std::strerror(errno));
  sockaddr_storage ss{};
  socklen_t len = sizeof(ss);
  M4T_CHECK(getsockname(fd, reinterpret_cast<sockaddr*>(&ss), &len) == 0, "getsockname() failed");
  *port_out = ss.ss_family == AF_INET6 ? ntohs(reinterpret_cast<sockaddr_in6*>(&ss)->sin6_port)
                                       : ntohs(reinterpret_cast<sockaddr_in*>(&ss)->sin_port);
  return fd;
}

NetEngine::NetEngine(int rank, int size, int listen_fd, const std::vector<std::string>& addrs, double timeout_s,
                     uint32_t job_key)
    : rank_(rank), size_(size), timeout_s_(timeout_s), peers_(static_cast<size_t>(size)) {
  M4T_CHECK(static_cast<int>(addrs.size()) == size, "need one address per rank");
  const uint64_t deadline = now_ns() + static_cast<uint64_t>(timeout_s * 1e9);
  struct Hello {
    uint32_t magic;
    int32_t rank;
    uint32_t job_key;
    uint32_t reserved;
  };
  for (int p = 0; p < rank; ++p) {
    const int fd = connect_to(addrs[static_cast<size_t>(p)], deadline);
    Hello h{kNetMagic, rank, job_key, 0};
    write_full(fd, &h, sizeof(h), deadline);
    peers_[static_cast<size_t>(p)].fd = fd;
  }
  for (int k = rank + 1; k < size; ++k) {
    int fd = -1;
    for (;;) {
      pollfd pf{listen_fd, POLLIN, 0};
      const int pr = ::poll(&pf, 1, 200);
      M4T_CHECK(now_ns() < deadline, "timed out waiting for " << (size - k) << " higher rank(s) to connect");
      if (pr <= 0) continue;
      fd = ::accept(listen_fd, nullptr, nullptr);
      if (fd >= 0) break;
    }
    Hello h{};
    bool ok = true;
    try {
      read_full(fd, &h, sizeof(h), std::min<uint64_t>(deadline, now_ns() + 5000000000ull));
    } catch (const std::exception&) {
      ok = false;  // a connection that never greets (port scanner, health check)
    }
    ok = ok && h.magic == kNetMagic && h.job_key == job_key && h.rank > rank && h.rank < size &&
         peers_[static_cast<size_t>(h.rank)].fd < 0;
    if (!ok) {  // not one of ours: turn it away and keep waiting for the real peer
      ::close(fd);
      --k;
      M4T_CHECK(now_ns() < deadline, "timed out waiting for the higher ranks to connect");
      continue;
    }
    peers_[static_cast<size_t>(h.rank)].fd = fd;
  }
  ::close(listen_fd);
  for (int p = 0; p < size; ++p) {
    if (p == rank) continue;
    tune_socket(peers_[static_cast<size_t>(p)].fd);
    set_nonblocking(peers_[static_cast<size_t>(p)].fd);
  }
}

NetEngine::~NetEngine() {
  for (auto& p : peers_)
    if (p.fd >= 0) ::close(p.fd);
}

void NetEngine::fail(const std::string& what) const {
  throw std::runtime_error("mpi4torch_b200 (tcp transport, rank " + std::to_string(rank_) + "): " + what);
}

void NetEngine::deliver_unexpected(Peer& pr, std::list<Unexpected>::iterator it, uint64_t opid) {
  Op& o = ops_.at(opid);
  if (it->h.bytes > o.cap)
    fail("message truncated: " + std::to_string(it->h.bytes) + " bytes (tag " + std::to_string(it->h.tag) + ") into a " +
         std::to_string(o.cap) + "-byte receive buffer");
  o.bytes = it->h.bytes;
  if (o.engine_buffer) {
    o.owned = std::move(it->data);
  } else if (o.bytes) {
    std::memcpy(o.data, it->data.data(), o.bytes);
  }
  o.done = true;
  pr.ux.erase(it);
}

uint64_t NetEngine::post_send(int peer, uint32_t comm, uint32_t kind, int64_t tag, const void* data, size_t bytes,
                              bool copy) {
  M4T_CHECK(peer >= 0 && peer < size_, "tcp transport: peer " << peer << " out of range");
  const uint64_t id = next_op_++;
  Op o;
  o.peer = peer;
  o.comm = comm;
  o.kind = kind;
  o.tag = tag;
  o.bytes = bytes;
  Header h{kNetMagic, comm, kind, 0, tag, static_cast<uint64_t>(bytes)};
  Peer& pr = peers_[static_cast<size_t>(peer)];
  if (peer == rank_) {
    // local hand-over: into a posted receive, else buffered
    o.done = true;
    ops_[id] = std::move(o);
    for (auto it = pr.posted.begin(); it != pr.posted.end(); ++it) {
      Op& r = ops_.at(*it);
      if (!matches(h, r)) continue;
      if (bytes > r.cap)
        fail("message truncated: " + std::to_string(bytes) + " bytes (tag " + std::to_string(tag) + ") into a " +
             std::to_string(r.cap) + "-byte receive buffer");
      r.bytes = bytes;
      if (r.engine_buffer) r.owned.assign(data, bytes);
      else if (bytes) std::memcpy(r.data, data, bytes);
      r.done = true;
      pr.posted.erase(it);
      return id;
    }
    Unexpected u;
    u.h = h;
    u.data.assign(data, bytes);
    u.complete = true;
    pr.ux.push_back(std::move(u));
    return id;
  }
  if (pr.closed) fail("rank " + std::to_string(peer) + " has closed its connection");
  if (copy && bytes) o.owned.assign(data, bytes);
  ops_[id] = std::move(o);
  const char* src = (copy && bytes) ? ops_.at(id).owned.data() : static_cast<const char*>(data);
  pr.sendq.push_back(SendItem{h, src, 0, id});
  pump_out(peer);  // eager: whatever fits into the socket buffer leaves now
  return id;
}

uint64_t NetEngine::post_recv(int peer, uint32_t comm, uint32_t kind, int64_t tag, void* data, size_t cap) {
  M4T_CHECK(peer >= 0 && peer < size_, "tcp transport: peer " << peer << " out of range");
  const uint64_t id = next_op_++;
  Op o;
  o.is_recv = true;
  o.peer = peer;
  o.comm = comm;
  o.kind = kind;
  o.tag = tag;
  o.data = static_cast<char*>(data);
  o.cap = cap;
  o.engine_buffer = data == nullptr;
  ops_[id] = std::move(o);
  Peer& pr = peers_[static_cast<size_t>(peer)];
  for (auto it = pr.ux.begin(); it != pr.ux.end(); ++it) {
    if (it->claimed || !matches(it->h, ops_.at(id))) continue;
    if (it->complete) deliver_unexpected(pr, it, id);
    else it->claimed = id;  // the rest of the frame is still on the wire
    return id;
  }
  pr.posted.push_back(id);
  return id;
}

void NetEngine::frame_started(int p) {
  Peer& pr = peers_[static_cast<size_t>(p)];
  const Header& h = pr.in_h;
  if (h.magic != kNetMagic) fail("corrupt frame header from rank " + std::to_string(p));
  pr.in_got = 0;
  pr.in_op = 0;
  pr.in_ux = nullptr;
  pr.in_active = true;
  for (auto it = pr.posted.begin(); it != pr.posted.end(); ++it) {
    Op& o = ops_.at(*it);
    if (!matches(h, o)) continue;
    if (h.bytes > o.cap)
      fail("message truncated: " + std::to_string(h.bytes) + " bytes sent by rank " + std::to_string(p) + " (tag " +
           std::to_string(h.tag) + ") into a " + std::to_string(o.cap) + "-byte receive buffer");
    if (o.engine_buffer) {
      o.owned.allocate(h.bytes);
      o.data = o.owned.data();
    }
    o.bytes = h.bytes;
    pr.in_op = *it;
    pr.in_dst = o.data;
    pr.posted.erase(it);
    if (h.bytes == 0) frame_finished(p);
    return;
  }
  pr.ux.emplace_back();
  Unexpected& u = pr.ux.back();
  u.h = h;
  u.data.allocate(h.bytes);
  pr.in_ux = &u;
  pr.in_dst = u.data.data();
  if (h.bytes == 0) frame_finished(p);
}

void NetEngine::frame_finished(int p) {
  Peer& pr = peers_[static_cast<size_t>(p)];
  pr.in_active = false;
  pr.in_hdr = 0;
  if (pr.in_op) {
    Op& o = ops_.at(pr.in_op);
    if (o.engine_buffer) o.data = nullptr;
    o.done = true;
  } else if (pr.in_ux) {
    pr.in_ux->complete = true;
    if (pr.in_ux->claimed) {
      const uint64_t opid = pr.in_ux->claimed;
      for (auto it = pr.ux.begin(); it != pr.ux.end(); ++it)
        if (&*it == pr.in_ux) {
          deliver_unexpected(pr, it, opid);
          break;
        }
    }
  }
  pr.in_op = 0;
  pr.in_ux = nullptr;
}

void NetEngine::pump_in(int p) {
  Peer& pr = peers_[static_cast<size_t>(p)];
  size_t budget = kPumpBudget;
  while (budget > 0 && !pr.closed) {
    if (!pr.in_active) {
      const ssize_t r = ::recv(pr.fd, reinterpret_cast<char*>(&pr.in_h) + pr.in_hdr, sizeof(Header) - pr.in_hdr, 0);
      if (r > 0) {
        pr.in_hdr += static_cast<size_t>(r);
        if (pr.in_hdr == sizeof(Header)) frame_started(p);
        continue;
      }
      if (r == 0) pr.closed = true;
      else if (errno == EINTR) continue;
      else if (errno != EAGAIN && errno != EWOULDBLOCK) pr.closed = true;
      return;
    }
    const size_t want = std::min<size_t>(pr.in_h.bytes - pr.in_got, budget);
    const ssize_t r = ::recv(pr.fd, pr.in_dst + pr.in_got, want, 0);
    if (r > 0) {
      pr.in_got += static_cast<size_t>(r);
      budget -= static_cast<size_t>(r);
      if (pr.in_got == pr.in_h.bytes) frame_finished(p);
      continue;
    }
    if (r == 0) pr.closed = true;
    else if (errno == EINTR) continue;
    else if (errno != EAGAIN && errno != EWOULDBLOCK) pr.closed = true;
    return;
  }
}

void NetEngine::pump_out(int p) {
  Peer& pr = peers_[static_cast<size_t>(p)];
  size_t budget = kPumpBudget;
  while (!pr.sendq.empty() && budget > 0 && !pr.closed) {
    SendItem& s = pr.sendq.front();
    const size_t total = sizeof(Header) + s.h.bytes;
    const char* src;
    size_t n;
    if (s.off < sizeof(Header)) {
      src = reinterpret_cast<const char*>(&s.h) + s.off;
      n = sizeof(Header) - s.off;
    } else {
      src = s.data + (s.off - sizeof(Header));
      n = std::min<size_t>(total - s.off, budget);
    }
    const ssize_t w = ::send(pr.fd, src, n, MSG_NOSIGNAL);
    if (w > 0) {
      s.off += static_cast<size_t>(w);
      budget -= static_cast<size_t>(w);
      if (s.off == total) {
        ops_.at(s.op).done = true;
        pr.sendq.pop_front();
      }
      continue;
    }
    if (w < 0 && errno == EINTR) continue;
    if (w < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) return;
    pr.closed = true;
    return;
  }
}

void NetEngine::progress(int timeout_ms) {
  auto& fds = poll_fds_;  // members: the busy part of wait() calls this hundreds of times per message
  auto& who = poll_who_;
  fds.clear();
  who.clear();
  for (int p = 0; p < size_; ++p) {
    Peer& pr = peers_[static_cast<size_t>(p)];
    if (p == rank_ || pr.fd < 0 || pr.closed) continue;
    pollfd pf{pr.fd, POLLIN, 0};
    if (!pr.sendq.empty()) pf.events |= POLLOUT;
    fds.push_back(pf);
    who.push_back(p);
  }
  if (fds.empty()) return;
  const int n = ::poll(fds.data(), fds.size(), timeout_ms);
  if (n <= 0) return;
  for (size_t i = 0; i < fds.size(); ++i) {
    const int p = who[i];
    if (fds[i].revents & (POLLIN | POLLHUP | POLLERR)) pump_in(p);
    if (fds[i].revents & POLLOUT) pump_out(p);
  }
}

void NetEngine::check_peer_alive(const Op& o) const {
  if (o.peer != rank_ && peers_[static_cast<size_t>(o.peer)].closed && !o.done)
    fail("rank " + std::to_string(o.peer) + " closed its connection (the process ended or raised) while this rank was " +
         (o.is_recv ? "waiting for a message from it" : "sending to it"));
}

size_t NetEngine::wait(uint64_t op, NetBuffer* owned) {
  auto it = ops_.find(op);
  M4T_CHECK(it != ops_.end(), "tcp transport: unknown or already completed operation");
  uint64_t start = 0;
  int spins = 0;
  while (!it->second.done) {
    check_peer_alive(it->second);
    // answers of a peer that is already inside the matching call arrive within microseconds: look a few times without
    // sleeping before paying a scheduler wake-up
    progress(spins < 200 ? 0 : 50);
    ++spins;
    if (!it->second.done) {
      if (start == 0) start = now_ns();
      if (static_cast<double>(now_ns() - start) > timeout_s_ * 1e9)
        fail(std::string("timed out after ") + std::to_string(static_cast<long>(timeout_s_)) + " s " +
             (it->second.is_recv ? "waiting for a message from rank " : "sending to rank ") +
             std::to_string(it->second.peer) + " (mismatched collective order across ranks?)");
    }
  }
  const size_t bytes = it->second.bytes;
  if (owned) *owned = std::move(it->second.owned);
  ops_.erase(it);
  return bytes;
}

void NetEngine::wait_all(const std::vector<uint64_t>& ops) {
  for (uint64_t op : ops) wait(op);
}

// ---------------------------------------------------------------------------------------------------------------------

NetLink::NetLink(std::shared_ptr<NetEngine> eng, uint32_t comm_id, std::vector<int> members, int my_index)
    : eng_(std::move(eng)), comm_(comm_id), members_(std::move(members)), rank_(my_index) {
  M4T_CHECK(rank_ >= 0 && rank_ < size() && members_[static_cast<size_t>(rank_)] == eng_->rank(),
            "tcp transport: inconsistent member list");
}

void NetLink::barrier() {
  const int P = size();
  const int64_t seq = ++ctrl_seq_;
  int round = 0;
  for (int d = 1; d < P; d <<= 1, ++round) {
    const int64_t tag = seq * 64 + round;
    const uint64_t r = recv((rank_ - d + P) % P, kNetCtrl, tag, nullptr, 0);
    const uint64_t s = send((rank_ + d) % P, kNetCtrl, tag, nullptr, 0);
    eng_->wait(s);
    eng_->wait(r);
  }
}

void NetLink::allgather_i64(const int64_t* mine, int k, int64_t* all) {
  const int P = size();
  const int64_t tag = (++ctrl_seq_) * 64;
  const size_t bytes = static_cast<size_t>(k) * sizeof(int64_t);
  std::vector<uint64_t> ops;
  for (int p = 0; p < P; ++p) {
    if (p == rank_) continue;
    ops.push_back(recv(p, kNetCtrl, tag, all + static_cast<size_t>(p) * k, bytes));
  }
  for (int p = 0; p < P; ++p) {
    if (p == rank_) continue;
    ops.push_back(send(p, kNetCtrl, tag, mine, bytes));
  }
  std::memcpy(all + static_cast<size_t>(rank_) * k, mine, bytes);
  for (uint64_t op : ops) {
    const size_t got = eng_->wait(op);
    M4T_CHECK(got == bytes, "tcp transport: metadata exchange size mismatch (" << got << " vs " << bytes << " bytes)");
  }
}

void NetLink::bcast_i64(int64_t* data, int k, int root) {
  const int P = size();
  const int64_t tag = (++ctrl_seq_) * 64;
  const size_t bytes = static_cast<size_t>(k) * sizeof(int64_t);
  if (rank_ == root) {
    std::vector<uint64_t> ops;
    for (int p = 0; p < P; ++p)
      if (p != root) ops.push_back(send(p, kNetCtrl, tag, data, bytes));
    eng_->wait_all(ops);
  } else {
    eng_->wait(recv(root, kNetCtrl, tag, data, bytes));
  }
}

bool NetLink::quiesce(double timeout_s) noexcept {
  const double saved = eng_->timeout_s();
  bool ok = true;
  try {
    eng_->set_timeout_s(std::min(saved, timeout_s));
    barrier();
  } catch (...) {
    ok = false;
  }
  eng_->set_timeout_s(saved);
  return ok;
}

}  // namespace m4t
