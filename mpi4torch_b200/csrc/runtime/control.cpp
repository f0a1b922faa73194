#include "control.h"

#include <signal.h>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

namespace m4t {

namespace {
constexpr uint64_t kMagic = 0x6d34745f62323030ull;  // "m4t_b200"

std::string sanitize(const std::string& s) {
  std::string out;
  for (char c : s) out.push_back((std::isalnum(static_cast<unsigned char>(c)) || c == '_') ? c : '_');
  return out;
}
}  // namespace

uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return static_cast<uint64_t>(ts.tv_sec) * 1000000000ull + static_cast<uint64_t>(ts.tv_nsec);
}

WorldEnv world_env_from_environment() {
  WorldEnv w;
  const char* r = std::getenv("RANK");
  const char* s = std::getenv("WORLD_SIZE");
  if (!r) r = std::getenv("M4T_RANK");
  if (!s) s = std::getenv("M4T_WORLD_SIZE");
  if (r && s) {
    w.rank = std::atoi(r);
    w.size = std::atoi(s);
  }
  const char* lr = std::getenv("LOCAL_RANK");
  w.local_rank = lr ? std::atoi(lr) : w.rank;
  M4T_CHECK(w.size >= 1 && w.size <= kMaxRanks, "WORLD_SIZE " << w.size << " outside [1," << kMaxRanks << "]");
  M4T_CHECK(w.rank >= 0 && w.rank < w.size, "RANK " << w.rank << " outside [0," << w.size << ")");
  const char* job = std::getenv("M4T_JOB_ID");
  if (job && *job) {
    w.job_id = sanitize(job);
  } else if (w.size > 1) {
    // All workers of one torchrun / launcher invocation share the parent pid.
    const char* port = std::getenv("MASTER_PORT");
    w.job_id = sanitize(std::string("p") + (port ? port : "0") + "_" + std::to_string(getppid()));
  } else {
    w.job_id = "solo_" + std::to_string(getpid());
  }
  return w;
}

Control::Control(int rank, int size, const std::string& job_id)
    : rank_(rank), size_(size), job_(job_id) {
  timeout_s_ = static_cast<double>(env_i64("M4T_TIMEOUT_S", 300));
  seg_ = "/m4t_" + job_ + "_ctl";
  const size_t bytes = sizeof(ControlBlock);
  int fd = -1;
  if (rank_ == 0) {
    shm_unlink(seg_.c_str());  // stale segment from a crashed job with the same id
    fd = shm_open(seg_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    M4T_CHECK(fd >= 0, "shm_open(create " << seg_ << ") failed: " << std::strerror(errno));
    M4T_CHECK(ftruncate(fd, static_cast<off_t>(bytes)) == 0, "ftruncate failed: " << std::strerror(errno));
  } else {
    uint64_t start = now_ns();
    uint64_t spins = 0;
    for (;;) {
      fd = shm_open(seg_.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && static_cast<size_t>(st.st_size) >= bytes) break;
        close(fd);
        fd = -1;
      }
      backoff(spins);
      M4T_CHECK((now_ns() - start) * 1e-9 < timeout_s_,
                "rank " << rank_ << " timed out waiting for rank 0 to create " << seg_);
    }
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  M4T_CHECK(p != MAP_FAILED, "mmap of control block failed: " << std::strerror(errno));
  cb_ = static_cast<ControlBlock*>(p);
  if (rank_ == 0) {
    // tmpfs pages are zero-filled; only the header needs explicit values.
    cb_->world.store(static_cast<uint32_t>(size_), std::memory_order_relaxed);
    cb_->magic.store(kMagic, std::memory_order_release);
  } else {
    wait_until([&] { return cb_->magic.load(std::memory_order_acquire) == kMagic; }, "control block magic");
    M4T_CHECK(cb_->world.load() == static_cast<uint32_t>(size_),
              "WORLD_SIZE mismatch: rank 0 says " << cb_->world.load() << ", rank " << rank_ << " says " << size_);
  }
  cb_->slots[rank_].pid.store(static_cast<int32_t>(getpid()));
  cb_->slots[rank_].state.store(1, std::memory_order_release);
  cb_->attached.fetch_add(1, std::memory_order_acq_rel);
  wait_until([&] { return cb_->attached.load(std::memory_order_acquire) >= static_cast<uint32_t>(size_); },
             "all ranks to attach");
  barrier();
  // Everyone holds a mapping now: the name can go, the memory lives until the
  // last process exits (robust cleanup even if a rank crashes later).
  if (rank_ == 0) shm_unlink(seg_.c_str());
  M4T_LOG("rank %d/%d attached to job %s", rank_, size_, job_.c_str());
}

Control::~Control() {
  if (cb_) {
    cb_->slots[rank_].state.store(2, std::memory_order_release);
    munmap(cb_, sizeof(ControlBlock));
  }
}

void Control::backoff(uint64_t& spins) {
  ++spins;
  if (idle_hook_ && !in_quiesce_ && (spins & 0x3f) == 0) idle_hook_();
  if (spins < 2000) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  } else if (spins < 20000) {
    sched_yield();
  } else {
    timespec ts{0, 50000};
    nanosleep(&ts, nullptr);
  }
}

void Control::check_abort_timeout(uint64_t start_ns, const char* what) {
  if (cb_ && cb_->abort_flag.load(std::memory_order_acquire)) {
    throw std::runtime_error(std::string("mpi4torch_b200: rank ") +
                             std::to_string(cb_->abort_rank.load()) + " aborted while rank " +
                             std::to_string(rank_) + " was waiting for " + what);
  }
  if ((now_ns() - start_ns) * 1e-9 > timeout_s_) {
    signal_abort();
    throw std::runtime_error(std::string("mpi4torch_b200: rank ") + std::to_string(rank_) +
                             " timed out after " + std::to_string(timeout_s_) + " s waiting for " + what +
                             " (mismatched collective order across ranks?)");
  }
}

void Control::signal_abort() noexcept {
  if (cb_) {
    cb_->abort_rank.store(static_cast<uint32_t>(rank_));
    cb_->abort_flag.store(1, std::memory_order_release);
  }
}

bool Control::aborted() const { return cb_ && cb_->abort_flag.load(std::memory_order_acquire) != 0; }

void Control::barrier() {
  const uint64_t my = ++barrier_count_;
  cb_->slots[rank_].arrive.store(my, std::memory_order_release);
  for (int p = 0; p < size_; ++p) {
    if (p == rank_) continue;
    auto& a = cb_->slots[p].arrive;
    if (a.load(std::memory_order_acquire) >= my) continue;
    wait_until([&] { return a.load(std::memory_order_acquire) >= my; }, "host barrier");
  }
}

bool Control::quiesce(double timeout_s) noexcept {
  if (!cb_ || size_ <= 1) return true;
  in_quiesce_ = true;  // noexcept path: the idle hook (which may raise) stays out of it
  cb_->departing.fetch_add(1, std::memory_order_acq_rel);
  const uint64_t start = now_ns();
  uint64_t spins = 0;
  while (cb_->departing.load(std::memory_order_acquire) < static_cast<uint32_t>(size_)) {
    if (cb_->abort_flag.load(std::memory_order_acquire)) return false;
    if ((now_ns() - start) * 1e-9 > timeout_s) return false;
    if ((spins & 0xfff) == 0xfff) {
      // a peer that died without detaching will never arrive
      for (int p = 0; p < size_; ++p) {
        const int32_t pid = cb_->slots[p].pid.load(std::memory_order_relaxed);
        if (p != rank_ && pid > 0 && kill(pid, 0) != 0 && errno == ESRCH) return false;
      }
    }
    backoff(spins);
  }
  return true;
}

void Control::allgather_i64(const int64_t* mine, int k, int64_t* all) {
  M4T_CHECK(k >= 0 && k <= kMetaWords, "metadata exchange of " << k << " words exceeds " << kMetaWords);
  const int par = static_cast<int>(meta_seq_++ & 1);
  std::memcpy(cb_->slots[rank_].meta[par], mine, sizeof(int64_t) * static_cast<size_t>(k));
  barrier();
  for (int p = 0; p < size_; ++p)
    std::memcpy(all + static_cast<size_t>(p) * k, cb_->slots[p].meta[par], sizeof(int64_t) * static_cast<size_t>(k));
}

void Control::bcast_i64(int64_t* data, int k, int root) {
  M4T_CHECK(k >= 0 && k <= kMetaWords, "metadata broadcast of " << k << " words exceeds " << kMetaWords);
  const int par = static_cast<int>(meta_seq_++ & 1);
  if (rank_ == root) std::memcpy(cb_->slots[rank_].meta[par], data, sizeof(int64_t) * static_cast<size_t>(k));
  barrier();
  if (rank_ != root) std::memcpy(data, cb_->slots[root].meta[par], sizeof(int64_t) * static_cast<size_t>(k));
}

std::string Control::seg_name(const std::string& suffix) const { return "/m4t_" + job_ + "_" + suffix; }

// ---------------------------------------------------------------------------
// fd exchange (SCM_RIGHTS over abstract unix sockets)
// ---------------------------------------------------------------------------
namespace {
socklen_t abstract_addr(sockaddr_un& addr, const std::string& name) {
  std::memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  M4T_CHECK(name.size() + 1 < sizeof(addr.sun_path), "socket name too long: " << name);
  addr.sun_path[0] = '\0';
  std::memcpy(addr.sun_path + 1, name.data(), name.size());
  return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

void send_fds(int sock, int from_rank, const std::vector<int>& fds) {
  int32_t payload[2] = {from_rank, static_cast<int32_t>(fds.size())};
  iovec iov{payload, sizeof(payload)};
  std::vector<char> ctrl(CMSG_SPACE(sizeof(int) * fds.size()));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl.data();
  msg.msg_controllen = ctrl.size();
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int) * fds.size());
  std::memcpy(CMSG_DATA(c), fds.data(), sizeof(int) * fds.size());
  ssize_t n;
  do { n = sendmsg(sock, &msg, 0); } while (n < 0 && errno == EINTR);
  M4T_CHECK(n == static_cast<ssize_t>(sizeof(payload)), "sendmsg(SCM_RIGHTS) failed: " << std::strerror(errno));
}

int recv_fds(int sock, std::vector<int>& fds, size_t expect) {
  int32_t payload[2] = {-1, 0};
  iovec iov{payload, sizeof(payload)};
  std::vector<char> ctrl(CMSG_SPACE(sizeof(int) * expect));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl.data();
  msg.msg_controllen = ctrl.size();
  ssize_t n;
  do { n = recvmsg(sock, &msg, MSG_WAITALL); } while (n < 0 && errno == EINTR);
  M4T_CHECK(n == static_cast<ssize_t>(sizeof(payload)), "recvmsg(SCM_RIGHTS) failed: " << std::strerror(errno));
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  M4T_CHECK(c && c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS, "no SCM_RIGHTS control message");
  size_t nfd = (c->cmsg_len - CMSG_LEN(0)) / sizeof(int);
  M4T_CHECK(nfd == expect && static_cast<size_t>(payload[1]) == expect,
            "expected " << expect << " fds, got " << nfd);
  fds.resize(nfd);
  std::memcpy(fds.data(), CMSG_DATA(c), sizeof(int) * nfd);
  return payload[0];
}
}  // namespace

std::vector<std::vector<int>> Control::exchange_fds(const std::vector<int>& mine) {
  std::vector<std::vector<int>> out(static_cast<size_t>(size_));
  out[rank_] = mine;
  if (size_ == 1) return out;
  const uint64_t round = fd_round_++;
  auto name_of = [&](int r) { return "m4t_" + job_ + "_fd" + std::to_string(r) + "_" + std::to_string(round); };

  int lsock = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  M4T_CHECK(lsock >= 0, "socket() failed: " << std::strerror(errno));
  sockaddr_un addr;
  socklen_t alen = abstract_addr(addr, name_of(rank_));
  M4T_CHECK(bind(lsock, reinterpret_cast<sockaddr*>(&addr), alen) == 0, "bind failed: " << std::strerror(errno));
  M4T_CHECK(listen(lsock, kMaxRanks) == 0, "listen failed: " << std::strerror(errno));
  barrier();  // every listener exists

  std::vector<int> csocks;
  for (int p = 0; p < size_; ++p) {
    if (p == rank_) continue;
    int s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    M4T_CHECK(s >= 0, "socket() failed: " << std::strerror(errno));
    sockaddr_un pa;
    socklen_t plen = abstract_addr(pa, name_of(p));
    int rc;
    do { rc = connect(s, reinterpret_cast<sockaddr*>(&pa), plen); } while (rc < 0 && errno == EINTR);
    M4T_CHECK(rc == 0, "connect to rank " << p << " failed: " << std::strerror(errno));
    send_fds(s, rank_, mine);
    csocks.push_back(s);
  }
  for (int i = 0; i < size_ - 1; ++i) {
    int s;
    do { s = accept(lsock, nullptr, nullptr); } while (s < 0 && errno == EINTR);
    M4T_CHECK(s >= 0, "accept failed: " << std::strerror(errno));
    std::vector<int> fds;
    int from = recv_fds(s, fds, mine.size());
    M4T_CHECK(from >= 0 && from < size_ && from != rank_, "bad sender rank " << from);
    out[from] = fds;
    close(s);
  }
  barrier();  // all transfers done before sockets close
  for (int s : csocks) close(s);
  close(lsock);
  return out;
}

}  // namespace m4t
