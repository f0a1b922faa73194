// Process-wide world: control plane + backends.  The analogue of MPI's global
// state after MPI_Init_thread (reference csrc/extension.cpp:1306-1394), built
// from the launcher environment instead of an MPI runtime.
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "backend.h"
#include "control.h"
#include "cpu_backend.h"
#include "net_backend.h"
#include "net_link.h"

namespace m4t {

class CudaBackend;

// Everything one communicator needs: its own control segment (barriers,
// metadata exchange, p2p descriptor rings), CPU backend and - when CUDA is up -
// its own symmetric heap + device counters.  COMM_WORLD's context is created
// from the launcher environment; Split() creates further contexts over a subset
// of ranks (the reference obtains such communicators through mpi4py,
// src/__init__.py:247-261).
class CommContext {
 public:
  CommContext(int rank, int size, const std::string& job_id);
  // Job that spans nodes: the communicator lives on the TCP mesh (net_link.h).  The shared-memory control segment
  // then only holds this process (nothing is shared through it) and there is no CUDA backend.
  // With local_size > 1 (the same number of ranks on every node, ranks numbered node by node) the ranks of one node
  // additionally share a control segment and Allreduce becomes hierarchical (HierBackend, net_backend.h).
  CommContext(std::shared_ptr<NetLink> link, const std::string& job_id, int local_rank = 0, int local_size = 1);
  ~CommContext();
  int rank() const { return net_ ? net_->rank() : ctl_->rank(); }
  int size() const { return net_ ? net_->size() : ctl_->size(); }
  const std::string& job_id() const { return job_id_; }
  Control& control() { return *ctl_; }
  CpuBackend& cpu() { return *cpu_; }
  // the backend that moves host tensors: POSIX shared memory on one node, TCP across nodes
  Backend& host() {
    if (hier_) return *hier_;
    return net_ ? static_cast<Backend&>(*netbe_) : static_cast<Backend&>(*cpu_);
  }
  bool hierarchical() const { return hier_ != nullptr; }
  // ranks per node when every node holds the same number of consecutive world ranks, else 0
  int ranks_per_node() const { return ranks_per_node_; }
  bool over_network() const { return net_ != nullptr; }
  const std::shared_ptr<NetLink>& net() const { return net_; }
  // host control operations of the communicator (shared-memory flags or TCP messages)
  void barrier();
  void allgather_i64(const int64_t* mine, int k, int64_t* all);
  // Collective over the context's ranks.  stage_mb / symm_mb < 0 = environment defaults.
  void init_cuda(int device, int64_t stage_mb = -1, int64_t symm_mb = -1);
  void shutdown_cuda();
  // Collective teardown (handshake, then backends, then the control segment).  Idempotent;
  // also run by the destructor.  World::finalize() calls it on every context that is still
  // registered, even if Python objects keep the context object itself alive.
  void shutdown();
  bool alive() const { return ctl_ != nullptr; }
  bool cuda_ready() const { return cuda_ != nullptr; }
  CudaBackend* cuda() { return cuda_.get(); }
  uint64_t next_split_id() { return ++split_seq_; }

 private:
  std::string job_id_;
  std::unique_ptr<Control> ctl_;
  std::unique_ptr<CpuBackend> cpu_;
  std::unique_ptr<CudaBackend> cuda_;
  std::shared_ptr<NetLink> net_;
  std::unique_ptr<NetBackend> netbe_;
  std::unique_ptr<HierBackend> hier_;
  int ranks_per_node_ = 0;
  uint64_t split_seq_ = 0;
};

class World {
 public:
  // Lazily created from the environment (RANK / WORLD_SIZE / LOCAL_RANK / M4T_JOB_ID).
  static World& instance();
  static bool initialised();
  // Multi-node start-up (before the first instance()): the mesh this process is part of.  The world communicator is
  // then created on it instead of on a shared-memory segment.
  static void set_network(std::shared_ptr<NetEngine> engine);
  // Tears the world down (idempotent); called from Python's atexit.
  static void finalize();

  int rank() const { return env_.rank; }
  int size() const { return env_.size; }
  int local_rank() const { return env_.local_rank; }
  const std::string& job_id() const { return env_.job_id; }

  // the world communicator's context
  const std::shared_ptr<CommContext>& ctx() { return ctx_; }
  Control& control() { return ctx_->control(); }
  CpuBackend& cpu() { return ctx_->cpu(); }

  // Collective: every rank must call it (done eagerly at import when CUDA is
  // visible, or explicitly via init_cuda()).  Idempotent.
  void init_cuda(int device);
  bool cuda_ready() const { return ctx_->cuda_ready(); }
  CudaBackend* cuda() { return ctx_->cuda(); }
  // Sub-communicator contexts live until Free() or finalize(): like MPI_Comm_free, releasing
  // a communicator is collective, so dropping the last Python reference must not tear down
  // segments that slower members of the group are still reading.  Torn down in creation
  // order (identical on all members), before the world.
  void register_child(const std::shared_ptr<CommContext>& c) { children_.push_back(c); }
  void release_child(const CommContext* c);

  // Runtime toggle mirroring deactivate_cuda_aware_mpi_support() (reference
  // csrc/extension.cpp:54-59): CUDA tensors are staged through host memory and
  // the CPU shared-memory backend.
  // Device this rank would use for NVLink communicators (-1: none).  In a job that spans nodes the world communicator
  // has no CUDA backend, but a sub-communicator whose members all live on one node gets one.
  void set_node_cuda_device(int device) { node_cuda_device_ = device; }
  int node_cuda_device() const { return node_cuda_device_; }
  void set_host_staging(bool on) { host_staging_ = on; }
  bool host_staging() const { return host_staging_; }

  // One lock for the forward thread and the autograd engine thread: both draw
  // from the same op order (survey 7.4 "two host threads issue collectives").
  std::recursive_mutex& mutex() { return mu_; }

 private:
  World();
  ~World();
  WorldEnv env_;
  std::shared_ptr<CommContext> ctx_;
  std::vector<std::shared_ptr<CommContext>> children_;
  bool host_staging_ = false;
  int node_cuda_device_ = -1;
  std::recursive_mutex mu_;
};

}  // namespace m4t
