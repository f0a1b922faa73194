// Process-wide world: control plane + backends.  The analogue of MPI's global
// state after MPI_Init_thread (reference csrc/extension.cpp:1306-1394), built
// from the launcher environment instead of an MPI runtime.
#pragma once
#include <memory>
#include <mutex>
#include <string>

#include "backend.h"
#include "control.h"
#include "cpu_backend.h"

namespace m4t {

class CudaBackend;

class World {
 public:
  // Lazily created from the environment (RANK / WORLD_SIZE / LOCAL_RANK / M4T_JOB_ID).
  static World& instance();
  static bool initialised();
  // Tears the world down (idempotent); called from Python's atexit.
  static void finalize();

  int rank() const { return env_.rank; }
  int size() const { return env_.size; }
  int local_rank() const { return env_.local_rank; }
  const std::string& job_id() const { return env_.job_id; }

  Control& control() { return *ctl_; }
  CpuBackend& cpu() { return *cpu_; }

  // Collective: every rank must call it (done eagerly at import when CUDA is
  // visible, or explicitly via init_cuda()).  Idempotent.
  void init_cuda(int device);
  bool cuda_ready() const { return cuda_ != nullptr; }
  CudaBackend* cuda() { return cuda_.get(); }

  // Runtime toggle mirroring deactivate_cuda_aware_mpi_support() (reference
  // csrc/extension.cpp:54-59): CUDA tensors are staged through host memory and
  // the CPU shared-memory backend.
  void set_host_staging(bool on) { host_staging_ = on; }
  bool host_staging() const { return host_staging_; }

  // One lock for the forward thread and the autograd engine thread: both draw
  // from the same op order (survey 7.4 "two host threads issue collectives").
  std::recursive_mutex& mutex() { return mu_; }

 private:
  World();
  ~World();
  WorldEnv env_;
  std::unique_ptr<Control> ctl_;
  std::unique_ptr<CpuBackend> cpu_;
  std::unique_ptr<CudaBackend> cuda_;
  bool host_staging_ = false;
  std::recursive_mutex mu_;
};

}  // namespace m4t
