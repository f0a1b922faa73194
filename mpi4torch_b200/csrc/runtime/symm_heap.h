// Symmetric heap: one peer-mapped (and, when the fabric allows, NVSwitch
// multicast-bound) arena per rank.  Replaces the reference's "hand the device
// pointer to a CUDA-aware MPI or stage through host memory" helper
// (MPIDeviceHelper, csrc/extension.cpp:61-104) and its CUDA-awareness probe
// (:28-59): here the probe is a capability ladder
//     VMM + multicast (NVLS)  ->  VMM unicast  ->  cudaIpc unicast
// negotiated across ranks at start-up, with M4T_NVLS=0 / M4T_VMM=0 as manual
// overrides.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "control.h"

namespace m4t {

enum class HeapMode : int { IPC = 0, VMM = 1, VMM_MULTICAST = 2 };

struct HeapCaps {
  bool vmm = false;        // CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED
  bool posix_fd = false;   // CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED
  bool multicast = false;  // CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
  int sm_count = 0;
  size_t total_mem = 0;
  std::string name;
};

HeapCaps probe_heap_caps(int device);

class SymmHeap {
 public:
  // Collective over all ranks of `ctl`.  `bytes` is rounded up to the mapping granularity.
  SymmHeap(Control& ctl, int device, size_t bytes);
  ~SymmHeap();
  SymmHeap(const SymmHeap&) = delete;
  SymmHeap& operator=(const SymmHeap&) = delete;

  HeapMode mode() const { return mode_; }
  size_t bytes() const { return bytes_; }
  char* local() const { return peers_[rank_]; }
  char* peer(int p) const { return peers_[p]; }
  char* multicast() const { return mc_; }  // nullptr unless mode() == VMM_MULTICAST
  const HeapCaps& caps() const { return caps_; }
  std::string describe() const;

 private:
  void setup_vmm(bool want_multicast);
  void setup_ipc();
  void teardown();

  Control& ctl_;
  int device_;
  int rank_, size_;
  size_t bytes_;
  HeapMode mode_ = HeapMode::IPC;
  HeapCaps caps_;
  std::vector<char*> peers_;
  char* mc_ = nullptr;
  // opaque driver handles (CUmemGenericAllocationHandle = unsigned long long)
  unsigned long long local_handle_ = 0;
  std::vector<unsigned long long> peer_handles_;
  unsigned long long mc_handle_ = 0;
  bool mc_bound_ = false;
  void* ipc_base_ = nullptr;
};

}  // namespace m4t
