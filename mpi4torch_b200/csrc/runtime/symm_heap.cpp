#include "symm_heap.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstring>
#include <sstream>

namespace m4t {

namespace {

// Driver entry points are resolved at run time so the extension links against
// libcudart only (the build box has no libcuda.so.1).
struct Driver {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  bool ok = false;
  bool has_multicast = false;
};

template <typename F> bool resolve(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

Driver& driver() {
  static Driver d = [] {
    Driver x;
    bool ok = true;
    ok &= resolve("cuGetErrorString", x.GetErrorString);
    ok &= resolve("cuDeviceGet", x.DeviceGet);
    ok &= resolve("cuDeviceGetAttribute", x.DeviceGetAttribute);
    ok &= resolve("cuMemGetAllocationGranularity", x.MemGetAllocationGranularity);
    ok &= resolve("cuMemCreate", x.MemCreate);
    ok &= resolve("cuMemRelease", x.MemRelease);
    ok &= resolve("cuMemExportToShareableHandle", x.MemExportToShareableHandle);
    ok &= resolve("cuMemImportFromShareableHandle", x.MemImportFromShareableHandle);
    ok &= resolve("cuMemAddressReserve", x.MemAddressReserve);
    ok &= resolve("cuMemAddressFree", x.MemAddressFree);
    ok &= resolve("cuMemMap", x.MemMap);
    ok &= resolve("cuMemUnmap", x.MemUnmap);
    ok &= resolve("cuMemSetAccess", x.MemSetAccess);
    x.ok = ok;
    bool mc = true;
    mc &= resolve("cuMulticastCreate", x.MulticastCreate);
    mc &= resolve("cuMulticastAddDevice", x.MulticastAddDevice);
    mc &= resolve("cuMulticastBindMem", x.MulticastBindMem);
    mc &= resolve("cuMulticastGetGranularity", x.MulticastGetGranularity);
    mc &= resolve("cuMulticastUnbind", x.MulticastUnbind);
    x.has_multicast = ok && mc;
    return x;
  }();
  return d;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  return std::string(s ? s : "unknown") + " (" + std::to_string(static_cast<int>(r)) + ")";
}

#define M4T_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t m4t_e_ = (expr);                                                         \
    M4T_CHECK(m4t_e_ == cudaSuccess, #expr << " failed: " << cudaGetErrorString(m4t_e_)); \
  } while (0)

// Driver call inside a "try" section: records the first failure instead of throwing.
#define M4T_CU_TRY(ok, why, expr)                             \
  do {                                                        \
    if (ok) {                                                 \
      CUresult m4t_r_ = (expr);                               \
      if (m4t_r_ != CUDA_SUCCESS) {                           \
        ok = false;                                           \
        why = std::string(#expr) + ": " + cu_err(m4t_r_);     \
      }                                                       \
    }                                                         \
  } while (0)

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

HeapCaps probe_heap_caps(int device) {
  HeapCaps c;
  M4T_CUDA(cudaSetDevice(device));
  M4T_CUDA(cudaFree(nullptr));  // create the primary context
  cudaDeviceProp prop;
  M4T_CUDA(cudaGetDeviceProperties(&prop, device));
  c.sm_count = prop.multiProcessorCount;
  c.total_mem = prop.totalGlobalMem;
  c.name = prop.name;
  Driver& d = driver();
  if (!d.ok) return c;
  CUdevice dev;
  if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return c;
  int v = 0;
  if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev) == CUDA_SUCCESS) c.vmm = v != 0;
  v = 0;
  if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS) c.posix_fd = v != 0;
  v = 0;
  if (d.has_multicast && d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS) c.multicast = v != 0;
  return c;
}

namespace {
// All ranks pass their local verdict; returns true iff every rank said true.
bool all_agree(Control& ctl, bool mine) {
  std::vector<int64_t> all(static_cast<size_t>(ctl.size()));
  int64_t m = mine ? 1 : 0;
  ctl.allgather_i64(&m, 1, all.data());
  for (int64_t v : all)
    if (!v) return false;
  return true;
}
}  // namespace

SymmHeap::SymmHeap(Control& ctl, int device, size_t bytes)
    : ctl_(ctl), device_(device), rank_(ctl.rank()), size_(ctl.size()), bytes_(bytes) {
  M4T_CHECK(size_ <= 16, "the NVLink backend supports up to 16 ranks per communicator (got " << size_ << ")");
  caps_ = probe_heap_caps(device);
  peers_.assign(static_cast<size_t>(size_), nullptr);
  peer_handles_.assign(static_cast<size_t>(size_), 0ull);
  const bool env_vmm = env_i64("M4T_VMM", 1) != 0;
  const bool env_nvls = env_i64("M4T_NVLS", 1) != 0;
  const bool try_vmm = all_agree(ctl_, caps_.vmm && caps_.posix_fd && env_vmm && driver().ok);
  const bool try_mc = try_vmm && size_ > 1 && all_agree(ctl_, caps_.multicast && env_nvls);
  bool done = false;
  if (try_vmm) {
    try {
      setup_vmm(try_mc);
      done = true;
    } catch (const std::exception& e) {
      // setup_vmm only throws at points where every rank throws
      if (rank_ == 0) std::fprintf(stderr, "[m4t] VMM heap unavailable (%s); falling back to cudaIpc\n", e.what());
      teardown();
    }
  }
  if (!done) setup_ipc();
  M4T_CUDA(cudaMemset(peers_[rank_], 0, bytes_));
  M4T_CUDA(cudaDeviceSynchronize());
  ctl_.barrier();
  M4T_LOG("rank %d heap: %s", rank_, describe().c_str());
}

SymmHeap::~SymmHeap() { teardown(); }

std::string SymmHeap::describe() const {
  std::ostringstream o;
  o << (mode_ == HeapMode::VMM_MULTICAST ? "vmm+multicast(NVLS)" : mode_ == HeapMode::VMM ? "vmm" : "cudaIpc")
    << " " << (bytes_ >> 20) << " MiB x " << size_ << " ranks on " << caps_.name;
  return o.str();
}

void SymmHeap::setup_vmm(bool want_mc) {
  Driver& d = driver();
  CUdevice dev;
  M4T_CHECK(d.DeviceGet(&dev, device_) == CUDA_SUCCESS, "cuDeviceGet failed");
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;

  CUmulticastObjectProp mcprop;
  std::memset(&mcprop, 0, sizeof(mcprop));
  mcprop.numDevices = static_cast<unsigned int>(size_);
  mcprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;

  bool ok = true;
  std::string why;
  size_t gran = 0, mcgran = 0;
  M4T_CU_TRY(ok, why, d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (ok && gran == 0) gran = 2u << 20;
  size_t sz = round_up(bytes_, gran ? gran : (2u << 20));
  if (want_mc) {
    mcprop.size = sz;
    bool mok = true;
    std::string mwhy;
    M4T_CU_TRY(mok, mwhy, d.MulticastGetGranularity(&mcgran, &mcprop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (mok && mcgran) sz = round_up(sz, mcgran);
    else mok = false;
    want_mc = all_agree(ctl_, mok);  // entered by every rank (want_mc is globally agreed on entry)
  }
  // the rounded size must be identical everywhere
  {
    std::vector<int64_t> all(static_cast<size_t>(size_));
    int64_t mine = static_cast<int64_t>(sz);
    ctl_.allgather_i64(&mine, 1, all.data());
    for (int64_t v : all) sz = std::max(sz, static_cast<size_t>(v));
  }
  bytes_ = sz;
  mcprop.size = sz;

  CUmemGenericAllocationHandle h = 0;
  int heap_fd = -1;
  M4T_CU_TRY(ok, why, d.MemCreate(&h, sz, &prop, 0));
  if (ok) local_handle_ = h;
  M4T_CU_TRY(ok, why, d.MemExportToShareableHandle(&heap_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  if (!all_agree(ctl_, ok)) {
    if (heap_fd >= 0) close(heap_fd);
    M4T_CHECK(false, "cuMemCreate/export failed on some rank: " << why);
  }

  // rank 0 creates the multicast object
  int mc_fd = -1;
  bool mc_ok = want_mc;
  if (want_mc && rank_ == 0) {
    std::string mwhy;
    CUmemGenericAllocationHandle mh = 0;
    M4T_CU_TRY(mc_ok, mwhy, d.MulticastCreate(&mh, &mcprop));
    if (mc_ok) mc_handle_ = mh;
    M4T_CU_TRY(mc_ok, mwhy, d.MemExportToShareableHandle(&mc_fd, mh, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    if (!mc_ok) std::fprintf(stderr, "[m4t] multicast object creation failed (%s); NVLS disabled\n", mwhy.c_str());
  }
  {
    std::vector<int64_t> all(static_cast<size_t>(size_));
    int64_t mine = mc_ok ? 1 : 0;
    ctl_.allgather_i64(&mine, 1, all.data());
    mc_ok = want_mc && all[0] != 0;
  }

  std::vector<int> mine_fds{heap_fd, (mc_ok && rank_ == 0) ? mc_fd : heap_fd};
  auto fds = ctl_.exchange_fds(mine_fds);

  CUmemAccessDesc access;
  std::memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = device_;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;

  for (int p = 0; p < size_; ++p) {
    CUmemGenericAllocationHandle ph = 0;
    if (p == rank_) {
      ph = local_handle_;
    } else {
      M4T_CU_TRY(ok, why, d.MemImportFromShareableHandle(&ph, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[p][0])),
                                                        CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      if (ok) peer_handles_[p] = ph;
    }
    CUdeviceptr va = 0;
    M4T_CU_TRY(ok, why, d.MemAddressReserve(&va, sz, gran, 0, 0));
    M4T_CU_TRY(ok, why, d.MemMap(va, sz, 0, ph, 0));
    M4T_CU_TRY(ok, why, d.MemSetAccess(va, sz, &access, 1));
    if (ok) peers_[p] = reinterpret_cast<char*>(va);
  }
  const bool mapped = all_agree(ctl_, ok);
  if (!mapped) {
    for (int p = 0; p < size_; ++p)
      for (int fd : fds[p])
        if (p != rank_ && fd >= 0) close(fd);
    if (heap_fd >= 0) close(heap_fd);
    if (mc_fd >= 0) close(mc_fd);
    M4T_CHECK(false, "peer mapping failed on some rank: " << why);
  }
  mode_ = HeapMode::VMM;

  if (mc_ok) {
    bool mok = true;
    std::string mwhy;
    if (rank_ != 0) {
      CUmemGenericAllocationHandle mh = 0;
      M4T_CU_TRY(mok, mwhy, d.MemImportFromShareableHandle(&mh, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[0][1])),
                                                          CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      if (mok) mc_handle_ = mh;
    }
    M4T_CU_TRY(mok, mwhy, d.MulticastAddDevice(mc_handle_, dev));
    bool all_added = all_agree(ctl_, mok);  // every device must be added before any bind
    if (all_added) {
      M4T_CU_TRY(mok, mwhy, d.MulticastBindMem(mc_handle_, 0, local_handle_, 0, sz, 0));
      if (mok) mc_bound_ = true;
      CUdeviceptr mva = 0;
      M4T_CU_TRY(mok, mwhy, d.MemAddressReserve(&mva, sz, mcgran ? mcgran : gran, 0, 0));
      M4T_CU_TRY(mok, mwhy, d.MemMap(mva, sz, 0, mc_handle_, 0));
      M4T_CU_TRY(mok, mwhy, d.MemSetAccess(mva, sz, &access, 1));
      if (mok) mc_ = reinterpret_cast<char*>(mva);
      all_added = all_agree(ctl_, mok);
    }
    if (all_added) {
      mode_ = HeapMode::VMM_MULTICAST;
    } else {
      if (rank_ == 0) std::fprintf(stderr, "[m4t] multicast bind/map failed (%s); NVLS disabled\n", mwhy.c_str());
      mc_ = nullptr;  // mapping (if any) is leaked deliberately: unicast path stays valid
    }
  }
  for (int p = 0; p < size_; ++p)
    for (int fd : fds[p])
      if (p != rank_ && fd >= 0) close(fd);
  if (heap_fd >= 0) close(heap_fd);
  if (mc_fd >= 0) close(mc_fd);
}

void SymmHeap::setup_ipc() {
  mode_ = HeapMode::IPC;
  M4T_CUDA(cudaSetDevice(device_));
  bytes_ = round_up(bytes_, 2u << 20);
  {
    std::vector<int64_t> all(static_cast<size_t>(size_));
    int64_t mine = static_cast<int64_t>(bytes_);
    ctl_.allgather_i64(&mine, 1, all.data());
    for (int64_t v : all) bytes_ = std::max(bytes_, static_cast<size_t>(v));
  }
  M4T_CUDA(cudaMalloc(&ipc_base_, bytes_));
  peers_[rank_] = static_cast<char*>(ipc_base_);
  if (size_ == 1) return;
  cudaIpcMemHandle_t hnd;
  M4T_CUDA(cudaIpcGetMemHandle(&hnd, ipc_base_));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected cudaIpcMemHandle_t size");
  int64_t words[8];
  std::memcpy(words, &hnd, 64);
  std::vector<int64_t> all(static_cast<size_t>(size_) * 8);
  ctl_.allgather_i64(words, 8, all.data());
  for (int p = 0; p < size_; ++p) {
    if (p == rank_) continue;
    cudaIpcMemHandle_t ph;
    std::memcpy(&ph, all.data() + static_cast<size_t>(p) * 8, 64);
    void* ptr = nullptr;
    M4T_CUDA(cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess));
    peers_[p] = static_cast<char*>(ptr);
  }
}

void SymmHeap::teardown() {
  Driver& d = driver();
  if (ipc_base_) {
    for (int p = 0; p < size_; ++p)
      if (p != rank_ && peers_[p]) cudaIpcCloseMemHandle(peers_[p]);
    cudaFree(ipc_base_);
    ipc_base_ = nullptr;
  } else if (d.ok) {
    if (mc_) {
      d.MemUnmap(reinterpret_cast<CUdeviceptr>(mc_), bytes_);
      d.MemAddressFree(reinterpret_cast<CUdeviceptr>(mc_), bytes_);
    }
    if (mc_bound_) {
      CUdevice dev;
      if (d.DeviceGet(&dev, device_) == CUDA_SUCCESS) d.MulticastUnbind(mc_handle_, dev, 0, bytes_);
    }
    if (mc_handle_) d.MemRelease(mc_handle_);
    for (int p = 0; p < size_; ++p) {
      if (peers_[p]) {
        d.MemUnmap(reinterpret_cast<CUdeviceptr>(peers_[p]), bytes_);
        d.MemAddressFree(reinterpret_cast<CUdeviceptr>(peers_[p]), bytes_);
      }
      if (p != rank_ && peer_handles_[p]) d.MemRelease(peer_handles_[p]);
    }
    if (local_handle_) d.MemRelease(local_handle_);
  }
  for (auto& p : peers_) p = nullptr;
  for (auto& h : peer_handles_) h = 0;
  mc_ = nullptr;
  mc_handle_ = 0;
  mc_bound_ = false;
  local_handle_ = 0;
}

}  // namespace m4t
