// Point-to-point transport for Isend / Irecv (reference csrc/extension.cpp:
// 1071-1157 hands the buffer to MPI_Isend/MPI_Irecv and blocks the host in
// MPI_Wait).  Here each directed pair owns a slot ring in the SENDER's
// symmetric heap; both kernels run on side streams so the transfer overlaps the
// compute stream, and Wait is a cudaStreamWaitEvent:
//
//   send kernel (sender GPU)  : wait slot free -> copy chunk into the ring ->
//                               st.release.sys head flag in the RECEIVER's pad
//   recv kernel (receiver GPU): poll head flag locally -> pull the chunk over
//                               NVLink straight into the destination tensor ->
//                               st.release.sys tail flag in the SENDER's pad
//
// Flags are per slot and hold (chunk index + 1), monotone per slot, so blocks
// may own different chunks concurrently (block b takes chunks b, b+G, ...).
#include <algorithm>

#include "kernels.h"

namespace m4t {

namespace {

constexpr int kThreads = 512;

struct P2pArgs {
  SyncCtx sync;
  char* slots;           // send: local ring; recv: the sender's ring through my mapping
  uint32_t* wait_flags;  // send: tail flags (local); recv: head flags (local)
  uint32_t* post_flags;  // send: head flags (in receiver's pad); recv: tail flags (in sender's pad)
  char* user;            // send: source buffer; recv: destination buffer
  int64_t bytes;
  int64_t slot_bytes;
  int64_t nchunks;       // >= 1: a zero-byte message is one empty chunk (flag only)
  unsigned long long first_chunk;
  int nslots;
  int user_aligned;
};

// Block-cooperative copy with kCopyUnroll independent 16-byte requests in
// flight per thread: a peer load takes ~2 us over NVLink, so bandwidth is set by
// memory-level parallelism (512 threads x 8 x 16 B = 64 KiB in flight per CTA).
constexpr int kCopyUnroll = 8;

__device__ __forceinline__ void copy_bytes(char* dst, const char* src, int64_t bytes, bool vec, bool src_remote) {
  if (vec) {
    const int64_t nvec = bytes / 16;
    int64_t i = threadIdx.x;
    for (; i + (kCopyUnroll - 1) * kThreads < nvec; i += kCopyUnroll * kThreads) {
      Vec16 v[kCopyUnroll];
#pragma unroll
      for (int u = 0; u < kCopyUnroll; ++u) {
        const char* p = src + (i + u * kThreads) * 16;
        v[u] = src_remote ? ld_vec_sys(p) : ld_vec_stream(p);
      }
#pragma unroll
      for (int u = 0; u < kCopyUnroll; ++u) st_vec(dst + (i + u * kThreads) * 16, v[u]);
    }
    for (; i < nvec; i += kThreads) {
      const Vec16 v = src_remote ? ld_vec_sys(src + i * 16) : ld_vec_stream(src + i * 16);
      st_vec(dst + i * 16, v);
    }
    for (int64_t j = nvec * 16 + threadIdx.x; j < bytes; j += kThreads)
      dst[j] = *reinterpret_cast<const volatile char*>(src + j);
  } else {
    for (int64_t j = threadIdx.x; j < bytes; j += kThreads) dst[j] = *reinterpret_cast<const volatile char*>(src + j);
  }
}

__global__ void __launch_bounds__(kThreads) p2p_send_kernel(const P2pArgs a) {
  for (int64_t k = blockIdx.x; k < a.nchunks; k += gridDim.x) {
    const unsigned long long cidx = a.first_chunk + static_cast<unsigned long long>(k);
    const int s = static_cast<int>(cidx % static_cast<unsigned long long>(a.nslots));
    if (threadIdx.x == 0 && cidx >= static_cast<unsigned long long>(a.nslots))
      wait_flag_ge(a.wait_flags + s, static_cast<uint32_t>(cidx - a.nslots + 1ull), a.sync);
    __syncthreads();
    const int64_t off = k * a.slot_bytes;
    const int64_t len = max(static_cast<int64_t>(0), min(a.slot_bytes, a.bytes - off));
    copy_bytes(a.slots + static_cast<int64_t>(s) * a.slot_bytes, a.user + off, len, a.user_aligned != 0, false);
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys_u32(a.post_flags + s, static_cast<uint32_t>(cidx + 1ull));
  }
}

__global__ void __launch_bounds__(kThreads) p2p_recv_kernel(const P2pArgs a) {
  for (int64_t k = blockIdx.x; k < a.nchunks; k += gridDim.x) {
    const unsigned long long cidx = a.first_chunk + static_cast<unsigned long long>(k);
    const int s = static_cast<int>(cidx % static_cast<unsigned long long>(a.nslots));
    if (threadIdx.x == 0) wait_flag_ge(a.wait_flags + s, static_cast<uint32_t>(cidx + 1ull), a.sync);
    __syncthreads();
    const int64_t off = k * a.slot_bytes;
    const int64_t len = max(static_cast<int64_t>(0), min(a.slot_bytes, a.bytes - off));
    copy_bytes(a.user + off, a.slots + static_cast<int64_t>(s) * a.slot_bytes, len, a.user_aligned != 0, true);
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys_u32(a.post_flags + s, static_cast<uint32_t>(cidx + 1ull));
  }
}

}  // namespace

int64_t p2p_num_chunks(int64_t bytes, int64_t slot_bytes) {
  return bytes <= 0 ? 1 : (bytes + slot_bytes - 1) / slot_bytes;  // zero-byte messages still carry a flag
}

namespace {

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, what << " launch failed: " << cudaGetErrorString(e));
  note_kernel_launch();
}

P2pArgs make_args(const SyncCtx& sync, const P2pChannel& ch, void* user, int64_t bytes,
                  unsigned long long first_chunk, bool is_send) {
  P2pArgs a;
  a.sync = sync;
  a.slots = ch.slots;
  a.wait_flags = is_send ? ch.tail_flags : ch.head_flags;
  a.post_flags = is_send ? ch.head_flags : ch.tail_flags;
  a.user = static_cast<char*>(user);
  a.bytes = bytes;
  a.slot_bytes = ch.slot_bytes;
  a.nchunks = p2p_num_chunks(bytes, ch.slot_bytes);
  a.first_chunk = first_chunk;
  a.nslots = ch.nslots;
  a.user_aligned = (reinterpret_cast<uintptr_t>(user) & 15u) == 0 && (ch.slot_bytes & 15) == 0;
  return a;
}

}  // namespace

// With CUDA's lazy module loading the first launch of a kernel loads it, which
// can need the device to drain - fatal if the other kernel of the pair is
// already spinning on the ring.  Load both before any message is sent.
void preload_p2p_kernels() {
  cudaFuncAttributes attr;
  cudaError_t e = cudaFuncGetAttributes(&attr, p2p_send_kernel);
  M4T_CHECK(e == cudaSuccess, "loading p2p_send_kernel failed: " << cudaGetErrorString(e));
  e = cudaFuncGetAttributes(&attr, p2p_recv_kernel);
  M4T_CHECK(e == cudaSuccess, "loading p2p_recv_kernel failed: " << cudaGetErrorString(e));
}

void launch_p2p_send(const SyncCtx& sync, const P2pChannel& ch, const void* src, int64_t bytes,
                     unsigned long long first_chunk, int blocks, cudaStream_t stream) {
  P2pArgs a = make_args(sync, ch, const_cast<void*>(src), bytes, first_chunk, true);
  const int64_t nchunks = p2p_num_chunks(bytes, ch.slot_bytes);
  blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(blocks, nchunks)));
  p2p_send_kernel<<<blocks, kThreads, 0, stream>>>(a);
  check_launch("p2p_send");
}

void launch_p2p_recv(const SyncCtx& sync, const P2pChannel& ch, void* dst, int64_t bytes,
                     unsigned long long first_chunk, int blocks, cudaStream_t stream) {
  P2pArgs a = make_args(sync, ch, dst, bytes, first_chunk, false);
  const int64_t nchunks = p2p_num_chunks(bytes, ch.slot_bytes);
  blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(blocks, nchunks)));
  p2p_recv_kernel<<<blocks, kThreads, 0, stream>>>(a);
  check_launch("p2p_recv");
}

}  // namespace m4t
