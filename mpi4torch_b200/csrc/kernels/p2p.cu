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
//
// Large messages take the COPY-ENGINE path instead (launch_p2p_*_ce): the same ring, the same flags and
// the same chunk numbering, but the payload moves with cudaMemcpyAsync - user buffer -> ring on the
// sender's DMA engine, ring -> destination over NVLink on the receiver's DMA engine - in groups of up to
// half a ring, and the flags are waited for / posted by one-warp kernels between the copies.  No SM
// copies a byte, so a transfer overlaps a GEMM on the compute stream without taking SMs from it (the
// reference's MPI_Isend/Irecv progress on the host; here the DMA engines do).
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

// One warp; lane i handles slot s0 + i of a group of g <= 32 consecutive slots.
// wait: flag[s0+i] >= target0 + i (skipped for lanes whose target is <= 0, i.e. first use of the slot)
__global__ void __launch_bounds__(32) p2p_flag_wait_kernel(const uint32_t* flags, int s0, int g, long long target0,
                                                            SyncCtx sync) {
  const int i = threadIdx.x;
  if (i < g) {
    const long long tgt = target0 + i;
    if (tgt > 0) wait_flag_ge(flags + s0 + i, static_cast<uint32_t>(tgt), sync);
  }
}
// post: flag[s0+i] = value0 + i, released at system scope (the DMA copy before this kernel in stream
// order has completed, so the payload is visible before the flag)
__global__ void __launch_bounds__(32) p2p_flag_post_kernel(uint32_t* flags, int s0, int g, unsigned long long value0) {
  const int i = threadIdx.x;
  if (i < g) st_release_sys_u32(flags + s0 + i, static_cast<uint32_t>(value0 + static_cast<unsigned long long>(i)));
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
  e = cudaFuncGetAttributes(&attr, p2p_flag_wait_kernel);
  M4T_CHECK(e == cudaSuccess, "loading p2p_flag_wait_kernel failed: " << cudaGetErrorString(e));
  e = cudaFuncGetAttributes(&attr, p2p_flag_post_kernel);
  M4T_CHECK(e == cudaSuccess, "loading p2p_flag_post_kernel failed: " << cudaGetErrorString(e));
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

// ---------------------------------------------------------------------------
// copy-engine path (large messages)
// ---------------------------------------------------------------------------
namespace {

// Walks a message in groups of consecutive chunks that neither wrap the ring nor exceed half of it
// (so the sender can refill one half while the receiver drains the other) nor 32 slots (one lane each).
template <typename F> void for_each_group(const P2pChannel& ch, int64_t bytes, unsigned long long first_chunk, F&& f) {
  const int64_t nchunks = p2p_num_chunks(bytes, ch.slot_bytes);
  const int64_t gmax = std::max<int64_t>(1, std::min<int64_t>(ch.nslots / 2, 32));
  for (int64_t k0 = 0; k0 < nchunks;) {
    const unsigned long long cidx0 = first_chunk + static_cast<unsigned long long>(k0);
    const int s0 = static_cast<int>(cidx0 % static_cast<unsigned long long>(ch.nslots));
    const int64_t g = std::min<int64_t>(std::min<int64_t>(gmax, nchunks - k0), ch.nslots - s0);
    const int64_t off = k0 * ch.slot_bytes;
    const int64_t len = std::max<int64_t>(0, std::min<int64_t>(g * ch.slot_bytes, bytes - off));
    f(cidx0, s0, static_cast<int>(g), off, len);
    k0 += g;
  }
}

}  // namespace

void launch_p2p_send_ce(const SyncCtx& sync, const P2pChannel& ch, const void* src, int64_t bytes,
                        unsigned long long first_chunk, cudaStream_t stream) {
  for_each_group(ch, bytes, first_chunk, [&](unsigned long long cidx0, int s0, int g, int64_t off, int64_t len) {
    // the slots must have been drained by the receiver: tail >= (chunk - nslots) + 1
    const long long target0 = static_cast<long long>(cidx0) - ch.nslots + 1;
    if (target0 + g - 1 > 0) {
      p2p_flag_wait_kernel<<<1, 32, 0, stream>>>(ch.tail_flags, s0, g, target0, sync);
      check_launch("p2p_flag_wait");
    }
    if (len > 0) {
      cudaError_t e = cudaMemcpyAsync(ch.slots + static_cast<int64_t>(s0) * ch.slot_bytes, static_cast<const char*>(src) + off,
                                      static_cast<size_t>(len), cudaMemcpyDeviceToDevice, stream);
      M4T_CHECK(e == cudaSuccess, "p2p send copy failed: " << cudaGetErrorString(e));
    }
    p2p_flag_post_kernel<<<1, 32, 0, stream>>>(ch.head_flags, s0, g, cidx0 + 1ull);
    check_launch("p2p_flag_post");
  });
}

void launch_p2p_recv_ce(const SyncCtx& sync, const P2pChannel& ch, void* dst, int64_t bytes,
                        unsigned long long first_chunk, cudaStream_t stream) {
  for_each_group(ch, bytes, first_chunk, [&](unsigned long long cidx0, int s0, int g, int64_t off, int64_t len) {
    p2p_flag_wait_kernel<<<1, 32, 0, stream>>>(ch.head_flags, s0, g, static_cast<long long>(cidx0) + 1, sync);
    check_launch("p2p_flag_wait");
    if (len > 0) {
      // the ring is the peer's memory (or ours in push mode) seen through the symmetric mapping: the
      // receiver's DMA engine pulls it over NVLink
      cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + off, ch.slots + static_cast<int64_t>(s0) * ch.slot_bytes,
                                      static_cast<size_t>(len), cudaMemcpyDeviceToDevice, stream);
      M4T_CHECK(e == cudaSuccess, "p2p receive copy failed: " << cudaGetErrorString(e));
    }
    p2p_flag_post_kernel<<<1, 32, 0, stream>>>(ch.tail_flags, s0, g, cidx0 + 1ull);
    check_launch("p2p_flag_post");
  });
}

}  // namespace m4t
