// CTA-pair (cta_group::2) tcgen05 GEMM:  C[M,N] = A[M,K] * B[N,K]^T, bf16.
//
// Two SMs of one TPC form a cluster and share a 256x256 output tile: each CTA
// stages its own 128 rows of A and HALF of the B tile (128 of 256 rows), the
// leader CTA issues tcgen05.mma.cta_group::2 (UMMA 256x256x16) that reads both
// halves, and each CTA keeps its 128x256 half of the accumulator in its own
// TMEM.  Per FLOP this moves 2/3 of the L2->SMEM bytes of the single-CTA kernel
// (gemm_tcgen05.cu) and a stage shrinks to 32 KiB, so the TMA ring is 6 deep
// instead of 4 - the single-CTA kernel is L2/latency bound at ~1.4 PFLOP/s.
//
//   warp 0 (both CTAs)   TMA producer: cp.async.bulk.tensor ... cta_group::2,
//                        transaction bytes credited to the LEADER's mbarrier
//   warp 1 (leader)      MMA issuer; tcgen05.commit multicasts to both CTAs
//   warps 2..5 (both)    epilogue on the CTA's own TMEM lanes; the peer's
//                        "accumulator drained" arrive goes to the leader's
//                        mbarrier through a mapa'd shared::cluster address
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "kernels.h"
#include "fused_comm.cuh"
#include "tcgen05_ptx.cuh"
#include "tma_host.h"
#include "vec_ops.cuh"

namespace m4t {

namespace {

constexpr int BMC = 128;        // rows of A per CTA
constexpr int BM2 = 2 * BMC;    // cluster tile rows
constexpr int BN = 256;         // cluster tile cols
constexpr int BNH = BN / 2;     // B rows staged per CTA
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr int kABytes = BMC * BK * 2;  // 16 KiB
constexpr int kBBytes = BNH * BK * 2;  // 16 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kWarps = 6;
constexpr int kCommWarps = 4;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;

struct Gemm2Args {
  void* C;
  int M, N, K;
  int ldc;
  const void* T;
  int ldt;
  float* loss_acc;
  float loss_scale, grad_scale;
  const uint32_t* panel_flags;  // fused mode: per-panel counters (local HBM, written through multicast)
  uint32_t panel_target;
  int prefetch_target;          // M4T_EPI_PREFETCH=1 (experimental): pull the MSE target tile into L2 while the MMA runs
};

struct __align__(8) Bars {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return static_cast<uint32_t>(float_to_bf16_bits(lo)) | (static_cast<uint32_t>(float_to_bf16_bits(hi)) << 16);
}

// B operand stored [K, N] row-major (the contraction dimension is the STRIDED one, as for the weight in
// dgrad = gy @ W): TMA boxes of {64 contiguous n, BK k-rows}, MN-major SWIZZLE_128B canonical layout
// (LBO = one 64-wide chunk, SBO = one 8-row swizzle atom), instruction-descriptor bit 16.
constexpr int kBoxBytesMN = 64 * BK * 2;  // 8 KiB
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>(kBoxBytesMN >> 4) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1u) << 46;
  d |= static_cast<uint64_t>(2u) << 61;
  return d;
}

template <bool FUSED, int EPI, bool BMN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kWarps + (FUSED ? kCommWarps : 0)) * 32, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const Gemm2Args g, const CommArgs cm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = tc::cluster_ctarank();
  const bool leader = cta == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int m_tiles = (g.M + BM2 - 1) / BM2;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bars->full[s], 1);
      tc::mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      tc::mbar_init(&bars->tmem_full[a], 1);
      tc::mbar_init(&bars->tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs (used in the leader)
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_2sm<kTmemCols>(&bars->tmem_base);
  tc::tcgen05_fence_before();
  tc::cluster_sync();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int n_blk = t / m_tiles;
        const int m_blk = t - n_blk * m_tiles;
        if (FUSED) {
          // the weight panel (both halves) must have been all-reduced by every rank
          const uint32_t* flag = g.panel_flags + n_blk;
          // bounded like every other cross-rank wait: a peer that never publishes
          // must end in an error, not in a hung GPU
          if (static_cast<int32_t>(ld_acquire_sys_u32(flag) - g.panel_target) < 0) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned int spins = 0;
            while (static_cast<int32_t>(ld_acquire_sys_u32(flag) - g.panel_target) < 0) {
              if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > cm.sync.timeout_ns) {
                *reinterpret_cast<volatile int*>(cm.sync.err_flag) = kErrTimeout;
                __threadfence_system();
                __trap();
              }
            }
          }
          tc::fence_proxy_async();
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&bars->empty[stage], phase ^ 1);  // own copy, signalled by the multicast commit
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) tc::mbar_arrive_expect_tx(&bars->full[stage], 2 * kStageBytes);  // both CTAs' bytes
          tc::tma_load_2d_2sm(sa, &tmap_a, &bars->full[stage], kb * BK, m_blk * BM2 + static_cast<int>(cta) * BMC);
          if (BMN) {
#pragma unroll
            for (int ch = 0; ch < BNH / 64; ++ch)
              tc::tma_load_2d_2sm(sb + ch * kBoxBytesMN, &tmap_b, &bars->full[stage],
                                  n_blk * BN + static_cast<int>(cta) * BNH + ch * 64, kb * BK);
          } else {
            tc::tma_load_2d_2sm(sb, &tmap_b, &bars->full[stage], kb * BK, n_blk * BN + static_cast<int>(cta) * BNH);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM2, BN) | (BMN ? (1u << 16) : 0u);  // bit 16: B MN-major
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);  // both CTAs drained this accumulator
        tc::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&bars->full[stage], phase);
          tc::tcgen05_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t adesc = tc::make_smem_desc_k_sw128(sa);
          const uint64_t bdesc = BMN ? make_smem_desc_mn_sw128(sb) : tc::make_smem_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);
            // MN-major: 16 k-rows = two 8-row swizzle atoms = 2048 bytes further into every chunk
            const uint64_t koff_b = BMN ? static_cast<uint64_t>((k * UMMA_K * 128) >> 4) : koff;
            tc::umma_bf16_ss_2sm(tmem_d, adesc + koff, bdesc + koff_b, idesc, (kb | k) ? 1u : 0u);
          }
          tc::umma_commit_2sm(&bars->empty[stage]);  // frees the stage in BOTH CTAs
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::umma_commit_2sm(&bars->tmem_full[acc]);  // accumulator ready in BOTH CTAs
      }
    }
  } else if (warp < kWarps) {
    // ===================== epilogue (both CTAs) =========================
    const int q = warp & 3;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const int n_blk = t / m_tiles;
      const int m_blk = t - n_blk * m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_blk * BM2 + static_cast<int>(cta) * BMC + q * 32 + lane;
      if (EPI == 1 && g.prefetch_target && row < g.M) {
        // the epilogue warps idle until the accumulator is complete: use the time to bring this
        // lane's 512 bytes of the target into L2, so the loads below do not pay HBM latency
        const char* tp = static_cast<const char*>(g.T) + (static_cast<int64_t>(row) * g.ldt + n_blk * BN) * 2;
#pragma unroll
        for (int i = 0; i < BN / 64; ++i)
          if (n_blk * BN + (i + 1) * 64 <= g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + i * 128));
      }
      tc::mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc::tcgen05_fence_after();
      uint16_t* crow = static_cast<uint16_t*>(g.C) + static_cast<int64_t>(row) * g.ldc + n_blk * BN;
      const uint16_t* trow = nullptr;
      float loss_part = 0.f;
      if (EPI == 1) trow = static_cast<const uint16_t*>(g.T) + static_cast<int64_t>(row) * g.ldt + n_blk * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + c * 32);
        const int col0 = n_blk * BN + c * 32;
        // all four target vectors of this chunk are requested up front (and before the TMEM
        // load): one HBM round trip per chunk instead of four dependent ones
        Vec16 tv[4];
        if (EPI == 1 && row < g.M) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (col0 + v * 8 + 8 <= g.N) tv[v] = ld_vec_stream(trow + c * 32 + v * 8);
        }
        tc::tmem_ld_32x32b_x32(taddr, r);
        tc::tmem_ld_wait();
        if (row < g.M) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (col0 + v * 8 + 8 <= g.N) {
              Vec16 o;
              if (EPI == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float d0 = __uint_as_float(r[v * 8 + 2 * e]) - bf16_bits_to_float(static_cast<uint16_t>(tv[v].w[e] & 0xffffu));
                  const float d1 = __uint_as_float(r[v * 8 + 2 * e + 1]) - bf16_bits_to_float(static_cast<uint16_t>(tv[v].w[e] >> 16));
                  loss_part = fmaf(d0, d0, loss_part);
                  loss_part = fmaf(d1, d1, loss_part);
                  o.w[e] = pack2(d0 * g.grad_scale, d1 * g.grad_scale);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  o.w[e] = pack2(__uint_as_float(r[v * 8 + 2 * e]), __uint_as_float(r[v * 8 + 2 * e + 1]));
              }
              st_vec(crow + c * 32 + v * 8, o);
            }
          }
        }
      }
      if (EPI == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
        if (lane == 0) atomicAdd(g.loss_acc, loss_part * g.loss_scale);
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        // "drained" is collected by the leader's barrier (count 8)
        if (leader) tc::mbar_arrive(&bars->tmem_empty[acc]);
        else tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bars->tmem_empty[acc]), 0));
      }
    }
  }

  else if (FUSED) {
    // ===================== communication (both CTAs) ====================
    comm_allreduce_panels<kCommWarps, BN>(cm, kWarps * 32, n_tiles, g.K);
  }

  tc::tcgen05_fence_before();
  tc::cluster_sync();
  if (warp == 1) tc::tmem_dealloc_2sm<kTmemCols>(tmem_base);
}

CUtensorMap make_tmap2(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return make_tmap_bf16_sw128(base, rows, cols, ld, BK, box_rows);
}

template <bool FUSED, int EPI, bool BMN = false> void configure2() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<FUSED, EPI, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    M4T_CHECK(e == cudaSuccess, "cudaFuncSetAttribute(smem) failed: " << cudaGetErrorString(e));
  });
}

}  // namespace

bool gemm_bf16_nn_supported(int64_t M, int64_t N, int64_t K, const void* A, const void* B, const void* C, int64_t lda,
                            int64_t ldb, int64_t ldc) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && al(A) && al(B) && al(C) && lda % 8 == 0 && ldb % 8 == 0 &&
         ldc % 8 == 0 && lda >= K && ldb >= N && ldc >= N && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31);
}

// C[M,N] = A[M,K] * B[K,N], both row-major (the dgrad shape: gy @ W).
void launch_gemm_bf16_nn_2cta(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                              int64_t ldb, int64_t ldc, int sm_count, cudaStream_t stream) {
  M4T_CHECK(gemm_bf16_nn_supported(M, N, K, A, B, C, lda, ldb, ldc), "unsupported GEMM shape/alignment for the tcgen05 NN path");
  const CUtensorMap ta = make_tmap2(A, M, K, lda, BMC);
  const CUtensorMap tb = make_tmap_bf16_sw128(B, K, N, ldb, 64, BK);  // boxes of {64 n, BK k-rows}
  Gemm2Args g{};
  g.C = C;
  g.M = static_cast<int>(M);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldc = static_cast<int>(ldc);
  const int tiles = static_cast<int>(((M + BM2 - 1) / BM2) * ((N + BN - 1) / BN));
  const int clusters = std::max(1, std::min(tiles, sm_count / 2));
  configure2<false, 0, true>();
  gemm_bf16_tn_2cta_kernel<false, 0, true><<<2 * clusters, kWarps * 32, kSmemBytes, stream>>>(ta, tb, g, CommArgs{});
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, "gemm_bf16_nn_2cta launch failed: " << cudaGetErrorString(e));
  note_kernel_launch("gemm_2cta_nn");
}

void launch_gemm_bf16_tn_2cta(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                              int64_t ldb, int64_t ldc, int sm_count, cudaStream_t stream, const MseEpilogue* mse) {
  M4T_CHECK(gemm_bf16_tn_supported(M, N, K, A, B, C, lda, ldb, ldc), "unsupported GEMM shape/alignment for the tcgen05 path");
  const CUtensorMap ta = make_tmap2(A, M, K, lda, BMC);
  const CUtensorMap tb = make_tmap2(B, N, K, ldb, BNH);
  Gemm2Args g{};
  g.C = C;
  g.M = static_cast<int>(M);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldc = static_cast<int>(ldc);
  const int tiles = static_cast<int>(((M + BM2 - 1) / BM2) * ((N + BN - 1) / BN));
  const int clusters = std::max(1, std::min(tiles, sm_count / 2));
  static const int epi_prefetch = static_cast<int>(env_i64("M4T_EPI_PREFETCH", 0));
  g.prefetch_target = epi_prefetch;
  if (mse) {
    g.T = mse->target;
    g.ldt = static_cast<int>(mse->ldt);
    g.loss_acc = mse->loss_acc;
    g.loss_scale = mse->loss_scale;
    g.grad_scale = mse->grad_scale;
    configure2<false, 1>();
    gemm_bf16_tn_2cta_kernel<false, 1><<<2 * clusters, kWarps * 32, kSmemBytes, stream>>>(ta, tb, g, CommArgs{});
  } else {
    configure2<false, 0>();
    gemm_bf16_tn_2cta_kernel<false, 0><<<2 * clusters, kWarps * 32, kSmemBytes, stream>>>(ta, tb, g, CommArgs{});
  }
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, "gemm_bf16_tn_2cta launch failed: " << cudaGetErrorString(e));
  note_kernel_launch(mse ? "gemm_2cta_mse_epilogue" : "gemm_2cta");
}

// Fused Allreduce->GEMM on CTA pairs.  Same contract as launch_fused_allreduce_gemm.
void launch_fused_allreduce_gemm_2cta(const DeviceComm& dc, const void* x, void* y, int64_t M, int64_t N, int64_t K,
                                      int64_t ldx, int64_t ldy, int64_t w_off, int64_t wavg_off, int64_t flags_off,
                                      uint32_t panel_target, float scale, cudaStream_t stream, const MseEpilogue* mse) {
  M4T_CHECK(dc.mc_heap != nullptr, "the fused Allreduce->GEMM kernel needs the NVLS multicast mapping");
  const int P = dc.sync.size;
  M4T_CHECK(BN % P == 0 && N % BN == 0 && (K * 2) % 16 == 0, "fused Allreduce->GEMM: N must be a multiple of 256 and 256 % ranks == 0");
  const char* wavg = dc.heap[dc.sync.rank] + wavg_off;
  M4T_CHECK(gemm_bf16_tn_supported(M, N, K, x, wavg, y, ldx, K, ldy), "unsupported GEMM shape/alignment for the fused path");
  const CUtensorMap ta = make_tmap2(x, M, K, ldx, BMC);
  const CUtensorMap tb = make_tmap2(wavg, N, K, K, BNH);
  Gemm2Args g{};
  g.C = y;
  g.M = static_cast<int>(M);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldc = static_cast<int>(ldy);
  g.panel_flags = reinterpret_cast<const uint32_t*>(dc.heap[dc.sync.rank] + flags_off);
  g.panel_target = panel_target;
  static const int epi_prefetch = static_cast<int>(env_i64("M4T_EPI_PREFETCH", 0));
  g.prefetch_target = epi_prefetch;
  CommArgs cm{};
  cm.sync = dc.sync;
  cm.mc_heap = dc.mc_heap;
  cm.my_heap = dc.heap[dc.sync.rank];
  cm.w_off = w_off;
  cm.wavg_off = wavg_off;
  cm.flags_off = flags_off;
  cm.scale = scale;
  cm.do_barrier = 1;
  static const int fused_debug = static_cast<int>(env_i64("M4T_FUSED_DEBUG", 0));  // read once
  cm.debug_skip = fused_debug;
  if (cm.debug_skip & 2) g.M = 0;  // timing experiment: communication only, no GEMM tiles
  const int grid = fused_gemm_grid(dc);  // identical on every rank, even (whole CTA pairs)
  if (mse) {
    g.T = mse->target;
    g.ldt = static_cast<int>(mse->ldt);
    g.loss_acc = mse->loss_acc;
    g.loss_scale = mse->loss_scale;
    g.grad_scale = mse->grad_scale;
    configure2<true, 1>();
    gemm_bf16_tn_2cta_kernel<true, 1><<<grid, (kWarps + kCommWarps) * 32, kSmemBytes, stream>>>(ta, tb, g, cm);
  } else {
    configure2<true, 0>();
    gemm_bf16_tn_2cta_kernel<true, 0><<<grid, (kWarps + kCommWarps) * 32, kSmemBytes, stream>>>(ta, tb, g, cm);
  }
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, "fused_allreduce_gemm_2cta launch failed: " << cudaGetErrorString(e));
  note_kernel_launch(mse ? "fused_allreduce_gemm_2cta_mse" : "fused_allreduce_gemm_2cta");
}

}  // namespace m4t
