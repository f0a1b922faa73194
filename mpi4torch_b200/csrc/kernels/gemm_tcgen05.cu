// tcgen05 GEMM for the linear layer:  C[M,N] = A[M,K] * B[N,K]^T   (bf16 in,
// fp32 accumulate in TMEM, bf16 out) - and its fused form, where the B operand
// is the rank-average of every rank's weight and is produced INSIDE the same
// kernel by communication warps (fused_allreduce_gemm, below).
//
// Structure (one persistent CTA per SM, 128x256 output tiles, BK = 64):
//   warp 0      TMA producer   : cp.async.bulk.tensor A/B tiles -> 4-stage smem ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (UMMA 128x256x16),
//                                accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2..5  epilogue       : tcgen05.ld -> bf16 -> 16-byte global stores,
//                                overlapping the next tile's MMAs
//   warps 6..9  communication  : (fused kernel only) NVLS reduce of weight panels:
//                                multimem.ld_reduce over all ranks' weights ->
//                                x 1/size -> multimem.st into every rank's W_avg ->
//                                multimem.red on the panel counter.  The TMA
//                                producer polls that counter before loading a
//                                panel, so the all-reduce of panel p+1.. overlaps
//                                the MMAs on panel p.
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "fused_comm.cuh"
#include "tcgen05_ptx.cuh"
#include "tma_host.h"
#include "vec_ops.cuh"

namespace m4t {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr int kABytes = BM * BK * 2;   // 16 KiB
constexpr int kBBytes = BN * BK * 2;   // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmWarps = 6;
constexpr int kCommWarps = 4;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct GemmArgs {
  void* C;
  int M, N, K;
  int ldc;
  // fused MSE epilogue (EPI == 1): C receives dL/dy = grad_scale * (y - T),
  // loss_acc accumulates loss_scale * sum((y - T)^2); y itself is never stored
  const void* T;
  int ldt;
  float* loss_acc;
  float loss_scale, grad_scale;
  // fused mode ----------------------------------------------------------------
  const uint32_t* panel_flags;  // [N / BN] counters in local HBM (written through multicast)
  uint32_t panel_target;        // a panel is ready when its counter >= target (wrap-safe)
};


struct __align__(8) SharedBarriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  return static_cast<uint32_t>(float_to_bf16_bits(lo)) | (static_cast<uint32_t>(float_to_bf16_bits(hi)) << 16);
}

template <bool FUSED, int EPI>
__global__ void __launch_bounds__((kGemmWarps + (FUSED ? kCommWarps : 0)) * 32, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmArgs g, const CommArgs cm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SharedBarriers* bars = reinterpret_cast<SharedBarriers*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (g.M + BM - 1) / BM;
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
      tc::mbar_init(&bars->tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<kTmemCols>(&bars->tmem_base);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_blk = t / m_tiles;  // M fastest: a wave works on few weight panels
        const int m_blk = t - n_blk * m_tiles;
        if (FUSED) {
          // the weight panel must have been all-reduced (by every rank) first
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
          tc::fence_proxy_async();  // generic-proxy writes (multimem.st) -> async-proxy reads (TMA)
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          tc::mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
          tc::tma_load_2d(sa, &tmap_a, &bars->full[stage], kb * BK, m_blk * BM);
          tc::tma_load_2d(sb, &tmap_b, &bars->full[stage], kb * BK, n_blk * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          tc::mbar_wait(&bars->full[stage], phase);
          tc::tcgen05_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t adesc = tc::make_smem_desc_k_sw128(sa);
          const uint64_t bdesc = tc::make_smem_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle row
            const uint64_t koff = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);
            tc::umma_bf16_ss(tmem_d, adesc + koff, bdesc + koff, idesc, (kb | k) ? 1u : 0u);
          }
          tc::umma_commit(&bars->empty[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::umma_commit(&bars->tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp < kGemmWarps) {
    // =========================== epilogue ===============================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int n_blk = t / m_tiles;
      const int m_blk = t - n_blk * m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      tc::mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc::tcgen05_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
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
                  o.w[e] = pack_bf16x2(d0 * g.grad_scale, d1 * g.grad_scale);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  o.w[e] = pack_bf16x2(__uint_as_float(r[v * 8 + 2 * e]), __uint_as_float(r[v * 8 + 2 * e + 1]));
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
      if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[acc]);
    }
  } else if (FUSED) {
    // =========================== communication ==========================
    comm_allreduce_panels<kCommWarps, BN>(cm, kGemmWarps * 32, n_tiles, g.K);
  }

  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// Row-major [rows, cols] bf16 matrix, box = [box_rows, 64 cols], 128-byte swizzle.
CUtensorMap make_tmap(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return make_tmap_bf16_sw128(base, rows, cols, ld, BK, box_rows);
}

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, what << " launch failed: " << cudaGetErrorString(e));
  note_kernel_launch(what);
}

template <bool FUSED, int EPI> void configure_once() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<FUSED, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    M4T_CHECK(e == cudaSuccess, "cudaFuncSetAttribute(smem) failed: " << cudaGetErrorString(e));
  });
}

}  // namespace

bool gemm_bf16_tn_supported(int64_t M, int64_t N, int64_t K, const void* A, const void* B, const void* C, int64_t lda,
                            int64_t ldb, int64_t ldc) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return M > 0 && N > 0 && K > 0 && (K % 8) == 0 && (N % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (ldc % 8) == 0 &&
         al16(A) && al16(B) && al16(C) && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31);
}

void launch_gemm_bf16_tn(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                         int64_t ldb, int64_t ldc, int sm_count, cudaStream_t stream, const MseEpilogue* mse) {
  M4T_CHECK(gemm_bf16_tn_supported(M, N, K, A, B, C, lda, ldb, ldc), "unsupported GEMM shape/alignment for the tcgen05 path");
  const CUtensorMap ta = make_tmap(A, M, K, lda, BM);
  const CUtensorMap tb = make_tmap(B, N, K, ldb, BN);
  GemmArgs g{};
  g.C = C;
  g.M = static_cast<int>(M);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldc = static_cast<int>(ldc);
  g.panel_flags = nullptr;
  g.panel_target = 0;
  CommArgs cm{};
  const int tiles = static_cast<int>(((M + BM - 1) / BM) * ((N + BN - 1) / BN));
  const int grid = std::max(1, std::min(tiles, sm_count));
  if (mse) {
    g.T = mse->target;
    g.ldt = static_cast<int>(mse->ldt);
    g.loss_acc = mse->loss_acc;
    g.loss_scale = mse->loss_scale;
    g.grad_scale = mse->grad_scale;
    configure_once<false, 1>();
    gemm_bf16_tn_kernel<false, 1><<<grid, kGemmWarps * 32, kSmemBytes, stream>>>(ta, tb, g, cm);
  } else {
    configure_once<false, 0>();
    gemm_bf16_tn_kernel<false, 0><<<grid, kGemmWarps * 32, kSmemBytes, stream>>>(ta, tb, g, cm);
  }
  check_launch("gemm_bf16_tn");
}

// y = x @ mean_ranks(W)^T in ONE kernel.  `w_off`/`wavg_off`/`flags_off` are
// byte offsets inside the symmetric heap (identical on all ranks); the weight
// must already be staged at w_off on every rank.
void launch_fused_allreduce_gemm(const DeviceComm& dc, const void* x, void* y, int64_t M, int64_t N, int64_t K,
                                 int64_t ldx, int64_t ldy, int64_t w_off, int64_t wavg_off, int64_t flags_off,
                                 uint32_t panel_target, float scale, cudaStream_t stream, const MseEpilogue* mse) {
  M4T_CHECK(dc.mc_heap != nullptr, "the fused Allreduce->GEMM kernel needs the NVLS multicast mapping");
  const int P = dc.sync.size;
  M4T_CHECK(BN % P == 0 && N % BN == 0 && (K * 2) % 16 == 0, "fused Allreduce->GEMM: N must be a multiple of 256 and 256 % ranks == 0");
  const char* wavg = dc.heap[dc.sync.rank] + wavg_off;
  M4T_CHECK(gemm_bf16_tn_supported(M, N, K, x, wavg, y, ldx, K, ldy), "unsupported GEMM shape/alignment for the fused path");
  const CUtensorMap ta = make_tmap(x, M, K, ldx, BM);
  const CUtensorMap tb = make_tmap(wavg, N, K, K, BN);
  GemmArgs g{};
  g.C = y;
  g.M = static_cast<int>(M);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldc = static_cast<int>(ldy);
  g.panel_flags = reinterpret_cast<const uint32_t*>(dc.heap[dc.sync.rank] + flags_off);
  g.panel_target = panel_target;
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
  // the grid must be identical on every rank (per-block barrier + counter targets)
  const int grid = fused_gemm_grid(dc);
  if (mse) {
    g.T = mse->target;
    g.ldt = static_cast<int>(mse->ldt);
    g.loss_acc = mse->loss_acc;
    g.loss_scale = mse->loss_scale;
    g.grad_scale = mse->grad_scale;
    configure_once<true, 1>();
    gemm_bf16_tn_kernel<true, 1><<<grid, (kGemmWarps + kCommWarps) * 32, kSmemBytes, stream>>>(ta, tb, g, cm);
  } else {
    configure_once<true, 0>();
    gemm_bf16_tn_kernel<true, 0><<<grid, (kGemmWarps + kCommWarps) * 32, kSmemBytes, stream>>>(ta, tb, g, cm);
  }
  check_launch("fused_allreduce_gemm");
}

int fused_gemm_grid(const DeviceComm& dc) { return std::min(dc.sm_count, kMaxChannels) & ~1; }

}  // namespace m4t
