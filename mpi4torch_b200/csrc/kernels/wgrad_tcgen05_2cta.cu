// Weight-gradient GEMM on CTA pairs, optionally fused with the gradient
// all-reduce and the SGD update.
//
//   G[N,K] = dY[Mb,N]^T * X[Mb,K]          (bf16 in, fp32 accumulate in TMEM)
//
// The contraction runs over the batch dimension Mb, which is the STRIDED
// dimension of both row-major operands, so both UMMA operands are MN-major:
// TMA boxes are {64 contiguous elements (128 B swizzle span), BK batch rows} and
// the shared-memory descriptors use the MN-major SWIZZLE_128B canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (LBO = one 64-wide chunk =
// BK*128 B, SBO = one 8-row swizzle atom = 1024 B); instruction-descriptor bits
// 15/16 select MN-major A and B.  No transposed copy of dY or X is ever made.
//
// Fused mode (the backward of the data-parallel linear layer, where the
// reference does wgrad GEMM -> MPI_Allreduce -> optimizer step as three passes
// over the gradient, csrc/extension.cpp:197-260 + the example's SGD loop):
//   * tiles that fill whole waves of the 74 CTA pairs run over the full batch; only the
//     tiles of the last, partial wave are split in `ksplit` batch halves (256 tiles =
//     3 waves of 74 + 34 tiles -> 68 half units: 3.5 tile times instead of 4), so
//     only those tiles produce two partial buffers;
//   * the epilogue writes each partial tile as bf16 into this rank's symmetric
//     staging buffer and bumps the tile's counter ON THE OWNER rank (tile t is
//     owned by rank t % P) with a release-scoped remote red;
//   * four communication warps per CTA wait for owned tiles to be complete on
//     all ranks, pull the sum through the NVSwitch (multimem.ld_reduce), apply
//     W -= lr/P * sum in fp32 and multicast the new bf16 weights into every
//     rank's copy (multimem.st) - reduce-scatter + update + all-gather, hidden
//     under the GEMM of later tiles;
//   * a multicast "done" counter makes kernel completion imply that every
//     rank's weights are final.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <mutex>

#include "kernels.h"
#include "tcgen05_ptx.cuh"
#include "tma_host.h"
#include "vec_ops.cuh"

namespace m4t {

namespace {

constexpr int BMC = 128;        // rows of G (columns of dY) per CTA
constexpr int BM2 = 2 * BMC;    // cluster tile rows (n)
constexpr int BN = 256;         // cluster tile cols (k)
constexpr int BNH = BN / 2;     // X columns staged per CTA
constexpr int BK = 64;          // batch rows per pipeline stage
constexpr int UMMA_K = 16;
constexpr int kChunk = 64;      // elements per 128-byte swizzle span
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr int kBoxBytes = kChunk * BK * 2;          // 8 KiB: one TMA box
constexpr int kABytes = (BMC / kChunk) * kBoxBytes;  // 16 KiB
constexpr int kBBytes = (BNH / kChunk) * kBoxBytes;  // 16 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kWarps = 6;
constexpr int kCommWarps = 8;   // owner role: in-flight round trips scale with warps x rows (registers hold the data)
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
// peer-store mode: every epilogue warp stages {32 rows x 64 columns} bf16 boxes (4 KiB, 128-byte swizzled rows)
// in shared memory, double buffered, and hands them to the TMA store engine
constexpr int kEpiBoxBytes = 32 * 128;
constexpr int kEpiSmemBytes = 4 * 2 * kEpiBoxBytes;  // 4 epilogue warps x 2 buffers = 32 KiB
// layout: [TMA ring][barriers, 1 KiB][epilogue boxes, 1024-byte aligned swizzle atoms] (+1 KiB alignment slack)
constexpr int kSmemBytesFused = kStages * kStageBytes + 1024 + kEpiSmemBytes + 1024;
constexpr int kSignalsPerUnit = 8;  // 4 epilogue warps x 2 CTAs arrive on the tile counter
constexpr int kMaxUnicastRanks = 3;  // peer-load mode keeps ranks x rows x partial buffers requests in registers

struct WgradArgs {
  void* out;            // plain: G [N, K]; fused: this rank's staging buffer(s)
  int64_t out_split_stride;  // bytes between the ksplit partial buffers
  int Mb, N, K;
  int ldo;              // leading dimension of out (elements)
  int ksplit;
  const float* gscale;  // optional device scalar multiplied into every output element (the upstream
                        // gradient of the loss: autograd hands it over as a tensor, never through the host)
  float axpy;           // plain mode with axpy != 0: out <- out + axpy * gscale * G  (single-rank SGD epilogue)
};

struct WgradComm {
  SyncCtx sync;
  char* heap[kMaxGpuPeers];
  char* mc_heap;
  int64_t stage_off;     // partial-gradient staging (same offset on every rank)
  int64_t stage_stride;  // bytes between the ksplit partial buffers
  int64_t src_stride;    // peer-store mode: bytes between the per-source-rank copies of the staging area in the
                         // OWNER's heap (the epilogue pushes its partial tile there; the owner only reads locally)
  int64_t w_off;         // bf16 weights [N, K] (contiguous) inside every rank's heap
  int64_t cnt_off;       // u32 tile counters [tiles]
  int64_t done_off;      // u32 completion counter
  uint32_t tile_target;  // counters are monotonic: targets of THIS call when epoch_off < 0 ...
  uint32_t done_target;
  int64_t epoch_off;     // ... otherwise per-call increments; the call index is read from this local
                         // device word (and advanced by the kernel), which makes the launch graph-capturable
  float scale;           // W += scale * sum_ranks G   (scale = -lr / P)
  // optional: all-reduce the UPDATED weights for the next forward while the GEMM is still running
  int prefetch;          // 1: W_avg[tile] = avg_scale * sum_ranks W[tile] after the update of the tile
  int64_t wavg_off;      // bf16 [N, K] buffer receiving the averaged weights on every rank
  float avg_scale;       // 1 / P
  int unicast;           // 1: peer loads / peer stores instead of multimem (no NVLS, or too few ranks for it to pay)
  int debug;             // timing experiments (M4T_WGRAD_DEBUG): 1 no comm data movement, 2 no GEMM,
                         // 4 local loads instead of multimem.ld_reduce, 8 local stores instead of multimem.st,
                         // 16 no tile signals and no owner work, 32 no completion barrier
};

// peer-store mode: tensor maps of this rank's staging area inside every OWNER's heap
// ([ksplit * N rows, K columns] bf16, boxes of {64 columns, 32 rows}, 128-byte swizzle)
struct PushMaps {
  CUtensorMap owner[kMaxUnicastRanks];
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

// MN-major operand tile stored by TMA with SWIZZLE_128B as [chunk][BK rows][128 B].
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>(kBoxBytes >> 4) << 16;  // LBO: next 64-element chunk along M/N
  d |= static_cast<uint64_t>(1024u >> 4) << 32;      // SBO: next 8 rows along K
  d |= static_cast<uint64_t>(1u) << 46;              // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2u) << 61;              // SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc_bf16_f32_mn(uint32_t umma_m, uint32_t umma_n) {
  return tc::make_idesc_bf16_f32(umma_m, umma_n) | (1u << 15) | (1u << 16);  // A and B MN-major
}

// Work-unit schedule shared by every role (and by the host for the counter targets): tiles
// [0, whole) run over the full batch, tiles [whole, num_tiles) are split in `split` batch slices.
struct UnitSched {
  int whole, split, num_units;
};
__host__ __device__ inline UnitSched make_sched(int num_tiles, int num_clusters, int ksplit) {
  UnitSched s;
  s.split = ksplit;
  s.whole = ksplit > 1 ? (num_tiles / num_clusters) * num_clusters : num_tiles;
  s.num_units = s.whole + (num_tiles - s.whole) * ksplit;
  return s;
}
// unit u -> tile t, slice h of `parts` slices
__host__ __device__ inline void unit_to_tile(const UnitSched& s, int u, int& t, int& h, int& parts) {
  if (u < s.whole) {
    t = u;
    h = 0;
    parts = 1;
  } else {
    const int j = u - s.whole;
    t = s.whole + j / s.split;
    h = j - (j / s.split) * s.split;
    parts = s.split;
  }
}

__device__ __forceinline__ void bounded_wait_ge(const uint32_t* flag, uint32_t target, const SyncCtx& c) {
  if (static_cast<int32_t>(ld_acquire_sys_u32(flag) - target) >= 0) return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned int spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys_u32(flag) - target) < 0) {
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > c.timeout_ns) {
      *reinterpret_cast<volatile int*>(c.err_flag) = kErrTimeout;
      __threadfence_system();
      __trap();
    }
  }
}

// Owner side of the fused mode; runs on kCommWarps warps of every CTA.
// MC: reduce through the NVSwitch (multimem.ld_reduce / multimem.st); otherwise NSRC peer loads in rank
// order and one store per peer (2-3 ranks, or no multicast mapping).
template <bool MC, int NSRC>
__device__ __forceinline__ void comm_reduce_update(const WgradComm& wc, int first_thread, int cluster_id,
                                                   int num_clusters, uint32_t cta, int num_tiles, int k_tiles, int K,
                                                   const UnitSched sched) {
  const SyncCtx& c = wc.sync;
  const int P = c.size, r = c.rank;
  const int ct = threadIdx.x - first_thread;
  const int lane = ct & 31;
  const int cw = ct >> 5;
  constexpr int kCommThreads = kCommWarps * 32;
  constexpr int kU = NSRC <= 2 ? 4 : 2;  // rows in flight per warp (x partial buffers x peers independent round trips)
  constexpr int kRowSplit = 2;    // CTA pairs sharing one owned tile: shortens the tail after the last GEMM wave
                                  // (4 was measured slower on 8 GPUs: 0.303 vs 0.266 ms with the read-back)
  constexpr int kRows = BMC / kRowSplit;  // rows of this CTA's half handled per work item
  const uint32_t* my_cnt = reinterpret_cast<const uint32_t*>(wc.heap[r] + wc.cnt_off);
  const int64_t row_bytes = static_cast<int64_t>(K) * 2;
  // device-resident call counter: every communication thread reads it before any CTA can pass the
  // completion barrier below, and only CTA 0 advances it after that barrier
  uint32_t* epoch_ptr = wc.epoch_off >= 0 ? reinterpret_cast<uint32_t*>(wc.heap[r] + wc.epoch_off) : nullptr;
  const uint32_t epoch = epoch_ptr ? *reinterpret_cast<volatile uint32_t*>(epoch_ptr) : 0u;
  const uint32_t calls = epoch_ptr ? epoch + 1u : 1u;
  const uint32_t done_target = epoch_ptr ? calls * wc.done_target : wc.done_target;
  const bool skip = (wc.debug & 1) != 0;
  // Work item w = (j-th owned tile, row slice s); item w is served by cluster w % num_clusters.
  // Owned tiles (t = r + j*P) finish in GEMM order, so consecutive items land on different CTA
  // pairs and their round trips overlap.
  const int owned = r < num_tiles ? (num_tiles - r + P - 1) / P : 0;
  for (int w_item = cluster_id; w_item < owned * kRowSplit && !(wc.debug & 16); w_item += num_clusters) {
    const int j = w_item / kRowSplit;
    const int slice = w_item - j * kRowSplit;
    const int t = r + j * P;
    const int parts = t < sched.whole ? 1 : sched.split;
    // every rank's `parts` partial tiles are complete: kSignalsPerUnit signals per unit and rank, per call
    const uint32_t tile_target = (epoch_ptr ? calls : 1u) * wc.tile_target * static_cast<uint32_t>(parts);
    if (lane == 0) bounded_wait_ge(my_cnt + t, tile_target, c);
    __syncwarp();
    const int n_blk = t / k_tiles;
    const int k_blk = t - n_blk * k_tiles;
    // this CTA's rows of the item: kRows rows x 512 bytes, one row per warp pass
    const int64_t tile_off = (static_cast<int64_t>(n_blk) * BM2 + static_cast<int64_t>(cta) * BMC + slice * kRows) * row_bytes +
                             static_cast<int64_t>(k_blk) * BN * 2 + lane * 16;
    for (int row0 = cw; row0 < kRows && !skip; row0 += kCommWarps * kU) {
      // all loads of the kU rows (x partial buffers x peers) are issued before the first one is consumed
      constexpr int kSrc = NSRC;
      Vec16 s[kU][kSrc][2], w[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int row = row0 + u * kCommWarps;
        if (row < kRows) {
          const int64_t off = tile_off + row * row_bytes;
          if (MC) {
            if (wc.debug & 4) {
              s[u][0][0] = ld_vec(wc.heap[r] + wc.stage_off + off);
              if (parts > 1) s[u][0][1] = ld_vec(wc.heap[r] + wc.stage_off + wc.stage_stride + off);
            } else {
              s[u][0][0] = multimem_ld_reduce_vec<NvlsKind::ADD_BF16>(wc.mc_heap + wc.stage_off + off);
              if (parts > 1) s[u][0][1] = multimem_ld_reduce_vec<NvlsKind::ADD_BF16>(wc.mc_heap + wc.stage_off + wc.stage_stride + off);
            }
          } else {
#pragma unroll
            for (int p = 0; p < kSrc; ++p) {
              if (p < P) {
                // pushed here by rank p's epilogue: local reads
                const char* src = wc.heap[r] + wc.stage_off + static_cast<int64_t>(p) * wc.src_stride + off;
                s[u][p][0] = ld_vec(src);
                if (parts > 1) s[u][p][1] = ld_vec(src + wc.stage_stride);
              }
            }
          }
          w[u] = ld_vec(wc.heap[r] + wc.w_off + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int row = row0 + u * kCommWarps;
        if (row < kRows) {
          const int64_t off = tile_off + row * row_bytes;
          float a[8], b[8], wv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = 0.f;
#pragma unroll
          for (int p = 0; p < kSrc; ++p) {  // fixed rank order: every run adds in the same order
            if (MC || p < P) {
              VecOf<DType::BF16>::unpack(s[u][p][0], b);
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] += b[e];
              if (parts > 1) {
                VecOf<DType::BF16>::unpack(s[u][p][1], b);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] += b[e];
              }
            }
          }
          VecOf<DType::BF16>::unpack(w[u], wv);
#pragma unroll
          for (int e = 0; e < 8; ++e) wv[e] = fmaf(wc.scale, a[e], wv[e]);
          const Vec16 o = VecOf<DType::BF16>::pack(wv);
          if (MC && !(wc.debug & 8)) {
            multimem_st_vec(wc.mc_heap + wc.w_off + off, o);
          } else if (MC) {
            st_vec(wc.heap[r] + wc.w_off + off, o);
          } else {
#pragma unroll
            for (int p = 0; p < kSrc; ++p)
              if (p < P) st_vec(wc.heap[p] + wc.w_off + off, o);
          }
        }
      }
    }
    if (wc.prefetch && !skip) {
      // Next step's forward needs Allreduce(W) / P.  The rows this lane just wrote into every rank's
      // copy are the final values of this step, so the parameter all-reduce of the NEXT step can run here,
      // under the GEMM of later tiles.  Every address read below was written above by THIS lane:
      //  * peer-store mode: plain stores followed by loads of the same addresses - program order to the
      //    same location from one thread is coherent without a fence (and a MEMBAR.SYS per lane and item,
      //    waiting for every posted remote write to be acknowledged, was the most expensive instruction of
      //    this role);
      //  * multicast mode: the write fans out inside the switch, so the fence stays.
      if (MC) __threadfence_system();
      for (int row0 = cw; row0 < kRows; row0 += kCommWarps * kU) {
        constexpr int kSrc = NSRC;
        Vec16 x[kU][kSrc];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int row = row0 + u * kCommWarps;
          if (row < kRows) {
            const int64_t off = tile_off + row * row_bytes;
            if (MC) {
              x[u][0] = multimem_ld_reduce_vec<NvlsKind::ADD_BF16>(wc.mc_heap + wc.w_off + off);
            } else {
#pragma unroll
              for (int p = 0; p < kSrc; ++p)
                if (p < P) x[u][p] = ld_vec_sys(wc.heap[p] + wc.w_off + off);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int row = row0 + u * kCommWarps;
          if (row < kRows) {
            const int64_t off = tile_off + row * row_bytes;
            float a[8], b[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = 0.f;
#pragma unroll
            for (int p = 0; p < kSrc; ++p) {
              if (MC || p < P) {
                VecOf<DType::BF16>::unpack(x[u][p], b);
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] += b[e];
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] *= wc.avg_scale;
            const Vec16 o = VecOf<DType::BF16>::pack(a);
            if (MC) {
              multimem_st_vec(wc.mc_heap + wc.wavg_off + off, o);
            } else {
#pragma unroll
              for (int p = 0; p < kSrc; ++p)
                if (p < P) st_vec(wc.heap[p] + wc.wavg_off + off, o);
            }
          }
        }
      }
    }
  }
  // completion: every CTA of every rank reports once; leaving the kernel means all weights are final
  asm volatile("bar.sync 1, %0;" ::"n"(kCommThreads));
  if (ct == 0 && !(wc.debug & 32)) {
    if (MC) {
      asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(wc.mc_heap + wc.done_off), "r"(1u) : "memory");
    } else {
      __threadfence_system();
      for (int p = 0; p < P; ++p)
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(wc.heap[p] + wc.done_off), "r"(1u) : "memory");
    }
    bounded_wait_ge(reinterpret_cast<const uint32_t*>(wc.heap[r] + wc.done_off), done_target, c);
    if (epoch_ptr && blockIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(epoch_ptr) = epoch + 1u;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kCommThreads));
}

template <bool FUSED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kWarps + (FUSED ? kCommWarps : 0)) * 32, 1)
wgrad_bf16_nt_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const WgradArgs g, const WgradComm wc, const __grid_constant__ PushMaps push) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = tc::cluster_ctarank();
  const bool leader = cta == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int n_tiles = g.N / BM2;  // tiles along the rows of G
  const int k_tiles = g.K / BN;   // tiles along the columns of G
  const int num_tiles = n_tiles * k_tiles;
  const UnitSched sched = make_sched(num_tiles, num_clusters, g.ksplit);
  const int num_units = sched.num_units;
  const int kb_total = g.Mb / BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bars->full[s], 1);
      tc::mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      tc::mbar_init(&bars->tmem_full[a], 1);
      tc::mbar_init(&bars->tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs (collected by the leader)
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
    if (lane == 0 && !(FUSED && (wc.debug & 2))) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        int t, h, parts;
        unit_to_tile(sched, u, t, h, parts);
        const int kb_per_unit = kb_total / parts;
        const int n_blk = t / k_tiles;
        const int k_blk = t - n_blk * k_tiles;
        const int n0 = n_blk * BM2 + static_cast<int>(cta) * BMC;  // column of dY
        const int k0 = k_blk * BN + static_cast<int>(cta) * BNH;   // column of X
        for (int kb = 0; kb < kb_per_unit; ++kb) {
          const int mb0 = (h * kb_per_unit + kb) * BK;
          tc::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) tc::mbar_arrive_expect_tx(&bars->full[stage], 2 * kStageBytes);
#pragma unroll
          for (int ch = 0; ch < BMC / kChunk; ++ch)
            tc::tma_load_2d_2sm(sa + ch * kBoxBytes, &tmap_a, &bars->full[stage], n0 + ch * kChunk, mb0);
#pragma unroll
          for (int ch = 0; ch < BNH / kChunk; ++ch)
            tc::tma_load_2d_2sm(sb + ch * kBoxBytes, &tmap_b, &bars->full[stage], k0 + ch * kChunk, mb0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =================
    if (leader && lane == 0 && !(FUSED && (wc.debug & 2))) {
      constexpr uint32_t idesc = make_idesc_bf16_f32_mn(BM2, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        const int kb_per_unit = u < sched.whole ? kb_total : kb_total / sched.split;
        for (int kb = 0; kb < kb_per_unit; ++kb) {
          tc::mbar_wait(&bars->full[stage], phase);
          tc::tcgen05_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t adesc = make_smem_desc_mn_sw128(sa);
          const uint64_t bdesc = make_smem_desc_mn_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // 16 batch rows = two 8-row swizzle atoms = 2048 bytes further into every chunk
            const uint64_t koff = static_cast<uint64_t>((k * UMMA_K * 128) >> 4);
            tc::umma_bf16_ss_2sm(tmem_d, adesc + koff, bdesc + koff, idesc, (kb | k) ? 1u : 0u);
          }
          tc::umma_commit_2sm(&bars->empty[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc::umma_commit_2sm(&bars->tmem_full[acc]);
      }
    }
  } else if (warp < kWarps) {
    // ===================== epilogue (both CTAs) =========================
    const int q = warp & 3;
    int it = 0;
    for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
      int t, h, parts;
      unit_to_tile(sched, u, t, h, parts);
      const int n_blk = t / k_tiles;
      const int k_blk = t - n_blk * k_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (FUSED && (wc.debug & 2)) {  // timing experiment: no GEMM, only the tile signals
        __syncwarp();
        if (lane == 0) {
          uint32_t* cnt = reinterpret_cast<uint32_t*>(wc.heap[t % wc.sync.size] + wc.cnt_off) + t;
          asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
        }
        continue;
      }
      tc::mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc::tcgen05_fence_after();
      const int row = n_blk * BM2 + static_cast<int>(cta) * BMC + q * 32 + lane;
      // plain / multicast mode: this rank's own staging buffer.  Peer-store mode: the OWNER's staging area for
      // this source rank - posted writes over NVLink, spread over the whole GEMM, so the owner never has to
      // pull (no read round trips on the reduction path).
      uint16_t* orow = reinterpret_cast<uint16_t*>(static_cast<char*>(g.out) + h * g.out_split_stride) +
                       static_cast<int64_t>(row) * g.ldo + k_blk * BN;
      const float gs = (g.gscale ? __ldg(g.gscale) : 1.0f) * (g.axpy != 0.0f ? g.axpy : 1.0f);
      const bool axpy = !FUSED && g.axpy != 0.0f;
      if (FUSED && wc.unicast) {
        // Peer-store mode: push this warp's 32 x 256 piece of the partial tile into the OWNER's staging area for
        // this source rank.  TMEM -> registers -> swizzled shared memory -> TMA bulk tensor store: whole 128-byte
        // rows cross NVLink as posted writes issued by the copy engine of the SM, the warp itself never waits
        // for them (only for its shared-memory buffer to be read), and the owner later reads everything locally.
        uint8_t* ebuf = smem + kStages * kStageBytes + 1024 + q * (2 * kEpiBoxBytes);
        const CUtensorMap* omap = &push.owner[t % wc.sync.size];
        const int row0 = h * g.N + n_blk * BM2 + static_cast<int>(cta) * BMC + q * 32;  // first row of the box
#pragma unroll 1
        for (int c2 = 0; c2 < BN / 64; ++c2) {
          uint8_t* buf = ebuf + (c2 & 1) * kEpiBoxBytes;
          // the store that last used this buffer (two boxes ago) must have finished reading it
          if (lane == 0) tc::tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN + c2 * 64 + half * 32);
            tc::tmem_ld_32x32b_x32(taddr, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              Vec16 o;
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o.w[e] = pack2(gs * __uint_as_float(r[v * 8 + 2 * e]), gs * __uint_as_float(r[v * 8 + 2 * e + 1]));
              // row `lane`, 16-byte chunk j of the 128-byte row, stored at chunk j ^ (row & 7) (SWIZZLE_128B)
              const int j = half * 4 + v;
              *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(o.w[0], o.w[1], o.w[2], o.w[3]);
            }
          }
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tc::tma_store_2d(omap, buf, k_blk * BN + c2 * 64, row0);
            tc::tma_store_commit();
          }
        }
        tc::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) tc::mbar_arrive(&bars->tmem_empty[acc]);  // the accumulator is drained: the MMA warp may reuse it
          else tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bars->tmem_empty[acc]), 0));
          if (!(wc.debug & 16)) {
            tc::tma_store_wait<0>();  // all four boxes have been written (performed), then publish them to the owner
            uint32_t* cnt = reinterpret_cast<uint32_t*>(wc.heap[t % wc.sync.size] + wc.cnt_off) + t;
            asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        Vec16 prev[4];
        if (axpy) {  // the four 16-byte pieces of this lane's row are requested before the TMEM load
#pragma unroll
          for (int v = 0; v < 4; ++v) prev[v] = ld_vec(orow + c * 32 + v * 8);
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + c * 32);
        tc::tmem_ld_32x32b_x32(taddr, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          Vec16 o;
          if (axpy) {
            float wv[8];
            VecOf<DType::BF16>::unpack(prev[v], wv);
#pragma unroll
            for (int e = 0; e < 8; ++e) wv[e] = fmaf(gs, __uint_as_float(r[v * 8 + e]), wv[e]);
            o = VecOf<DType::BF16>::pack(wv);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o.w[e] = pack2(gs * __uint_as_float(r[v * 8 + 2 * e]), gs * __uint_as_float(r[v * 8 + 2 * e + 1]));
          }
          st_vec(orow + c * 32 + v * 8, o);
        }
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) tc::mbar_arrive(&bars->tmem_empty[acc]);
        else tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bars->tmem_empty[acc]), 0));
        if (FUSED && !(wc.debug & 16)) {
          // this warp's 32 rows of the partial tile are in local HBM: tell the owner.  The release
          // (a system-scope fence + the add) is cumulative over the warp's stores through __syncwarp.
          uint32_t* cnt = reinterpret_cast<uint32_t*>(wc.heap[t % wc.sync.size] + wc.cnt_off) + t;
          asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
        }
      }
    }
  } else if (FUSED) {
    // ===================== reduce + update (both CTAs) ==================
    if (!wc.unicast) comm_reduce_update<true, 1>(wc, kWarps * 32, cluster_id, num_clusters, cta, num_tiles, k_tiles, g.K, sched);
    else if (wc.sync.size <= 2) comm_reduce_update<false, 2>(wc, kWarps * 32, cluster_id, num_clusters, cta, num_tiles, k_tiles, g.K, sched);
    else comm_reduce_update<false, kMaxUnicastRanks>(wc, kWarps * 32, cluster_id, num_clusters, cta, num_tiles, k_tiles, g.K, sched);
  }

  tc::tcgen05_fence_before();
  tc::cluster_sync();
  if (warp == 1) tc::tmem_dealloc_2sm<kTmemCols>(tmem_base);
}

// Row-major [rows, cols] bf16 tensor, boxes of {64 columns, BK rows}.
CUtensorMap make_tmap_mn(const void* base, int64_t rows, int64_t cols, int64_t ld) {
  return make_tmap_bf16_sw128(base, rows, cols, ld, kChunk, BK);
}

template <bool FUSED> void configure_w() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaError_t e = cudaFuncSetAttribute(wgrad_bf16_nt_2cta_kernel<FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FUSED ? kSmemBytesFused : kSmemBytes);
    M4T_CHECK(e == cudaSuccess, "cudaFuncSetAttribute(smem) failed: " << cudaGetErrorString(e));
  });
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// timing experiments only (scripts/wgrad_diag.py): initial value from M4T_WGRAD_DEBUG, changeable at run time
static std::atomic<int> g_wgrad_debug{static_cast<int>(env_i64("M4T_WGRAD_DEBUG", 0))};
void set_wgrad_debug(int mask) { g_wgrad_debug.store(mask, std::memory_order_relaxed); }

bool wgrad_bf16_supported(int64_t Mb, int64_t N, int64_t K, const void* dy, const void* x, const void* g, int64_t ldy,
                          int64_t ldx, int64_t ldg) {
  return Mb > 0 && N > 0 && K > 0 && N % BM2 == 0 && K % BN == 0 && Mb % BK == 0 && aligned16(dy) && aligned16(x) &&
         aligned16(g) && ldy % 8 == 0 && ldx % 8 == 0 && ldg % 8 == 0 && ldy >= N && ldx >= K && ldg >= K &&
         Mb < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31);
}

void launch_wgrad_bf16(const void* dy, const void* x, void* gout, int64_t Mb, int64_t N, int64_t K, int64_t ldy,
                       int64_t ldx, int64_t ldg, int sm_count, cudaStream_t stream, const float* gscale, float axpy) {
  M4T_CHECK(wgrad_bf16_supported(Mb, N, K, dy, x, gout, ldy, ldx, ldg),
            "unsupported wgrad shape/alignment for the tcgen05 path (N % 256, K % 256, batch % 64)");
  const CUtensorMap ta = make_tmap_mn(dy, Mb, N, ldy);
  const CUtensorMap tb = make_tmap_mn(x, Mb, K, ldx);
  WgradArgs g{};
  g.out = gout;
  g.out_split_stride = 0;
  g.Mb = static_cast<int>(Mb);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldo = static_cast<int>(ldg);
  g.ksplit = 1;
  g.gscale = gscale;
  g.axpy = axpy;
  const int tiles = static_cast<int>((N / BM2) * (K / BN));
  const int clusters = std::max(1, std::min(tiles, sm_count / 2));
  configure_w<false>();
  wgrad_bf16_nt_2cta_kernel<false><<<2 * clusters, kWarps * 32, kSmemBytes, stream>>>(ta, tb, g, WgradComm{}, PushMaps{});
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, "wgrad_bf16 launch failed: " << cudaGetErrorString(e));
  note_kernel_launch(axpy != 0.0f ? "wgrad_2cta_sgd_epilogue" : "wgrad_2cta");
}

int64_t fused_wgrad_tiles(int64_t N, int64_t K) { return (N / BM2) * (K / BN); }
int fused_wgrad_signals_per_unit() { return kSignalsPerUnit; }
int fused_wgrad_max_unicast_ranks() { return kMaxUnicastRanks; }

void launch_fused_wgrad_update(const DeviceComm& dc, const void* dy, const void* x, int64_t Mb, int64_t N, int64_t K,
                               int64_t ldy, int64_t ldx, int64_t w_off, int64_t stage_off, int64_t stage_stride,
                               int64_t cnt_off, int64_t done_off, int ksplit, uint32_t tile_target,
                               uint32_t done_target, float scale, int64_t wavg_off, cudaStream_t stream,
                               int64_t epoch_off, const float* gscale, bool use_multicast, int64_t src_stride) {
  M4T_CHECK(use_multicast ? dc.mc_heap != nullptr : dc.sync.size <= kMaxUnicastRanks,
            "the fused wgrad->Allreduce->SGD kernel needs the NVLS multicast mapping beyond " << kMaxUnicastRanks << " ranks");
  M4T_CHECK(ksplit == 1 || ksplit == 2, "ksplit must be 1 or 2");
  M4T_CHECK((Mb / BK) % ksplit == 0, "batch / 64 must be divisible by ksplit");
  char* stage = dc.heap[dc.sync.rank] + stage_off;
  M4T_CHECK(wgrad_bf16_supported(Mb, N, K, dy, x, stage, ldy, ldx, K), "unsupported wgrad shape/alignment for the fused path");
  const CUtensorMap ta = make_tmap_mn(dy, Mb, N, ldy);
  const CUtensorMap tb = make_tmap_mn(x, Mb, K, ldx);
  WgradArgs g{};
  g.out = stage;
  g.out_split_stride = stage_stride;
  g.Mb = static_cast<int>(Mb);
  g.N = static_cast<int>(N);
  g.K = static_cast<int>(K);
  g.ldo = static_cast<int>(K);
  g.ksplit = ksplit;
  g.gscale = gscale;
  g.axpy = 0.0f;
  WgradComm wc{};
  wc.sync = dc.sync;
  for (int p = 0; p < dc.sync.size; ++p) wc.heap[p] = dc.heap[p];
  wc.mc_heap = dc.mc_heap;
  wc.stage_off = stage_off;
  wc.stage_stride = stage_stride;
  wc.src_stride = src_stride;
  wc.w_off = w_off;
  wc.cnt_off = cnt_off;
  wc.done_off = done_off;
  wc.tile_target = tile_target;
  wc.done_target = done_target;
  wc.epoch_off = epoch_off;
  wc.scale = scale;
  wc.prefetch = wavg_off >= 0 ? 1 : 0;
  wc.wavg_off = wavg_off >= 0 ? wavg_off : 0;
  wc.avg_scale = 1.0f / static_cast<float>(dc.sync.size);
  wc.unicast = use_multicast ? 0 : 1;
  wc.debug = g_wgrad_debug.load(std::memory_order_relaxed);
  const int grid = fused_gemm_grid(dc);  // identical on every rank, whole CTA pairs
  configure_w<true>();
  PushMaps push{};
  if (!use_multicast) {
    M4T_CHECK(stage_stride == N * K * 2, "peer-store mode expects densely packed partial buffers");
    for (int p = 0; p < dc.sync.size; ++p)  // my staging area inside rank p's heap: rows [h * N + n] of K columns
      push.owner[p] = make_tmap_bf16_sw128(dc.heap[p] + stage_off + static_cast<int64_t>(dc.sync.rank) * src_stride,
                                           static_cast<int64_t>(ksplit) * N, K, K, 64, 32);
  }
  wgrad_bf16_nt_2cta_kernel<true><<<grid, (kWarps + kCommWarps) * 32, kSmemBytesFused, stream>>>(ta, tb, g, wc, push);
  cudaError_t e = cudaGetLastError();
  M4T_CHECK(e == cudaSuccess, "fused_wgrad_update launch failed: " << cudaGetErrorString(e));
  note_kernel_launch(wc.prefetch ? "fused_wgrad_reduce_scatter_sgd_prefetch" : "fused_wgrad_reduce_scatter_sgd");
}

}  // namespace m4t
