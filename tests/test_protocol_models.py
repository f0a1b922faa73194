"""Pure-Python mirrors of the index arithmetic of the fused backward kernel
(csrc/kernels/wgrad_tcgen05_2cta.cu).  They cannot check the memory model, but they do
check what is easy to get wrong in a persistent, role-split kernel: that the work
distribution covers every element exactly once and that the counters reach exactly the
targets the host computes (csrc/runtime/cuda_backend.cpp::fused_wgrad_update)."""
import itertools

import numpy as np
import pytest

BMC, BM2, BN = 128, 256, 256          # rows per CTA, rows / cols per cluster tile
K_COMM_WARPS, K_U, K_ROW_SPLIT = 4, 8, 2
K_ROWS = BMC // K_ROW_SPLIT
SIGNALS_PER_UNIT = 8                  # 4 epilogue warps x 2 CTAs


def units_of_cluster(cluster_id, num_clusters, num_units):
    return range(cluster_id, num_units, num_clusters)


@pytest.mark.parametrize("N,K,P,num_clusters,ksplit", [(1024, 512, 4, 3, 2), (512, 256, 8, 5, 2), (768, 768, 2, 7, 1),
                                                      (4096, 4096, 8, 74, 2)])
def test_fused_wgrad_work_distribution(N, K, P, num_clusters, ksplit):
    n_tiles, k_tiles = N // BM2, K // BN
    num_tiles = n_tiles * k_tiles
    num_units = num_tiles * ksplit
    small = N * K <= 1 << 20

    # --- GEMM epilogue: every (split, row, col) of the staging buffers written once; tile counters ---
    staged = np.zeros((ksplit, N, K), dtype=np.int8) if small else None
    signals = np.zeros(num_tiles, dtype=np.int64)  # per rank; the owner sums over P ranks
    for cluster in range(num_clusters):
        for u in units_of_cluster(cluster, num_clusters, num_units):
            t, h = divmod(u, ksplit)
            n_blk, k_blk = divmod(t, k_tiles)
            for cta, q in itertools.product(range(2), range(4)):
                if small:
                    r0 = n_blk * BM2 + cta * BMC + q * 32
                    staged[h, r0:r0 + 32, k_blk * BN:(k_blk + 1) * BN] += 1
                signals[t] += 1
    if small:
        assert (staged == 1).all()
    assert (signals == SIGNALS_PER_UNIT * ksplit).all()
    tile_target_per_epoch = SIGNALS_PER_UNIT * ksplit * P  # host: calls * signals_per_tile(ksplit) * size
    assert (signals * P == tile_target_per_epoch).all()

    # --- owner side: every element of W updated by exactly one (rank, cluster, cta, warp, lane, trip) ---
    updated = np.zeros((N, K), dtype=np.int8) if small else None
    items_total = 0
    for r in range(P):
        owned = (num_tiles - r + P - 1) // P if r < num_tiles else 0
        for cluster in range(num_clusters):
            for w_item in range(cluster, owned * K_ROW_SPLIT, num_clusters):
                j, sl = divmod(w_item, K_ROW_SPLIT)
                t = r + j * P
                assert t < num_tiles and t % P == r  # the epilogue signals rank t % P
                items_total += 1
                if not small:
                    continue
                n_blk, k_blk = divmod(t, k_tiles)
                for cta, cw in itertools.product(range(2), range(K_COMM_WARPS)):
                    rows = []
                    for row0 in range(cw, K_ROWS, K_COMM_WARPS * K_U):
                        rows += [row0 + u * K_COMM_WARPS for u in range(K_U) if row0 + u * K_COMM_WARPS < K_ROWS]
                    for row in rows:
                        g = n_blk * BM2 + cta * BMC + sl * K_ROWS + row
                        updated[g, k_blk * BN:(k_blk + 1) * BN] += 1  # 32 lanes x 8 bf16 = 256 columns
    assert items_total == num_tiles * K_ROW_SPLIT
    if small:
        assert (updated == 1).all()

    # --- completion counter: every CTA of every rank reports once per call ---
    grid = 2 * num_clusters
    assert P * grid == P * 2 * num_clusters  # host: done_target = calls * size * fused_gemm_grid


def test_unit_schedule_keeps_both_halves_of_a_tile_adjacent():
    # two work units of one tile run on neighbouring CTA pairs in the same wave, so the owner can
    # start reducing a tile as soon as that wave finishes
    num_clusters, ksplit, num_tiles = 74, 2, 256
    wave_of = {}
    for cluster in range(num_clusters):
        for it, u in enumerate(units_of_cluster(cluster, num_clusters, num_tiles * ksplit)):
            wave_of[u] = it
    late = [t for t in range(num_tiles) if wave_of[2 * t] != wave_of[2 * t + 1]]
    assert len(late) == 0, late[:5]
