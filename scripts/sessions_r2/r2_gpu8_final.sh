#!/bin/bash
# Final multi-GPU confirmation of the defaults: suites, flagship bench with the self-check block, ring with default rings.
set -u
NP=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1 M4T_TEST_EXPERIMENTAL=1
echo "=== suites np=$NP"
M4T_TEST_DEVICE=cuda timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py "spmd_[gcn]*.py" > $OUT/z_spmd_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/z_spmd_np$NP.log | tail -5 | cut -c1-400
echo "=== bench ours np=$NP"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 > $OUT/z_bench_ours_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/z_bench_ours_n$NP.log | tail -2 | cut -c1-3800
echo "=== wgrad diag np=$NP"
timeout 150 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep "^{" | tee $OUT/z_wgrad_diag_np$NP.jsonl
echo "=== p2p ring np=$NP (defaults)"
timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/z_ring_np$NP.json 2>&1 | grep "^{" | cut -c1-600
