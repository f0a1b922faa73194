#!/bin/bash
# np=2: p2p copy-engine path + fused backward with epilogue push
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1 M4T_TEST_EXPERIMENTAL=1
echo "=== gpu + nonblocking + stress suites np=$NP"
M4T_TEST_DEVICE=cuda timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py "spmd_[gns]*.py" > $OUT/h_spmd_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/h_spmd_np$NP.log | tail -6 | cut -c1-400
echo "=== wgrad diag np=$NP"
timeout 150 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep "^{" | tee $OUT/h_wgrad_diag_np$NP.jsonl
echo "=== bench ours np=$NP"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras --no-checks > $OUT/h_bench_ours_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/h_bench_ours_n$NP.log | tail -1 | cut -c1-400
echo "=== ring np=$NP: kernel path / copy-engine path (16 MiB ring) / copy-engine path (64 MiB ring)"
M4T_P2P_CE_MIN_KB=-1 timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/h_ring_kernel_np$NP.json 2>&1 | grep "^{" | cut -c1-600
timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/h_ring_ce16_np$NP.json 2>&1 | grep "^{" | cut -c1-600
M4T_P2P_SLOTS=64 timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/h_ring_ce64_np$NP.json 2>&1 | grep "^{" | cut -c1-600
