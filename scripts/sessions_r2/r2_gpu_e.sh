#!/bin/bash
# Round-2 session E (np GPUs): fused backward after the tail-split / unicast rewrite.
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1 M4T_TEST_EXPERIMENTAL=1
echo "=== fused wgrad numerics np=$NP"
M4T_TEST_DEVICE=cuda timeout 200 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_gpu.py > $OUT/e_spmd_gpu_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/e_spmd_gpu_np$NP.log | tail -6 | cut -c1-400
echo "=== wgrad diag np=$NP"
timeout 150 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep "^{" | tee $OUT/e_wgrad_diag_np$NP.jsonl
echo "=== step breakdown np=$NP"
timeout 120 python -m mpi4torch_b200.launch -np $NP scripts/step_breakdown.py 2>&1 | grep "^{" | tee $OUT/e_step_breakdown_np$NP.jsonl
echo "=== bench ours np=$NP"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/e_bench_ours_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/e_bench_ours_n$NP.log | tail -2 | cut -c1-2500
