#!/bin/bash
# Round-2 session D (np GPUs): fused-backward timing experiments, new autograd-path tests, bench both arms.
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1 M4T_TEST_EXPERIMENTAL=1
echo "=== wgrad diag np=$NP"
timeout 150 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep "^{" | tee $OUT/d_wgrad_diag_np$NP.jsonl
echo "=== spmd_gpu suite np=$NP"
M4T_TEST_DEVICE=cuda timeout 240 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_gpu.py > $OUT/d_spmd_gpu_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/d_spmd_gpu_np$NP.log | tail -25 | cut -c1-400
echo "=== step breakdown np=$NP"
timeout 120 python -m mpi4torch_b200.launch -np $NP scripts/step_breakdown.py 2>&1 | grep "^{" | tee $OUT/d_step_breakdown_np$NP.jsonl
echo "=== bench ours np=$NP"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 > $OUT/d_bench_ours_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/d_bench_ours_n$NP.log | tail -3 | cut -c1-3000
echo "=== bench reference np=$NP"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $NP --steps 10 --warmup 3 > $OUT/d_bench_ref_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/d_bench_ref_n$NP.log | tail -3 | cut -c1-3000
