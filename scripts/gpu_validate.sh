#!/bin/bash
set -u
NP=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1
echo "=== spmd_gpu np=$NP"; M4T_TEST_DEVICE=cuda timeout 400 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py "spmd_gpu.py" > $OUT/spmd_gpu_np$NP.log 2>&1; echo "exit=$?"; grep -v "^W0" $OUT/spmd_gpu_np$NP.log | tail -5 | cut -c1-300
echo "=== step breakdown np=$NP"; timeout 200 python -m mpi4torch_b200.launch -np $NP scripts/step_breakdown.py 2>&1 | grep -v "^W0" | tail -6
echo "=== bench N=$NP"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/bench_n$NP.log 2>&1; echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n$NP.log | tail -1 | cut -c1-1500
