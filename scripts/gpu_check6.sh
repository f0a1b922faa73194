#!/bin/bash
# usage: gpu_check6.sh <np>
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1
echo "=== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit=$?"; grep -v "Deprecat\|warnings.warn\|^$" $OUT/pytest_gpu.log | tail -40 | cut -c1-600
echo "=== sweep np=$NP"; timeout 900 python -m mpi4torch_b200.launch -np $NP benchmarks/allreduce_sweep.py --raw --full --no-staged --out $OUT/sweep_np$NP.json 2>&1 | grep -v "^W0" | tail -14
echo "=== bench N=$NP"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 10 --warmup 3 > $OUT/bench_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n$NP.log | tail -6
echo "=== bench N=$NP unfused"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 10 --warmup 3 --unfused --no-extras > $OUT/bench_n${NP}_unfused.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n${NP}_unfused.log | tail -4
