#!/bin/bash
# usage: gpu_check7.sh <np>
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1
for n in 1 $NP; do
echo "=== spmd_gpu np=$n"; M4T_TEST_DEVICE=cuda timeout 600 python -m mpi4torch_b200.launch -np $n tests/spmd/run_all.py "spmd_gpu.py" > $OUT/spmd_gpu_np$n.log 2>&1; echo "exit=$?"; grep -v "^W0" $OUT/spmd_gpu_np$n.log | tail -12 | cut -c1-400
done
echo "=== bench N=1"; timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > $OUT/bench_n1.log 2>&1; echo "exit=$?"; grep -v "^W0" $OUT/bench_n1.log | tail -2 | cut -c1-1500
echo "=== bench N=$NP fused"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/bench_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n$NP.log | tail -3 | cut -c1-1500
echo "=== bench N=$NP unfused-forward (fast path, separate allreduce)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 20 --warmup 5 --unfused --no-extras > $OUT/bench_n${NP}_unfused.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n${NP}_unfused.log | tail -3 | cut -c1-1500
