#!/bin/bash
# usage: gpu_check8.sh <np> [quick]
set -u
NP=${1:-8}
STAGED=""; if [ "$NP" -gt 2 ]; then STAGED="--no-staged"; fi
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1
echo "=== spmd suite np=$NP"; M4T_TEST_DEVICE=cuda timeout 900 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py > $OUT/spmd_np$NP.log 2>&1; echo "exit=$?"; grep -v "^W0" $OUT/spmd_np$NP.log | tail -8 | cut -c1-400
echo "=== fuzz np=$NP: device digests must equal the host digests"
for seed in 1 2 3; do
  D=$(M4T_TEST_DEVICE=cuda timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/fuzz_ops.py $seed 150 2>&1 | grep "^FUZZ")
  H=$(M4T_CUDA=0 timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/fuzz_ops.py $seed 150 2>&1 | grep "^FUZZ")
  if [ -n "$D" ] && [ "$D" = "$H" ]; then echo "seed $seed: identical"; else echo "seed $seed: DIFFERENT"; echo " device: $D"; echo " host:   $H"; fi
done
echo "=== bench N=$NP fused"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 20 --warmup 5 > $OUT/bench_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n$NP.log | tail -2 | cut -c1-2000
echo "=== bench N=$NP unfused-forward"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 20 --warmup 5 --unfused --no-extras > $OUT/bench_n${NP}_unfused.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n${NP}_unfused.log | tail -2 | cut -c1-700
if [ "${2:-}" != "quick" ]; then
echo "=== tune np=$NP"; TUNE_SIZES=33554432,268435456 TUNE_BLOCKS=64,148 TUNE_CHUNKS_KB=32768,1048576 timeout 600 python -m mpi4torch_b200.launch -np $NP benchmarks/tune_allreduce.py > $OUT/tune_np$NP.log 2>&1; grep -v "^W0" $OUT/tune_np$NP.log | tail -30
echo "=== sweep np=$NP"; timeout 900 python -m mpi4torch_b200.launch -np $NP benchmarks/allreduce_sweep.py --raw --full $STAGED --out $OUT/sweep_np$NP.json 2>&1 | grep -v "^W0" | tail -14 | cut -c1-600
echo "=== collectives np=$NP"; timeout 600 python -m mpi4torch_b200.launch -np $NP benchmarks/collectives_bench.py --max-mb 64 --out $OUT/collectives_np$NP.json 2>&1 | grep -v "^W0" | tail -8 | cut -c1-700
echo "=== ring overlap np=$NP"; timeout 300 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/ring_np$NP.json 2>&1 | grep -v "^W0" | tail -3
fi
