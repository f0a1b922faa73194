#!/bin/bash
# Round-2 session B (np GPUs): bring-up of the fused backward, CUDA Split, zero-copy, graph capture.
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1 M4T_TEST_EXPERIMENTAL=1
echo "=== fused wgrad suite np=$NP"
M4T_TEST_DEVICE=cuda M4T_FUSED_WGRAD=2 M4T_ZERO_COPY_IN=1 timeout 500 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_gpu.py > $OUT/b_fwgrad_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/b_fwgrad_np$NP.log | tail -25 | cut -c1-300
echo "=== sub-communicators on CUDA np=$NP"
M4T_TEST_DEVICE=cuda M4T_TEST_SPLIT_CUDA=1 timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_split.py > $OUT/b_split_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/b_split_np$NP.log | tail -8 | cut -c1-300
echo "=== bench: default / fused backward / fused backward + prefetch"
for v in "default:" "fwgrad:M4T_FUSED_WGRAD=2" "prefetch:M4T_FUSED_WGRAD=2 M4T_WAVG_PREFETCH=1"; do
  name=${v%%:*}; kv=${v#*:}
  env $kv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/b_bench_${name}_n$NP.log 2>&1
  echo "$name exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/b_bench_${name}_n$NP.log | tail -1 | cut -c1-700
done
