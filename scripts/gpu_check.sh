#!/bin/bash
# First-contact GPU check: bring-up diagnostics + SPMD suites + a short bench.
# Usage (through gpurun): bash scripts/gpu_check.sh <max_np>
set -u
NP=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
export M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=15
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "=== build check"; timeout 900 python __graft_entry__.py 2>&1 | tail -3
echo "=== pytest gpu-marked (single process part)"
timeout 600 python -m pytest tests/test_gpu_spmd.py -x -q -k "native" 2>&1 | tail -5
for n in 1 $NP; do
  echo "=== SPMD suite np=$n (cuda)"
  M4T_TEST_DEVICE=cuda M4T_DEBUG_SEGV=1 timeout 900 python -m mpi4torch_b200.launch -np $n --timeout 800 tests/spmd/run_all.py > $OUT/spmd_np$n.log 2>&1
  echo "exit=$?"; grep -v "^W0" $OUT/spmd_np$n.log | tail -40
done
echo "=== bench N=$NP"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 5 --warmup 3 > $OUT/bench_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/bench_n$NP.log | tail -20
