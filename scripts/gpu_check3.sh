#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=5 M4T_DEBUG_SEGV=1
NP=${1:-2}
echo "=== ring debug np=1"; M4T_DEBUG=1 timeout 200 python -m mpi4torch_b200.launch -np 1 scripts/debug_ring.py > $OUT/ring1.log 2>&1; grep -v "^W0" $OUT/ring1.log | grep -v "attached\|heap:" | tail -40
echo "=== ring debug np=$NP"; timeout 200 python -m mpi4torch_b200.launch -np $NP scripts/debug_ring.py 2>&1 | grep -v "^W0" | tail -24
echo "=== allreduce sweep np=$NP"; timeout 600 python -m mpi4torch_b200.launch -np $NP benchmarks/allreduce_sweep.py --raw --out $OUT/sweep_np$NP.json 2>&1 | grep -v "^W0" | tail -12
