#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=5 M4T_DEBUG_SEGV=1 M4T_TEST_DEVICE=cuda
echo "=== only TestLargeP2P np=1"; M4T_DEBUG=1 timeout 200 python -m mpi4torch_b200.launch -np 1 tests/spmd/run_all.py "spmd_gpu.py" > $OUT/gpu_mod.log 2>&1; grep -v "^W0" $OUT/gpu_mod.log | grep -v "attached\|heap:" | tail -60
