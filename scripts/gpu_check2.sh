#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=5 M4T_DEBUG_SEGV=1
echo "=== build"; timeout 900 python __graft_entry__.py 2>&1 | tail -2
echo "=== p2p debug np=1 blocks=16"; timeout 200 python -m mpi4torch_b200.launch -np 1 scripts/debug_p2p.py 2>&1 | grep -v "^W0" | tail -20
echo "=== p2p debug np=1 blocks=1"; M4T_P2P_BLOCKS=1 timeout 200 python -m mpi4torch_b200.launch -np 1 scripts/debug_p2p.py 2>&1 | grep -v "^W0" | tail -20
echo "=== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -s 2>&1 | grep -v Deprecat | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
