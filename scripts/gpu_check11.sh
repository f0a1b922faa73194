#!/bin/bash
set -u
NP=${1:-2}
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10
run() { timeout 200 python -m mpi4torch_b200.launch -np $NP scripts/fused_diag.py 2>&1 | grep -v "^W0" | tail -2; }
VARIANT=1cta run
VARIANT=1cta_nocomm M4T_FUSED_DEBUG=1 run
VARIANT=1cta_nogemm M4T_FUSED_DEBUG=2 run
VARIANT=2cta M4T_FUSED_2CTA=1 run
VARIANT=2cta_nocomm M4T_FUSED_2CTA=1 M4T_FUSED_DEBUG=1 run
echo "=== 2cta fused correctness"; M4T_FUSED_2CTA=1 M4T_TEST_DEVICE=cuda timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py "spmd_gpu.py" 2>&1 | grep -v "^W0" | tail -4 | cut -c1-300
