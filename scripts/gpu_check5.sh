#!/bin/bash
# full GPU test pass + tuning data.  usage: gpu_check5.sh <np>
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "Deprecat\|warnings.warn\|^$" | tail -15
echo "=== tune allreduce np=$NP"; timeout 600 python -m mpi4torch_b200.launch -np $NP benchmarks/tune_allreduce.py > $OUT/tune_np$NP.log 2>&1; grep -v "^W0" $OUT/tune_np$NP.log | tail -50
for mask in 1 2 4 6 5 3; do
  echo "=== skip mask $mask (256MB nvls blocks=128 chunk=8MB)"
  M4T_AR_DEBUG_SKIP=$mask TUNE_SIZES=268435456 TUNE_BLOCKS=128 TUNE_CHUNKS_KB=8192 TUNE_ALGOS=3 timeout 120 python -m mpi4torch_b200.launch -np $NP benchmarks/tune_allreduce.py 2>&1 | grep -v "^W0" | tail -1
done
