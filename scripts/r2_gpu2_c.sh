#!/bin/bash
# Round-2 session C: why is the fused backward slow?  (np GPUs)
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_FUSED_WGRAD=2
for v in "base:0" "full:0" "nocomm:1" "nogemm:2" "local_ld:4" "local_st:8" "local_ldst:12" "nogemm_local:14"; do
  name=${v%%:*}; dbg=${v#*:}
  VARIANT=$name M4T_WGRAD_DEBUG=$dbg timeout 200 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep -v "^W0" | tail -1 | tee -a $OUT/c_wgrad_diag_np$NP.jsonl
done
echo "=== ksplit=1"
VARIANT=ksplit1 M4T_WGRAD_KSPLIT=1 timeout 200 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep -v "^W0" | tail -1 | tee -a $OUT/c_wgrad_diag_np$NP.jsonl
