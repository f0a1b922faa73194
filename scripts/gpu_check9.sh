#!/bin/bash
set -u
NP=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10
echo "=== step breakdown np=$NP"; timeout 300 python -m mpi4torch_b200.launch -np $NP scripts/step_breakdown.py 2>&1 | grep -v "^W0" | tail -8
echo "=== bench N=$NP"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/bench_n$NP.log 2>&1; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/bench_n$NP.log | tail -1 | cut -c1-600
echo "=== gemm 2cta tests (1 gpu)"; CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -s -k "2cta or speed" 2>&1 | grep -v "Deprecat\|warnings.warn\|^$" | tail -12 | cut -c1-300
