#!/bin/bash
# Round-2 first 1-GPU session: wgrad bring-up, ncu captures, N=1 bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1 M4T_TEST_EXPERIMENTAL=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "=== wgrad MN-major numerics (small shapes first)"
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k wgrad > $OUT/a_wgrad.log 2>&1; echo "exit=$?"; tail -30 $OUT/a_wgrad.log | cut -c1-300
echo "=== ncu full: gemm 2cta / mse / wgrad"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 3 -c 1 -f -o $OUT/prof_gemm python scripts/run_gemm_once.py gemm 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 3 -c 1 -f -o $OUT/prof_gemm_mse python scripts/run_gemm_once.py mse 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf16_nt -s 3 -c 1 -f -o $OUT/prof_wgrad python scripts/run_gemm_once.py wgrad 2>&1 | tail -3
for r in prof_gemm prof_gemm_mse prof_wgrad; do
  [ -f $OUT/$r.ncu-rep ] && ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/$r.raw.csv 2>/dev/null
done
echo "=== gemm speed"
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -s -k "speed" 2>&1 | grep "\[gemm\]"
echo "=== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/a_bench_n1.log 2>&1; grep -v "^W0" $OUT/a_bench_n1.log | tail -2 | cut -c1-1500
