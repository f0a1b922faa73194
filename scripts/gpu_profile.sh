#!/bin/bash
# 1 GPU: ncu capture of the tcgen05 GEMM + launch list of a short bench run.
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD
echo "=== ncu full: gemm_bf16_tn"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 3 -c 1 -f -o $OUT/prof_gemm python scripts/run_gemm_once.py 2>&1 | tail -5
echo "=== ncu full: GEMM + fused MSE epilogue (forward of the flagship step)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 3 -c 1 -f -o $OUT/prof_gemm_mse python scripts/run_gemm_once.py mse 2>&1 | tail -5
if [ "${M4T_TEST_EXPERIMENTAL:-0}" = "1" ]; then
echo "=== ncu full: MN-major wgrad GEMM (experimental)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf16 -s 3 -c 1 -f -o $OUT/prof_wgrad python scripts/run_gemm_once.py wgrad 2>&1 | tail -5
fi
for r in prof_gemm prof_gemm_mse prof_wgrad; do
  [ -f $OUT/$r.ncu-rep ] && ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/$r.raw.csv 2>/dev/null
done
echo "=== launch list of bench.py (N=1)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $OUT/launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-extras > $OUT/bench_under_ncu.log 2>&1
tail -3 $OUT/bench_under_ncu.log | cut -c1-300
echo "=== bench N=1 (clean)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.log 2>&1; grep -v "^W0" $OUT/bench_n1.log | tail -3
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
