#!/bin/bash
# 1 GPU: full GPU test tier (as the driver runs it), smoke, N=1 bench of both arms, ncu captures of the new kernels,
# compute-sanitizer on the single-GPU kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/k_pytest_gpu.log 2>&1; echo "exit=$?"; tail -4 $OUT/k_pytest_gpu.log | cut -c1-300
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2 | cut -c1-300
echo "=== bench N=1 ours"
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/k_bench_ours_n1.log 2>&1; grep -v "^W0" $OUT/k_bench_ours_n1.log | tail -1 | cut -c1-2500
echo "=== bench N=1 reference"
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/k_bench_ref_n1.log 2>&1; grep -v "^W0" $OUT/k_bench_ref_n1.log | tail -1 | cut -c1-2000
echo "=== ncu: wgrad with SGD epilogue, NN dgrad GEMM"
cat > /tmp/run_new.py <<'PY'
import sys, torch, mpi4torch_b200 as m4t
m4t.COMM_WORLD
M, N, K = 8192, 4096, 4096
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
g = torch.ones(1, device="cuda")
for _ in range(4):
    if sys.argv[1] == "sgd": torch.ops.mpi4torch_b200.wgrad_sgd_(w, dy, x, -1e-6, g)
    else: out = torch.ops.mpi4torch_b200.gemm_bf16_nn(dy, w)
torch.cuda.synchronize(); print("done")
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf16_nt -s 3 -c 1 -f -o $OUT/prof_wgrad_sgd python /tmp/run_new.py sgd 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_2cta -s 3 -c 1 -f -o $OUT/prof_gemm_nn python /tmp/run_new.py nn 2>&1 | tail -2
for r in prof_wgrad_sgd prof_gemm_nn; do [ -f $OUT/$r.ncu-rep ] && ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/$r.raw.csv 2>/dev/null; done
echo "=== launch list of the N=1 bench (ncu, cold, serialised: shares only)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 60 --csv --log-file $OUT/k_launches_bench_n1.csv python bench.py --steps 3 --warmup 3 --no-extras > $OUT/k_bench_under_ncu.log 2>&1; tail -1 $OUT/k_bench_under_ncu.log | cut -c1-200
echo "=== compute-sanitizer"
timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/k_memcheck_gemm.log python -m pytest tests/test_gpu_gemm.py -x -q -k "shape0 or sgd_epilogue or nn_dgrad and shape1" 2>&1 | tail -2
timeout 300 compute-sanitizer --tool racecheck --log-file $OUT/k_racecheck_gemm.log python -m pytest tests/test_gpu_gemm.py -x -q -k "test_gemm_2cta_matches_fp32_reference and shape0 or wgrad_mn_major and shape0" 2>&1 | tail -2
M4T_TEST_DEVICE=cuda timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/k_memcheck_np1.log python -m mpi4torch_b200.launch -np 1 tests/spmd/run_all.py "spmd_collectives.py" 2>&1 | tail -2
grep -h "ERROR SUMMARY" $OUT/k_memcheck_*.log $OUT/k_racecheck_*.log
