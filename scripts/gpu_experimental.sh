#!/bin/bash
# Brings up the kernels / paths that are compiled but have not run on hardware yet
# (DESIGN.md section 10).  usage:  gpurun --gpus <np> --timeout 900 -- bash scripts/gpu_experimental.sh <np>
set -u
NP=${1:-1}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=90 M4T_DEVICE_TIMEOUT_S=10 M4T_DEBUG_SEGV=1 M4T_TEST_EXPERIMENTAL=1
echo "=== [1 GPU] MN-major wgrad GEMM numerics"
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -k wgrad > $OUT/exp_wgrad.log 2>&1; echo "exit=$?"; tail -5 $OUT/exp_wgrad.log | cut -c1-400
if [ "$NP" -gt 1 ]; then
  echo "=== sub-communicators on CUDA np=$NP"
  M4T_TEST_DEVICE=cuda M4T_TEST_SPLIT_CUDA=1 timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_split.py > $OUT/exp_split_np$NP.log 2>&1
  echo "exit=$?"; grep -v "^W0" $OUT/exp_split_np$NP.log | tail -6 | cut -c1-400
  echo "=== fused wgrad -> reduce-scatter -> SGD -> multicast np=$NP"
  M4T_TEST_DEVICE=cuda M4T_FUSED_WGRAD=2 timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_gpu.py > $OUT/exp_fwgrad_np$NP.log 2>&1
  echo "exit=$?"; grep -v "^W0" $OUT/exp_fwgrad_np$NP.log | tail -6 | cut -c1-400
  echo "=== randomised traffic patterns np=$NP"
  M4T_TEST_DEVICE=cuda timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_stress.py 2>&1 | grep -v "^W0" | tail -3
  echo "=== side-stream bucketed gradient sync np=$NP"
  M4T_TEST_DEVICE=cuda timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_parallel.py > $OUT/exp_parallel_np$NP.log 2>&1
  echo "exit=$?"; grep -v "^W0" $OUT/exp_parallel_np$NP.log | tail -4 | cut -c1-400
  echo "=== bench with the fused backward np=$NP"
  M4T_FUSED_WGRAD=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/exp_bench_fwgrad_n$NP.log 2>&1
  echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/exp_bench_fwgrad_n$NP.log | tail -1 | cut -c1-900
  echo "=== fused forward diagnosis np=$NP (full / comm skipped / GEMM skipped / target prefetch)"
  for v in "full:" "nocomm:M4T_FUSED_DEBUG=1" "nogemm:M4T_FUSED_DEBUG=2" "prefetch:M4T_EPI_PREFETCH=1"; do
    name=${v%%:*}; kv=${v#*:}
    env VARIANT=$name $kv timeout 300 python -m mpi4torch_b200.launch -np $NP scripts/fused_diag.py 2>&1 | grep -v "^W0" | tail -1 | tee -a $OUT/exp_fused_diag_np$NP.jsonl
  done
  echo "=== step breakdown np=$NP"; timeout 200 python -m mpi4torch_b200.launch -np $NP scripts/step_breakdown.py 2>&1 | grep -v "^W0" | tail -6 | tee $OUT/exp_step_breakdown_np$NP.jsonl
  echo "=== collectives np=$NP (rotated pull plans)"; timeout 600 python -m mpi4torch_b200.launch -np $NP benchmarks/collectives_bench.py --max-mb 64 --out $OUT/exp_collectives_np$NP.json 2>&1 | grep -v "^W0" | tail -8 | cut -c1-700
  echo "=== p2p ring: pull (default) vs push np=$NP"
  for push in 0 1; do
    M4T_P2P_PUSH=$push timeout 300 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/exp_ring_push${push}_np$NP.json 2>&1 | grep -v "^W0" | tail -2
  done
  for cfg in "64 256 64" "32 512 64"; do
    set -- $cfg
    for push in 0 1; do
      echo "slots=$1 slot_kb=$2 blocks=$3 push=$push"
      M4T_P2P_SLOTS=$1 M4T_P2P_SLOT_KB=$2 M4T_P2P_BLOCKS=$3 M4T_P2P_PUSH=$push timeout 300 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 2>&1 | grep -v "^W0" | tail -1
    done
  done
  M4T_TEST_DEVICE=cuda M4T_P2P_PUSH=1 timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_nonblocking.py 2>&1 | grep -v "^W0" | tail -3
  echo "=== bench with fused backward + prefetched parameter average np=$NP"
  M4T_FUSED_WGRAD=2 M4T_WAVG_PREFETCH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > $OUT/exp_bench_prefetch_n$NP.log 2>&1
  echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/exp_bench_prefetch_n$NP.log | tail -1 | cut -c1-900
  echo "=== zero-copy symmetric allreduce inputs np=$NP"
  M4T_TEST_DEVICE=cuda M4T_ZERO_COPY_IN=1 timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_gpu.py 2>&1 | grep -v "^W0" | tail -3
  echo "=== multicast-push Allgather np=$NP"
  M4T_TEST_DEVICE=cuda M4T_AG_PUSH=1 timeout 600 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py spmd_collectives.py > $OUT/exp_agpush_np$NP.log 2>&1
  echo "exit=$?"; grep -v "^W0" $OUT/exp_agpush_np$NP.log | tail -4 | cut -c1-400
fi
