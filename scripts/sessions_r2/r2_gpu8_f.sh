#!/bin/bash
# Round-2 session F: the multi-GPU numbers (np = 4 or 8): suites, fused-backward experiments, bench both arms,
# axis collectives and the p2p ring with their variants.
set -u
NP=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1 M4T_TEST_EXPERIMENTAL=1
echo "=== suites np=$NP"
M4T_TEST_DEVICE=cuda timeout 300 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py "spmd_[gcn]*.py" > $OUT/f_spmd_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/f_spmd_np$NP.log | tail -5 | cut -c1-400
echo "=== wgrad diag np=$NP"
timeout 150 python -m mpi4torch_b200.launch -np $NP scripts/wgrad_diag.py 2>&1 | grep "^{" | tee $OUT/f_wgrad_diag_np$NP.jsonl
echo "=== step breakdown np=$NP"
timeout 120 python -m mpi4torch_b200.launch -np $NP scripts/step_breakdown.py 2>&1 | grep "^{" | tee $OUT/f_step_breakdown_np$NP.jsonl
echo "=== bench ours np=$NP"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NP --steps 20 --warmup 5 --full-sweep > $OUT/f_bench_ours_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/f_bench_ours_n$NP.log | tail -2 | cut -c1-3800
echo "=== bench reference np=$NP"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $NP --steps 6 --warmup 3 > $OUT/f_bench_ref_n$NP.log 2>&1
echo "exit=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" $OUT/f_bench_ref_n$NP.log | tail -2 | cut -c1-2500
echo "=== collectives np=$NP: default / multicast-push allgather"
for v in "pull:M4T_AG_PUSH=0" "push:M4T_AG_PUSH=1"; do
  name=${v%%:*}; kv=${v#*:}
  env $kv timeout 200 python -m mpi4torch_b200.launch -np $NP benchmarks/collectives_bench.py --max-mb 64 --out $OUT/f_collectives_${name}_np$NP.json 2>&1 | grep -v "^W0" | tail -7 | cut -c1-600
done
echo "=== p2p ring np=$NP: SM copy kernels / copy engines (16 MiB ring) / copy engines (64 MiB ring)"
M4T_P2P_CE_MIN_KB=-1 timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/f_ring_kernel_np$NP.json 2>&1 | grep "^{" | cut -c1-600
timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/f_ring_ce16_np$NP.json 2>&1 | grep "^{" | cut -c1-600
M4T_P2P_SLOTS=64 timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/f_ring_ce64_np$NP.json 2>&1 | grep "^{" | cut -c1-600
