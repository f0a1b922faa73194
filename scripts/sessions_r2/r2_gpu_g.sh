#!/bin/bash
# p2p copy-engine path: correctness + ring numbers
set -u
NP=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONPATH=$PWD M4T_TIMEOUT_S=60 M4T_DEVICE_TIMEOUT_S=10 M4T_NO_BUILD=1 M4T_TEST_EXPERIMENTAL=1
echo "=== nonblocking + stress suites (CE path for >= 2 MiB) np=$NP"
M4T_TEST_DEVICE=cuda timeout 200 python -m mpi4torch_b200.launch -np $NP tests/spmd/run_all.py "spmd_[ns]*.py" > $OUT/g_spmd_p2p_np$NP.log 2>&1
echo "exit=$?"; grep -v "^W0" $OUT/g_spmd_p2p_np$NP.log | tail -6 | cut -c1-400
echo "=== ring np=$NP: kernel path / copy-engine path (16 MiB ring) / copy-engine path (64 MiB ring)"
M4T_P2P_CE_MIN_KB=-1 timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/g_ring_kernel_np$NP.json 2>&1 | grep "^{" | cut -c1-600
timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/g_ring_ce16_np$NP.json 2>&1 | grep "^{" | cut -c1-600
M4T_P2P_SLOTS=64 timeout 100 python -m mpi4torch_b200.launch -np $NP benchmarks/ring_overlap.py --mb 64 --out $OUT/g_ring_ce64_np$NP.json 2>&1 | grep "^{" | cut -c1-600
