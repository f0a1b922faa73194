#!/bin/bash
# AddressSanitizer + UBSan run of the host runtime (control plane, CPU backend, plans,
# autograd layer) under the CPU SPMD suites.  Builds an instrumented copy of the extension
# out of tree (/tmp/m4t_asan); the in-tree release build is untouched.
#   usage: bash scripts/asan_cpu.sh [np ...]      (default: 2 5)
set -u
cd "$(dirname "$0")/.."
export M4T_LIB_DIR=/tmp/m4t_asan
export M4T_EXTRA_CXXFLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer -g -O1"
export M4T_EXTRA_LDFLAGS="-fsanitize=address,undefined"
ASAN_LIB=$(/usr/bin/g++ -print-file-name=libasan.so)
UBSAN_LIB=$(/usr/bin/g++ -print-file-name=libubsan.so)
export ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0:abort_on_error=0:halt_on_error=1
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0
mkdir -p $M4T_LIB_DIR
echo "=== building instrumented extension into $M4T_LIB_DIR"
# compile WITHOUT the preload (nvcc crashes under it); the final import of the instrumented
# module fails here for lack of the ASan runtime - expected, the objects are built by then
python -m mpi4torch_b200._build > $M4T_LIB_DIR/build.log 2>&1 || true
ls -la $M4T_LIB_DIR/_m4t_C.so || { tail -20 $M4T_LIB_DIR/build.log; exit 1; }
for np in ${@:-2 5}; do
  echo "=== SPMD suite np=$np under ASan/UBSan"
  LD_PRELOAD="$ASAN_LIB $UBSAN_LIB" timeout 1800 python -m mpi4torch_b200.launch --nproc $np tests/spmd/run_all.py 2>&1 \
    | grep -v "No CUDA runtime\|^W0" | grep -E "ERROR: AddressSanitizer|runtime error|SUMMARY|SPMD suite|FAILED|^OK|Ran " | head -40
done
