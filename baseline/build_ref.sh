#!/bin/bash
# Builds the reference arm: the shared-memory MPI shim (baseline/mpi_shim) and the UNMODIFIED
# reference (/root/reference, installed from a /tmp copy because the source tree is read-only)
# into baseline/_ref.  The reference's setup.py swaps compiler_so[0]/compiler_cxx[0] for
# mpicc/mpicxx, which newer setuptools no longer consult for C++ sources, so the MPI wrappers are
# also exported as CC/CXX — the usual way to build an MPI extension.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref_src="${M4T_REFERENCE_SRC:-/root/reference}"
make -C "$here/mpi_shim" >/dev/null
if [ -f "$here/_ref/mpi4torch/__init__.py" ] && ls "$here"/_ref/mpi4torch/_mpi*.so >/dev/null 2>&1 && [ "${1:-}" != "--force" ]; then
  echo "[build_ref] baseline/_ref already built"; exit 0
fi
[ -d "$ref_src" ] || { echo "[build_ref] $ref_src not found"; exit 3; }
tmp="$(mktemp -d /tmp/refcopy.XXXXXX)"
cp -r "$ref_src/." "$tmp/"
rm -rf "$here/_ref"
PATH="$here/mpi_shim/bin:$PATH" CC=mpicc CXX=mpicxx python -m pip install --no-index --no-build-isolation --no-deps \
  --find-links /opt/wheelhouse --target "$here/_ref" "$tmp" 2>&1 | tail -3
rm -rf "$tmp"
ls "$here"/_ref/mpi4torch/_mpi*.so
