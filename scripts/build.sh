#!/bin/bash
# Always build from the repository root.
cd "$(dirname "$0")/.." && python -m mpi4torch_b200._build 2>&1 | grep -E "error|FAILED|built"
