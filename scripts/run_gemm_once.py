"""Tiny driver for ncu: a few launches of each single-GPU tensor-core kernel at the flagship shape.

  python scripts/run_gemm_once.py [gemm|mse|wgrad|all]
"""
import sys

import torch
import mpi4torch_b200 as m4t

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
m4t.COMM_WORLD
M, N, K = 8192, 4096, 4096
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
t = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
out = None
if which in ("gemm", "all"):
    for _ in range(4):
        out = torch.ops.mpi4torch_b200.gemm_bf16_tn(x, w)
if which in ("mse", "all"):  # forward GEMM with the fused MSE / dL/dy epilogue
    for _ in range(4):
        out, loss, _ = torch.ops.mpi4torch_b200.linear_mse_forward(x, w, t, 1.0, 1.0 / M, 2.0 / M, False)
if which in ("wgrad", "all"):  # experimental MN-major weight-gradient GEMM
    dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        out = torch.ops.mpi4torch_b200.wgrad_bf16(dy, x)
torch.cuda.synchronize()
print("done", float(out.flatten()[0]))
