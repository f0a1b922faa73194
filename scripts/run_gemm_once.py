"""Tiny driver for ncu: a few launches of the CTA-pair tcgen05 GEMM at the flagship shape."""
import torch
import mpi4torch_b200 as m4t

m4t.COMM_WORLD
M, N, K = 8192, 4096, 4096
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    y = torch.ops.mpi4torch_b200.gemm_bf16_tn(x, w)
torch.cuda.synchronize()
print("done", float(y[0, 0]))
