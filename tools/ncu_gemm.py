"""Driver for one `ncu --set full` capture of the dense GEMM on prepared operands (the benchmarked shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_vae_b200 import ops

dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(0)
for (M, N, K, ta, tb) in [(14541, 1000, 500, False, False), (14541, 500, 500, False, True), (500, 1000, 14541, True, False),
                          (40914, 500, 500, False, True), (500, 500, 40914, True, False)]:
    a = torch.randn((K, M) if ta else (M, K), device=dev, generator=gen)
    b = torch.randn((N, K) if tb else (K, N), device=dev, generator=gen)
    out = torch.empty((M, N), device=dev)
    pa, pb = ops.Prepared(a), ops.Prepared(b)
    for _ in range(2):
        ops.gemm(pa, pb, out, trans_a=ta, trans_b=tb)
torch.cuda.synchronize()
print("done")
