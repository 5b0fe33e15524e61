"""Quick correctness + timing of kg_gemm_f32 at the benchmarked shapes (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcn_vae_b200 import ops

dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
for (M, N, K, ta, tb, full) in [(1906, 500, 500, False, False, True), (14541, 500, 500, False, False, True),
                                (14541, 1000, 500, False, False, True), (14541, 500, 1000, False, True, False),
                                (40914, 500, 500, False, True, True), (1024, 256, 64, False, False, False),
                                (131, 500, 500, False, False, True), (500, 1000, 14541, True, False, False),
                                (500, 500, 40914, True, False, False), (300, 200, 1000, True, True, False),
                                (333, 257, 129, True, False, True), (200, 136, 77, False, True, True),
                                (312500, 1000, 500, False, False, True)]:
    a = torch.randn((K, M) if ta else (M, K), device=dev, generator=gen)
    b = torch.randn((N, K) if tb else (K, N), device=dev, generator=gen)
    kw = {}
    if full:
        kw = dict(bias=torch.randn(N, device=dev, generator=gen), addend=torch.randn(M, N, device=dev, generator=gen),
                  relu=True, mask=(torch.rand(M, N, device=dev, generator=gen) < 0.8).float() / 0.8)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, b, out, trans_a=ta, trans_b=tb, **kw)
    torch.cuda.synchronize()
    am, bm = (a.t() if ta else a), (b.t() if tb else b)
    want = am.double() @ bm.double() if M * N * K < 3e11 else (am @ bm).double()
    if full:
        want = torch.relu(want + kw["bias"].double() + kw["addend"].double()) * kw["mask"].double()
    err = float((out.double() - want).abs().max() / want.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.gemm(a, b, out, trans_a=ta, trans_b=tb, **kw)
    e0.record()
    for _ in range(10):
        ops.gemm(a, b, out, trans_a=ta, trans_b=tb, **kw)
    e1.record(); e1.synchronize()
    print(f"{M}x{N}x{K} {'T' if ta else 'N'}{'T' if tb else 'N'} epilogue={full}: err {err:.2e}  {e0.elapsed_time(e1) / 10:.3f} ms  "
          f"finite={bool(torch.isfinite(out).all())}", flush=True)

# prepared operands reused: the product alone (what a layer pays after x, W and g are split once)
for (M, N, K, ta, tb) in [(14541, 500, 500, False, False), (14541, 1000, 500, False, False), (14541, 500, 1000, False, True),
                          (500, 1000, 14541, True, False), (40914, 500, 500, False, True), (500, 500, 40914, True, False)]:
    a = torch.randn((K, M) if ta else (M, K), device=dev, generator=gen)
    b = torch.randn((N, K) if tb else (K, N), device=dev, generator=gen)
    out = torch.empty((M, N), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pa, pb = ops.Prepared(a), ops.Prepared(b)
    e1.record(); e1.synchronize()
    t_prep = e0.elapsed_time(e1) / 10
    for _ in range(3):
        ops.gemm(pa, pb, out, trans_a=ta, trans_b=tb)
    e0.record()
    for _ in range(10):
        ops.gemm(pa, pb, out, trans_a=ta, trans_b=tb)
    e1.record(); e1.synchronize()
    print(f"{M}x{N}x{K} {'T' if ta else 'N'}{'T' if tb else 'N'}: prepare both {t_prep:.3f} ms, product alone {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
