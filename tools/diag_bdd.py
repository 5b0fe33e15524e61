"""Diagnostic: bdd message passing (C ABI) against a float64 torch evaluation on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_vae_b200 as K
from gcn_vae_b200 import ops, _lib as L

dev = "cuda:0"
torch.manual_seed(0)
for (n, e, r, B, si, so) in [(1900, 4000, 474, 100, 5, 5), (1900, 4000, 474, 100, 5, 10), (14541, 272114, 474, 100, 5, 5),
                             (14541, 272114, 474, 100, 5, 10), (300, 6000, 11, 50, 10, 10), (300, 6000, 11, 24, 4, 4)]:
    src = torch.randint(0, n, (e,), device=dev, dtype=torch.int32)
    dst = torch.randint(0, n, (e,), device=dev, dtype=torch.int32)
    et = torch.randint(0, r, (e,), device=dev, dtype=torch.int32)
    norm = torch.rand(e, device=dev) + 0.1
    gi = ops.graph_index(src, dst, et, norm, n, r)
    x = torch.randn(n, B * si, device=dev)
    w = torch.randn(r, B * si * so, device=dev) * 0.3
    w_fwd = torch.empty((r, si, B * so), device=dev); w_bwd = torch.empty((r, so, B * si), device=dev)
    L.call("kg_bdd_weight_layouts", L.f32(w), r, B, si, so, L.f32(w_fwd), L.f32(w_bwd), L.stream())
    agg = torch.zeros(n, B * so, device=dev)
    L.call("kg_bdd_rel_fwd", L.f32(x), None, 0, L.i32(gi.rel_pack), e, L.f32(w), L.f32(w_fwd), B, si, so, L.f32(agg), 0, L.stream())
    # float64 reference in chunks
    ref = torch.zeros(n, B * so, device=dev, dtype=torch.float64)
    dref_x = torch.zeros(n, B * si, device=dev, dtype=torch.float64)
    dref_w = torch.zeros(r, B, si, so, device=dev, dtype=torch.float64)
    g = torch.randn(n, B * so, device=dev)
    for c0 in range(0, e, 20000):
        sl = slice(c0, min(e, c0 + 20000))
        ws = w[et[sl].long()].double().view(-1, B, si, so)
        xs = x[src[sl].long()].double().view(-1, B, 1, si)
        msg = torch.matmul(xs, ws).view(-1, B * so) * norm[sl].double().view(-1, 1)
        ref.index_add_(0, dst[sl].long(), msg)
        gd = g[dst[sl].long()].double().view(-1, B, so, 1) * norm[sl].double().view(-1, 1, 1, 1)
        dref_x.index_add_(0, src[sl].long(), torch.matmul(ws, gd).view(-1, B * si))
        dref_w.index_add_(0, et[sl].long(), torch.matmul(xs.transpose(-1, -2), gd.transpose(-1, -2)))
    dx = torch.zeros(n, B * si, device=dev); dw = torch.zeros_like(w)
    L.call("kg_bdd_rel_bwd", L.f32(x), None, 0, L.f32(g), L.i32(gi.rel_pack), e, L.f32(w), L.f32(w_bwd), B, si, so, L.f32(dx), L.f32(dw), 0, L.stream())
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
    print(f"n={n} e={e} B={B} {si}x{so}: fwd {rel(agg, ref):.2e}  dx {rel(dx, dref_x):.2e}  dW {rel(dw.view(r, B, si, so), dref_w):.2e}", flush=True)
