"""Development aid: times the pieces a two-pass DistMult decoder would be made of, at the FB15k-237 step shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gcn_vae_b200 as K
from gcn_vae_b200 import ops, _lib as L
import bench

dev = torch.device("cuda:0")
data = K.datasets.synthetic_kg("FB15k-237", seed=0)
g, node_id, etype, node_norm, samples, labels = bench.sample_step(K.utils, data, len(data.train), seed=0)
N, h = len(node_id), 500
trip = torch.from_numpy(samples).to(torch.int32).to(dev)
lab = torch.from_numpy(labels).float().to(dev)
S = trip.shape[0]
z = torch.randn(N, h, device=dev) * 0.3
w = torch.randn(data.num_rels, h, device=dev) * 0.3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


idx = ops.TripletIndex(trip, N, data.num_rels, entity_index=True)
gsc = torch.empty(S, device=dev)
dw = torch.zeros_like(w)
dz = torch.zeros_like(z)
out = torch.empty(2, device=dev)
ws = L.workspace(L.lib().kg_distmult_bce_workspace_bytes(S), dev)


def fused(with_dz):
    L.call("kg_distmult_bce_fwd", L.f32(z), L.f32(w), L.i32(idx.rs_rec), L.f32(lab), S, h, None, None, L.f32(gsc),
           L.f32(dw), L.f32(dz) if with_dz else None, L.ptr(out[0:1]), L.ptr(out[1:2]), L.ptr(ws), ws.numel(), L.stream())


def gather_dz():
    L.call("kg_distmult_bwd_dz", L.f32(z), L.f32(w), L.f32(gsc), L.i32(idx.ent_ptr), L.i32(idx.ent_pack), N, h,
           L.f32(dz), L.stream())


print(f"S={S} N={N}")
print(f"fused pass with dz (today):      {timeit(lambda: fused(True)):.3f} ms")
print(f"fused pass without dz:           {timeit(lambda: fused(False)):.3f} ms")
print(f"(entity, r) gather of dz, 2S:    {timeit(gather_dz):.3f} ms")
print(f"triplet index, rs only:          {timeit(lambda: ops.TripletIndex(trip, N, data.num_rels, entity_index=False)):.3f} ms")
print(f"triplet index, rs + entity (2S): {timeit(lambda: ops.TripletIndex(trip, N, data.num_rels, entity_index=True)):.3f} ms")
