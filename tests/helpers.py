"""Shared helpers for the parity tests (oracle side only - never imported by the package)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import kgvae_oracle as O  # noqa: E402

RTOL = 1e-4   # north_star: embeddings, loss and scores within 1e-4 relative in fp32


def rel_err(a, b):
    """max |a-b| / max(|b|_max, tiny): error relative to the tensor's scale."""
    a, b = (t.detach().cpu() if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t)) for t in (a, b))
    a, b = a.to(torch.float64), b.to(torch.float64)
    if a.numel() == 1 and b.numel() == 1:      # the reference mixes [] and [1] scalars
        a, b = a.reshape(()), b.reshape(())
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def assert_close(a, b, rtol=RTOL, what=""):
    e = rel_err(a, b)
    assert e <= rtol, f"{what}: relative error {e:.3e} > {rtol:.1e}"


def golden_params(gv, requires_grad=False):
    p = {}
    for key, val in gv.items():
        if key.startswith("param/") and not key.endswith(".mask") and not key.endswith("pi"):
            t = torch.from_numpy(val.copy())
            if requires_grad:
                t.requires_grad_(True)
            p[key[len("param/"):]] = t
    return p


def golden_graph(gv, prefix="g_", etype_key="edge_type", norm_key="node_norm", n=None):
    dst = gv[prefix + "dst"].astype(np.int64)
    norm = gv[norm_key].astype(np.float32)
    return {"num_nodes": int(len(norm) if n is None else n),
            "src": gv[prefix + "src"].astype(np.int64), "dst": dst,
            "etype": gv[etype_key].astype(np.int64), "norm": norm,
            "edge_norm": norm[dst].reshape(-1, 1)}


def oracle_train_step(gv, requires_grad=True):
    """Run the oracle port on a golden case's inputs; returns (params, enc, loss dict)."""
    n_ent, n_rel, h, bases, k, n_flows, neg = (int(x) for x in gv["cfg"])
    reg_param, kl_param, dropout = (float(x) for x in gv["cfg_f"])
    params = golden_params(gv, requires_grad)
    graph = golden_graph(gv)
    enc = O.kgvae_encode(params, graph, gv["node_id"], torch.from_numpy(gv["eps"]), bases, n_flows,
                         (torch.from_numpy(gv["mask1"]), torch.from_numpy(gv["mask2"])))
    out = O.kgvae_loss(params, enc, gv["samples"], gv["labels"], reg_param, kl_param, n_flows)
    return params, enc, out
