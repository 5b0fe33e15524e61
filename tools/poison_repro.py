"""Hunt for reads of uninitialised device memory on the train / eval path.

Every scratch buffer of the path comes from ``torch.empty`` (PyTorch's caching allocator), so what an
unwritten element contains depends on what ran before - the signature of a failure that only shows up
inside a long test session.  This tool makes that deterministic: before every repetition the allocator's
cache is emptied and refilled with blocks whose bytes are all 0xFF (NaN as fp32 and fp16, -1 as int32), so
any element that is read without having been written turns the outputs into NaN (or trips an index check).

    python tools/poison_repro.py [reps] [pattern]     pattern: ff (default) | lo (small fp16-like garbage)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gcn_vae_b200 as K

DEV = torch.device("cuda:0")
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
PATTERN = sys.argv[2] if len(sys.argv) > 2 else "ff"


def poison(big_bytes=6 << 30, n_small=768):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    byte = 0xFF if PATTERN == "ff" else 0x11          # 0x1111 as fp16 = 6.5e-4, as fp32 = 1.1e-28
    big = torch.empty(big_bytes, dtype=torch.uint8, device=DEV).fill_(byte)
    mids = [torch.empty(3 << 20, dtype=torch.uint8, device=DEV).fill_(byte) for _ in range(64)]
    smalls = [torch.empty(1 << 19, dtype=torch.uint8, device=DEV).fill_(byte) for _ in range(n_small)]
    tiny = [torch.empty(512, dtype=torch.uint8, device=DEV).fill_(byte) for _ in range(4096)]
    torch.cuda.synchronize()
    del big, mids, smalls, tiny


def build(n_flows, sample_edges):
    n_ent, n_rel, h, bases, k = 14541, 237, 500, 100, 10
    data = K.datasets.synthetic_kg("FB15k-237", seed=0, scale=0.05)
    torch.manual_seed(0)
    model = K.LinkPredict(K.KGVAE, n_ent, h, n_rel, num_bases=bases, dropout=0.2, reg_param=0.01,
                          kl_param=1e-5, k=k, n_flows=n_flows).to(DEV)
    np.random.seed(0)
    g, node_id, edge_type, node_norm, samples, labels = K.utils.generate_sampled_graph_and_labels(
        data.train, sample_edges, 0.5, n_rel, None, None, 10, "uniform")
    n = len(node_id)
    eps = torch.randn(n, h)
    m1 = (torch.rand(n, h) < 0.8).float() / 0.8
    m2 = (torch.rand(n, 2 * h) < 0.8).float() / 0.8
    host = dict(g=g, node_id=node_id, edge_type=edge_type, node_norm=node_norm, samples=samples, labels=labels,
                eps=eps, m1=m1, m2=m2, test=data.train[:256] % np.array([n, n_rel, n]))
    return model, host


def step(model, host):
    """One train step + one rank evaluation with every input re-uploaded (fresh allocations)."""
    g = K.Graph()
    g.add_nodes(host["g"]._n)
    g.add_edges(host["g"]._src, host["g"]._dst)
    enc = model.encoder
    enc.preset_eps = host["eps"].to(DEV)
    enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = host["m1"].to(DEV), host["m2"].to(DEV)
    edge_norm = K.node_norm_to_edge_norm(g, torch.from_numpy(host["node_norm"]).view(-1, 1)).to(DEV)
    ids = torch.from_numpy(host["node_id"]).view(-1, 1).to(DEV)
    et = torch.from_numpy(host["edge_type"]).to(DEV)
    model.zero_grad(set_to_none=True)
    z = model(g, ids, et, edge_norm)
    loss, pred, kl, _ = model.get_loss(g, z, torch.from_numpy(host["samples"]).to(DEV),
                                       torch.from_numpy(host["labels"]).to(DEV))
    loss.backward()
    out = {"z": z.detach(), "loss": loss.detach().reshape(1), "kl": torch.as_tensor(kl).detach().reshape(1).float()}
    for name, p in model.named_parameters():
        if p.grad is not None:
            out["grad " + name] = p.grad.detach()
    mrr, ranks = K.utils.calc_mrr(z.detach(), model.w_relation, torch.from_numpy(host["test"]).to(DEV), eval_bz=128,
                                  flow_log_prob=model._flow_shift(), verbose=False, return_ranks=True)
    out["ranks"] = ranks.float()
    torch.cuda.synchronize()
    return {k_: v.cpu().clone() for k_, v in out.items()}


def main():
    bad_total = 0
    for n_flows, sample_edges in ((0, 2000), (1, 2000), (0, 8000)):
        model, host = build(n_flows, sample_edges)
        first = step(model, host)          # un-poisoned reference run
        for rep in range(REPS):
            poison()
            cur = step(model, host)
            msgs = []
            for key, ref in first.items():
                got = cur[key]
                n_nan = int((~torch.isfinite(got)).sum())
                scale = float(ref.abs().max().clamp_min(1e-30))
                dev = float((got.double() - ref.double()).abs().nan_to_num(0.0).max()) / scale
                lim = 0.0 if key == "ranks" else 2e-5     # atomics reorder fp32 sums: ~1e-6
                if n_nan or dev > lim:
                    msgs.append(f"{key}: {n_nan} non-finite, dev {dev:.2e}")
            if msgs:
                bad_total += 1
            print(f"[flows={n_flows} edges={sample_edges} rep {rep}] " + ("; ".join(msgs) if msgs else "clean"), flush=True)
    print("POISON RESULT:", "CLEAN" if bad_total == 0 else f"{bad_total} repetitions affected")
    return 0 if bad_total == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
