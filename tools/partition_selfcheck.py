"""Destination-partitioned training step against the single-GPU step on the same graph, noise and dropout
masks: the summed loss, the owned rows of z and every gradient must agree.

Used from two places: tests/test_gpu_partitioned.py (2 ranks spawned by pytest) and bench.py, which runs the
four variants inside every --gpus N > 1 job and ASSERTS them, so that they execute on whatever multi-GPU box
the benchmark runs on.  Every rank computes the single-GPU reference itself (the toy graph is small)."""
import numpy as np
import torch
import torch.distributed as dist

# (n_flows, mode, mmd_param, column chunks of the layer-input gathers; None = the Partition's own choice: 2 from 4 ranks up)
VARIANTS = [(0, "allgather", 0.0, None), (1, "allgather", 0.0, None), (0, "peer", 0.0, None), (1, "peer", 0.0, None),
            (1, "allgather", 1.0, None), (0, "allgather", 0.0, 2)]


def setup(K, dev, n_flows, mmd_param=0.0):
    rng = np.random.default_rng(0)
    n_ent, n_rel, h, bases, T, S = 211, 5, 40, 8, 900, 3000
    torch.manual_seed(0)
    model = K.LinkPredict(K.KGVAE, n_ent, h, n_rel, num_bases=bases, dropout=0.2, use_cuda=True, reg_param=0.01,
                          kl_param=1e-3, mmd_param=mmd_param, k=4, n_flows=n_flows).to(dev)
    src, rel, dst = rng.integers(0, n_ent, T), rng.integers(0, n_rel, T), rng.integers(0, n_ent, T)
    g, etype, node_norm = K.utils.build_graph_from_triplets(n_ent, n_rel, (src, rel, dst))
    trip = np.stack([rng.integers(0, n_ent, S), rng.integers(0, n_rel, S), rng.integers(0, n_ent, S)], 1)
    labels = (rng.random(S) < 0.2).astype(np.float32)
    eps = torch.from_numpy(rng.standard_normal((n_ent, h)).astype(np.float32))
    m1 = torch.from_numpy(((rng.random((n_ent, h)) < 0.8) / 0.8).astype(np.float32))
    m2 = torch.from_numpy(((rng.random((n_ent, 2 * h)) < 0.8) / 0.8).astype(np.float32))
    return model, g, etype, node_norm, trip, labels, eps, m1, m2, n_ent


def run(K, dev, rank, world, n_flows, mode, group=None, mmd_param=0.0, col_chunks=None):
    """Returns {"loss", "z", "grads": {name: rel err}} of this rank; raises AssertionError beyond 1e-4 / 2e-4.
    ``mmd_param`` > 0 adds the MMD term of the README configuration (kgvae/model.py:89-102): python's ``random``
    and the device generator are re-seeded before each of the two steps so that both draw the same 200 rows
    and the same prior noise."""
    import random
    from gcn_vae_b200 import parallel
    model, g, etype, node_norm, trip, labels, eps, m1, m2, N = setup(K, dev, n_flows, mmd_param)
    gtol = 2e-4 if mmd_param == 0 else 1e-3        # MMD gradients: a difference of three kernel means
    enc = model.encoder
    names = [n_ for n_, p in model.named_parameters() if p.requires_grad]
    params = [p for n_, p in model.named_parameters() if p.requires_grad]
    edge_norm = node_norm[g._dst].reshape(-1, 1).astype(np.float32)

    # ---- single-GPU reference step (every rank computes it; identical by construction) ----
    enc.preset_eps, enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = eps.to(dev), m1.to(dev), m2.to(dev)
    ids = torch.arange(N, device=dev).view(-1, 1)
    random.seed(7)
    torch.manual_seed(123)
    z = model(g, ids, torch.from_numpy(etype).to(dev), torch.from_numpy(edge_norm).to(dev))
    loss, pred, kl, _ = model.get_loss(g, z, torch.from_numpy(trip).to(dev), torch.from_numpy(labels).to(dev))
    loss.backward()
    want = {"loss": loss.detach().clone(), "z": z.detach().clone(), "grads": [p.grad.detach().clone() for p in params]}
    model.zero_grad(set_to_none=True)

    # ---- partitioned step ------------------------------------------------------------------
    peer = mode == "peer"
    parts = parallel.partition_by_destination(g._src, g._dst, etype, edge_norm, N, world, uniform=peer)
    mine = parts[rank]
    lo, hi = mine["lo"], mine["hi"]
    pg = K.Graph()
    pg.add_nodes(N)
    pg.add_edges(mine["src"], mine["dst"] - lo)
    pg.partition = parallel.Partition(lo, hi, N, group=group, peer_gather=peer, col_chunks=col_chunks)
    assert pg.partition.use_peer_gather(len(mine["src"])) == peer
    random.seed(7)
    torch.manual_seed(123)
    enc.preset_eps = eps[lo:hi].to(dev)
    enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = m1[lo:hi].to(dev), m2[lo:hi].to(dev)
    zl = model(pg, ids[lo:hi], torch.from_numpy(mine["etype"]).to(dev),
               torch.from_numpy(mine["norm"].reshape(-1, 1)).to(dev))
    s0, s1 = parallel.block_range(len(trip), rank, world)          # any split of the triplets works
    lp, _, _, _ = model.get_loss(pg, zl, torch.from_numpy(trip[s0:s1]).to(dev), torch.from_numpy(labels[s0:s1]).to(dev))
    lp.backward()
    parallel.allreduce_sum_grads(params, group=group)
    total = lp.detach().clone()
    dist.all_reduce(total, group=group)
    for cache in pg.partition._peer_rows.values():
        cache.close()

    res = {"loss": abs(float(total) - float(want["loss"])) / abs(float(want["loss"])),
           "z": float((zl.detach() - want["z"][lo:hi]).abs().max() / want["z"].abs().max()), "grads": {}}
    assert res["loss"] <= 1e-4, f"{mode} flows={n_flows}: summed loss off by {res['loss']:.2e}"
    assert res["z"] <= 1e-4, f"{mode} flows={n_flows}: owned rows of z off by {res['z']:.2e}"
    for name, p, ref in zip(names, params, want["grads"]):
        got = p.grad.detach()
        e = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))
        res["grads"][name] = e
        assert e <= gtol, f"{mode} flows={n_flows} mmd={mmd_param}: gradient of {name} off by {e:.2e}"
    return res


def run_entity(K, dev, rank, world, group=None):
    """Source-sharded / destination-partitioned entity classification (entity_classify.PartitionedEntityClassify)
    against the single-GPU EntityClassify on the same toy graph and parameters: owned rows of the logits, the
    summed loss, the owned rows of the basis-table gradient and every replicated gradient."""
    import torch.nn.functional as F
    from gcn_vae_b200 import entity_classify as EC
    data = EC.synthetic_graph("toy", seed=4)
    N, h, bases = data.num_nodes, 10, 4
    torch.manual_seed(0)
    full = EC.EntityClassify(N, h, data.num_classes, data.num_rels, num_bases=bases, num_hidden_layers=0, dropout=0.0,
                             use_self_loop=False, use_cuda=True).to(dev)
    g = K.Graph()
    g.add_nodes(N)
    g.add_edges(data.edge_src, data.edge_dst)
    et = torch.from_numpy(data.edge_type).to(dev)
    nm = torch.from_numpy(data.edge_norm).unsqueeze(1).to(dev)
    labels = torch.from_numpy(data.labels).to(dev)
    tr = torch.from_numpy(data.train_idx).to(dev)
    logits = full(g, full.create_features(), et, nm)
    loss = F.cross_entropy(logits[tr], labels[tr])
    loss.backward()
    from gcn_vae_b200 import parallel
    lo, hi = parallel.block_range(N, rank, world)
    torch.manual_seed(0)
    shard = EC.EntityClassify(hi - lo, h, data.num_classes, data.num_rels, num_bases=bases, num_hidden_layers=0,
                              dropout=0.0, use_self_loop=False, use_cuda=True).to(dev)
    sd = {k_: v.clone() for k_, v in full.state_dict().items()}
    sd["layers.0.weight"] = sd["layers.0.weight"][:, lo:hi, :].contiguous()
    shard.load_state_dict(sd)
    pe = EC.PartitionedEntityClassify(shard, data, rank, world, dev, group)
    lg = pe.logits()
    lp = pe.loss(lg)
    lp.backward()
    pe.reduce_grads()
    total = lp.detach().clone()
    dist.all_reduce(total, group=group)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
    res = {"loss": abs(float(total) - float(loss)) / abs(float(loss)), "logits": rel(lg.detach()[lo:hi], logits.detach()[lo:hi]),
           "grads": {}}
    fp, sp = dict(full.named_parameters()), dict(shard.named_parameters())
    for name, p in sp.items():
        want = fp[name].grad[:, lo:hi, :] if name == "layers.0.weight" else fp[name].grad
        res["grads"][name] = rel(p.grad, want)
    assert res["loss"] <= 1e-4 and res["logits"] <= 1e-4, res
    assert max(res["grads"].values()) <= 2e-4, res
    return res


def run_all(K, dev, rank, world, group=None):
    """All variants; returns a one-line summary (raises on the first mismatch)."""
    worst = 0.0
    for n_flows, mode, mmd, chunks in VARIANTS:
        r = run(K, dev, rank, world, n_flows, mode, group, mmd, chunks)
        worst = max(worst, r["loss"], r["z"], *r["grads"].values())
    e = run_entity(K, dev, rank, world, group)
    worst = max(worst, e["loss"], e["logits"], *e["grads"].values())
    return {"variants": len(VARIANTS) + 1, "world_size": world, "worst_rel_err": worst,
            "bars": "loss, z <= 1e-4; gradients <= 2e-4 (1e-3 with the MMD term), relative to the tensor's largest entry"}
