"""Destination-partitioned training on 2 GPUs against the single-GPU step
on the same graph, noise and dropout masks: loss, the owned rows of z and every gradient must
agree.  Needs 2 CUDA devices (skipped otherwise); NCCL over NVLink."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(K, dev, n_flows):
    rng = np.random.default_rng(0)
    n_ent, n_rel, h, bases, T, S = 211, 5, 40, 8, 900, 3000
    torch.manual_seed(0)
    model = K.LinkPredict(K.KGVAE, n_ent, h, n_rel, num_bases=bases, dropout=0.2, use_cuda=True, reg_param=0.01,
                          kl_param=1e-3, k=4, n_flows=n_flows).to(dev)
    src, rel, dst = rng.integers(0, n_ent, T), rng.integers(0, n_rel, T), rng.integers(0, n_ent, T)
    g, etype, node_norm = K.utils.build_graph_from_triplets(n_ent, n_rel, (src, rel, dst))
    trip = np.stack([rng.integers(0, n_ent, S), rng.integers(0, n_rel, S), rng.integers(0, n_ent, S)], 1)
    labels = (rng.random(S) < 0.2).astype(np.float32)
    eps = torch.from_numpy(rng.standard_normal((n_ent, h)).astype(np.float32))
    m1 = torch.from_numpy(((rng.random((n_ent, h)) < 0.8) / 0.8).astype(np.float32))
    m2 = torch.from_numpy(((rng.random((n_ent, 2 * h)) < 0.8) / 0.8).astype(np.float32))
    return model, g, etype, node_norm, trip, labels, eps, m1, m2, n_ent


def _worker(rank, port, n_flows, mode, out):
    import gcn_vae_b200 as K
    from gcn_vae_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        model, g, etype, node_norm, trip, labels, eps, m1, m2, N = _setup(K, dev, n_flows)
        enc = model.encoder
        params = [p for n_, p in model.named_parameters() if p.requires_grad]
        edge_norm = node_norm[g._dst].reshape(-1, 1).astype(np.float32)

        # ---- single-GPU reference step (every rank computes it; identical by construction) ----
        enc.preset_eps, enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = eps.to(dev), m1.to(dev), m2.to(dev)
        ids = torch.arange(N, device=dev).view(-1, 1)
        z = model(g, ids, torch.from_numpy(etype).to(dev), torch.from_numpy(edge_norm).to(dev))
        loss, pred, kl, _ = model.get_loss(g, z, torch.from_numpy(trip).to(dev), torch.from_numpy(labels).to(dev))
        loss.backward()
        want = {"loss": loss.detach().cpu(), "z": z.detach().cpu(),
                "grads": [p.grad.detach().cpu().clone() for p in params]}
        model.zero_grad(set_to_none=True)

        # ---- partitioned step ------------------------------------------------------------------
        peer = mode == "peer"
        parts = parallel.partition_by_destination(g._src, g._dst, etype, edge_norm, N, WORLD, uniform=peer)
        mine = parts[rank]
        lo, hi = mine["lo"], mine["hi"]
        pg = K.Graph()
        pg.add_nodes(N)
        pg.add_edges(mine["src"], mine["dst"] - lo)
        pg.partition = parallel.Partition(lo, hi, N, peer_gather=peer)
        assert pg.partition.use_peer_gather(len(mine["src"])) == peer
        enc.preset_eps = eps[lo:hi].to(dev)
        enc.rconv_layer_1.dropout_mask, enc.rconv_layer_2.dropout_mask = m1[lo:hi].to(dev), m2[lo:hi].to(dev)
        zl = model(pg, ids[lo:hi], torch.from_numpy(mine["etype"]).to(dev),
                   torch.from_numpy(mine["norm"].reshape(-1, 1)).to(dev))
        s0, s1 = parallel.block_range(len(trip), rank, WORLD)          # any split of the triplets works
        lp, _, _, _ = model.get_loss(pg, zl, torch.from_numpy(trip[s0:s1]).to(dev),
                                     torch.from_numpy(labels[s0:s1]).to(dev))
        lp.backward()
        parallel.allreduce_sum_grads(params)
        total = lp.detach().clone()
        dist.all_reduce(total)
        out[rank] = {"want": want, "loss": total.cpu(), "z": zl.detach().cpu(), "lo": lo, "hi": hi,
                     "grads": [p.grad.detach().cpu().clone() for p in params],
                     "names": [n_ for n_, p in model.named_parameters() if p.requires_grad]}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_flows,mode", [(0, "allgather"), (1, "allgather"), (0, "peer"), (1, "peer")])
def test_partitioned_step_matches_single_gpu(n_flows, mode):
    """mode "allgather": NCCL all-gather of every layer input; mode "peer": the message-passing kernels
    gather source rows from the owners' HBM over NVLink (CUDA IPC row blocks), reduce-scatter backward."""
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(port, n_flows, mode, out), nprocs=WORLD, join=True)
        res = [out[r] for r in range(WORLD)]
    for r in range(WORLD):
        w = res[r]["want"]
        assert abs(float(res[r]["loss"]) - float(w["loss"])) <= 1e-4 * abs(float(w["loss"]))
        zl, lo, hi = res[r]["z"], res[r]["lo"], res[r]["hi"]
        assert (zl - w["z"][lo:hi]).abs().max() <= 1e-4 * w["z"].abs().max()
        for name, got, ref in zip(res[r]["names"], res[r]["grads"], w["grads"]):
            scale = ref.abs().max().clamp_min(1e-12)
            assert (got - ref).abs().max() <= 2e-4 * scale, name
