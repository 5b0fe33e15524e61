"""Multi-process host logic on CPU: gloo backend, world_size 2.

Covers what does not need a GPU: the replica gradient all-reduce, entity-sharded rank counts
adding up to the single-process ranks (checked with the oracle's scores), destination-ownership
partitioning and the padded all-gather of node features."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O

from gcn_vae_b200 import parallel

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        res = {}
        # --- replicas: one flattened all-reduce averages the gradients ------------------------
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)),
                  torch.nn.Parameter(torch.randn(2, 2))]
        g = torch.Generator().manual_seed(100 + rank)
        params[0].grad = torch.randn(5, 3, generator=g)
        params[1].grad = torch.randn(7, generator=g)          # params[2] has no grad on any rank
        parallel.allreduce_mean_grads(params)
        res["grads"] = [p.grad.clone() for p in params]

        # --- replicas, bucketed: gradients are views of one flat buffer, buckets all-reduced from hooks ------
        torch.manual_seed(1)
        bp = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(6)), torch.nn.Parameter(torch.randn(2))]
        buckets = parallel.GradBuckets(bp, buckets=[[bp[1]], [bp[0]]])      # bp[2] lands in the trailing bucket
        for it in range(2):
            buckets.zero()
            if it == 1:                  # a caller that forgets zero() semantics and drops a gradient view
                bp[0].grad = None
            scale = float(rank + 1 + it)
            (bp[0].sum() * scale + (bp[1] * bp[1]).sum() * scale).backward()       # bp[2] unused: zeros
            buckets.finish()
        norm_before = float(buckets.flat.norm())
        buckets.clip_(0.5)
        res["bucket_grads"] = [p.grad.clone() for p in bp]
        res["bucket_views"] = all(p.grad.data_ptr() >= buckets.flat.data_ptr() and
                                  p.grad.data_ptr() < buckets.flat.data_ptr() + 4 * buckets.flat.numel() for p in bp)
        res["bucket_norms"] = (norm_before, float(buckets.flat.norm()), [x.detach().clone() for x in bp])

        # --- entity-sharded evaluation: shard counts add up ----------------------------------
        rng = np.random.default_rng(5)
        V, h, M = 101, 16, 37
        emb = torch.from_numpy(rng.integers(-4, 5, size=(V, h)).astype(np.float32) / 4)
        w = torch.from_numpy(rng.integers(-4, 5, size=(3, h)).astype(np.float32) / 4)
        a, r, b = (torch.from_numpy(rng.integers(0, n, M)) for n in (V, 3, V))
        score = O.eval_scores(emb, w, a, r)
        st = score.gather(1, b.view(-1, 1))
        col = torch.arange(V).view(1, -1)

        def count(lo, hi):
            ahead = (score > st) | ((score == st) & (col < b.view(-1, 1)))
            return ahead[:, lo:hi].sum(1).to(torch.int32)

        res["ranks"] = parallel.sharded_rank_counts(count, V)
        res["want_ranks"] = O.rank_of_target(score, b)

        # --- destination ownership + all-gather of the owned feature rows ---------------------
        N, E = 23, 200
        src, dst, et = rng.integers(0, N, E), rng.integers(0, N, E), rng.integers(0, 4, E)
        parts = parallel.partition_by_destination(src, dst, et, None, N, WORLD)
        mine = parts[rank]
        feats = torch.arange(N * 3, dtype=torch.float32).view(N, 3)
        full = parallel.allgather_rows(feats[mine["lo"]:mine["hi"]], N)
        res["gathered_ok"] = bool(torch.equal(full, feats))
        res["parts"] = [(p["lo"], p["hi"], p["edge_ids"]) for p in parts]
        res["edges"] = (src, dst)

        # --- differentiable collectives of the partitioned step ---------------------------------
        part = parallel.Partition(mine["lo"], mine["hi"], N)
        xl = (feats[mine["lo"]:mine["hi"]] * 0.1).clone().requires_grad_(True)
        wgt = torch.arange(N * 3, dtype=torch.float32).view(N, 3) * (rank + 1)      # rank-specific loss
        full2 = parallel.AllGatherRowsFn.apply(xl, part)
        s_loc = xl.sum()
        s_all = parallel.AllReduceSumFn.apply(s_loc, None)
        ((full2 * wgt).sum() + s_all * (rank + 2.0)).backward()
        res["gather_grad"] = xl.grad.clone()
        res["block"] = (mine["lo"], mine["hi"])

        # --- uniform blocks (what the peer-memory gather needs): ceil(N / P) rows each --------------
        uparts = parallel.partition_by_destination(src, dst, et, None, N, WORLD, uniform=True)
        um = uparts[rank]
        upart = parallel.Partition(um["lo"], um["hi"], N)
        # the gather / all-gather choice is COLLECTIVE: ranks with very different edge counts (and peer_gather
        # left to None) must come out with the same answer - average edges per rank < nodes -> peer gather
        few = parallel.Partition(um["lo"], um["hi"], N)
        many = parallel.Partition(um["lo"], um["hi"], N)
        d_few = few.use_peer_gather(5 if rank == 0 else 40)          # 45 < 2 * 23
        d_many = many.use_peer_gather(5 if rank == 0 else 60)        # 65 >= 46
        d_cached = few.use_peer_gather(10 ** 6)                      # cached: no second collective, same answer
        # blocks are uniform only if EVERY rank's block is: rank 1 hands over a shorter block here
        ragged = parallel.Partition(um["lo"], um["hi"] - (1 if rank == 1 else 0), N)
        res["collective"] = (d_few, d_many, d_cached, ragged.blk, ragged.use_peer_gather(1))
        res["uniform"] = (um["lo"], um["hi"], upart.blk, not part.use_peer_gather(1), part.blk)
        ufull = parallel.allgather_rows(feats[um["lo"]:um["hi"]], N, uniform=True)
        res["uniform_gathered_ok"] = bool(torch.equal(ufull, feats))
        contrib = torch.arange(WORLD * upart.blk * 3, dtype=torch.float32).view(-1, 3) * (rank + 1)
        res["rs_rows"] = parallel.reduce_scatter_rows(contrib.clone(), upart)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(port, out), nprocs=WORLD, join=True)
        res = [out[r] for r in range(WORLD)]
    # gradients: identical on both ranks and equal to the mean of the per-rank gradients
    gens = [torch.Generator().manual_seed(100 + r) for r in range(WORLD)]
    g0 = torch.stack([torch.randn(5, 3, generator=g) for g in gens]).mean(0)
    g1 = torch.stack([torch.randn(7, generator=g) for g in gens]).mean(0)
    for r in range(WORLD):
        assert torch.allclose(res[r]["grads"][0], g0) and torch.allclose(res[r]["grads"][1], g1)
        assert torch.equal(res[r]["grads"][2], torch.zeros(2, 2))
        assert torch.equal(res[r]["ranks"].long(), res[r]["want_ranks"])
        assert res[r]["gathered_ok"]
    # bucketed replicas: second iteration's scales are rank + 2 -> mean 2.5 over the two ranks
    for r in range(WORLD):
        nb, na, bp = res[r]["bucket_norms"]
        want = [torch.full((4, 3), 2.5), 2 * bp[1] * 2.5, torch.zeros(2)]
        coef = min(1.0, 0.5 / (nb + 1e-6))
        for got, w_ in zip(res[r]["bucket_grads"], want):
            assert torch.allclose(got, w_ * coef, rtol=1e-5, atol=1e-6)
        assert res[r]["bucket_views"] and abs(na - min(nb, 0.5)) < 1e-4
    # all-gather backward = sum over ranks of the gradient rows; all-reduce backward = sum of upstream grads
    wsum = torch.arange(23 * 3, dtype=torch.float32).view(23, 3) * sum(r + 1 for r in range(WORLD))
    for r in range(WORLD):
        lo, hi = res[r]["block"]
        want = wsum[lo:hi] + sum(q + 2.0 for q in range(WORLD))
        assert torch.allclose(res[r]["gather_grad"], want)
    # uniform blocks: [0, 12), [12, 23) for N = 23; owner(row) = row // 12; reduce-scatter keeps own rows
    total = torch.arange(WORLD * 12 * 3, dtype=torch.float32).view(-1, 3) * sum(r + 1 for r in range(WORLD))
    for r in range(WORLD):
        lo, hi, blk, peer_ok, blk_nonuniform = res[r]["uniform"]
        assert (lo, hi, blk) == ((0, 12, 12) if r == 0 else (12, 23, 12)) and peer_ok and blk_nonuniform is None
        assert res[r]["collective"] == (True, False, True, None, False), res[r]["collective"]
        assert res[r]["uniform_gathered_ok"]
        assert torch.equal(res[r]["rs_rows"], total[lo:hi])
    # partition: every edge exactly once, owned by the rank whose node block holds its destination
    src, dst = res[0]["edges"]
    seen = np.concatenate([ids for _, _, ids in res[0]["parts"]])
    assert np.array_equal(np.sort(seen), np.arange(len(dst)))
    for lo, hi, ids in res[0]["parts"]:
        assert ((dst[ids] >= lo) & (dst[ids] < hi)).all()
        assert np.all(np.diff(ids) > 0)           # original relative order kept
    assert res[0]["parts"][0][0] == 0 and res[0]["parts"][-1][1] == 23


def test_block_range_covers_everything():
    for n in (0, 1, 7, 14541):
        for ws in (1, 2, 3, 8):
            blocks = [parallel.block_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
