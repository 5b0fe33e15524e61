"""Diagnostic (2 GPUs): peer row blocks - torch copy, LDG kernel and TMA bulk-copy kernel on peer memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, torch.multiprocessing as mp


def worker(rank, ws):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", CUDA_LAUNCH_BLOCKING="1")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    import gcn_vae_b200 as K
    from gcn_vae_b200 import parallel, ops, _lib as L
    blk, width = 64, 500
    pr = parallel.PeerRows(blk, width, dev)
    x = torch.full((blk, width), float(rank + 1), device=dev) + torch.arange(blk, device=dev).view(-1, 1) * 0.001
    pr.publish(x)
    torch.cuda.synchronize()
    peer = 1 - rank
    print(rank, "views devices", [str(v.device) for v in pr.views], "ptrs", [hex(p) for p in pr.ptrs.tolist()], flush=True)
    # 1. torch copy from the peer view
    got = pr.views[peer].to(dev)
    torch.cuda.synchronize()
    print(rank, "torch copy ok:", float(got[0, 0]), float(got[5, 3]), flush=True)
    # 2. LDG kernel: embedding lookup with the peer block as the table
    ids = torch.arange(blk, device=dev, dtype=torch.int32)
    out = torch.empty((blk, width), device=dev)
    L.call("kg_embedding_fwd", pr.views[peer].data_ptr(), L.i32(ids), blk, width, L.f32(out), L.stream())
    torch.cuda.synchronize()
    print(rank, "LDG kernel on peer memory ok:", float(out[0, 0]), float(out[7, 1]), flush=True)
    # 3. bulk-copy kernel: bdd forward, identity-ish weights, sources on both ranks
    B, si, so, R, E = 100, 5, 5, 3, 200
    g = torch.Generator(device=dev).manual_seed(5)
    src = torch.randint(0, ws * blk, (E,), device=dev, dtype=torch.int32, generator=g)
    dst = torch.randint(0, blk, (E,), device=dev, dtype=torch.int32, generator=g)
    et = torch.randint(0, R, (E,), device=dev, dtype=torch.int32, generator=g)
    norm = torch.ones(E, device=dev)
    gi = ops.graph_index(src, dst, et, norm, ws * blk, R)
    w = torch.randn(R, B * si * so, device=dev, generator=g)
    agg = torch.zeros(blk, B * so, device=dev)
    L.call("kg_bdd_rel_fwd", None, L.ptr(pr.ptrs), blk, L.i32(gi.rel_pack), E, L.f32(w), None, B, si, so, L.f32(agg), 0, L.stream())
    torch.cuda.synchronize()
    full = torch.cat([pr.views[r].to(dev) for r in range(ws)])
    ref = torch.zeros_like(agg)
    msg = torch.matmul(full[src.long()].view(E, B, 1, si), w[et.long()].view(E, B, si, so)).view(E, B * so)
    ref.index_add_(0, dst.long(), msg)
    print(rank, "bulk-copy kernel on peer memory: max err", float((agg - ref).abs().max()), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    mp.spawn(worker, args=(2,), nprocs=2, join=True)
