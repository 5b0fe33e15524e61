"""Multi-GPU plumbing for the link-prediction path.  One process per GPU,
``torch.distributed`` (NCCL on the GPUs; the host logic below is backend-agnostic and is tested
with gloo, world_size 2, on CPU tensors).

The reference is single-process (kgvae/link_predict.py:113-115); everything here is new:

* small graphs (FB15k-237 / WN18 shapes): **replicas** - each rank samples its own subgraph with
  the reference's sampler, gradients are averaged with one all-reduce of the flattened gradient;
* evaluation: **entity-sharded** - every rank scores all queries against its slice of the
  candidates and the per-shard rank counts are summed;
* large graphs (wikikg2 shape): **destination-node ownership** - rank p owns a contiguous block
  of nodes and every edge whose destination is in it; layer inputs are all-gathered.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def block_range(n, rank, world_size):
    """Contiguous block [lo, hi) of ``n`` items owned by ``rank``; sizes differ by at most one."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


entity_shard = block_range


def allreduce_mean_grads(params, group=None):
    """Average ``p.grad`` over the group with ONE all-reduce of the flattened gradient (replicas).
    Parameters without a gradient contribute zeros so that every rank sends the same layout."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    ws = dist.get_world_size(group)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= ws
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def sharded_rank_counts(count_fn, num_entities, group=None):
    """Entity-sharded evaluation: ``count_fn(lo, hi)`` returns, for every query, the number of
    candidates in [lo, hi) ranked ahead of the target (an integer tensor); the shards' counts add
    up to the 0-indexed rank.  The sum is exact because every (query, candidate) decision is made
    on the same score on whichever rank holds the candidate."""
    rank, ws = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    lo, hi = block_range(num_entities, rank, ws)
    counts = count_fn(lo, hi)
    if ws > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def partition_by_destination(src, dst, etype, norm, num_nodes, world_size):
    """Destination-node ownership: returns a list (one entry per rank) of dicts with the owned node
    block ``[lo, hi)``, the rank's edges (every edge with dst in the block, original relative
    order kept - so per-destination summation order equals the single-GPU order) and
    ``needed_src``: the sorted unique source ids the rank must receive features for."""
    src, dst, etype = (np.asarray(a) for a in (src, dst, etype))
    norm = None if norm is None else np.asarray(norm).reshape(-1)
    # owner of an edge: the rank p with lo_p <= dst < hi_p
    bounds = np.array([block_range(num_nodes, p, world_size)[0] for p in range(world_size)] + [num_nodes])
    owner = np.searchsorted(bounds, dst, side="right") - 1
    parts = []
    for p in range(world_size):
        sel = np.nonzero(owner == p)[0]
        lo, hi = int(bounds[p]), int(bounds[p + 1])
        parts.append({"lo": lo, "hi": hi, "edge_ids": sel, "src": src[sel], "dst": dst[sel], "etype": etype[sel],
                      "norm": None if norm is None else norm[sel], "needed_src": np.unique(src[sel])})
    return parts


def allgather_rows(local_rows, num_rows, group=None):
    """All-gather of node features for destination-partitioned message passing: every rank holds
    rows [lo, hi) of a [num_rows, d] matrix and receives the full matrix.  Blocks may differ by one
    row, so the gather is padded to the largest block."""
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    sizes = [block_range(num_rows, p, ws)[1] - block_range(num_rows, p, ws)[0] for p in range(ws)]
    pad = max(sizes)
    buf = local_rows.new_zeros((pad,) + tuple(local_rows.shape[1:]))
    buf[:local_rows.shape[0]] = local_rows
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


# ---------------------------------------------------------------------------------------------
# destination-partitioned training: differentiable collectives and the partition descriptor
# ---------------------------------------------------------------------------------------------
class Partition:
    """Rank-local view of a destination-partitioned graph: this rank owns nodes [lo, hi) of
    ``n_global`` and every edge whose destination is in that block (edge sources keep global ids,
    destinations are local: dst - lo).  Attach to the graph object as ``g.partition``; RelGraphConv,
    KGVAE and LinkPredict then insert the all-gathers / reductions below."""

    def __init__(self, lo, hi, n_global, group=None):
        self.lo, self.hi, self.n_global, self.group = int(lo), int(hi), int(n_global), group
        self.n_local = self.hi - self.lo
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)


class AllGatherRowsFn(torch.autograd.Function):
    """x_local [n_local, d] -> x_full [n_global, d]; backward: reduce-scatter(sum) of the gradient -
    every rank's loss depends on every row it gathered."""

    @staticmethod
    def forward(ctx, x_local, part):
        ctx.part = part
        return allgather_rows(x_local.contiguous(), part.n_global, part.group)

    @staticmethod
    def backward(ctx, g_full):
        part = ctx.part
        g_full = g_full.contiguous()
        dist.all_reduce(g_full, op=dist.ReduceOp.SUM, group=part.group)   # blocks differ by a row: sum, then slice
        return g_full[part.lo:part.hi].clone(), None


class AllReduceSumFn(torch.autograd.Function):
    """sum over ranks of a tensor every rank then uses in its own loss; backward sums the upstream
    gradients of all ranks."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def allreduce_sum_grads(params, group=None):
    """Partitioned training: every rank holds a partial gradient of the replicated parameters;
    the total is their SUM (one flattened all-reduce)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
