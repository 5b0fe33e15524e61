"""Entity classification with the basis-regularised RGCN (kgvae/entity_classify.py; config 4).

``EntityClassify`` keeps the reference's constructor (through ``BaseRGCN``) and layer stack:
an input layer on integer node ids (``RelGraphConv(num_nodes, h, "basis")``), hidden layers,
and a softmax output layer.  ``synthetic_graph`` stands in for ``dgl.contrib.data.load_data``
(rdflib datasets are not available offline): it draws a typed multigraph of a named shape with
the loader's edge norm (1 / number of same-type edges into the destination) and random labels.
"""
import argparse
import os
import time
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from .model import BaseRGCN
from .nn import RelGraphConv

# name -> (nodes, relation types, directed edges, classes)        (AM: baselines/rgcn/README.md:35-38)
SHAPES = {"am": (1666764, 133, 5988321, 11), "aifb": (8285, 91, 58086, 4), "toy": (300, 7, 2500, 3)}


class EntityClassify(BaseRGCN):
    def create_features(self):
        features = torch.arange(self.num_nodes)
        return features.cuda() if self.use_cuda else features

    def build_input_layer(self):
        return RelGraphConv(self.num_nodes, self.h_dim, self.num_rels, "basis", self.num_bases,
                            activation=F.relu, self_loop=self.use_self_loop, dropout=self.dropout)

    def build_hidden_layer(self, idx):
        return RelGraphConv(self.h_dim, self.h_dim, self.num_rels, "basis", self.num_bases,
                            activation=F.relu, self_loop=self.use_self_loop, dropout=self.dropout)

    def build_output_layer(self):
        return RelGraphConv(self.h_dim, self.out_dim, self.num_rels, "basis", self.num_bases,
                            activation=partial(F.softmax, dim=1), self_loop=self.use_self_loop)


def synthetic_graph(name="am", seed=0, scale=1.0):
    n_nodes, n_rels, n_edges, n_classes = SHAPES[name]
    n_edges = max(1, int(n_edges * scale))
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_nodes, n_edges)
    dst = rng.integers(0, n_nodes, n_edges)
    etype = rng.integers(0, n_rels, n_edges)
    # loader's norm: 1 / |{edges into dst with the same type}|
    key = dst.astype(np.int64) * n_rels + etype
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    norm = (1.0 / cnt[inv]).astype(np.float32)
    labels = rng.integers(0, n_classes, n_nodes)
    idx = rng.permutation(n_nodes)
    n_train = max(1, n_nodes // 10)
    return SimpleNamespace(num_nodes=n_nodes, num_rels=n_rels, num_classes=n_classes, edge_src=src, edge_dst=dst,
                           edge_type=etype, edge_norm=norm, labels=labels, train_idx=idx[:n_train],
                           test_idx=idx[n_train:2 * n_train])


# --------------------------------------------------------------------------------------------
# multi-GPU form (new; the reference is single-process, kgvae/entity_classify.py:66-72)
# --------------------------------------------------------------------------------------------
class PartitionedEntityClassify:
    """Entity classification over the GPUs of one box (BASELINE.json configs[3]: "1/2/4/8 B200").

    The integer-id input layer is an embedding-style lookup into the basis table V [B, N, h] (2.67 GB at the AM
    shape) - so it is sharded by SOURCE-row ownership: rank p holds V[:, lo_p:hi_p, :] (and its
    Adam state) and every edge whose source it owns, computes that share of the messages for all destinations
    (h = 10: a 67 MB matrix) and the partial sums are all-reduced; its backward all-reduces the gradient of that
    matrix once, after which every row of dV is produced by its owner without any communication.  The output
    layer (dense basis conv h -> classes, softmax) is destination-partitioned on the all-reduced hidden state.
    Replicated parameters (coefficients, output-layer basis, biases) receive partial gradients: SUM over ranks.

    ``model``: an EntityClassify built with ``num_nodes = hi - lo`` (the table shard), two layers."""

    def __init__(self, model, data, rank, world, device, group=None):
        from . import parallel
        from .graph import Graph
        if len(model.layers) != 2:
            raise RuntimeError("partitioned entity classification covers the two-layer model (the reference's AM flags)")
        self.model, self.rank, self.world, self.group = model, rank, world, group
        self.n_global = int(data.num_nodes)
        self.lo, self.hi = parallel.block_range(self.n_global, rank, world)
        src, dst = np.asarray(data.edge_src), np.asarray(data.edge_dst)
        et, nm = np.asarray(data.edge_type), np.asarray(data.edge_norm, dtype=np.float32)
        own_src = (src >= self.lo) & (src < self.hi)
        own_dst = (dst >= self.lo) & (dst < self.hi)
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
        self.g_src = Graph()                 # edges by source owner: local source ids, global destinations
        self.g_src._n = self.n_global
        self.g_src._dev_edges[torch.device(device)] = (to(src[own_src] - self.lo, torch.int32), to(dst[own_src], torch.int32))
        self.et_src, self.nm_src = to(et[own_src], torch.int32), to(nm[own_src].reshape(-1, 1), torch.float32)
        self.g_dst = Graph()                 # edges by destination owner: global ids on both ends
        self.g_dst._n = self.n_global
        self.g_dst._dev_edges[torch.device(device)] = (to(src[own_dst], torch.int32), to(dst[own_dst], torch.int32))
        self.et_dst, self.nm_dst = to(et[own_dst], torch.int32), to(nm[own_dst].reshape(-1, 1), torch.float32)
        tr = np.asarray(data.train_idx)
        self.n_train = len(tr)
        self.train_own = to(tr[(tr >= self.lo) & (tr < self.hi)], torch.int64)
        self.labels = to(np.asarray(data.labels), torch.int64)
        self.sharded = [model.layers[0].weight]
        self.replicated = [p for p in model.parameters() if p.requires_grad and p is not model.layers[0].weight]

    def logits(self):
        from . import basis, parallel
        l1, l2 = self.model.layers[0], self.model.layers[1]
        if l1.self_loop:
            raise RuntimeError("partitioned entity classification: use_self_loop is not supported")
        gi = self.g_src.index_for(self.et_src, self.nm_src, l1.num_rels, node_major=True)
        part = basis.BasisIdSrcPartialFn.apply(l1.weight, l1.w_comp, gi, self.n_global)
        agg = parallel.AllReduceSumFn.apply(part, self.group)
        if l1.bias:          # added after the sum: the gradient reaching it is each rank's PARTIAL one, so SUM is right
            agg = agg + l1.h_bias
        h1 = F.relu(agg)
        return l2(self.g_dst, h1, self.et_dst, self.nm_dst)

    def loss(self, logits):
        """This rank's share of F.cross_entropy(logits[train_idx], labels[train_idx]) (entity_classify.py:110):
        the sum over its own training nodes divided by the GLOBAL count - the shares add up to the reference's loss."""
        idx = self.train_own
        if idx.numel() == 0:
            return logits.sum() * 0.0
        return F.cross_entropy(logits[idx], self.labels[idx], reduction="sum") / self.n_train

    def reduce_grads(self):
        from . import parallel
        if self.world > 1:
            parallel.allreduce_sum_grads(self.replicated, group=self.group)


# --------------------------------------------------------------------------------------------
# on-disk typed graphs (stands in for dgl.contrib.data.load_data, kgvae/entity_classify.py:47)
# --------------------------------------------------------------------------------------------
def _bfs_keep_edges(src, dst, seeds, num_nodes, levels):
    """Edges that can carry information to a labelled node within ``levels`` message-passing rounds: round 1
    keeps every edge INTO a seed, round i every edge into a source reached in round i - 1 (the loader's
    ``bfs_level`` pruning, kgvae/entity_classify.py:169: "pruning used nodes for memory")."""
    keep = np.zeros(len(src), dtype=bool)
    frontier = np.zeros(num_nodes, dtype=bool)
    frontier[np.asarray(seeds, dtype=np.int64)] = True
    seen = frontier.copy()
    for _ in range(int(levels)):
        hit = frontier[dst] & ~keep
        keep |= hit
        nxt = np.zeros(num_nodes, dtype=bool)
        nxt[src[hit]] = True
        frontier = nxt & ~seen
        seen |= nxt
        if not frontier.any():
            break
    return keep


def load_data(dataset, bfs_level=3, relabel=False):
    """Typed-graph dataset with the fields ``main`` reads (kgvae/entity_classify.py:47-56): ``num_nodes``,
    ``num_rels``, ``num_classes``, ``edge_src``, ``edge_dst``, ``edge_type``, ``edge_norm``, ``labels``,
    ``train_idx``, ``test_idx``.

    ``dataset``: a synthetic shape name (``am``, ``aifb``, ``toy``, optionally ``name:seed``) or a directory with
    ``edges.tsv`` (subject, relation, object per line; names or integers), ``trainingSet.tsv`` and ``testSet.tsv``
    (node, class label).  The directory form follows what DGL's RDF loader produces from the same triples
    [DGL, not in /root/reference; restated]: every triple (s, p, o) becomes the edges s -> o with type 2p and
    o -> s with type 2p + 1, every node gets a self-loop of type 2P, edges are sorted by (dst, src, type),
    ``edge_norm`` = 1 / (number of edges of the same type into the destination); ``bfs_level`` > 0 drops the edges
    that cannot reach a labelled node in that many rounds, ``relabel`` additionally drops untouched nodes."""
    if not os.path.isdir(dataset):
        parts = dataset.split(":")
        if parts[0] not in SHAPES:
            raise ValueError(f"unknown dataset {dataset!r}: a directory or one of {sorted(SHAPES)}")
        return synthetic_graph(parts[0], seed=int(parts[1]) if len(parts) > 1 else 0)
    ent, rel, cls = {}, {}, {}
    rows = []
    with open(os.path.join(dataset, "edges.tsv")) as f:
        for line in f:
            if not line.strip():
                continue
            s_, p_, o_ = line.rstrip("\n").split("\t") if "\t" in line else line.split()
            rows.append((ent.setdefault(s_, len(ent)), rel.setdefault(p_, len(rel)), ent.setdefault(o_, len(ent))))

    def read_labels(name):
        ids, labs = [], []
        path = os.path.join(dataset, name)
        if os.path.exists(path):
            with open(path) as f:
                for line in f:
                    if not line.strip():
                        continue
                    n_, c_ = line.rstrip("\n").split("\t") if "\t" in line else line.split()
                    ids.append(ent.setdefault(n_, len(ent)))
                    labs.append(cls.setdefault(c_, len(cls)))
        return np.asarray(ids, dtype=np.int64), np.asarray(labs, dtype=np.int64)

    train_idx, train_lab = read_labels("trainingSet.tsv")
    test_idx, test_lab = read_labels("testSet.tsv")
    t = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
    n, P = len(ent), len(rel)
    loops = np.arange(n, dtype=np.int64)
    src = np.concatenate([t[:, 0], t[:, 2], loops])
    dst = np.concatenate([t[:, 2], t[:, 0], loops])
    typ = np.concatenate([2 * t[:, 1], 2 * t[:, 1] + 1, np.full(n, 2 * P, dtype=np.int64)])
    labels = np.zeros(n, dtype=np.int64)
    labels[train_idx], labels[test_idx] = train_lab, test_lab
    if bfs_level and bfs_level > 0 and (len(train_idx) + len(test_idx)) > 0:
        keep = _bfs_keep_edges(src, dst, np.concatenate([train_idx, test_idx]), n, bfs_level)
        src, dst, typ = src[keep], dst[keep], typ[keep]
    if relabel:
        used = np.unique(np.concatenate([src, dst, train_idx, test_idx]))
        remap = np.full(n, -1, dtype=np.int64)
        remap[used] = np.arange(len(used))
        src, dst, train_idx, test_idx = remap[src], remap[dst], remap[train_idx], remap[test_idx]
        labels, n = labels[used], len(used)
    order = np.lexsort((typ, src, dst))
    src, dst, typ = src[order], dst[order], typ[order]
    key = dst * (2 * P + 1) + typ
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    norm = (1.0 / cnt[inv]).astype(np.float32)
    return SimpleNamespace(num_nodes=n, num_rels=2 * P + 1, num_classes=max(len(cls), 1), edge_src=src, edge_dst=dst,
                           edge_type=typ, edge_norm=norm, labels=labels, train_idx=train_idx, test_idx=test_idx)


# --------------------------------------------------------------------------------------------
# driver (kgvae/entity_classify.py:45-135; flags :138-170)
# --------------------------------------------------------------------------------------------
def main(args):
    from .graph import Graph
    data = load_data(args.dataset, bfs_level=args.bfs_level, relabel=args.relabel)
    num_nodes, num_rels, num_classes = data.num_nodes, data.num_rels, data.num_classes
    train_idx, test_idx = data.train_idx, data.test_idx
    if args.validation:
        val_idx = train_idx[:len(train_idx) // 5]
        train_idx = train_idx[len(train_idx) // 5:]
    else:
        val_idx = train_idx
    if not torch.cuda.is_available():
        raise RuntimeError("kgvae_b200.entity_classify runs on CUDA only (no CPU fallback)")
    device = torch.device("cuda", max(args.gpu, 0))
    torch.cuda.set_device(device)
    edge_type = torch.from_numpy(np.asarray(data.edge_type)).to(device)
    edge_norm = torch.from_numpy(np.asarray(data.edge_norm)).unsqueeze(1).to(device)
    labels = torch.from_numpy(np.asarray(data.labels)).view(-1).to(device)
    train_t, val_t, test_t = (torch.from_numpy(np.asarray(i, dtype=np.int64)).to(device) for i in (train_idx, val_idx, test_idx))
    g = Graph()
    g.add_nodes(num_nodes)
    g.add_edges(data.edge_src, data.edge_dst)
    model = EntityClassify(len(g), args.n_hidden, num_classes, num_rels, num_bases=args.n_bases,
                           num_hidden_layers=args.n_layers - 2, dropout=args.dropout,
                           use_self_loop=args.use_self_loop, use_cuda=True).to(device)
    feats = model.create_features()
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.l2norm, fused=True)
    print("start training...")
    forward_time, backward_time = [], []
    model.train()
    for epoch in range(args.n_epochs):
        optimizer.zero_grad()
        torch.cuda.synchronize()
        t0 = time.time()
        logits = model(g, feats, edge_type, edge_norm)
        loss = F.cross_entropy(logits[train_t], labels[train_t])
        torch.cuda.synchronize()
        t1 = time.time()
        loss.backward()
        optimizer.step()
        torch.cuda.synchronize()
        t2 = time.time()
        forward_time.append(t1 - t0)
        backward_time.append(t2 - t1)
        print("Epoch {:05d} | Train Forward Time(s) {:.4f} | Backward Time(s) {:.4f}".format(
            epoch, forward_time[-1], backward_time[-1]))
        with torch.no_grad():
            train_acc = (logits[train_t].argmax(dim=1) == labels[train_t]).float().mean().item()
            val_loss = F.cross_entropy(logits[val_t], labels[val_t])
            val_acc = (logits[val_t].argmax(dim=1) == labels[val_t]).float().mean().item()
        print("Train Accuracy: {:.4f} | Train Loss: {:.4f} | Validation Accuracy: {:.4f} | Validation loss: {:.4f}".format(
            train_acc, loss.item(), val_acc, val_loss.item()))
    print()
    model.eval()
    with torch.no_grad():
        logits = model.forward(g, feats, edge_type, edge_norm)
        test_loss = F.cross_entropy(logits[test_t], labels[test_t])
        test_acc = (logits[test_t].argmax(dim=1) == labels[test_t]).float().mean().item()
    print("Test Accuracy: {:.4f} | Test loss: {:.4f}".format(test_acc, test_loss.item()))
    print()
    print("Mean forward time: {:4f}".format(np.mean(forward_time[len(forward_time) // 4:])))
    print("Mean backward time: {:4f}".format(np.mean(backward_time[len(backward_time) // 4:])))
    return {"test_acc": test_acc, "test_loss": test_loss.item(), "train_loss": loss.item()}


def build_parser():
    parser = argparse.ArgumentParser(description="RGCN")
    parser.add_argument("--dropout", type=float, default=0)
    parser.add_argument("--n-hidden", type=int, default=16)
    parser.add_argument("--gpu", type=int, default=-1)
    parser.add_argument("--lr", type=float, default=1e-2)
    parser.add_argument("--n-bases", type=int, default=-1)
    parser.add_argument("--n-layers", type=int, default=2)
    parser.add_argument("-e", "--n-epochs", type=int, default=50)
    parser.add_argument("-d", "--dataset", type=str, required=True)
    parser.add_argument("--l2norm", type=float, default=0)
    parser.add_argument("--relabel", default=False, action="store_true")
    parser.add_argument("--use-self-loop", default=False, action="store_true")
    fp = parser.add_mutually_exclusive_group(required=False)
    fp.add_argument("--validation", dest="validation", action="store_true")
    fp.add_argument("--testing", dest="validation", action="store_false")
    parser.set_defaults(validation=True)
    return parser


if __name__ == "__main__":
    parsed = build_parser().parse_args()
    print(parsed)
    parsed.bfs_level = parsed.n_layers + 1      # pruning used nodes for memory (entity_classify.py:169)
    main(parsed)
