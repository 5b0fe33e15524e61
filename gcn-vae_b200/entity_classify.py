"""Entity classification with the basis-regularised RGCN (kgvae/entity_classify.py; config 4).

``EntityClassify`` keeps the reference's constructor (through ``BaseRGCN``) and layer stack:
an input layer on integer node ids (``RelGraphConv(num_nodes, h, "basis")``), hidden layers,
and a softmax output layer.  ``synthetic_graph`` stands in for ``dgl.contrib.data.load_data``
(rdflib datasets are not available offline): it draws a typed multigraph of a named shape with
the loader's edge norm (1 / number of same-type edges into the destination) and random labels.
"""
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
