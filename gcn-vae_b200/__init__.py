"""kgvae_b200: B200-native (sm_100a) implementation of the GCN-VAE link-prediction hot path.

Module layout mirrors the reference's flat ``kgvae/`` scripts so that call sites translate one
to one: ``model`` (KGVAE, EmbeddingLayer, RGCN), ``link_predict`` (LinkPredict,
node_norm_to_edge_norm, main), ``flow_network`` (MaskedLinear, PermuteLayer, MADE), ``utils``
(graph build, sampling, calc_mrr, probability helpers), ``nn`` (RelGraphConv with DGL's
signature), ``graph`` (the DGLGraph surface the reference uses).  ``ops`` holds the autograd
wrappers over the C ABI in ``include/kgvae_b200.h``; ``csrc/`` the CUDA kernels.
"""
from . import _lib, datasets, flow_network, graph, link_predict, model, nn, ops, parallel, utils  # noqa: F401
from .flow_network import MADE, MaskedLinear, PermuteLayer  # noqa: F401
from .graph import DGLGraph, Graph  # noqa: F401
from .link_predict import LinkPredict, node_norm_to_edge_norm  # noqa: F401
from .model import KGVAE, RGCN, BaseRGCN, EmbeddingLayer  # noqa: F401
from .nn import RelGraphConv  # noqa: F401

__all__ = ["KGVAE", "RGCN", "BaseRGCN", "EmbeddingLayer", "LinkPredict", "node_norm_to_edge_norm",
           "MADE", "MaskedLinear", "PermuteLayer", "RelGraphConv", "Graph", "DGLGraph"]
