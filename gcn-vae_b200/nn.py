"""``RelGraphConv`` with DGL's constructor and call signature, running on the sm_100a kernels.

Drop-in for ``dgl.nn.pytorch.RelGraphConv`` as used by the reference (kgvae/model.py:54-59,
209-211; kgvae/entity_classify.py:31-43): same parameter names, shapes and initialisers
(``weight``, ``w_comp``, ``h_bias``, ``loop_weight``), same ``forward(g, x, etypes, norm)``.
"""
import torch
import torch.nn as nn

from . import ops

_ACT_IDENTITY, _ACT_RELU = 0, 1


def _classify_activation(act):
    """Map the callables the reference passes to a fused epilogue; others run unfused."""
    if act is None or isinstance(act, nn.Identity):
        return _ACT_IDENTITY, None
    if isinstance(act, nn.ReLU) or act is torch.relu or act is torch.nn.functional.relu:
        return _ACT_RELU, None
    return _ACT_IDENTITY, act      # e.g. `lambda x: x` (kgvae/model.py:58) or softmax


class RelGraphConv(nn.Module):
    def __init__(self, in_feat, out_feat, num_rels, regularizer="basis", num_bases=None,
                 bias=True, activation=None, self_loop=False, dropout=0.0):
        super().__init__()
        self.in_feat, self.out_feat, self.num_rels = in_feat, out_feat, num_rels
        self.regularizer = regularizer
        if num_bases is None or num_bases > num_rels or num_bases < 0:
            num_bases = num_rels
        self.num_bases = num_bases
        self.bias, self.activation, self.self_loop = bias, activation, self_loop
        gain = nn.init.calculate_gain("relu")
        if regularizer == "basis":
            self.weight = nn.Parameter(torch.empty(num_bases, in_feat, out_feat))
            nn.init.xavier_uniform_(self.weight, gain=gain)
            if num_bases < num_rels:
                self.w_comp = nn.Parameter(torch.empty(num_rels, num_bases))
                nn.init.xavier_uniform_(self.w_comp, gain=gain)
        elif regularizer == "bdd":
            if in_feat % num_bases != 0 or out_feat % num_bases != 0:
                raise ValueError("Feature size must be a multiplier of num_bases.")
            self.submat_in = in_feat // num_bases
            self.submat_out = out_feat // num_bases
            self.weight = nn.Parameter(
                torch.empty(num_rels, num_bases * self.submat_in * self.submat_out))
            nn.init.xavier_uniform_(self.weight, gain=gain)
        else:
            raise ValueError("Regularizer must be either 'basis' or 'bdd'")
        if bias:
            self.h_bias = nn.Parameter(torch.zeros(out_feat))
        if self_loop:
            self.loop_weight = nn.Parameter(torch.empty(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=gain)
        self.dropout = nn.Dropout(dropout)
        # optional preset keep-mask [N, out] already scaled by 1/(1-p): lets a parity harness
        # feed the same mask to this layer and to the oracle (CUDA Philox != CPU MT)
        self.dropout_mask = None

    def _keep_mask(self, n, device, part=None):
        if self.dropout_mask is not None:
            return self.dropout_mask
        p = self.dropout.p
        if not self.training or p == 0.0:
            return None
        keep = 1.0 - p
        gen = None if part is None else part.noise_generator(device)     # per-rank stream (see KGVAE.forward)
        return torch.empty((n, self.out_feat), device=device).bernoulli_(keep, generator=gen).div_(keep)

    def forward(self, g, x, etypes, norm=None):
        if not x.is_cuda:
            raise RuntimeError("kgvae_b200.RelGraphConv runs on CUDA only (no CPU fallback)")
        act_code, post = _classify_activation(self.activation)
        part = getattr(g, "partition", None)
        mask = self._keep_mask(x.shape[0], x.device, part)
        # integer-id features walk node-major lists as well: have them built with the first index
        id_feats = x.dim() == 1 and x.dtype in (torch.int64, torch.int32)
        gi = g.index_for(etypes, norm, self.num_rels, node_major=id_feats)
        h_bias = self.h_bias if self.bias else None
        loop_w = self.loop_weight if self.self_loop else None
        if self.regularizer == "bdd":
            if x.dtype == torch.int64 and x.dim() == 1:
                raise TypeError("Block decomposition does not allow integer ID feature.")
            gather, peer = False, None
            if part is not None:      # destination-partitioned
                peer_ok = part.use_peer_gather(gi.n_edges) and not ops.L.lib().kg_bdd_layouts_needed(
                    self.num_bases, self.submat_in, self.submat_out)
                if peer_ok:           # fused: the kernel gathers source rows from the owners' HBM
                    peer = part.peer_rows(id(self), self.in_feat, x.device)
                else:                 # NCCL all-gather of every node's features, overlapped with the self-loop GEMM
                    gather = True
            if post is None:
                return ops.BddConvFn.apply(x, self.weight, loop_w, h_bias, gi, self.num_bases,
                                           act_code, mask, gather, peer, part)
            h = ops.BddConvFn.apply(x, self.weight, loop_w, h_bias, gi, self.num_bases,
                                    _ACT_IDENTITY, None, gather, peer, part)
            h = post(h)
            return h if mask is None else h * mask
        from . import basis   # entity-classification layers (config 4)
        return basis.forward(self, g, gi, x, h_bias, loop_w, act_code, post, mask)
