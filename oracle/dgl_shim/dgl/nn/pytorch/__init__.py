"""``dgl.nn.pytorch.RelGraphConv`` restated in plain PyTorch (DGL 0.4.x semantics).

TEST INFRASTRUCTURE ONLY (see ``oracle/dgl_shim/dgl/__init__.py``).

Restates ``python/dgl/nn/pytorch/conv/relgraphconv.py`` of DGL 0.4.x (not in
/root/reference): per-edge message = (block-diagonal or
basis-combined) relation weight applied to the source feature, scaled by the
edge norm, summed over incoming edges, then ``+ h_bias``, ``+ x @ loop_weight``,
activation, dropout.

One hook beyond DGL: ``dropout_mask`` (a preset ``[N, out]`` keep-mask already
scaled by ``1/(1-p)``) so that golden vectors are reproducible across devices.
"""
import torch
import torch.nn as nn


class RelGraphConv(nn.Module):
    def __init__(self, in_feat, out_feat, num_rels, regularizer="basis",
                 num_bases=None, bias=True, activation=None, self_loop=False,
                 dropout=0.0):
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
        self.dropout_mask = None  # golden-vector hook, see module docstring

    def _messages(self, g, x, etypes):
        h_src = x[g._src]
        if self.regularizer == "bdd":
            if x.dtype == torch.int64 and x.dim() == 1:
                raise TypeError("Block decomposition does not allow integer ID feature.")
            w = self.weight.index_select(0, etypes).view(-1, self.submat_in, self.submat_out)
            return torch.bmm(h_src.reshape(-1, 1, self.submat_in), w).view(-1, self.out_feat)
        if self.num_bases < self.num_rels:
            w = torch.matmul(self.w_comp, self.weight.view(self.num_bases, -1))
            w = w.view(self.num_rels, self.in_feat, self.out_feat)
        else:
            w = self.weight
        if h_src.dtype == torch.int64 and h_src.dim() == 1:
            return w.view(-1, self.out_feat).index_select(0, etypes * self.in_feat + h_src)
        return torch.bmm(h_src.unsqueeze(1), w.index_select(0, etypes)).squeeze(1)

    def forward(self, g, x, etypes, norm=None):
        msg = self._messages(g, x, etypes)
        if norm is not None:
            msg = msg * norm
        h = g.sum_messages(msg)
        if self.bias:
            h = h + self.h_bias
        if self.self_loop:
            if x.dtype == torch.int64 and x.dim() == 1:
                h = h + self.loop_weight.index_select(0, x)
            else:
                h = h + torch.matmul(x, self.loop_weight)
        if self.activation:
            h = self.activation(h)
        if self.dropout_mask is not None:
            return h * self.dropout_mask
        return self.dropout(h)
