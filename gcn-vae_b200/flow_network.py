"""IAF flow blocks with the reference's module surface (kgvae/flow_network.py).

``MaskedLinear`` / ``PermuteLayer`` / ``MADE`` keep constructor arguments, attribute names
(``net``, ``m``, ``mask``) and state-dict keys ``net.{0,2,...}.{weight,bias,mask}`` so reference
checkpoints load.  The arithmetic runs through the fused ops: masked weights are formed once per
``MADE.forward`` call (the reference re-multiplies ``mask * weight`` in each of its 30 linear
calls), every linear is one GEMM with bias + ReLU in the epilogue, and the per-pass element
update + log-det row-sum is one kernel.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops


class MaskedLinear(nn.Linear):
    """Linear layer whose weight is multiplied by a fixed 0/1 mask (flow_network.py:7-15)."""

    def __init__(self, input_size, output_size, mask):
        super().__init__(input_size, output_size)
        self.register_buffer("mask", mask)

    def masked_weight(self):
        return self.mask * self.weight

    def forward(self, x, relu=False, weight=None, prepared=None):
        w = self.masked_weight() if weight is None else weight
        return ops.LinearFn.apply(x, w, self.bias, relu, prepared)


class PermuteLayer(nn.Module):
    """Reverses the column order; log-det is zero (flow_network.py:18-34)."""

    def __init__(self, num_inputs):
        super().__init__()
        self.perm = np.array(np.arange(0, num_inputs)[::-1])

    def forward(self, inputs):
        return ops.ReverseColumnsFn.apply(inputs), torch.zeros(inputs.size(0), 1, device=inputs.device)

    def inverse(self, inputs):
        return self.forward(inputs)


class MADE(nn.Module):
    """Gaussian MADE used as one IAF step (flow_network.py:37-112)."""

    def __init__(self, input_size, hidden_size, n_hidden):
        super().__init__()
        self.input_size, self.hidden_size, self.n_hidden = input_size, hidden_size, n_hidden
        masks = self.create_masks()
        layers = [MaskedLinear(input_size, hidden_size, masks[0]), nn.ReLU(inplace=True)]
        for i in range(n_hidden):
            layers += [MaskedLinear(hidden_size, hidden_size, masks[i + 1]), nn.ReLU(inplace=True)]
        layers += [MaskedLinear(hidden_size, input_size * 2, masks[-1].repeat(2, 1))]
        self.net = nn.Sequential(*layers)

    def create_masks(self):
        """Sequential-order degrees; ``self.m`` doubles as the per-pass column lists
        (flow_network.py:65-83)."""
        D = self.input_size
        degrees = [torch.arange(D)]
        degrees += [torch.arange(self.hidden_size) % (D - 1) for _ in range(self.n_hidden + 1)]
        degrees += [torch.arange(D) % D - 1]
        self.m = degrees
        # per pass: how often each column occurs in the index list (0 = column is not rewritten)
        self._col_mult = [torch.bincount(cols % D, minlength=D).to(torch.int32) for cols in degrees]
        return [(hi.unsqueeze(-1) >= lo.unsqueeze(0)).float()
                for lo, hi in zip(degrees[:-1], degrees[1:])]

    def _linears(self):
        return [layer for layer in self.net if isinstance(layer, MaskedLinear)]

    def _run_net(self, x, weights, prepared=None):
        linears = self._linears()
        for i, (layer, w) in enumerate(zip(linears, weights)):
            x = layer(x, relu=(i + 1 < len(linears)), weight=w, prepared=prepared[i] if prepared else None)
        return x

    def forward(self, z):
        """z -> (x, log_det[N]).  One full pass per entry of ``self.m``: pass p rewrites the
        columns listed in ``self.m[p]`` (all columns for the first and last pass, all but the
        last column - with column 0 listed twice - in between) with z * exp(alpha + mu)
        (flow_network.py:85-98)."""
        weights = [layer.masked_weight() for layer in self._linears()]
        if self._col_mult[0].device != z.device:
            self._col_mult = [m.to(z.device) for m in self._col_mult]
        # every full pass multiplies by the same masked weights: split them for the tensor cores once
        prepared = [ops.prepare(w.contiguous(), z.shape[0]) for w in weights] if z.is_cuda else None
        if prepared is not None and not any(isinstance(p_, ops.Prepared) for p_ in prepared):
            prepared = None
        x = torch.zeros_like(z)
        log_det = None
        n_pass = len(self.m)
        for p, mult in enumerate(self._col_mult):
            if p == 0:
                # the reference starts from x = 0 (flow_network.py:91): every row of the first pass sees the
                # same input, so net(0) is ONE row - 5 one-row products instead of 5 N-row GEMMs, and in the
                # backward the expand turns the N-row gradient into its column sum before the network
                out = self._run_net(x[:1], weights).expand(z.shape[0], -1)
            else:
                out = self._run_net(x, weights, prepared)
            x, log_det = ops.IafUpdateFn.apply(z, out, x, mult, p + 1 == n_pass)
        return x, log_det

    def inverse(self, x):
        """x -> (z, log_det) (flow_network.py:100-112); generation only, plain tensor ops."""
        weights = [layer.masked_weight() for layer in self._linears()]
        mu, alpha = self._run_net(x, weights).chunk(2, dim=-1)
        return (x - mu) * torch.exp(-alpha), torch.sum(-alpha, dim=-1)
