"""RelGraphConv(regularizer="basis"): the entity-classification layers (config 4).

Mirrors DGL's basis message function as used by kgvae/entity_classify.py:30-43:
``W_r = sum_b w_comp[r, b] V_b``; with dense features ``msg = x[src] @ W_r``; with 1-D integer
features (node ids) ``msg = W_r[id]`` - an embedding-style lookup.  Kernels: csrc/rgcn_basis.cu.
"""
import torch

from . import _lib as L
from . import ops

_c = ops._c


class MatmulFn(torch.autograd.Function):
    """a @ b through kg_gemm_f32 (composition W = w_comp @ V and its backward)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        out = torch.empty((a.shape[0], b.shape[1]), dtype=torch.float32, device=a.device)
        ops.gemm(a, b, out)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _c(g)
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            ops.gemm(g, b, da, trans_b=True)
        if ctx.needs_input_grad[1]:
            db = torch.empty_like(b)
            ops.gemm(a, g, db, trans_a=True)
        return da, db


def _tail(agg, h_bias, act_code, mask):
    out = torch.empty_like(agg)
    ops.epilogue_only(out, bias=None if h_bias is None else _c(h_bias), addend=agg, relu=(act_code == 1),
                      mask=None if mask is None else _c(mask))
    return out


def _tail_bwd(g, out, mask, act_code):
    gpre = torch.empty_like(out)
    L.call("kg_act_dropout_bwd", L.f32(_c(g)), L.f32(out), L.f32(mask), act_code, out.numel(), L.f32(gpre), L.stream())
    return gpre


_ARANGE_CACHE = {}


def _is_arange(ids32):
    """ids == arange(len(ids))?  (the reference's featureless input, entity_classify.py:63).  One
    device comparison per distinct feature tensor; the answer is cached on (storage, version)."""
    key = (ids32.data_ptr(), ids32.numel(), ids32._version, ids32.device)
    hit = _ARANGE_CACHE.get(key)
    if hit is None:
        n = ids32.numel()
        hit = bool(n > 0 and torch.equal(ids32, torch.arange(n, dtype=ids32.dtype, device=ids32.device)))
        if len(_ARANGE_CACHE) > 64:
            _ARANGE_CACHE.clear()
        _ARANGE_CACHE[key] = (hit, ids32)          # the tensor is kept alive so its address is not reused
        return hit
    return hit[0]


class BasisIdConvFn(torch.autograd.Function):
    """Integer-id features: out = dropout(act(sum_e norm_e sum_b coef[r_e,b] V[b, id_src, :] + h_bias
    + loop_weight[id]))."""

    @staticmethod
    def forward(ctx, ids, V, coef, loop_weight, h_bias, gi, act, drop_mask):
        V = _c(V)
        NB, n_in, out_f = V.shape
        n = gi.n_nodes
        dev = V.device
        ids32 = ops.as_i32(ids.reshape(-1), dev)
        if loop_weight is not None:                      # matmul_maybe_select: loop_weight[ids]
            agg = ops.EmbeddingFn.apply(loop_weight.detach(), ids32).contiguous()
        else:
            agg = torch.zeros((n, out_f), dtype=torch.float32, device=dev)
        cf = None if coef is None else _c(coef)
        gi.ensure_node_major()
        # featureless input (ids = arange) with a composed basis: source-tiled kernels, V read once
        src_tiled = bool(cf is not None and n_in == n and ids32.numel() == n and
                         L.lib().kg_basis_id_src_eligible(cf.shape[0], NB, out_f) and _is_arange(ids32))
        if src_tiled:
            L.call("kg_basis_id_src_fwd", L.f32(V), L.f32(cf), L.i32(gi.col_ptr), L.i32(gi.bwd_pack), n,
                   cf.shape[0], NB, out_f, L.f32(agg), L.stream())
        else:
            L.call("kg_basis_id_fwd", L.f32(V), L.f32(cf), L.i32(ids32), L.i32(gi.row_ptr), L.i32(gi.fwd_pack), n,
                   n_in, NB, out_f, L.f32(agg), L.stream())
        mask = None if drop_mask is None else _c(drop_mask)
        out = _tail(agg, h_bias, act, mask)
        ctx.save_for_backward(V, cf, ids32, out, mask)
        ctx.gi, ctx.act, ctx.src_tiled = gi, act, src_tiled
        ctx.loop_shape = None if loop_weight is None else loop_weight.shape
        ctx.has_bias = h_bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        V, cf, ids32, out, mask = ctx.saved_tensors
        gi = ctx.gi
        NB, n_in, out_f = V.shape
        gpre = _tail_bwd(g, out, mask, ctx.act)
        dcoef = None if cf is None else torch.zeros_like(cf)
        if ctx.src_tiled:
            dV = torch.empty_like(V)                      # every row is written exactly once
            L.call("kg_basis_id_src_bwd", L.f32(V), L.f32(cf), L.f32(gpre), L.i32(gi.col_ptr), L.i32(gi.bwd_pack),
                   n_in, cf.shape[0], NB, out_f, L.f32(dV), L.f32(dcoef), L.stream())
        else:
            dV = torch.zeros_like(V)
            L.call("kg_basis_id_bwd", L.f32(V), L.f32(cf), L.i32(ids32), L.f32(gpre), L.i32(gi.rel_pack), gi.n_edges,
                   n_in, NB, out_f, L.f32(dV), L.f32(dcoef), L.stream())
        dloop = None
        if ctx.loop_shape is not None:
            dloop = torch.zeros(ctx.loop_shape, dtype=torch.float32, device=V.device)
            L.call("kg_embedding_bwd", L.f32(gpre), L.i32(ids32), ids32.numel(), out_f, L.f32(dloop), L.stream())
        dbias = ops.colsum(gpre) if ctx.has_bias else None
        return None, dV, dcoef, dloop, dbias, None, None, None


class BasisIdSrcPartialFn(torch.autograd.Function):
    """Source-sharded form of the integer-id input layer (multi-GPU entity classification): this rank holds the
    rows V[:, lo:hi, :] of the basis table and every edge whose SOURCE it owns (local source ids, global
    destination ids); the result is this rank's PARTIAL sum of messages for all ``n_global`` destinations - the
    caller sums the partials over the ranks (parallel.AllReduceSumFn, whose backward all-reduces the gradient).
    dV needs no communication at all: every row of the table is written by its owner."""

    @staticmethod
    def forward(ctx, V, coef, gi, n_global):
        V, cf = _c(V), _c(coef)
        NB, n_local, out_f = V.shape
        if not L.lib().kg_basis_id_src_eligible(cf.shape[0], NB, out_f):
            raise RuntimeError("source-sharded basis layer: shape not covered (num_bases <= 64, out_feat <= 16)")
        gi.ensure_node_major()
        agg = torch.zeros((n_global, out_f), dtype=torch.float32, device=V.device)
        L.call("kg_basis_id_src_fwd", L.f32(V), L.f32(cf), L.i32(gi.col_ptr), L.i32(gi.bwd_pack), n_local,
               cf.shape[0], NB, out_f, L.f32(agg), L.stream())
        ctx.save_for_backward(V, cf)
        ctx.gi = gi
        return agg

    @staticmethod
    def backward(ctx, g):
        V, cf = ctx.saved_tensors
        gi = ctx.gi
        NB, n_local, out_f = V.shape
        dV = torch.empty_like(V)
        dcoef = torch.zeros_like(cf)
        L.call("kg_basis_id_src_bwd", L.f32(V), L.f32(cf), L.f32(_c(g)), L.i32(gi.col_ptr), L.i32(gi.bwd_pack),
               n_local, cf.shape[0], NB, out_f, L.f32(dV), L.f32(dcoef), L.stream())
        return dV, dcoef, None, None


class BasisDenseConvFn(torch.autograd.Function):
    """Dense features with composed per-relation weights W [R, in, out]."""

    @staticmethod
    def forward(ctx, x, W, loop_weight, h_bias, gi, act, drop_mask):
        x, W = _c(x), _c(W)
        R, in_f, out_f = W.shape
        n = x.shape[0]
        agg = torch.zeros((n, out_f), dtype=torch.float32, device=x.device)
        L.call("kg_basis_dense_fwd", L.f32(x), L.i32(gi.rel_pack), gi.n_edges, L.f32(W), R, in_f, out_f, L.f32(agg),
               L.stream())
        mask = None if drop_mask is None else _c(drop_mask)
        out = torch.empty_like(agg)
        bias = None if h_bias is None else _c(h_bias)
        if loop_weight is not None:
            loop_weight = _c(loop_weight)
            ops.gemm(x, loop_weight, out, bias=bias, addend=agg, relu=(act == 1), mask=mask)
        else:
            ops.epilogue_only(out, bias=bias, addend=agg, relu=(act == 1), mask=mask)
        ctx.save_for_backward(x, W, loop_weight, out, mask)
        ctx.gi, ctx.act, ctx.has_bias = gi, act, h_bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, W, loop_weight, out, mask = ctx.saved_tensors
        gi = ctx.gi
        R, in_f, out_f = W.shape
        gpre = _tail_bwd(g, out, mask, ctx.act)
        dx = torch.zeros_like(x) if ctx.needs_input_grad[0] else None
        dW = torch.zeros_like(W)
        L.call("kg_basis_dense_bwd", L.f32(x), L.f32(gpre), L.i32(gi.rel_pack), gi.n_edges, L.f32(W), R, in_f, out_f,
               L.f32(dx), L.f32(dW), L.stream())
        dloop = None
        if loop_weight is not None:
            if dx is not None:
                ops.gemm(gpre, loop_weight, dx, trans_b=True, accumulate=True)
            dloop = torch.empty_like(loop_weight)
            ops.gemm(x, gpre, dloop, trans_a=True)
        dbias = ops.colsum(gpre) if ctx.has_bias else None
        return dx, dW, dloop, dbias, None, None, None


def forward(layer, g, gi, x, h_bias, loop_w, act_code, post, mask):
    """Called by nn.RelGraphConv.forward for regularizer == "basis"."""
    V = layer.weight                                             # [NB, in, out]
    coef = layer.w_comp if layer.num_bases < layer.num_rels else None
    fused_act, fused_mask = (act_code, mask) if post is None else (0, None)
    if x.dtype in (torch.int64, torch.int32) and x.dim() == 1:
        h = BasisIdConvFn.apply(x, V, coef, loop_w, h_bias, gi, fused_act, fused_mask)
    else:
        if coef is not None:
            W = MatmulFn.apply(coef, V.reshape(layer.num_bases, -1)).view(layer.num_rels, layer.in_feat, layer.out_feat)
        else:
            W = V
        h = BasisDenseConvFn.apply(x, W, loop_w, h_bias, gi, fused_act, fused_mask)
    if post is not None:
        h = post(h)
        if mask is not None:
            h = h * mask
    return h
