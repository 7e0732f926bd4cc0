"""Autograd layer over the kernel-level ops (torch.autograd.Function).

Mirrors upstream DGL v0.6.1 python/dgl/backend/pytorch/sparse.py (GSpMM / GSDDMM / EdgeSoftmax):
the backward of an SpMM is an SpMM on the reversed graph plus an SDDMM, the backward of an SDDMM is
an SpMM with `copy_rhs` / `mul`, `sub` and `div` are rewritten as `add(-rhs)` / `mul(1/rhs)` before
the Function, and gradients w.r.t. broadcast operands are reduced back to the operand's shape
(SURVEY.md Appendix A.4).  Two things are new relative to upstream: EdgeSoftmax is one fused
kernel per direction instead of a 4+1 / 2+2 launch composite, and GATFused keeps the whole
attention (scores, softmax, dropout, aggregation) in one forward kernel + two backward kernels.
"""
import torch

from . import sparse as K
from ._capi import DGLError


def _reduce_grad(grad, shape):
    """Sum `grad` over the dims that were broadcast so that it matches `shape` (operand shape)."""
    grad_shape = grad.shape[1:]
    in_shape = tuple(shape[1:])
    if in_shape == tuple(grad_shape):
        return grad.view(shape) if grad.shape != tuple(shape) else grad
    num_to_squeeze = len(grad_shape) - len(in_shape)
    in_shape = (1,) * num_to_squeeze + in_shape
    dims = [i + 1 for i, (a, b) in enumerate(zip(grad_shape, in_shape)) if a != b]
    if dims:
        grad = grad.sum(dim=dims, keepdim=True)
    return grad.reshape((-1,) + tuple(shape[1:]))


def _need_reduce_last_dim(ufeat, efeat):
    """True for the (N,..,F) x (E,..,1) head-broadcast case: d efeat is a dot over the last dim."""
    ushp, eshp = ufeat.shape, efeat.shape
    return len(ushp) == len(eshp) and ushp[1:-1] == eshp[1:-1] and eshp[-1] == 1 and ushp[-1] > 1


def _expand(x, shape):
    return x.expand(-1, *shape)


class GSpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gidx, op, reduce_op, X, Y, row_scale, zero_inf=False):
        out, (argX, argY) = K._gspmm(gidx, op, reduce_op, X, Y, row_scale, zero_inf=zero_inf)
        ctx.backward_cache = gidx, op, reduce_op
        ctx.has_scale = row_scale is not None
        ctx.zero_inf = bool(zero_inf) and reduce_op in ("max", "min")
        ctx.save_for_backward(X, Y, argX, argY, row_scale)
        return out

    @staticmethod
    def backward(ctx, dZ):
        gidx, op, reduce_op = ctx.backward_cache
        X, Y, argX, argY, row_scale = ctx.saved_tensors
        dZ = dZ.contiguous()
        if row_scale is not None:  # fused mean: out = sum / deg  =>  d(sum) = dZ / deg
            dZ = (dZ / row_scale.view((-1,) + (1,) * (dZ.dim() - 1))).to(dZ.dtype)
        if ctx.zero_inf and gidx.n_edges > 0:
            # upstream applies where(isinf(out), 0, out) as a differentiable post-pass OUTSIDE GSpMM, which
            # zeroes the gradient of rows without in-edges before the arg-scatter; the kernel folds that
            # post-pass into its store (arg = 0 for such rows), so the same mask has to be applied here or
            # node 0 / edge 0 would collect the gradient of every isolated destination row
            empty = gidx.csc().degrees() == 0
            dZ = dZ.masked_fill(empty.view((-1,) + (1,) * (dZ.dim() - 1)), 0)
        dX = dY = None
        if op != "copy_rhs" and ctx.needs_input_grad[3]:
            g_rev = gidx.reverse()
            if reduce_op == "sum":
                if op == "mul":
                    dX = K._gspmm(g_rev, "mul", "sum", dZ, Y)[0]
                else:  # add, copy_lhs
                    dX = K._gspmm(g_rev, "copy_lhs", "sum", dZ, None)[0]
            else:  # max / min: route dZ through the recorded arg indices
                dX = torch.zeros((X.shape[0],) + dZ.shape[1:], dtype=X.dtype, device=X.device)
                if op == "mul":
                    grad = _expand(Y, dZ.shape[1:]).gather(0, argY.long()) * dZ
                    dX.scatter_add_(0, argX.long(), grad)
                else:
                    dX.scatter_add_(0, argX.long(), dZ)
            dX = _reduce_grad(dX, X.shape)
        if op != "copy_lhs" and ctx.needs_input_grad[4]:
            if reduce_op == "sum":
                if op == "mul" and _need_reduce_last_dim(X, Y):
                    dY = K._gsddmm(gidx, "dot", X, dZ)
                elif op == "mul":
                    dY = K._gsddmm(gidx, "mul", X, dZ)
                else:  # add, copy_rhs
                    dY = K._gsddmm(gidx, "copy_rhs", X, dZ)
            else:
                dY = torch.zeros((Y.shape[0],) + dZ.shape[1:], dtype=Y.dtype, device=Y.device)
                if op == "mul":
                    grad = _expand(X, dZ.shape[1:]).gather(0, argX.long()) * dZ
                    dY.scatter_add_(0, argY.long(), grad)
                else:
                    dY.scatter_add_(0, argY.long(), dZ)
            dY = _reduce_grad(dY, Y.shape)
        return None, None, None, dX, dY, None, None


class GSDDMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gidx, op, X, Y, lhs_target, rhs_target):
        out = K._gsddmm(gidx, op, X, Y, lhs_target, rhs_target)
        ctx.backward_cache = gidx, op, lhs_target, rhs_target
        ctx.save_for_backward(X, Y)
        return out

    @staticmethod
    def backward(ctx, dZ):
        gidx, op, lhs_target, rhs_target = ctx.backward_cache
        X, Y = ctx.saved_tensors
        dZ = dZ.contiguous()
        dX = dY = None
        if op != "copy_rhs" and ctx.needs_input_grad[2]:
            if lhs_target in ("u", "v"):
                _g = gidx if lhs_target == "v" else gidx.reverse()
                if op in ("add", "copy_lhs"):
                    dX = K._gspmm(_g, "copy_rhs", "sum", None, dZ)[0]
                else:  # mul, dot
                    if rhs_target == lhs_target:
                        dX = K._gspmm(_g, "copy_rhs", "sum", None, dZ)[0] * Y
                    elif rhs_target == "e":
                        dX = K._gspmm(_g, "copy_rhs", "sum", None, dZ * Y)[0]
                    else:  # the other endpoint
                        dX = K._gspmm(_g, "mul", "sum", Y, dZ)[0]
            else:  # lhs lives on edges
                if op in ("add", "copy_lhs"):
                    dX = dZ
                else:
                    dX = K._gsddmm(gidx, "mul", dZ, Y, "e", rhs_target)
            dX = _reduce_grad(dX, X.shape)
        if op != "copy_lhs" and ctx.needs_input_grad[3]:
            if rhs_target in ("u", "v"):
                _g = gidx if rhs_target == "v" else gidx.reverse()
                if op in ("add", "copy_rhs"):
                    dY = K._gspmm(_g, "copy_rhs", "sum", None, dZ)[0]
                else:
                    if lhs_target == rhs_target:
                        dY = K._gspmm(_g, "copy_rhs", "sum", None, dZ)[0] * X
                    elif lhs_target == "e":
                        dY = K._gspmm(_g, "copy_rhs", "sum", None, dZ * X)[0]
                    else:
                        dY = K._gspmm(_g, "mul", "sum", X, dZ)[0]
            else:
                if op in ("add", "copy_rhs"):
                    dY = dZ
                else:
                    dY = K._gsddmm(gidx, "mul", dZ, X, "e", lhs_target)
            dY = _reduce_grad(dY, Y.shape)
        return None, None, dX, dY, None, None


class EdgeSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gidx, score, norm_by):
        if norm_by == "src":
            gidx = gidx.reverse()
        out = K._edge_softmax_fwd(gidx, score)
        ctx.backward_cache = gidx
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        gidx = ctx.backward_cache
        out, = ctx.saved_tensors
        return None, K._edge_softmax_bwd(gidx, out, grad_out), None


class GATFused(torch.autograd.Function):
    """rst[v,h,:] = sum_{u->v} dropout(softmax_v(leaky_relu(el[u,h] + er[v,h]))) * ft[u,h,:]"""

    @staticmethod
    def forward(ctx, gidx, ft, el, er, negative_slope, dropout_p, seed):
        rst, row_max, row_sum, _ = K._gat_fwd(gidx, ft, el, er, negative_slope, dropout_p, seed)
        ctx.backward_cache = gidx, negative_slope, dropout_p, seed
        ctx.save_for_backward(ft, el, er, row_max, row_sum)
        return rst

    @staticmethod
    def backward(ctx, grad_rst):
        gidx, slope, dropout_p, seed = ctx.backward_cache
        ft, el, er, row_max, row_sum = ctx.saved_tensors
        grad_ft, grad_el, grad_er = K._gat_bwd(gidx, ft.contiguous(), el.contiguous(), er.contiguous(), row_max,
                                               row_sum, grad_rst, slope, dropout_p, seed)
        return None, grad_ft, grad_el, grad_er, None, None, None


def gspmm(gidx, op, reduce_op, lhs_data, rhs_data, row_scale=None, zero_inf=False):
    """zero_inf (max / min): rows without in-edges come back as 0 instead of -/+inf (the replacement
    upstream's dgl.ops.gspmm applies afterwards, done in the kernel's store instead of two more passes)."""
    if op == "sub":
        op, rhs_data = "add", -rhs_data
    if op == "div":
        op, rhs_data = "mul", 1.0 / rhs_data
    return GSpMM.apply(gidx, op, reduce_op, lhs_data, rhs_data, row_scale, zero_inf)


def gsddmm(gidx, op, lhs_data, rhs_data, lhs_target="u", rhs_target="v"):
    if op == "sub":
        op, rhs_data = "add", -rhs_data
    if op == "div":
        op, rhs_data = "mul", 1.0 / rhs_data
    return GSDDMM.apply(gidx, op, lhs_data, rhs_data, lhs_target, rhs_target)


def edge_softmax(gidx, logits, eids=None, norm_by="dst"):
    if norm_by not in ("dst", "src"):
        raise DGLError("norm_by must be 'src' or 'dst'")
    if eids is not None:
        # upstream (python/dgl/backend/pytorch/sparse.py::edge_softmax): softmax inside the edge-induced subgraph that
        # keeps every node -- `logits` has one row per listed edge, in the order of `eids`
        from .graph_index import GraphIndex
        eids = torch.as_tensor(eids, device=gidx.src.device).long().view(-1)
        if logits.shape[0] != eids.shape[0]:
            raise DGLError("edge_softmax: expect %d logit rows for the listed edges, got %d" % (eids.shape[0], logits.shape[0]))
        gidx = GraphIndex(gidx.src[eids].contiguous(), gidx.dst[eids].contiguous(), gidx.n_src, gidx.n_dst, gidx.idtype)
    return EdgeSoftmax.apply(gidx, logits, norm_by)


def gat_fused(gidx, ft, el, er, negative_slope=0.2, dropout_p=0.0, seed=0):
    return GATFused.apply(gidx, ft, el, er, float(negative_slope), float(dropout_p), int(seed))


class GCNMsgSum(torch.autograd.Function):
    """Fused message + reduce of the graph-classification GCN layer (main_dgl_molhiv_gcn.py:46,50-52):
    sum over in-edges of (c[u] c[v]) relu(x[u] + w[e]).  The (E, D) message tensor is never materialised; the backward
    recomputes the ReLU mask from x and w.  The norm vectors get no gradient (they are functions of the degrees)."""

    @staticmethod
    def forward(ctx, gidx, x, w, c_src, c_dst):
        out = K._gcn_msg_sum_fwd(gidx, x, w, c_src, c_dst)
        ctx.gidx = gidx
        ctx.save_for_backward(x, w, c_src, c_dst)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, w, c_src, c_dst = ctx.saved_tensors
        gx, gw = K._gcn_msg_sum_bwd(ctx.gidx, x, w, c_src, c_dst, grad_out)
        return None, gx if ctx.needs_input_grad[1] else None, gw if ctx.needs_input_grad[2] else None, None, None


def gcn_msg_sum(gidx, x, w, c_src, c_dst):
    if c_src.requires_grad or c_dst.requires_grad:
        raise DGLError("gcn_norm_relu_sum: the normalisation vectors are not differentiable inputs")
    return GCNMsgSum.apply(gidx, x, w, c_src, c_dst)


class CatEmbedSum(torch.autograd.Function):
    """out[i] = sum_k table[off[k] + x[i,k]] (AtomEncoder / BondEncoder of the OGB molecule scripts) as one kernel each
    way; the backward is deterministic (no atomics, no sort)."""

    @staticmethod
    def forward(ctx, x, table, offsets):
        _capi = K._capi
        _capi.require_cuda(x, table)
        out = _capi.call(_capi.ops().cat_embed_sum_fwd, x.contiguous(), table.contiguous(), offsets)
        _capi.count_launch(1)
        ctx.offsets = offsets
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        _capi = K._capi
        x, = ctx.saved_tensors
        gt = _capi.call(_capi.ops().cat_embed_sum_bwd, x.contiguous(), grad_out.contiguous(), ctx.offsets)
        _capi.count_launch(1)
        return None, gt, None
