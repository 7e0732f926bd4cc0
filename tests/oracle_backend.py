"""TEST INFRASTRUCTURE: swaps the kernel-level entry points of the dgl stand-in for the CPU oracle so
that the host-side logic (autograd wiring, update_all translation, nn layers, the unchanged reference
scripts) can be exercised in this GPU-less container.  Only tests may import this module; the
product never routes through it (dgl/sparse.py raises on CPU tensors)."""
import contextlib

import numpy as np
import torch

from oracle import dgl_ref as R


def _og(gidx):
    key = "_oracle_graph_rev" if gidx._rev else "_oracle_graph"
    if key not in gidx._c:
        gidx._c[key] = R.OracleGraph(gidx.src.cpu().numpy(), gidx.dst.cpu().numpy(), gidx.n_src, gidx.n_dst)
    return gidx._c[key]


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _gspmm(gidx, op, reduce_op, u, e, row_scale=None, out=None, zero_inf=False):
    prev = out
    out, (au, ae) = R._gspmm(_og(gidx), op, reduce_op, _np(u), _np(e))
    out = torch.from_numpy(np.ascontiguousarray(out))
    if zero_inf and reduce_op in ("max", "min"):
        out = torch.where(torch.isinf(out), torch.zeros((), dtype=out.dtype), out)
    if row_scale is not None:
        out = out / row_scale.view((-1,) + (1,) * (out.dim() - 1))
    if prev is not None:
        prev += out
        out = prev
    cv = lambda a: None if a is None else torch.from_numpy(a).to(gidx.idtype)
    return out, (cv(au), cv(ae))


def _gsddmm(gidx, op, lhs, rhs, lhs_target="u", rhs_target="v"):
    return torch.from_numpy(np.ascontiguousarray(R._gsddmm(_og(gidx), op, _np(lhs), _np(rhs), lhs_target, rhs_target)))


def _edge_softmax_fwd(gidx, logits):
    return torch.from_numpy(R.edge_softmax(_og(gidx), _np(logits)))


def _edge_softmax_bwd(gidx, out, grad_out):
    return torch.from_numpy(R.edge_softmax_backward(_og(gidx), _np(out), _np(grad_out)))


def _gat_fwd(gidx, ft, el, er, slope, dropout_p, seed, want_scores=False):
    assert dropout_p == 0.0, "oracle backend has no dropout replay"
    og = _og(gidx)
    H = ft.shape[1]
    e = R._gsddmm(og, "add", _np(el).reshape(-1, H, 1), _np(er).reshape(-1, H, 1))
    e = np.where(e > 0, e, e * np.float32(slope)).astype(np.float32)
    a = R.edge_softmax(og, e)
    rst, _ = R._gspmm(og, "mul", "sum", _np(ft), a)
    mx, _ = R._gspmm(og, "copy_rhs", "max", None, e)
    return torch.from_numpy(rst), torch.from_numpy(mx[:, :, 0]), torch.from_numpy(a[:, :, 0].copy()), \
        (torch.from_numpy(a[:, :, 0].copy()) if want_scores else None)


class _MeanDivisor:
    pass


@contextlib.contextmanager
def installed():
    """Route dgl's kernel-level calls to the oracle for the duration of the context (CPU tensors)."""
    from dgl import sparse as K, _capi
    from dgl import graph_index as GI
    from dgl.nn.pytorch import conv
    saved = (K._gspmm, K._gsddmm, K._edge_softmax_fwd, K._edge_softmax_bwd, K._gat_fwd, conv.GATConv.fused,
             GI.GraphIndex.csc)
    K._gspmm, K._gsddmm = _gspmm, _gsddmm
    K._edge_softmax_fwd, K._edge_softmax_bwd, K._gat_fwd = _edge_softmax_fwd, _edge_softmax_bwd, _gat_fwd
    conv.GATConv.fused = False  # the oracle has no fused backward: use upstream's composition

    class _CscStub:
        def __init__(self, gidx):
            self._g = gidx

        def mean_divisor(self):
            return self._g.in_degrees().clamp(min=1).to(torch.float32)

        def degrees(self):
            return self._g.in_degrees()

    GI.GraphIndex.csc = lambda self: _CscStub(self)
    try:
        yield
    finally:
        (K._gspmm, K._gsddmm, K._edge_softmax_fwd, K._edge_softmax_bwd, K._gat_fwd, conv.GATConv.fused,
         GI.GraphIndex.csc) = saved
