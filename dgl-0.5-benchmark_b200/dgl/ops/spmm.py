"""dgl.ops.gspmm -- generalized SpMM (same signature and post-processing as upstream DGL v0.6.1
python/dgl/ops/spmm.py; called by kernel/dgl-new.py:20)."""
import sys

import torch

from .. import backend as B
from .._capi import DGLError

__all__ = ["gspmm"]


def reshape_lhs_rhs(lhs_data, rhs_data):
    """Give both operands the same number of dims by inserting 1s after the leading dim, so that
    e.g. (N, D) and (E,) -> (N, D) and (E, 1)."""
    lhs_shape, rhs_shape = lhs_data.shape, rhs_data.shape
    if len(lhs_shape) != len(rhs_shape):
        max_ndims = max(len(lhs_shape), len(rhs_shape))
        lhs_data = lhs_data.reshape((lhs_shape[0],) + (1,) * (max_ndims - len(lhs_shape)) + tuple(lhs_shape[1:]))
        rhs_data = rhs_data.reshape((rhs_shape[0],) + (1,) * (max_ndims - len(rhs_shape)) + tuple(rhs_shape[1:]))
    return lhs_data, rhs_data


def _gidx(g):
    gi = getattr(g, "_graph", None)
    if gi is None:
        raise DGLError("expected a DGLGraph, got %r" % type(g))
    return gi


def gspmm(g, op, reduce_op, lhs_data, rhs_data):
    r"""Generalized Sparse Matrix Multiplication: for every edge compute a message
    ``op(lhs_data[src], rhs_data[eid])`` and aggregate the messages at the destination node with
    ``reduce_op``.

    op : 'add' | 'sub' | 'mul' | 'div' | 'copy_lhs' | 'copy_rhs'
    reduce_op : 'sum' | 'max' | 'min' | 'mean'
    lhs_data : node features (N_src, ...) or None;  rhs_data : edge features (E, ...) or None.
    Returns a tensor of shape (N_dst, ...).  'mean' is sum / clamp(in_degree, 1); max/min give 0
    for nodes without in-edges.
    """
    gidx = _gidx(g)
    if op not in ("copy_lhs", "copy_rhs"):
        if lhs_data is None or rhs_data is None:
            raise DGLError("gspmm: op %s needs both operands" % op)
        lhs_data, rhs_data = reshape_lhs_rhs(lhs_data, rhs_data)
    if reduce_op == "mean":
        # sum with the IEEE divide by float(clamp(in_deg, 1)) fused into the kernel epilogue
        if gidx.n_edges == 0:
            return B.gspmm(gidx, op, "sum", lhs_data, rhs_data)
        return B.gspmm(gidx, op, "sum", lhs_data, rhs_data, gidx.csc().mean_divisor())
    # max / min: upstream replaces +-inf (nodes without in-edges) by 0 in a Python post-pass; here the
    # kernel's store does it (flag DGLB_SPMM_ZERO_INF), saving two full passes over the output
    return B.gspmm(gidx, op, reduce_op, lhs_data, rhs_data, zero_inf=reduce_op in ("min", "max"))


def _attach_shorthands():
    """copy_u_sum, u_mul_e_max, copy_e_mean, ... (upstream generates the same names)."""
    mod = sys.modules[__name__]
    for binary in ("add", "sub", "mul", "div"):
        for reduce_op in ("sum", "max", "min", "mean"):
            name = "u_{}_e_{}".format(binary, reduce_op)

            def fn(g, x, y, _b=binary, _r=reduce_op):
                return gspmm(g, _b, _r, x, y)
            fn.__name__ = name
            fn.__doc__ = "gspmm(g, '%s', '%s', x, y)" % (binary, reduce_op)
            setattr(mod, name, fn)
            __all__.append(name)
    for reduce_op in ("sum", "max", "min", "mean"):
        def cu(g, x, _r=reduce_op):
            return gspmm(g, "copy_lhs", _r, x, None)

        def ce(g, x, _r=reduce_op):
            return gspmm(g, "copy_rhs", _r, None, x)
        cu.__name__ = "copy_u_" + reduce_op
        ce.__name__ = "copy_e_" + reduce_op
        setattr(mod, cu.__name__, cu)
        setattr(mod, ce.__name__, ce)
        __all__.extend([cu.__name__, ce.__name__])


_attach_shorthands()
