"""dgl.ops.gsddmm -- generalized SDDMM (same signature as upstream DGL v0.6.1
python/dgl/ops/sddmm.py; called by kernel/dgl-new.py:39)."""
import sys

from .. import backend as B
from .._capi import DGLError
from .spmm import _gidx, reshape_lhs_rhs

__all__ = ["gsddmm", "copy_u", "copy_v", "copy_e"]


def gsddmm(g, op, lhs_data, rhs_data, lhs_target="u", rhs_target="v"):
    r"""Generalized Sampled-Dense-Dense Matrix Multiplication: for every edge compute
    ``op(lhs_data[sel(lhs_target)], rhs_data[sel(rhs_target)])`` where a target is the edge's source
    node ('u'), the edge itself ('e') or its destination node ('v').

    op : 'add' | 'sub' | 'mul' | 'div' | 'dot' | 'copy_lhs' | 'copy_rhs'
    Returns a tensor of shape (E, ...) in edge-id order ('dot' reduces the last dim to size 1).
    """
    gidx = _gidx(g)
    if op not in ("copy_lhs", "copy_rhs"):
        if lhs_data is None or rhs_data is None:
            raise DGLError("gsddmm: op %s needs both operands" % op)
        lhs_data, rhs_data = reshape_lhs_rhs(lhs_data, rhs_data)
    return B.gsddmm(gidx, op, lhs_data, rhs_data, lhs_target, rhs_target)


def copy_u(g, x):
    return gsddmm(g, "copy_lhs", x, None)


def copy_v(g, x):
    return gsddmm(g, "copy_rhs", None, x)


def copy_e(g, x):
    return x


def _attach_shorthands():
    """u_add_v, u_dot_v, e_mul_v, v_sub_u, ..."""
    mod = sys.modules[__name__]
    for lhs in ("u", "v", "e"):
        for rhs in ("u", "v", "e"):
            if lhs == rhs:
                continue
            for binary in ("add", "sub", "mul", "div", "dot"):
                name = "{}_{}_{}".format(lhs, binary, rhs)

                def fn(g, x, y, _b=binary, _l=lhs, _r=rhs):
                    return gsddmm(g, _b, x, y, lhs_target=_l, rhs_target=_r)
                fn.__name__ = name
                fn.__doc__ = "gsddmm(g, '%s', x, y, '%s', '%s')" % (binary, lhs, rhs)
                setattr(mod, name, fn)
                __all__.append(name)


_attach_shorthands()
