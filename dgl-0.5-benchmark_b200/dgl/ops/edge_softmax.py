"""dgl.ops.edge_softmax (same signature as upstream DGL v0.6.1 python/dgl/ops/edge_softmax.py)."""
from .. import backend as B
from .spmm import _gidx

__all__ = ["edge_softmax"]

ALL = None


def edge_softmax(graph, logits, eids=ALL, norm_by="dst"):
    r"""Softmax of the edge values over the edges that share a destination (``norm_by='dst'``) or a
    source (``'src'``):  a_ij = exp(z_ij) / sum_{k in N(i)} exp(z_ik).  `logits` is (E, *, 1) or
    (E, *) in edge-id order; the result has the same shape.  `eids` (a tensor of edge ids) restricts the softmax to
    the edge-induced subgraph: `logits` then has one row per listed edge, in the order of `eids`."""
    return B.edge_softmax(_gidx(graph), logits, eids=eids, norm_by=norm_by)
