"""Edge-list transforms used before training (integer-only, run once on the host or device with
torch index ops): dgl.add_self_loop / to_bidirected (main_dgl_arxiv_gat.py:130-131),
remove_self_loop, add_reverse_edges, reverse.  Semantics follow upstream DGL v0.6.1
python/dgl/transform.py: self loops are APPENDED after the existing edges (so they get the largest
edge ids), to_bidirected = add the reverse edges, then drop duplicate (src, dst) pairs.
"""
import torch

from ._capi import DGLError
from .graph_index import GraphIndex
from .heterograph import DGLHeteroGraph, Frame


def _new_graph(g, src, dst, copy_ndata=True, edge_frame=None):
    gi = g._graph
    new = DGLHeteroGraph(GraphIndex(src.contiguous(), dst.contiguous(), gi.n_src, gi.n_dst, gi.idtype))
    if copy_ndata:
        new._src_frame = new._dst_frame = g._src_frame.clone()
    if edge_frame is not None:
        new._edge_frame = edge_frame
    return new


def add_self_loop(g):
    """Append one (i, i) edge per node after the existing edges.  Existing edge features are kept
    and zero-padded for the new edges."""
    if g.is_block:
        raise DGLError("add_self_loop expects a graph with one node set")
    gi = g._graph
    loops = torch.arange(gi.n_src, dtype=gi.idtype, device=gi.device)
    ef = Frame(gi.n_edges + gi.n_src)
    for k, v in g.edata.items():
        ef[k] = torch.cat([v, v.new_zeros((gi.n_src,) + tuple(v.shape[1:]))], 0)
    return _new_graph(g, torch.cat([gi.src, loops]), torch.cat([gi.dst, loops]), True, ef)


def remove_self_loop(g):
    gi = g._graph
    keep = gi.src != gi.dst
    ef = Frame(int(keep.sum().item()))
    for k, v in g.edata.items():
        ef[k] = v[keep]
    return _new_graph(g, gi.src[keep], gi.dst[keep], True, ef)


def add_reverse_edges(g, copy_ndata=True, copy_edata=False):
    gi = g._graph
    ef = None
    if copy_edata:
        ef = Frame(2 * gi.n_edges)
        for k, v in g.edata.items():
            ef[k] = torch.cat([v, v], 0)
    return _new_graph(g, torch.cat([gi.src, gi.dst]), torch.cat([gi.dst, gi.src]), copy_ndata, ef)


def to_bidirected(g, copy_ndata=False):
    """Add the reverse of every edge and keep one copy of each distinct (src, dst) pair
    (edges end up ordered by (src, dst))."""
    gi = g._graph
    if gi.n_src != gi.n_dst:
        raise DGLError("to_bidirected expects a graph with one node set")
    src = torch.cat([gi.src, gi.dst]).long()
    dst = torch.cat([gi.dst, gi.src]).long()
    key = torch.unique(src * gi.n_dst + dst)  # sorted, exact (int64)
    return _new_graph(g, (key // gi.n_dst).to(gi.idtype), (key % gi.n_dst).to(gi.idtype), copy_ndata)


def reverse(g, copy_ndata=True, copy_edata=False):
    return g.reverse(copy_ndata=copy_ndata, copy_edata=copy_edata)
