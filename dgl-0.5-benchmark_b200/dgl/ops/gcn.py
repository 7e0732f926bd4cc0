"""Fused message passing of the graph-classification GCN layer (new relative to upstream, which runs the Python UDF
`message` of main_dgl_molhiv_gcn.py:50-52 with torch ops on (E, D) tensors and then update_all(copy_e, sum), :46)."""
from .. import backend as B
from .spmm import _gidx

__all__ = ["gcn_norm_relu_sum", "categorical_embedding_sum"]


def gcn_norm_relu_sum(graph, x, w, c_src, c_dst=None):
    r"""h[v] = \sum_{e=(u \to v)} (c_src[u] c_dst[v]) \, relu(x[u] + w[e])

    x (N_src, D) node data, w (E, D) edge data in edge-id order, c_src (N_src[,1]) / c_dst (N_dst[,1]) the per-node
    normalisation (deg^-1/2); c_dst defaults to c_src on a homogeneous graph.  Equals
    ``update_all(lambda e: {'m': e.src['c'] * e.dst['c'] * relu(e.src['x'] + e.data['w'])}, fn.sum('m', 'h'))``
    bit for bit in the forward; differentiable w.r.t. x and w."""
    gidx = _gidx(graph)
    if c_dst is None:
        c_dst = c_src
    return B.gcn_msg_sum(gidx, x, w, c_src.reshape(-1), c_dst.reshape(-1))


def categorical_embedding_sum(x, table, dims):
    r"""out[i] = \sum_k table_k[x[i, k]] for the stacked tables `table` ((sum(dims), D); table_k holds dims[k] rows):
    the AtomEncoder / BondEncoder of ogb.graphproppred.mol_encoder (main_dgl_molhiv_gcn.py:28,72) in one kernel, with a
    deterministic backward.  x: (n, len(dims)) int64 codes."""
    offsets = [0]
    for d in dims:
        offsets.append(offsets[-1] + int(d))
    return B.CatEmbedSum.apply(x, table, offsets)
