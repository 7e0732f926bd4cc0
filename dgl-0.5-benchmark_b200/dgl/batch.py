"""dgl.batch -- disjoint union of small graphs (graph classification: main_dgl_molhiv_gcn.py:163
through GraphDataLoader).  Node / edge ids of graph i are shifted by the sizes of graphs 0..i-1;
features are concatenated; per-graph sizes are kept for the readout."""
import torch

from ._capi import DGLError
from .graph_index import GraphIndex
from .heterograph import DGLHeteroGraph, Frame


def batch(graphs, ndata=None, edata=None):
    if len(graphs) == 0:
        raise DGLError("The input list of graphs cannot be empty.")
    idtype, device = graphs[0].idtype, graphs[0].device
    n_nodes = torch.tensor([g.number_of_nodes() for g in graphs], dtype=torch.int64)
    n_edges = torch.tensor([g.number_of_edges() for g in graphs], dtype=torch.int64)
    offsets = torch.cumsum(n_nodes, 0) - n_nodes
    src = torch.cat([g._graph.src.to(torch.int64) + int(o) for g, o in zip(graphs, offsets)]).to(idtype)
    dst = torch.cat([g._graph.dst.to(torch.int64) + int(o) for g, o in zip(graphs, offsets)]).to(idtype)
    total = int(n_nodes.sum())
    bg = DGLHeteroGraph(GraphIndex(src, dst, total, total, idtype))
    for k in graphs[0].ndata.keys():
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata.keys():
        bg.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    bg._batch_max_nodes = int(n_nodes.max())
    bg._batch_num_nodes = n_nodes.to(device)
    bg._batch_num_edges = n_edges.to(device)
    return bg
