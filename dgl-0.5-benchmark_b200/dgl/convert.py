"""Graph construction from other libraries (dgl.from_networkx: main_dgl_citation_sage.py:190)."""
import torch

from ._capi import DGLError
from .heterograph import graph


def from_networkx(nx_graph, node_attrs=None, edge_attrs=None, idtype=None, device=None):
    """Nodes are relabelled 0..N-1 in sorted order; an undirected graph contributes both
    directions of every edge; edge ids follow networkx's edge iteration order."""
    import networkx as nx
    if node_attrs or edge_attrs:
        raise DGLError("node_attrs / edge_attrs are not supported")
    if not nx_graph.is_directed():
        nx_graph = nx_graph.to_directed()
    nodes = sorted(nx_graph.nodes())
    if nodes != list(range(len(nodes))):
        nx_graph = nx.relabel_nodes(nx_graph, {n: i for i, n in enumerate(nodes)})
    edges = list(nx_graph.edges())
    src = torch.tensor([e[0] for e in edges], dtype=torch.int64)
    dst = torch.tensor([e[1] for e in edges], dtype=torch.int64)
    return graph((src, dst), num_nodes=len(nodes), idtype=idtype, device=device)


def from_scipy(sp_mat, idtype=None, device=None):
    coo = sp_mat.tocoo()
    if coo.shape[0] != coo.shape[1]:
        raise DGLError("from_scipy expects a square matrix")
    return graph((torch.as_tensor(coo.row).long(), torch.as_tensor(coo.col).long()), num_nodes=coo.shape[0],
                 idtype=idtype, device=device)


def to_networkx(g):
    import networkx as nx
    src, dst = g.edges()
    nxg = nx.MultiDiGraph()
    nxg.add_nodes_from(range(g.number_of_nodes()))
    nxg.add_edges_from(zip(src.tolist(), dst.tolist()))
    return nxg
