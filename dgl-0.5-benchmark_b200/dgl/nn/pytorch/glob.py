"""Graph readout layers for batched graph classification (main_dgl_molhiv_gcn.py:75,93; main_dgl_enzymes_gcn.py).

Upstream implements these with its segment_reduce kernels (dmlc/dgl@0.6.1 src/array/cuda/segment_reduce.cu, SURVEY.md
2.2 U16).  A per-graph sum / mean / max over the graph's nodes is exactly gspmm(copy_lhs, reduce) on the bipartite
membership relation "node -> the member graph it belongs to", whose CSC (rows = member graphs, entries = their nodes in
order) and CSR (rows = nodes, one entry each) can be written down directly from the batch's node counts -- no sort.
So the readouts run on the same sm_100a row kernels as every other aggregation (deterministic, hub segments for very
large member graphs, fused mean divide) and their backward is the reverse-graph gspmm of dgl/backend.py, instead of
torch index_add_ / scatter_reduce with atomics.
"""
import torch
from torch import nn

from ... import backend as B
from ...graph_index import CSRView, GraphIndex


def _membership(graph):
    """(GraphIndex of the nodes -> member graphs relation, float node counts clamped to >= 1); cached on the graph."""
    if graph._readout_index is None:
        counts = graph.batch_num_nodes()
        dev = counts.device
        n_graphs, n_nodes = int(counts.shape[0]), graph.number_of_nodes()
        nodes = torch.arange(n_nodes, dtype=torch.int32, device=dev)
        ids = torch.repeat_interleave(torch.arange(n_graphs, dtype=torch.int32, device=dev), counts, output_size=n_nodes)
        gi = GraphIndex(nodes, ids, n_nodes, n_graphs, torch.int32)
        indptr = torch.zeros(n_graphs + 1, dtype=torch.int32, device=dev)
        indptr[1:] = torch.cumsum(counts, 0)
        csc = CSRView(n_graphs, n_nodes, indptr, nodes, None)
        csc.max_deg = graph._batch_max_nodes if graph._batch_max_nodes is not None else None
        csr = CSRView(n_nodes, n_graphs, torch.arange(n_nodes + 1, dtype=torch.int32, device=dev), ids, None)
        csr.max_deg = 1
        gi._c["csc"], gi._c["csr"], gi._c["coo32"] = csc, csr, (nodes, ids)
        gi._c["dst_sorted"] = gi._c["src_sorted"] = True
        graph._readout_index = (gi, counts.clamp(min=1).to(torch.float32))
    return graph._readout_index


class SumPooling(nn.Module):
    def forward(self, graph, feat):
        gi, _ = _membership(graph)
        return B.gspmm(gi, "copy_lhs", "sum", feat, None)


class AvgPooling(nn.Module):
    def forward(self, graph, feat):
        gi, counts = _membership(graph)
        return B.gspmm(gi, "copy_lhs", "sum", feat, None, counts)     # mean: IEEE divide fused into the kernel's store


class MaxPooling(nn.Module):
    def forward(self, graph, feat):
        gi, _ = _membership(graph)
        return B.gspmm(gi, "copy_lhs", "max", feat, None)            # member graphs without nodes: -inf, as before
