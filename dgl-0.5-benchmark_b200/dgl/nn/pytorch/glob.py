"""Graph readout layers for batched graph classification (main_dgl_molhiv_gcn.py:75,93).
Upstream implements these with its segment_reduce kernels (out of scope for the
native path this round); here they are torch index ops over the per-graph node counts."""
import torch
from torch import nn


def _graph_ids(graph):
    counts = graph.batch_num_nodes()
    return torch.repeat_interleave(torch.arange(counts.shape[0], device=counts.device), counts), counts


class SumPooling(nn.Module):
    def forward(self, graph, feat):
        ids, counts = _graph_ids(graph)
        out = torch.zeros((counts.shape[0],) + tuple(feat.shape[1:]), dtype=feat.dtype, device=feat.device)
        return out.index_add_(0, ids.to(feat.device), feat)


class AvgPooling(nn.Module):
    def forward(self, graph, feat):
        ids, counts = _graph_ids(graph)
        out = torch.zeros((counts.shape[0],) + tuple(feat.shape[1:]), dtype=feat.dtype, device=feat.device)
        out = out.index_add_(0, ids.to(feat.device), feat)
        return out / counts.to(feat).clamp(min=1).view((-1,) + (1,) * (feat.dim() - 1))


class MaxPooling(nn.Module):
    def forward(self, graph, feat):
        ids, counts = _graph_ids(graph)
        out = torch.full((counts.shape[0],) + tuple(feat.shape[1:]), float("-inf"), dtype=feat.dtype, device=feat.device)
        idx = ids.to(feat.device).view((-1,) + (1,) * (feat.dim() - 1)).expand_as(feat)
        return out.scatter_reduce(0, idx, feat, "amax", include_self=True)
