"""Full-graph GraphSAGE / GAT training on synthetic graphs of the reference datasets' shapes, written
against the drop-in `dgl` package.  The models restate the reference scripts' architectures and
hyper-parameters (main_dgl_citation_sage.py:20-120, main_dgl_product_sage.py:33-100,
main_dgl_arxiv_gat.py:14-63, main_dgl_reddit_gat.py) so epoch times are comparable in shape with
README.md:36-46; the epoch timer follows the scripts (skip the first 3 epochs) but always
synchronises the device.  With a RowPartition the SAGE model trains row-partitioned on N GPUs
(all-gather of the layer input forward, of the output gradient backward; dense grads all-reduced).
"""
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import dgl
import dgl.function as fn
from dgl.data import synthetic
from dgl.nn.pytorch import GATConv


class SAGELayer(nn.Module):
    """h' = W_self h + W_neigh aggregate(h)  (aggregate BEFORE the projection, as the hand-written
    layer of main_dgl_citation_sage.py:44-86 does)."""

    # aggregate AFTER the neighbour projection when that makes the aggregated rows narrower (in > out): the mean / sum
    # is linear, so fc_neigh(aggregate(h)) == aggregate(h W^T) + b up to rounding (upstream's nn.SAGEConv does the same,
    # `lin_before_mp`); products layer 1 aggregates 64-wide instead of 100-wide rows, cora 16 instead of 1 433.
    # False reproduces the hand-written layer's order of operations exactly.
    project_first = True

    def __init__(self, in_feats, out_feats, aggr="mean", feat_drop=0.0, activation=None):
        super().__init__()
        self.aggr, self.activation = aggr, activation
        self.in_feats, self.out_feats = in_feats, out_feats
        self.feat_drop = nn.Dropout(feat_drop)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(in_feats, out_feats)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, graph, feat):
        h = self.feat_drop(feat)
        # project first when it shrinks the sparse work: with a trainable input both the forward and the backward
        # aggregation narrow from `in` to `out`; with a constant input (first layer) the input-width backward
        # aggregation would have been skipped anyway, so the projection has to more than halve the width to pay
        # (products layer 1, 100 -> 64: 5.2 ms before vs 2.7 + 2.7 ms after; reddit layer 1, 602 -> 16: 3.9 vs 0.2 ms)
        need = self.out_feats if h.requires_grad else 2 * self.out_feats
        if SAGELayer.project_first and self.in_feats > need:
            w = self.fc_neigh.weight
            pad = (-self.out_feats) % 4      # keep the aggregated rows 16-byte aligned (47 classes -> 48 columns)
            if pad:
                w = torch.nn.functional.pad(w, (0, 0, 0, pad))
            z = torch.nn.functional.linear(h, w)
            if hasattr(graph, "copy_u_sum"):
                zn = graph.copy_u_sum(z, self.aggr)
            else:
                g = graph.local_var()
                g.srcdata["h"] = z
                g.update_all(fn.copy_src("h", "m"), fn.mean("m", "neigh") if self.aggr == "mean" else fn.sum("m", "neigh"))
                zn = g.dstdata["neigh"]
            if pad:
                zn = zn[:, :self.out_feats]
            rst = self.fc_self(h) + zn + self.fc_neigh.bias
            return self.activation(rst) if self.activation is not None else rst
        if hasattr(graph, "copy_u_sum"):          # RowPartition: collective + local kernel
            h_neigh = graph.copy_u_sum(h, self.aggr)
        else:
            g = graph.local_var()
            g.srcdata["h"] = h
            g.update_all(fn.copy_src("h", "m"), fn.mean("m", "neigh") if self.aggr == "mean" else fn.sum("m", "neigh"))
            h_neigh = g.dstdata["neigh"]
        rst = self.fc_self(h) + self.fc_neigh(h_neigh)
        return self.activation(rst) if self.activation is not None else rst


class GraphSAGE(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers=2, aggr="mean", dropout=0.5):
        super().__init__()
        dims = [in_feats] + [n_hidden] * (n_layers - 1) + [n_classes]
        self.layers = nn.ModuleList(
            SAGELayer(dims[i], dims[i + 1], aggr, feat_drop=(dropout if i > 0 else 0.0),
                      activation=(F.relu if i < n_layers - 1 else None)) for i in range(n_layers))

    def forward(self, graph, h):
        for layer in self.layers:
            h = layer(graph, h)
        return h


class GAT(nn.Module):
    """3-layer GAT of main_dgl_arxiv_gat.py:14-63 (heads e.g. [4,4,4]; hidden layers flatten heads,
    the last layer averages them; log-softmax output)."""

    def __init__(self, in_feats, n_hidden, n_classes, heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2):
        super().__init__()
        n = len(heads)
        self.layers = nn.ModuleList()
        self.layers.append(GATConv(in_feats, n_hidden, heads[0], 0.0, 0.0, negative_slope, activation=F.elu))
        for l in range(n - 2):
            self.layers.append(GATConv(n_hidden * heads[l], n_hidden, heads[l + 1], feat_drop, attn_drop,
                                       negative_slope, activation=F.elu))
        self.layers.append(GATConv(n_hidden * heads[-2], n_classes, heads[-1], feat_drop, attn_drop, negative_slope))

    def forward(self, g, h):
        for layer in self.layers[:-1]:
            h = layer(g, h).flatten(1)
        return self.layers[-1](g, h).mean(1).log_softmax(dim=-1)


class PartGATLayer(nn.Module):
    """GATConv math (fc -> el/er -> fused attention) on a RowPartition: the layer input holds the rank's
    rows only; attention dropout seeds are derived from a step counter shared by all ranks."""

    def __init__(self, in_feats, out_feats, heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2, activation=None):
        super().__init__()
        self.H, self.F, self.slope, self.activation = heads, out_feats, negative_slope, activation
        self.fc = nn.Linear(in_feats, out_feats * heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, heads, out_feats))
        self.feat_drop, self.attn_p = nn.Dropout(feat_drop), attn_drop
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        self.calls = 0

    def forward(self, part, h):
        ft = self.fc(self.feat_drop(h)).view(-1, self.H, self.F)
        el = (ft * self.attn_l).sum(-1)
        er = (ft * self.attn_r).sum(-1)
        self.calls += 1
        p = self.attn_p if self.training else 0.0
        rst = part.gat_attention(ft, el, er, self.slope, p, seed=1000003 * self.calls + 17)
        return self.activation(rst) if self.activation is not None else rst


class PartGAT(nn.Module):
    """The GAT of main_dgl_arxiv_gat.py:14-63 on a row partition."""

    def __init__(self, in_feats, n_hidden, n_classes, heads, feat_drop=0.0, attn_drop=0.0):
        super().__init__()
        n = len(heads)
        self.layers = nn.ModuleList([PartGATLayer(in_feats, n_hidden, heads[0], 0.0, 0.0, activation=F.elu)])
        for l in range(n - 2):
            self.layers.append(PartGATLayer(n_hidden * heads[l], n_hidden, heads[l + 1], feat_drop, attn_drop, activation=F.elu))
        self.layers.append(PartGATLayer(n_hidden * heads[-2], n_classes, heads[-1], feat_drop, attn_drop))

    def forward(self, part, h):
        for layer in self.layers[:-1]:
            h = layer(part, h).flatten(1)
        return self.layers[-1](part, h).mean(1).log_softmax(dim=-1)


def synthetic_task(name, device, seed=0, degree="uniform", self_loops=False, edges=None):
    """(graph on device, features, labels, train index) with the shape of dataset `name`."""
    n, e, d, c = synthetic.SHAPES[name]
    if edges is not None:
        e = edges
    src, dst = synthetic.random_edges(n, n, e, seed=seed, degree=degree)
    if self_loops:
        loops = np.arange(n)
        src, dst = np.concatenate([src, loops]), np.concatenate([dst, loops])
    gen = torch.Generator().manual_seed(seed)
    feats = torch.rand(n, d, generator=gen)
    labels = torch.randint(0, c, (n,), generator=gen)
    train_idx = torch.randperm(n, generator=gen)[: max(1, n // 10)]
    return (n, src, dst), feats, labels, train_idx, c


def time_epochs(step, epochs, skip=3):
    """Mean wall time of step() over epochs after `skip` warm-up epochs (device synchronised)."""
    dur = []
    for ep in range(epochs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        if ep >= skip:
            dur.append(time.perf_counter() - t0)
    return float(np.mean(dur)), dur
