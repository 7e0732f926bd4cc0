"""The GCN of main_dgl_molhiv_gcn.py:20-93 (AtomEncoder -> L x [Linear, in-degree norm, message = norm * relu(x_src +
bond_emb), sum, root term] with BatchNorm / ReLU / dropout between layers -> mean readout -> Linear) in three forms that
share parameters and arithmetic:

  * `fused=False`  the script's own formulation: a Python message UDF on (E, D) tensors + update_all(copy_e, sum);
  * `fused=True`   dgl.ops.gcn_norm_relu_sum (one forward / one backward kernel per layer, no (E, D) message tensor);
  * `padded=True`  for dgl.StaticBatch graphs: BatchNorm statistics run over the real nodes only (mask + count), so a
                   padded fixed-size batch gives the same numbers as the unpadded one and the step can be captured in
                   a CUDA graph.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import dgl
import dgl.function as fn
from dgl.nn import AvgPooling
from ogb.graphproppred.mol_encoder import AtomEncoder, BondEncoder


class FusedEncoder(nn.Module):
    """AtomEncoder / BondEncoder with the per-column tables stacked into one parameter and the lookup-and-sum done by
    dgl.ops.categorical_embedding_sum (one kernel forward, one deterministic kernel backward instead of ~40 launches)."""

    def __init__(self, dims, dim):
        super().__init__()
        self.dims = list(dims)
        self.weight = nn.Parameter(torch.empty(sum(self.dims), dim))
        lo = 0
        for d in self.dims:
            nn.init.xavier_uniform_(self.weight.data[lo:lo + d])
            lo += d

    def load_columns(self, embeddings):
        """Copy the tables of an ogb-style encoder (ModuleList of nn.Embedding)."""
        with torch.no_grad():
            self.weight.copy_(torch.cat([e.weight for e in embeddings], 0))

    def forward(self, x):
        return dgl.ops.categorical_embedding_sum(x, self.weight, self.dims)


ATOM_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_DIMS = [5, 6, 2]


class GCNLayer(nn.Module):
    def __init__(self, dim, fused=False):
        super().__init__()
        self.fc = nn.Linear(dim, dim, bias=False)
        self.root_emb = nn.Embedding(1, dim)
        self.bond_encoder = FusedEncoder(BOND_DIMS, dim) if fused else BondEncoder(dim)
        self.fused = fused

    def forward(self, g, feat, bond, deg=None):
        g = g.local_var()
        x = self.fc(feat)
        if deg is None:
            deg = g.in_degrees().float().unsqueeze(1) + 1
        c = deg.pow(-0.5)
        w = self.bond_encoder(bond)
        if self.fused:
            h = dgl.ops.gcn_norm_relu_sum(g, x, w, c)
        else:
            g.ndata["c"], g.ndata["x"], g.edata["w"] = c, x, w
            g.update_all(lambda e: {"m": e.src["c"] * e.dst["c"] * F.relu(e.src["x"] + e.data["w"])}, fn.sum("m", "h"))
            h = g.ndata["h"]
        return h + F.relu(x + self.root_emb.weight) * 1. / deg


class MaskedBatchNorm1d(nn.BatchNorm1d):
    """BatchNorm1d whose batch statistics run over the rows with mask == 1 (`count` of them, a 0-d device tensor):
    same mean / biased variance / running-stat updates as nn.BatchNorm1d applied to those rows alone."""

    def forward(self, x, mask=None, count=None):
        if mask is None:
            return super().forward(x)
        if self.training:
            mean = (x * mask).sum(0) / count
            d = (x - mean) * mask
            var = (d * d).sum(0) / count
            with torch.no_grad():
                m = self.momentum
                self.running_mean.mul_(1 - m).add_(mean.detach() * m)
                self.running_var.mul_(1 - m).add_(var.detach() * (count / (count - 1).clamp(min=1)) * m)
                self.num_batches_tracked.add_(1)
        else:
            mean, var = self.running_mean, self.running_var
        return (x - mean) * torch.rsqrt(var + self.eps) * self.weight + self.bias


class GCN(nn.Module):
    def __init__(self, dim=256, layers=5, dropout=0.5, fused=False):
        super().__init__()
        self.atom = FusedEncoder(ATOM_DIMS, dim) if fused else AtomEncoder(dim)
        self.fused = fused
        self.layers = nn.ModuleList(GCNLayer(dim, fused) for _ in range(layers))
        self.norms = nn.ModuleList(MaskedBatchNorm1d(dim) for _ in range(layers - 1))
        self.pool = AvgPooling()
        self.out = nn.Linear(dim, 1, bias=False)
        self.dropout = dropout

    def forward(self, g, atom, bond, mask=None, count=None):
        h = self.atom(atom)
        # the fused form computes the degree term once per batch (the script recomputes it in every layer)
        deg = (g.in_degrees().float().unsqueeze(1) + 1) if self.fused else None
        for i, layer in enumerate(self.layers):
            h = layer(g, h, bond, deg)
            if i < len(self.layers) - 1:
                h = F.dropout(F.relu(self.norms[i](h, mask, count)), self.dropout, self.training)
        return self.out(self.pool(g, h))


def copy_parameters(ref, fast):
    """Load the parameters of an unfused GCN into a fused one (stacked encoder tables)."""
    sd = {k: v for k, v in ref.state_dict().items() if ".embs." not in k}
    fast.load_state_dict(sd, strict=False)
    fast.atom.load_columns(ref.atom.embs)
    for lr, lf in zip(ref.layers, fast.layers):
        lf.bond_encoder.load_columns(lr.bond_encoder.embs)
