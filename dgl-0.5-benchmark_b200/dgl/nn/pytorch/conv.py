"""Graph convolution layers with the constructor / forward signatures of upstream DGL v0.6.1
(python/dgl/nn/pytorch/conv/{gatconv,sageconv,graphconv}.py).  The dense parts (Linear, bias,
activation) stay on torch/cuBLAS; the message passing goes through dgl.ops into the sm_100a kernels.
GATConv's attention (u_add_v -> leaky_relu -> edge_softmax -> attn_drop -> u_mul_e_sum) runs as the
fused kernel; `GATConv.fused = False` selects the op-by-op composition upstream uses (kept as the
parity reference for the fused path).
"""
import torch
from torch import nn

from ... import function as fn
from ... import ops
from ..._capi import DGLError
from ...utils import expand_as_pair


def _has_zero_in_degree(graph):
    cache = graph._graph._c
    key = "zero_in_deg_rev" if graph._graph._rev else "zero_in_deg"
    if key not in cache:
        cache[key] = bool((graph.in_degrees() == 0).any().item())
    return cache[key]


_ZERO_DEG_MSG = ("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                 "This is harmful for some applications, causing silent performance regression. "
                 "Adding self-loop on the input graph by calling `g = dgl.add_self_loop(g)` will resolve "
                 "the issue. Setting ``allow_zero_in_degree`` to be `True` when constructing this module "
                 "will suppress the check and let the code run.")


class Identity(nn.Module):
    def forward(self, x):
        return x


class GATConv(nn.Module):
    r"""Graph attention layer (Velickovic et al.):  h_i' = sum_j alpha_ij W h_j with
    alpha_ij = softmax_i(LeakyReLU(a^T [W h_i || W h_j]))."""

    fused = True  # class-wide switch; set False to run upstream's unfused composition
    # Fused path only: compute el / er as two skinny GEMMs h @ (W^T a_l), h @ (W^T a_r) instead of the elementwise
    # product + reduction over ft upstream does (el = (ft * attn_l).sum(-1)), and let the projection GEMM emit ft with
    # the per-head width already padded to a multiple of 4.  Same math, different association (~1e-6 relative): on the
    # arxiv / products GAT epochs the product + reduction and the padding copy were 28 % / 21 % of the device time
    # (profiles/r02_epoch_*.txt).  False restores upstream's order of operations exactly.
    fold_attention = True

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=False):
        super().__init__()
        self._num_heads = num_heads
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self._negative_slope = negative_slope
        if isinstance(in_feats, tuple):
            self.fc_src = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=False)
            self.fc_dst = nn.Linear(self._in_dst_feats, out_feats * num_heads, bias=False)
        else:
            self.fc = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.attn_r = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        if residual:
            if self._in_dst_feats != out_feats:
                self.res_fc = nn.Linear(self._in_dst_feats, num_heads * out_feats, bias=False)
            else:
                self.res_fc = Identity()
        else:
            self.register_buffer("res_fc", None)
        self.reset_parameters()
        self.activation = activation

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if hasattr(self, "fc"):
            nn.init.xavier_normal_(self.fc.weight, gain=gain)
        else:
            nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
            nn.init.xavier_normal_(self.fc_dst.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, get_attention=False):
        if not self._allow_zero_in_degree and _has_zero_in_degree(graph):
            raise DGLError(_ZERO_DEG_MSG)
        H, F = self._num_heads, self._out_feats
        if (GATConv.fused and GATConv.fold_attention and H <= 8 and not get_attention and not isinstance(feat, tuple)
                and hasattr(self, "fc")):
            return self._forward_folded(graph, feat)
        if isinstance(feat, tuple):
            h_src = self.feat_drop(feat[0])
            h_dst = self.feat_drop(feat[1])
            if not hasattr(self, "fc_src"):
                feat_src = self.fc(h_src).view(-1, H, F)
                feat_dst = self.fc(h_dst).view(-1, H, F)
            else:
                feat_src = self.fc_src(h_src).view(-1, H, F)
                feat_dst = self.fc_dst(h_dst).view(-1, H, F)
        else:
            h_src = h_dst = self.feat_drop(feat)
            feat_src = feat_dst = self.fc(h_src).view(-1, H, F)
            if graph.is_block:
                feat_dst = feat_src[:graph.number_of_dst_nodes()]
                h_dst = h_dst[:graph.number_of_dst_nodes()]
        el = (feat_src * self.attn_l).sum(dim=-1).unsqueeze(-1)
        er = (feat_dst * self.attn_r).sum(dim=-1).unsqueeze(-1)
        a = None
        if GATConv.fused and H <= 8 and not get_attention:
            p = self.attn_drop.p if self.training else 0.0
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0  # CPU generator: no device sync
            rst = ops.gat_attention(graph, feat_src, el, er, self._negative_slope, p, seed)
        else:
            e = self.leaky_relu(ops.u_add_v(graph, el, er))
            a = self.attn_drop(ops.edge_softmax(graph, e))
            rst = ops.u_mul_e_sum(graph, feat_src, a)
        if self.res_fc is not None:
            rst = rst + self.res_fc(h_dst).view(h_dst.shape[0], -1, F)
        if self.activation:
            rst = self.activation(rst)
        return (rst, a) if get_attention else rst


    def _forward_folded(self, graph, feat):
        from ... import backend as B
        H, F = self._num_heads, self._out_feats
        h_src = h_dst = self.feat_drop(feat)
        if graph.is_block:
            h_dst = h_src[:graph.number_of_dst_nodes()]
        W = self.fc.weight.view(H, F, -1)                                        # (H, F, in)
        Fp = (F + 3) // 4 * 4
        Wp = W if Fp == F else torch.nn.functional.pad(W, (0, 0, 0, Fp - F))     # zero rows -> zero padded columns
        ft = torch.matmul(h_src, Wp.reshape(H * Fp, -1).t()).view(-1, H, Fp)
        el = torch.matmul(h_src, (W * self.attn_l.view(H, F, 1)).sum(1).t())     # (N_src, H)
        er = torch.matmul(h_dst, (W * self.attn_r.view(H, F, 1)).sum(1).t())     # (N_dst, H)
        p = self.attn_drop.p if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0      # CPU generator: no device sync
        rst = B.gat_fused(ops.spmm._gidx(graph), ft, el, er, self._negative_slope, p, seed)
        if Fp != F:
            rst = rst[..., :F]
        if self.res_fc is not None:
            rst = rst + self.res_fc(h_dst).view(h_dst.shape[0], -1, F)
        if self.activation:
            rst = self.activation(rst)
        return rst


class SAGEConv(nn.Module):
    r"""GraphSAGE layer: h_i' = W_self h_i + W_neigh aggregate({h_j}).  Aggregators: mean, gcn, pool
    ('lstm' is not provided: it is not a sparse-kernel path and no in-scope script uses it)."""

    def __init__(self, in_feats, out_feats, aggregator_type, feat_drop=0., bias=True, norm=None, activation=None):
        super().__init__()
        if aggregator_type not in ("mean", "gcn", "pool"):
            raise DGLError("Invalid aggregator_type. Must be one of mean, gcn, pool. But got {!r}."
                           .format(aggregator_type))
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._aggre_type = aggregator_type
        self.norm = norm
        self.feat_drop = nn.Dropout(feat_drop)
        self.activation = activation
        if aggregator_type == "pool":
            self.fc_pool = nn.Linear(self._in_src_feats, self._in_src_feats)
        if aggregator_type != "gcn":
            self.fc_self = nn.Linear(self._in_dst_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(self._in_src_feats, out_feats, bias=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if self._aggre_type == "pool":
            nn.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        if self._aggre_type != "gcn":
            nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, graph, feat):
        graph = graph.local_var()
        if isinstance(feat, tuple):
            feat_src = self.feat_drop(feat[0])
            feat_dst = self.feat_drop(feat[1])
        else:
            feat_src = feat_dst = self.feat_drop(feat)
            if graph.is_block:
                feat_dst = feat_src[:graph.number_of_dst_nodes()]
        h_self = feat_dst
        if graph.number_of_edges() == 0:
            graph.dstdata["neigh"] = torch.zeros(feat_dst.shape[0], self._in_src_feats, device=feat_dst.device)
        # aggregate AFTER the projection when that narrows the rows the SpMM has to gather
        lin_before_mp = self._in_src_feats > self._out_feats
        if self._aggre_type == "mean":
            graph.srcdata["h"] = self.fc_neigh(feat_src) if lin_before_mp else feat_src
            graph.update_all(fn.copy_u("h", "m"), fn.mean("m", "neigh"))
            h_neigh = graph.dstdata["neigh"]
            if not lin_before_mp:
                h_neigh = self.fc_neigh(h_neigh)
        elif self._aggre_type == "gcn":
            graph.srcdata["h"] = self.fc_neigh(feat_src) if lin_before_mp else feat_src
            graph.dstdata["h"] = (self.fc_neigh(feat_dst) if lin_before_mp else feat_dst) if graph.is_block \
                else graph.srcdata["h"]
            graph.update_all(fn.copy_u("h", "m"), fn.sum("m", "neigh"))
            degs = graph.in_degrees().to(feat_dst)
            h_neigh = (graph.dstdata["neigh"] + graph.dstdata["h"]) / (degs.unsqueeze(-1) + 1)
            if not lin_before_mp:
                h_neigh = self.fc_neigh(h_neigh)
        else:  # pool
            graph.srcdata["h"] = torch.relu(self.fc_pool(feat_src))
            graph.update_all(fn.copy_u("h", "m"), fn.max("m", "neigh"))
            h_neigh = self.fc_neigh(graph.dstdata["neigh"])
        rst = h_neigh if self._aggre_type == "gcn" else self.fc_self(h_self) + h_neigh
        if self.bias is not None:
            rst = rst + self.bias
        if self.activation is not None:
            rst = self.activation(rst)
        if self.norm is not None:
            rst = self.norm(rst)
        return rst


class GraphConv(nn.Module):
    r"""GCN layer (Kipf & Welling): h' = D_in^{-1/2} A D_out^{-1/2} h W  (norm='both')."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm not in ("none", "both", "right"):
            raise DGLError('Invalid norm value. Must be either "none", "both" or "right". But got "{}".'.format(norm))
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        if weight:
            self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        else:
            self.register_parameter("weight", None)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, weight=None):
        if not self._allow_zero_in_degree and _has_zero_in_degree(graph):
            raise DGLError(_ZERO_DEG_MSG)
        graph = graph.local_var()
        feat_src, feat_dst = expand_as_pair(feat, graph)
        if self._norm == "both":
            degs = graph.out_degrees().to(feat_src).clamp(min=1)
            norm = torch.pow(degs, -0.5).view((-1,) + (1,) * (feat_src.dim() - 1))
            feat_src = feat_src * norm
        if weight is not None and self.weight is not None:
            raise DGLError("External weight is provided while at the same time the module has defined its own "
                           "weight parameter. Please create the module with flag weight=False.")
        weight = self.weight if weight is None else weight
        if self._in_feats > self._out_feats:
            if weight is not None:
                feat_src = torch.matmul(feat_src, weight)
            graph.srcdata["h"] = feat_src
            graph.update_all(fn.copy_src("h", "m"), fn.sum("m", "h"))
            rst = graph.dstdata["h"]
        else:
            graph.srcdata["h"] = feat_src
            graph.update_all(fn.copy_src("h", "m"), fn.sum("m", "h"))
            rst = graph.dstdata["h"]
            if weight is not None:
                rst = torch.matmul(rst, weight)
        if self._norm != "none":
            degs = graph.in_degrees().to(feat_dst).clamp(min=1)
            norm = torch.pow(degs, -0.5) if self._norm == "both" else 1.0 / degs
            rst = rst * norm.view((-1,) + (1,) * (feat_dst.dim() - 1))
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst
