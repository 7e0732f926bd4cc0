"""Fused GAT attention (new relative to upstream): one forward kernel for
u_add_v -> leaky_relu -> edge_softmax -> attn_drop -> u_mul_e_sum; used by dgl.nn.pytorch.GATConv."""
import torch.nn.functional as F_

from .. import backend as B
from .spmm import _gidx


def pad_head_dim(ft):
    """Zero-pad the per-head width to a multiple of 4 so the kernels gather with 128-bit loads
    (F = 41 / 47 / 7 class-count layers would otherwise fall back to 4-byte loads; the copy is one
    cheap pass over ft, the padded columns contribute zeros and are sliced off the result)."""
    F = ft.shape[-1]
    Fp = (F + 3) // 4 * 4
    return (ft if Fp == F else F_.pad(ft, (0, Fp - F))), F


def gat_attention(graph, ft, el, er, negative_slope=0.2, dropout_p=0.0, seed=0):
    """ft (N_src,H,F), el (N_src,H[,1]), er (N_dst,H[,1]) -> (N_dst,H,F)."""
    H = ft.shape[1]
    ftp, F = pad_head_dim(ft)
    rst = B.gat_fused(_gidx(graph), ftp, el.reshape(-1, H), er.reshape(-1, H), negative_slope, dropout_p, seed)
    return rst if rst.shape[-1] == F else rst[..., :F]
