"""Fused GAT attention (new relative to upstream): one forward kernel for
u_add_v -> leaky_relu -> edge_softmax -> attn_drop -> u_mul_e_sum; used by dgl.nn.pytorch.GATConv."""
from .. import backend as B
from .spmm import _gidx


def gat_attention(graph, ft, el, er, negative_slope=0.2, dropout_p=0.0, seed=0):
    """ft (N_src,H,F), el (N_src,H[,1]), er (N_dst,H[,1]) -> (N_dst,H,F)."""
    H = ft.shape[1]
    return B.gat_fused(_gidx(graph), ft, el.reshape(-1, H), er.reshape(-1, H), negative_slope, dropout_p, seed)
