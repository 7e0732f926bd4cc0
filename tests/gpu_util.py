"""Helpers shared by the -m gpu parity tests: the same seeded graph as an oracle graph (CPU) and a
dgl graph on the CUDA device."""
import numpy as np
import torch

import dgl
from conftest import make_edges


def graphs(oracle, n_src, n_dst, n_edges, seed, kind="uniform", order="shuffled", device="cuda", self_loops=False):
    src, dst = make_edges(n_src, n_dst, n_edges, seed, kind, order)
    if self_loops:
        assert n_src == n_dst
        src = np.concatenate([src, np.arange(n_src)])
        dst = np.concatenate([dst, np.arange(n_dst)])
    og = oracle.OracleGraph(src, dst, n_src, n_dst)
    if n_src == n_dst:
        g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n_src).int().to(device)
    else:
        g = dgl.create_block((torch.from_numpy(src), torch.from_numpy(dst)), n_src, n_dst).int().to(device)
    return og, g, src, dst


def t(a, device="cuda"):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)


def n(x):
    return x.detach().cpu().numpy()


def abs_sum_scale_spmm(src, dst, n_dst, msg_abs):
    """sum over in-edges of |message| per destination (the scale of the 1e-5 bound)."""
    out = np.zeros((n_dst,) + msg_abs.shape[1:], np.float64)
    np.add.at(out, dst, msg_abs.astype(np.float64))
    return out
