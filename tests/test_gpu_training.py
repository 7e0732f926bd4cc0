"""-m gpu integration: a few epochs of full-graph SAGE / GAT training through the drop-in layers on the
GPU kernels reproduce the loss trajectory of the same model run on the CPU with the kernel-level calls
routed to the oracle ("loss-trajectory comparison CPU-oracle vs GPU"), and the fused
GAT path matches upstream's op-by-op composition."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import dgl
import oracle_backend
from conftest import make_edges
from examples.full_graph import GAT, GraphSAGE

pytestmark = pytest.mark.gpu


def _task(n=600, e=7000, d=24, c=5, seed=0):
    src, dst = make_edges(n, n, e, seed=seed, kind="powerlaw")
    src = np.concatenate([src, np.arange(n)])
    dst = np.concatenate([dst, np.arange(n)])
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int()
    gen = torch.Generator().manual_seed(seed)
    return g, torch.rand(n, d, generator=gen), torch.randint(0, c, (n,), generator=gen), torch.arange(0, n, 3)


def _train(model, g, x, y, idx, steps, log_softmax_out):
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        out = model(g, x)
        loss = F.nll_loss(out[idx], y[idx]) if log_softmax_out else F.cross_entropy(out[idx], y[idx])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return np.array(losses)


def test_sage_loss_trajectory_gpu_vs_cpu_oracle(oracle, cuda):
    g, x, y, idx = _task()
    torch.manual_seed(1)
    model = GraphSAGE(24, 16, 5, n_layers=2, aggr="mean", dropout=0.0)
    ref_model = copy.deepcopy(model)
    with oracle_backend.installed():
        want = _train(ref_model, g, x, y, idx, 8, False)
    got = _train(model.to(cuda), g.to(cuda), x.to(cuda), y.to(cuda), idx.to(cuda), 8, False)
    np.testing.assert_allclose(got, want, rtol=2e-4)
    assert got[-1] < got[0]


@pytest.mark.parametrize("heads", [[2, 2, 1], [4, 4, 4]])
def test_gat_loss_trajectory_fused_vs_unfused_vs_cpu_oracle(oracle, cuda, heads):
    from dgl.nn.pytorch import GATConv
    g, x, y, idx = _task(seed=2)
    torch.manual_seed(3)
    model = GAT(24, 8, 5, heads, feat_drop=0.0, attn_drop=0.0)
    ref_model, unf_model = copy.deepcopy(model), copy.deepcopy(model)
    with oracle_backend.installed():
        want = _train(ref_model, g, x, y, idx, 6, True)
    gd, xd, yd, idd = g.to(cuda), x.to(cuda), y.to(cuda), idx.to(cuda)
    assert GATConv.fused
    fused = _train(model.to(cuda), gd, xd, yd, idd, 6, True)
    try:
        GATConv.fused = False
        unfused = _train(unf_model.to(cuda), gd, xd, yd, idd, 6, True)
    finally:
        GATConv.fused = True
    np.testing.assert_allclose(unfused, want, rtol=3e-4)
    np.testing.assert_allclose(fused, want, rtol=3e-4)


def test_gat_with_attention_dropout_trains(cuda):
    g, x, y, idx = _task(seed=4)
    torch.manual_seed(5)
    model = GAT(24, 8, 5, [2, 2, 1], feat_drop=0.1, attn_drop=0.3).to(cuda)
    losses = _train(model, g.to(cuda), x.to(cuda), y.to(cuda), idx.to(cuda), 30, True)
    assert np.isfinite(losses).all() and losses[-5:].mean() < losses[:5].mean()
