"""-m gpu parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle would
take too long at 11.6 M edges x 602 features): exact integer identities, linearity, checksums and
agreement with independent torch scatter ops, on the reddit-shaped benchmark graph itself."""
import numpy as np
import pytest
import torch

import dgl
from dgl.data import synthetic

pytestmark = pytest.mark.gpu

N, E = 232965, 11606919


@pytest.fixture(scope="module")
def reddit(cuda):
    src, dst = synthetic.random_edges(N, N, E, seed=0)
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=N).int().to(cuda)
    return g, torch.from_numpy(src).to(cuda), torch.from_numpy(dst).to(cuda)


def test_structure_matches_torch_sort(reddit):
    g, src, dst = reddit
    csc = g._graph.csc()
    perm = torch.sort(dst, stable=True).indices          # independent stable sort
    assert torch.equal(csc.eids.long(), perm)
    assert torch.equal(csc.indices.long(), src[perm])
    assert torch.equal(csc.indptr.long()[1:], torch.cumsum(torch.bincount(dst, minlength=N), 0))
    assert torch.equal(g.in_degrees().long(), torch.bincount(dst, minlength=N))
    assert torch.equal(g.out_degrees().long(), torch.bincount(src, minlength=N))


@pytest.mark.parametrize("D", [64, 602])
def test_copy_u_sum_of_ones_is_the_in_degree(reddit, D):
    g, src, dst = reddit
    out = dgl.ops.gspmm(g, "copy_lhs", "sum", torch.ones(N, D, device=src.device), None)
    deg = torch.bincount(dst, minlength=N).float()
    assert torch.equal(out, deg[:, None].expand(N, D))           # small integers: exact in fp32
    mean = dgl.ops.gspmm(g, "copy_lhs", "mean", torch.ones(N, D, device=src.device), None)
    assert torch.equal(mean, (deg > 0).float()[:, None].expand(N, D))


def test_copy_u_sum_integer_features_match_index_add_exactly(reddit):
    g, src, dst = reddit
    X = torch.randint(0, 8, (N, 64), device=src.device).float()   # sums < 2^24: order-independent, exact
    want = torch.zeros(N, 64, device=src.device).index_add_(0, dst, X[src])
    assert torch.equal(dgl.ops.gspmm(g, "copy_lhs", "sum", X, None), want)


def test_copy_u_sum_linearity_and_checksum(reddit):
    g, src, dst = reddit
    torch.manual_seed(0)
    A, B = torch.rand(N, 128, device=src.device), torch.rand(N, 128, device=src.device)
    ya = dgl.ops.gspmm(g, "copy_lhs", "sum", A, None)
    yb = dgl.ops.gspmm(g, "copy_lhs", "sum", B, None)
    yab = dgl.ops.gspmm(g, "copy_lhs", "sum", A + 2 * B, None)
    torch.testing.assert_close(yab, ya + 2 * yb, rtol=1e-5, atol=1e-4)
    # checksum of checksums: sum_v out[v] == sum_u outdeg[u] * X[u]   (fp64 on both sides)
    outdeg = torch.bincount(src, minlength=N).double()
    torch.testing.assert_close(ya.double().sum(0), (outdeg[:, None] * A.double()).sum(0), rtol=1e-6, atol=0)


def test_copy_u_max_and_argmax_are_exact(reddit):
    from dgl import sparse as K
    g, src, dst = reddit
    ids = torch.arange(N, device=src.device).float()               # node id as the feature: exact in fp32
    X = torch.stack([ids, -ids], 1).contiguous()
    out, (arg_u, _) = K._gspmm(g._graph, "copy_lhs", "max", X, None)
    want = torch.full((N, 2), float("-inf"), device=src.device)
    want = want.scatter_reduce(0, dst[:, None].expand(E, 2), X[src], "amax", include_self=True)
    assert torch.equal(out, want)
    has = torch.bincount(dst, minlength=N) > 0
    assert torch.equal(arg_u[has, 0].long(), want[has, 0].long())     # arg of the max id is that id
    assert torch.equal(arg_u[has, 1].long(), (-want[has, 1]).long())
    assert torch.equal(arg_u[~has], torch.zeros_like(arg_u[~has]))


def test_u_dot_v_properties(reddit):
    g, src, dst = reddit
    dev = src.device
    U = torch.randint(0, 4, (N, 64), device=dev).float()
    V = torch.randint(0, 4, (N, 64), device=dev).float()
    got = dgl.ops.gsddmm(g, "dot", U, V)
    assert got.shape == (E, 1)
    # exact: integer-valued products, sums < 2^24; check a 1M-edge sample against gathers
    idx = torch.randperm(E, device=dev)[:1_000_000]
    want = (U[src[idx]] * V[dst[idx]]).sum(-1, keepdim=True)
    assert torch.equal(got[idx], want)
    # dot with ones == row sums of U gathered by source: every edge, exact
    ones = torch.ones(N, 64, device=dev)
    assert torch.equal(dgl.ops.gsddmm(g, "dot", U, ones)[:, 0], U.sum(1)[src])
    # symmetry of the reversed graph: u_dot_v(g, U, V)[e] == u_dot_v(g_rev, V, U)[e]
    grev = g.reverse()
    assert torch.equal(dgl.ops.gsddmm(grev, "dot", V, U), got)


def test_edge_softmax_rows_sum_to_one_and_round_trip(reddit):
    g, src, dst = reddit
    dev = src.device
    z = torch.randn(E, 4, device=dev)
    a = dgl.ops.edge_softmax(g, z)
    sums = torch.zeros(N, 4, device=dev).index_add_(0, dst, a)
    has = torch.bincount(dst, minlength=N) > 0
    torch.testing.assert_close(sums[has], torch.ones_like(sums[has]), rtol=0, atol=2e-5)
    # shift invariance per destination (idempotent normalisation): softmax(z + c[dst]) == softmax(z)
    c = torch.randn(N, 4, device=dev)
    torch.testing.assert_close(dgl.ops.edge_softmax(g, z + c[dst]), a, rtol=2e-4, atol=1e-7)


def test_fused_gat_equals_unfused_composition_at_scale(cuda):
    n, e = 169343, 2315598                                          # ogbn-arxiv shape (+ self loops)
    src, dst = synthetic.random_edges(n, n, e, seed=1)
    src, dst = np.concatenate([src, np.arange(n)]), np.concatenate([dst, np.arange(n)])
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(cuda)
    torch.manual_seed(1)
    ft = torch.randn(n, 4, 16, device=cuda, requires_grad=True)
    el = torch.randn(n, 4, 1, device=cuda, requires_grad=True)
    er = torch.randn(n, 4, 1, device=cuda, requires_grad=True)
    gout = torch.randn(n, 4, 16, device=cuda)
    fused = dgl.ops.gat_attention(g, ft, el, er, 0.2)
    fused.backward(gout)
    gf = [t.grad.clone() for t in (ft, el, er)]
    for t in (ft, el, er):
        t.grad = None
    e_ = torch.nn.functional.leaky_relu(dgl.ops.u_add_v(g, el, er), 0.2)
    unf = dgl.ops.u_mul_e_sum(g, ft, dgl.ops.edge_softmax(g, e_))
    unf.backward(gout)
    torch.testing.assert_close(fused, unf, rtol=1e-4, atol=1e-5)
    for a, b in zip(gf, (ft.grad, el.grad, er.grad)):
        torch.testing.assert_close(a, b, rtol=1e-3, atol=2e-4)
