"""CPU tests of the oracle itself: the known-answer vector of SURVEY.md Appendix A.6 and
independent fp64 restatements (scipy CSR matmul, numpy scatter ops).  The reference repository
has no tests or golden vectors for this path (parity unpinned), so these are what pins the oracle.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_edges

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_a6.json")))


@pytest.fixture(scope="module")
def a6(oracle):
    return oracle.OracleGraph(GOLD["src"], GOLD["dst"], GOLD["n"], GOLD["n"])


def test_a6_csc_structure(oracle, a6):
    indptr, indices, data = a6.csc
    assert indptr.tolist() == GOLD["csc_indptr"]
    assert indices.tolist() == GOLD["csc_indices"]
    assert data.tolist() == GOLD["csc_data"]
    assert a6.in_degrees().tolist() == GOLD["in_deg"]


def test_a6_gspmm(oracle, a6):
    X = np.array(GOLD["X"], np.float32)
    assert oracle.gspmm(a6, "copy_lhs", "sum", X, None).tolist() == GOLD["copy_u_sum"]
    assert oracle.gspmm(a6, "copy_lhs", "mean", X, None).tolist() == GOLD["copy_u_mean"]
    assert oracle.gspmm(a6, "copy_lhs", "max", X, None).tolist() == GOLD["copy_u_max"]
    _, (au, ae) = oracle.gspmm_with_args(a6, "copy_lhs", "max", X, None, both_args=True)
    assert au.tolist() == GOLD["arg_u"]
    assert ae.tolist() == GOLD["arg_e"]


def test_a6_gsddmm_and_softmax(oracle, a6):
    X = np.array(GOLD["X"], np.float32)
    assert oracle.gsddmm(a6, "dot", X, X).reshape(-1).tolist() == GOLD["u_dot_v"]
    logits = np.array(GOLD["edge_softmax_logits_div4"], np.float32) / 4
    np.testing.assert_allclose(oracle.edge_softmax(a6, logits), GOLD["edge_softmax"], atol=5e-7)


@pytest.mark.parametrize("order", ["shuffled", "dst_sorted"])
def test_coo_to_csr_is_stable_sort(oracle, order):
    src, dst = make_edges(300, 200, 5000, seed=3, order=order)
    indptr, indices, data = oracle.coo_to_csr(200, dst, src)
    perm = np.argsort(dst, kind="stable")
    assert np.array_equal(data, perm.astype(np.int32))
    assert np.array_equal(indices, src[perm].astype(np.int32))
    assert np.array_equal(indptr, np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=200))]))
    if order == "dst_sorted":
        assert np.array_equal(data, np.arange(5000))


def test_coo_to_csr_empty(oracle):
    indptr, indices, data = oracle.coo_to_csr(4, np.zeros(0, np.int64), np.zeros(0, np.int64))
    assert indptr.tolist() == [0, 0, 0, 0, 0] and len(indices) == 0 and len(data) == 0


@pytest.mark.parametrize("D", [1, 3, 16, 100])
def test_copy_u_sum_vs_scipy_fp64(oracle, D):
    rng = np.random.default_rng(D)
    src, dst = make_edges(120, 90, 3000, seed=D)
    g = oracle.OracleGraph(src, dst, 120, 90)
    X = rng.random((120, D), dtype=np.float32)
    indptr, indices, _ = g.csc
    A = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(90, 120))  # keeps duplicates
    want = A @ X.astype(np.float64)
    got = oracle.gspmm(g, "copy_lhs", "sum", X, None)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    deg = np.maximum(np.diff(indptr), 1)[:, None]
    np.testing.assert_allclose(oracle.gspmm(g, "copy_lhs", "mean", X, None), want / deg, rtol=1e-5)


def test_u_mul_e_and_max_vs_numpy(oracle):
    rng = np.random.default_rng(0)
    src, dst = make_edges(50, 40, 600, seed=1)
    g = oracle.OracleGraph(src, dst, 50, 40)
    X = rng.standard_normal((50, 4, 8)).astype(np.float32)
    W = rng.standard_normal((600, 4, 1)).astype(np.float32)
    msg = X[src].astype(np.float64) * W.astype(np.float64)
    want = np.zeros((40, 4, 8))
    np.add.at(want, dst, msg)
    np.testing.assert_allclose(oracle.gspmm(g, "mul", "sum", X, W), want, rtol=1e-4, atol=1e-5)
    wantmax = np.full((40, 4, 8), -np.inf)
    np.maximum.at(wantmax, dst, (X[src] * W).astype(np.float32))
    wantmax[np.isinf(wantmax)] = 0
    assert np.array_equal(oracle.gspmm(g, "mul", "max", X, W), wantmax.astype(np.float32))


def test_max_args_point_at_first_maximum(oracle):
    rng = np.random.default_rng(5)
    src, dst = make_edges(30, 20, 400, seed=2)
    g = oracle.OracleGraph(src, dst, 30, 20)
    X = rng.integers(0, 4, size=(30, 5)).astype(np.float32)  # many ties
    out, (au, ae) = oracle.gspmm_with_args(g, "copy_lhs", "max", X, None, both_args=True)
    for v in range(20):
        eids = np.nonzero(dst == v)[0]  # increasing edge id == CSC row order
        for k in range(5):
            if len(eids) == 0:
                assert au[v, k] == 0 and ae[v, k] == 0 and np.isinf(out[v, k])
                continue
            vals = X[src[eids], k]
            first = eids[np.argmax(vals)]  # argmax returns the first maximum
            assert ae[v, k] == first and au[v, k] == src[first] and out[v, k] == vals.max()


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "dot"])
def test_gsddmm_vs_gather(oracle, op):
    """Written-out twin in the reference: kernel/pyg.py:47-49 + kernel/utils.py:8-16."""
    rng = np.random.default_rng(7)
    src, dst = make_edges(40, 35, 500, seed=4)
    g = oracle.OracleGraph(src, dst, 40, 35)
    U = (rng.random((40, 6)) + 0.5).astype(np.float32)
    V = (rng.random((35, 6)) + 0.5).astype(np.float32)
    table = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y,
             "div": lambda x, y: x / y, "dot": lambda x, y: (x * y).sum(-1, keepdims=True)}
    want = table[op](U[src].astype(np.float64), V[dst].astype(np.float64))
    np.testing.assert_allclose(oracle.gsddmm(g, op, U, V), want, rtol=2e-6)


def test_edge_softmax_and_backward_vs_fp64(oracle):
    rng = np.random.default_rng(11)
    src, dst = make_edges(30, 25, 300, seed=6)
    g = oracle.OracleGraph(src, dst, 30, 25)
    z = rng.standard_normal((300, 3, 1)).astype(np.float32)
    a = oracle.edge_softmax(g, z)
    z64 = z.astype(np.float64)
    m = np.full((25, 3, 1), -np.inf)
    np.maximum.at(m, dst, z64)
    ex = np.exp(z64 - m[dst])
    s = np.zeros((25, 3, 1))
    np.add.at(s, dst, ex)
    want = ex / s[dst]
    np.testing.assert_allclose(a, want, rtol=2e-6)
    gout = rng.standard_normal((300, 3, 1)).astype(np.float32)
    acc = np.zeros((25, 3, 1))
    np.add.at(acc, dst, want * gout)
    np.testing.assert_allclose(oracle.edge_softmax_backward(g, a, gout), want * gout - want * acc[dst],
                               rtol=1e-4, atol=1e-6)


def test_gat_forward_matches_written_out_math(oracle):
    """GAT attention as spelled out in main_pyg_arxiv_gat.py:98-111."""
    rng = np.random.default_rng(13)
    n = 40
    src, dst = make_edges(n, n, 300, seed=8)
    src = np.concatenate([src, np.arange(n)])
    dst = np.concatenate([dst, np.arange(n)])
    g = oracle.OracleGraph(src, dst, n, n)
    ft = rng.standard_normal((n, 2, 5)).astype(np.float32)
    al = rng.standard_normal((1, 2, 5)).astype(np.float32)
    ar = rng.standard_normal((1, 2, 5)).astype(np.float32)
    rst, a, el, er = oracle.gat_forward(g, ft, al, ar, 0.2)
    f64 = ft.astype(np.float64)
    e = (f64 * al).sum(-1)[src] + (f64 * ar).sum(-1)[dst]
    e = np.where(e > 0, e, 0.2 * e)
    m = np.full((n, 2), -np.inf)
    np.maximum.at(m, dst, e)
    ex = np.exp(e - m[dst])
    s = np.zeros((n, 2))
    np.add.at(s, dst, ex)
    alpha = ex / s[dst]
    want = np.zeros((n, 2, 5))
    np.add.at(want, dst, alpha[:, :, None] * f64[src])
    np.testing.assert_allclose(rst, want, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(a[:, :, 0], alpha, rtol=1e-4, atol=1e-6)


def test_broadcast_offsets_match_numpy(oracle):
    bc = oracle.calc_bcast("mul", (4, 8), (4, 1))
    assert bc["use_bcast"] and bc["out_len"] == 32
    assert bc["rhs_off"].tolist() == [k // 8 for k in range(32)]
    assert bc["lhs_off"].tolist() == list(range(32))
    bc = oracle.calc_bcast("dot", (4, 8), (4, 8))
    assert not bc["use_bcast"] and bc["out_len"] == 4 and bc["reduce_size"] == 8
    with pytest.raises(ValueError):
        oracle.calc_bcast("add", (3, 2), (4, 2))


def test_gcn_message_sum_vs_torch_autograd_fp64(oracle):
    """main_dgl_molhiv_gcn.py:50-52 + :46 restated (oracle.gcn_message_sum) against the same math through torch autograd
    in float64: forward within float32 rounding, backward formulas exact."""
    import torch
    rng = np.random.default_rng(5)
    n, e, D = 60, 400, 9
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    g = oracle.OracleGraph(src, dst, n, n)
    x, w = rng.standard_normal((n, D)).astype(np.float32), rng.standard_normal((e, D)).astype(np.float32)
    c = ((g.in_degrees() + 1) ** -0.5).astype(np.float32)
    gout = rng.standard_normal((n, D)).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    ct = torch.tensor(c, dtype=torch.float64)
    s, d = torch.from_numpy(src), torch.from_numpy(dst)
    m = (ct[s] * ct[d])[:, None] * torch.relu(xt[s] + wt)
    h = torch.zeros(n, D, dtype=torch.float64).index_add_(0, d, m)
    h.backward(torch.tensor(gout, dtype=torch.float64))
    np.testing.assert_allclose(oracle.gcn_message_sum(g, x, w, c, c), h.detach().numpy(), rtol=1e-5, atol=1e-6)
    gx, gw = oracle.gcn_message_sum_backward(g, x, w, c, c, gout)
    np.testing.assert_allclose(gx, xt.grad.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(gw, wt.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_batch_graphs_shifts_ids_and_keeps_member_order(oracle):
    a = (np.array([0, 1, 2]), np.array([1, 2, 0]), 3)
    b = (np.array([], dtype=np.int64), np.array([], dtype=np.int64), 2)       # a member without edges
    c = (np.array([1, 0]), np.array([0, 0]), 2)
    g, n_off, e_off = oracle.batch_graphs([a, b, c])
    assert n_off.tolist() == [0, 3, 5, 7] and e_off.tolist() == [0, 3, 3, 5]
    assert g.src.tolist() == [0, 1, 2, 6, 5] and g.dst.tolist() == [1, 2, 0, 5, 5]
    indptr, indices, data = g.csc
    assert indptr.tolist() == [0, 1, 2, 3, 3, 3, 5, 5]
    assert indices.tolist() == [2, 0, 1, 6, 5] and data.tolist() == [2, 0, 1, 3, 4]
