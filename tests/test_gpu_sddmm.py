"""-m gpu parity: dgl.ops.gsddmm through the C-ABI vs the CPU oracle (SDDMMCoo restatement).
Elementwise ops are bit-exact (one rounding, same operation); dot is within 1e-5 * sum|terms|
(tree vs sequential reduction order)."""
import numpy as np
import pytest
import torch

import dgl
from conftest import assert_close_sumscaled
from gpu_util import graphs, n, t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [1, 2, 3, 4, 8, 16, 33, 64, 128, 256, 602, 1433])
def test_u_dot_v(oracle, cuda, D):
    og, g, src, dst = graphs(oracle, 300, 250, 5000, seed=D)
    rng = np.random.default_rng(D)
    U = rng.standard_normal((300, D)).astype(np.float32)
    V = rng.standard_normal((250, D)).astype(np.float32)
    want = oracle.gsddmm(og, "dot", U, V)
    got = n(dgl.ops.gsddmm(g, "dot", t(U), t(V)))
    scale = (np.abs(U[src]).astype(np.float64) * np.abs(V[dst])).sum(-1, keepdims=True)
    assert_close_sumscaled(got, want, scale, rtol=1e-5, what="u_dot_v D=%d" % D)


@pytest.mark.parametrize("H,F", [(4, 16), (8, 8), (4, 40), (1, 41), (3, 4), (2, 64), (8, 1)])
def test_u_dot_v_multi_head(oracle, cuda, H, F):
    og, g, src, dst = graphs(oracle, 200, 200, 3000, seed=H * F)
    rng = np.random.default_rng(H)
    U = rng.standard_normal((200, H, F)).astype(np.float32)
    V = rng.standard_normal((200, H, F)).astype(np.float32)
    want = oracle.gsddmm(og, "dot", U, V)
    got = n(dgl.ops.gsddmm(g, "dot", t(U), t(V)))
    assert got.shape == (3000, H, 1)
    scale = (np.abs(U[src]).astype(np.float64) * np.abs(V[dst])).sum(-1, keepdims=True)
    assert_close_sumscaled(got, want, scale, rtol=1e-5, what="u_dot_v (%d,%d)" % (H, F))


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
@pytest.mark.parametrize("D", [1, 4, 6, 64, 100, 602])
def test_u_op_v_elementwise(oracle, cuda, op, D):
    og, g, src, dst = graphs(oracle, 150, 130, 2500, seed=D)
    rng = np.random.default_rng(D)
    U = (rng.random((150, D)) + 0.5).astype(np.float32)
    V = (rng.random((130, D)) + 0.5).astype(np.float32)
    want = oracle.gsddmm(og, op, U, V)
    got = n(dgl.ops.gsddmm(g, op, t(U), t(V)))
    if op in ("add", "mul"):
        assert np.array_equal(got, want)
    else:  # sub -> add(-v), div -> mul(1/v) exactly like upstream's Python layer; 1/v rounding is shared
        np.testing.assert_allclose(got, want, rtol=2e-7)


def test_gat_style_u_add_v(oracle, cuda):
    og, g, src, dst = graphs(oracle, 300, 300, 4000, seed=3)
    rng = np.random.default_rng(3)
    el = rng.standard_normal((300, 4, 1)).astype(np.float32)
    er = rng.standard_normal((300, 4, 1)).astype(np.float32)
    assert np.array_equal(n(dgl.ops.u_add_v(g, t(el), t(er))), oracle.gsddmm(og, "add", el, er))


@pytest.mark.parametrize("lt,rt", [("e", "v"), ("e", "u"), ("v", "u"), ("u", "e"), ("v", "e")])
@pytest.mark.parametrize("op", ["add", "mul", "dot"])
def test_other_targets(oracle, cuda, lt, rt, op):
    og, g, src, dst = graphs(oracle, 80, 80, 900, seed=5)
    rng = np.random.default_rng(5)
    rows = {"u": 80, "v": 80, "e": 900}
    L = rng.standard_normal((rows[lt], 3, 4)).astype(np.float32)
    R = rng.standard_normal((rows[rt], 3, 4)).astype(np.float32)
    want = oracle.gsddmm(og, op, L, R, lt, rt)
    got = n(dgl.ops.gsddmm(g, op, t(L), t(R), lhs_target=lt, rhs_target=rt))
    if op == "dot":
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    else:
        assert np.array_equal(got, want)


def test_broadcast_and_copy(oracle, cuda):
    og, g, src, dst = graphs(oracle, 70, 70, 800, seed=6)
    rng = np.random.default_rng(6)
    U = rng.standard_normal((70, 3, 1)).astype(np.float32)
    V = rng.standard_normal((70, 1, 5)).astype(np.float32)
    assert np.array_equal(n(dgl.ops.gsddmm(g, "mul", t(U), t(V))), oracle.gsddmm(og, "mul", U, V))
    X = rng.standard_normal((70, 9)).astype(np.float32)
    assert np.array_equal(n(dgl.ops.copy_u(g, t(X))), X[src])
    assert np.array_equal(n(dgl.ops.copy_v(g, t(X))), X[dst])


def test_formats_without_coo_use_the_csr_kernels(oracle, cuda):
    """main_dgl_product_sage.py:158 drops COO; targets other than (u,v) then take the CSR form."""
    og, g, src, dst = graphs(oracle, 90, 90, 1000, seed=7)
    g2 = g.formats(["csr", "csc"])
    rng = np.random.default_rng(7)
    E = rng.standard_normal((1000, 2)).astype(np.float32)
    V = rng.standard_normal((90, 2)).astype(np.float32)
    assert np.array_equal(n(dgl.ops.gsddmm(g2, "mul", t(E), t(V), "e", "v")), oracle.gsddmm(og, "mul", E, V, "e", "v"))


def test_hub_rows(oracle, cuda, small_hub_threshold):
    og, g, src, dst = graphs(oracle, 3000, 3000, 200000, seed=5, kind="powerlaw")
    rng = np.random.default_rng(8)
    U = rng.standard_normal((3000, 602)).astype(np.float32)
    V = rng.standard_normal((3000, 602)).astype(np.float32)
    want = oracle.gsddmm(og, "dot", U, V)
    got = n(dgl.ops.gsddmm(g, "dot", t(U), t(V)))
    scale = (np.abs(U[src]).astype(np.float64) * np.abs(V[dst])).sum(-1, keepdims=True)
    assert_close_sumscaled(got, want, scale, rtol=1e-5, what="hub dot")
    A = rng.standard_normal((3000, 64)).astype(np.float32)
    assert np.array_equal(n(dgl.ops.gsddmm(g, "add", t(A), t(A))), oracle.gsddmm(og, "add", A, A))
