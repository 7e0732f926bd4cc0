"""-m gpu parity of the backward passes: gradients of dgl.ops.gspmm / gsddmm vs the oracle's
restatement of upstream's backward formulas and vs fp64 torch autograd of the
gather/scatter formulation on the CPU."""
import numpy as np
import pytest
import torch

import dgl
from conftest import assert_close_sumscaled
from gpu_util import graphs, n, t

pytestmark = pytest.mark.gpu


def _ref_spmm(src, dst, n_dst, op, reduce_op, X, W):
    s, d = torch.from_numpy(src).long(), torch.from_numpy(dst).long()
    if op == "copy_lhs":
        msg = X[s]
    elif op == "copy_rhs":
        msg = W
    elif op == "mul":
        msg = X[s] * W
    elif op == "add":
        msg = X[s] + W
    elif op == "sub":
        msg = X[s] - W
    else:
        msg = X[s] / W
    out = torch.zeros((n_dst,) + msg.shape[1:], dtype=torch.float64)
    if reduce_op in ("sum", "mean"):
        out = out.index_add_(0, d, msg)
        if reduce_op == "mean":
            deg = torch.bincount(d, minlength=n_dst).clamp(min=1).double()
            out = out / deg.view((-1,) + (1,) * (out.dim() - 1))
        return out
    idx = d.view((-1,) + (1,) * (msg.dim() - 1)).expand_as(msg)
    red = "amax" if reduce_op == "max" else "amin"
    out = torch.full_like(out, -float("inf") if reduce_op == "max" else float("inf"))
    out = out.scatter_reduce(0, idx, msg, red, include_self=True)
    return torch.where(torch.isinf(out), torch.zeros_like(out), out)


@pytest.mark.parametrize("op,ls,rs", [("copy_lhs", (64,), None), ("copy_lhs", (602,), None), ("copy_rhs", None, (16,)),
                                      ("mul", (4, 16), (4, 1)), ("mul", (32,), (32,)), ("mul", (100,), (1,)),
                                      ("add", (8,), (8,)), ("sub", (8,), (8,)), ("div", (8,), (8,)), ("add", (3, 1), (1, 5))])
@pytest.mark.parametrize("reduce_op", ["sum", "mean"])
def test_gspmm_gradients(oracle, cuda, op, ls, rs, reduce_op):
    og, g, src, dst = graphs(oracle, 200, 170, 3000, seed=3)
    rng = np.random.default_rng(3)
    X = (rng.random((200,) + ls) + 0.5).astype(np.float32) if ls else None
    W = (rng.random((3000,) + rs) + 0.5).astype(np.float32) if rs else None
    Xt = t(X).requires_grad_(True) if X is not None else None
    Wt = t(W).requires_grad_(True) if W is not None else None
    out = dgl.ops.gspmm(g, op, reduce_op, Xt, Wt)
    gout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(t(gout))
    X64 = torch.tensor(X, dtype=torch.float64, requires_grad=True) if X is not None else None
    W64 = torch.tensor(W, dtype=torch.float64, requires_grad=True) if W is not None else None
    ref = _ref_spmm(src, dst, 170, op, reduce_op, X64, W64)
    np.testing.assert_allclose(n(out), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    ref.backward(torch.tensor(gout, dtype=torch.float64))
    if X is not None:
        np.testing.assert_allclose(n(Xt.grad), X64.grad.numpy(), rtol=1e-4, atol=2e-5)
    if W is not None:
        np.testing.assert_allclose(n(Wt.grad), W64.grad.numpy(), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("op", ["copy_lhs", "mul"])
@pytest.mark.parametrize("reduce_op", ["max", "min"])
def test_gspmm_cmp_gradients(oracle, cuda, op, reduce_op):
    og, g, src, dst = graphs(oracle, 120, 100, 1500, seed=8)
    rng = np.random.default_rng(8)
    X = rng.standard_normal((120, 6)).astype(np.float32)   # continuous values: no ties
    W = (rng.random((1500, 6)) + 0.5).astype(np.float32)
    Xt, Wt = t(X).requires_grad_(True), t(W).requires_grad_(True)
    out = dgl.ops.gspmm(g, op, reduce_op, Xt, Wt if op == "mul" else None)
    gout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(t(gout))
    X64 = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    W64 = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    ref = _ref_spmm(src, dst, 100, op, reduce_op, X64, W64)
    ref.backward(torch.tensor(gout, dtype=torch.float64))
    np.testing.assert_allclose(n(Xt.grad), X64.grad.numpy(), rtol=1e-5, atol=1e-6)
    if op == "mul":
        np.testing.assert_allclose(n(Wt.grad), W64.grad.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("op", ["copy_lhs", "mul", "copy_rhs"])
@pytest.mark.parametrize("reduce_op", ["max", "min"])
def test_gspmm_cmp_gradients_with_isolated_rows(oracle, cuda, op, reduce_op):
    """Destination rows without in-edges get out = 0 (upstream's where(isinf, 0) post-pass) and must not
    leak their incoming gradient into node 0 / edge 0 through the arg = 0 the kernel records for them."""
    rng = np.random.default_rng(11)
    n_src, n_dst, n_e = 40, 60, 90               # ~22 % of the destination rows stay empty
    src = rng.integers(0, n_src, n_e)
    dst = rng.integers(0, n_dst, n_e)
    dst[dst % 4 == 0] = 1                        # rows 0, 4, 8, ... are isolated for sure
    g = dgl.create_block((torch.from_numpy(src), torch.from_numpy(dst)), n_src, n_dst).int().to("cuda")
    assert int((g.in_degrees() == 0).sum()) >= 15
    X = rng.standard_normal((n_src, 5)).astype(np.float32)
    W = (rng.random((n_e, 5)) + 0.5).astype(np.float32)
    Xt, Wt = t(X).requires_grad_(True), t(W).requires_grad_(True)
    out = dgl.ops.gspmm(g, op, reduce_op, Xt if op != "copy_rhs" else None, Wt if op != "copy_lhs" else None)
    gout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(t(gout))
    X64 = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    W64 = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    ref = _ref_spmm(src, dst, n_dst, op, reduce_op, X64, W64)
    np.testing.assert_allclose(n(out), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    ref.backward(torch.tensor(gout, dtype=torch.float64))
    if op != "copy_rhs":
        np.testing.assert_allclose(n(Xt.grad), X64.grad.numpy(), rtol=1e-5, atol=1e-6)
    if op != "copy_lhs":
        np.testing.assert_allclose(n(Wt.grad), W64.grad.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("op,lt,rt,shape", [("dot", "u", "v", (64,)), ("dot", "u", "v", (4, 16)), ("add", "u", "v", (4, 1)),
                                            ("mul", "u", "v", (8,)), ("mul", "e", "v", (3,)), ("add", "e", "u", (3,)),
                                            ("sub", "u", "v", (5,)), ("div", "u", "v", (5,))])
def test_gsddmm_gradients(oracle, cuda, op, lt, rt, shape):
    og, g, src, dst = graphs(oracle, 90, 90, 1200, seed=4)
    rng = np.random.default_rng(4)
    rows = {"u": 90, "v": 90, "e": 1200}
    L = (rng.random((rows[lt],) + shape) + 0.5).astype(np.float32)
    R = (rng.random((rows[rt],) + shape) + 0.5).astype(np.float32)
    Lt, Rt = t(L).requires_grad_(True), t(R).requires_grad_(True)
    out = dgl.ops.gsddmm(g, op, Lt, Rt, lhs_target=lt, rhs_target=rt)
    gout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(t(gout))
    L64 = torch.tensor(L, dtype=torch.float64, requires_grad=True)
    R64 = torch.tensor(R, dtype=torch.float64, requires_grad=True)
    s, d = torch.from_numpy(src).long(), torch.from_numpy(dst).long()
    sel = {"u": lambda x: x[s], "v": lambda x: x[d], "e": lambda x: x}
    a, b = sel[lt](L64), sel[rt](R64)
    ref = {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b, "dot": (a * b).sum(-1, keepdim=True)}[op]
    np.testing.assert_allclose(n(out), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    ref.backward(torch.tensor(gout, dtype=torch.float64))
    np.testing.assert_allclose(n(Lt.grad), L64.grad.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(n(Rt.grad), R64.grad.numpy(), rtol=1e-4, atol=2e-5)


def test_backward_matches_upstream_decomposition_bitwise(oracle, cuda):
    """dX of copy_u_sum must be the copy_u_sum of dZ on the REVERSED graph (CSR traversal), i.e.
    the exact sequential sum the oracle computes on og.reverse()."""
    og, g, src, dst = graphs(oracle, 150, 150, 2500, seed=6)
    rng = np.random.default_rng(6)
    X = rng.random((150, 64), dtype=np.float32)
    Xt = t(X).requires_grad_(True)
    out = dgl.ops.gspmm(g, "copy_lhs", "sum", Xt, None)
    gout = rng.standard_normal((150, 64)).astype(np.float32)
    out.backward(t(gout))
    dX, _ = oracle.gspmm_sum_backward(og, "copy_lhs", X, None, gout)
    assert np.array_equal(n(Xt.grad), dX)


def test_update_all_and_apply_edges_route_to_the_same_kernels(oracle, cuda):
    import dgl.function as fn
    og, g, src, dst = graphs(oracle, 100, 100, 1500, seed=7)
    rng = np.random.default_rng(7)
    h = rng.random((100, 16), dtype=np.float32)
    w = rng.random((1500, 1), dtype=np.float32)
    g = g.local_var()
    g.srcdata["h"] = t(h)
    g.edata["w"] = t(w)
    g.update_all(fn.copy_src("h", "m"), fn.mean("m", "neigh"))           # main_dgl_citation_sage.py:77
    assert np.array_equal(n(g.dstdata["neigh"]), oracle.gspmm(og, "copy_lhs", "mean", h, None))
    g.update_all(fn.u_mul_e("h", "w", "m"), fn.sum("m", "s"))             # main_dgl_proteins_rgcn_for.py:52
    assert np.array_equal(n(g.dstdata["s"]), oracle.gspmm(og, "mul", "sum", h, w))
    g.apply_edges(fn.u_dot_v("h", "h", "score"))                          # gcmc_dgl/model.py:342
    np.testing.assert_allclose(n(g.edata["score"]), oracle.gsddmm(og, "dot", h, h), rtol=1e-5)
    # UDF message + builtin reduce (main_dgl_molhiv_gcn.py:46,50-52)
    g.update_all(lambda edges: {"m": edges.src["h"] * edges.data["w"]}, fn.sum("m", "udf"))
    np.testing.assert_allclose(n(g.dstdata["udf"]), oracle.gspmm(og, "mul", "sum", h, w), rtol=1e-6)
