"""-m gpu parity: dgl.ops.gspmm through the C-ABI vs the CPU oracle on the same seeded inputs.
Bit-exact: structure-derived outputs (arg_u/arg_e incl. first-wins ties, zero-degree handling) and,
for rows that are not split, the sums themselves (same sequential fp32 order as SpMMSumCsr).
Tolerance otherwise: |a-b| <= 1e-5 * sum|terms| (north_star: 1e-5 relative in fp32)."""
import numpy as np
import pytest
import torch

import dgl
from conftest import assert_close_sumscaled
from gpu_util import abs_sum_scale_spmm, graphs, n, t

pytestmark = pytest.mark.gpu

WIDTHS = [1, 2, 3, 4, 7, 16, 33, 64, 100, 128, 256, 602]


@pytest.mark.parametrize("D", WIDTHS)
@pytest.mark.parametrize("reduce_op", ["sum", "mean", "max", "min"])
def test_copy_u(oracle, cuda, D, reduce_op):
    og, g, src, dst = graphs(oracle, 300, 260, 6000, seed=D)
    X = np.random.default_rng(D).random((300, D), dtype=np.float32)
    want = oracle.gspmm(og, "copy_lhs", reduce_op, X, None)
    got = n(dgl.ops.gspmm(g, "copy_lhs", reduce_op, t(X), None))
    # no hub rows at this size: every row is accumulated sequentially in CSC order => bit-exact
    assert np.array_equal(got, want)


@pytest.mark.parametrize("D", [1433])
def test_copy_u_cora_width(oracle, cuda, D):
    og, g, src, dst = graphs(oracle, 2708, 2708, 10556, seed=0)
    X = np.random.default_rng(0).random((2708, D), dtype=np.float32)
    for r in ("sum", "mean"):
        assert np.array_equal(n(dgl.ops.gspmm(g, "copy_lhs", r, t(X), None)), oracle.gspmm(og, "copy_lhs", r, X, None))


@pytest.mark.parametrize("D", [1, 4, 8, 64, 256])
@pytest.mark.parametrize("reduce_op", ["sum", "max", "min", "mean"])
def test_copy_e(oracle, cuda, D, reduce_op):
    og, g, src, dst = graphs(oracle, 200, 150, 3000, seed=D + 1)
    W = np.random.default_rng(D).standard_normal((3000, D)).astype(np.float32)
    want = oracle.gspmm(og, "copy_rhs", reduce_op, None, W)
    got = n(dgl.ops.gspmm(g, "copy_rhs", reduce_op, None, t(W)))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape", [((64,), (64,)), ((4, 16), (4, 1)), ((8, 8), (8, 1)), ((1, 7), (1, 1)),
                                   ((100,), (1,)), ((4, 40), (4, 1)), ((3, 5), (3, 5)), ((6,), (1,))])
def test_u_mul_e_sum(oracle, cuda, shape):
    ls, rs = shape
    og, g, src, dst = graphs(oracle, 220, 220, 4000, seed=len(ls) + ls[-1])
    rng = np.random.default_rng(1)
    X = rng.standard_normal((220,) + ls).astype(np.float32)
    W = rng.standard_normal((4000,) + rs).astype(np.float32)
    want = oracle.gspmm(og, "mul", "sum", X, W)
    got = n(dgl.ops.gspmm(g, "mul", "sum", t(X), t(W)))
    assert np.array_equal(got, want)  # mul and add are not contracted: same bits as the CPU kernel


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
@pytest.mark.parametrize("reduce_op", ["sum", "max", "min"])
def test_generic_broadcast_path(oracle, cuda, op, reduce_op):
    """Shapes the vector kernels do not cover go through the generic kernel, same results."""
    og, g, src, dst = graphs(oracle, 90, 70, 1200, seed=17)
    rng = np.random.default_rng(2)
    X = (rng.random((90, 3, 1)) + 0.5).astype(np.float32)
    W = (rng.random((1200, 1, 5)) + 0.5).astype(np.float32)
    want = oracle.gspmm(og, op, reduce_op, X, W)
    got = n(dgl.ops.gspmm(g, op, reduce_op, t(X), t(W)))
    scale = abs_sum_scale_spmm(src, dst, 70, np.ones((1200, 3, 5)) * 4)
    assert_close_sumscaled(got, want, scale + 1, rtol=1e-6, what="%s/%s" % (op, reduce_op))


@pytest.mark.parametrize("op,D", [("copy_lhs", 5), ("copy_lhs", 64), ("copy_rhs", 16), ("mul", 12)])
def test_argmax_bit_exact_with_ties(oracle, cuda, op, D):
    from dgl import sparse as K
    og, g, src, dst = graphs(oracle, 150, 120, 5000, seed=D)
    rng = np.random.default_rng(D)
    X = rng.integers(0, 3, size=(150, D)).astype(np.float32)   # few distinct values => many ties
    W = rng.integers(1, 3, size=(5000, D)).astype(np.float32)
    for red in ("max", "min"):
        want, (wu, we) = oracle.gspmm_with_args(og, op, red, X if op != "copy_rhs" else None,
                                                W if op != "copy_lhs" else None)
        got, (gu, ge) = K._gspmm(g._graph, op, red, t(X) if op != "copy_rhs" else None,
                                 t(W) if op != "copy_lhs" else None)
        assert np.array_equal(n(got), want)
        if wu is not None:
            assert np.array_equal(n(gu), wu)
        if we is not None:
            assert np.array_equal(n(ge), we)


@pytest.mark.parametrize("D", [1, 4, 64, 256, 602, 1433])
@pytest.mark.parametrize("reduce_op", ["sum", "mean", "max"])
def test_hub_rows_power_law(oracle, cuda, D, reduce_op, small_hub_threshold):
    """Power-law graph whose hubs exceed the split threshold: split rows are within tolerance,
    unsplit rows stay bit-exact, arg-max stays exact (ties still resolve to the first CSC entry)."""
    from dgl import sparse as K
    from dgl import _capi
    og, g, src, dst = graphs(oracle, 3000, 3000, 200000, seed=5, kind="powerlaw")
    X = np.random.default_rng(3).integers(0, 5, size=(3000, D)).astype(np.float32) if reduce_op == "max" \
        else np.random.default_rng(3).random((3000, D), dtype=np.float32)
    deg = og.in_degrees()
    hub = deg > small_hub_threshold
    assert hub.sum() > 50 and (~hub).sum() > 50, "test graph must contain hub and ordinary rows"
    if reduce_op == "max":
        want, (wu, _) = oracle.gspmm_with_args(og, "copy_lhs", "max", X, None)
        got, (gu, _) = K._gspmm(g._graph, "copy_lhs", "max", t(X), None)
        assert np.array_equal(n(got), want) and np.array_equal(n(gu), wu)
        return
    want = oracle.gspmm(og, "copy_lhs", reduce_op, X, None)
    got = n(dgl.ops.gspmm(g, "copy_lhs", reduce_op, t(X), None))
    assert np.array_equal(got[~hub], want[~hub])
    scale = abs_sum_scale_spmm(src, dst, 3000, np.abs(X[src]))
    if reduce_op == "mean":
        scale = scale / np.maximum(deg, 1)[:, None]
    assert_close_sumscaled(got, want, scale, rtol=1e-5, what="hub rows")


def test_bipartite_block(oracle, cuda):
    og, g, src, dst = graphs(oracle, 500, 120, 4000, seed=9)
    X = np.random.default_rng(9).random((500, 32), dtype=np.float32)
    assert np.array_equal(n(dgl.ops.gspmm(g, "copy_lhs", "sum", t(X), None)), oracle.gspmm(og, "copy_lhs", "sum", X, None))


def test_zero_degree_rows_and_empty_graph(oracle, cuda):
    g = dgl.graph((torch.tensor([0, 0]), torch.tensor([3, 3])), num_nodes=5).int().to(cuda)
    X = torch.arange(10, dtype=torch.float32, device=cuda).reshape(5, 2)
    assert n(dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)).tolist() == [[0, 0], [0, 0], [0, 0], [0, 2], [0, 0]]
    assert n(dgl.ops.gspmm(g, "copy_lhs", "max", X, None)).tolist() == [[0, 0], [0, 0], [0, 0], [0, 1], [0, 0]]
    assert n(dgl.ops.gspmm(g, "copy_lhs", "mean", X, None)).tolist() == [[0, 0], [0, 0], [0, 0], [0, 1], [0, 0]]
    e = dgl.graph((torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)), num_nodes=3).int().to(cuda)
    assert n(dgl.ops.gspmm(e, "copy_lhs", "sum", X[:3], None)).tolist() == [[0, 0]] * 3


def test_scalar_features_are_squeezed_back(oracle, cuda):
    og, g, src, dst = graphs(oracle, 60, 60, 500, seed=2)
    x = np.random.default_rng(0).random(60, dtype=np.float32)
    w = np.random.default_rng(1).random(500, dtype=np.float32)
    got = dgl.ops.gspmm(g, "mul", "sum", t(x), t(w))
    assert got.shape == (60,)
    assert np.array_equal(n(got), oracle.gspmm(og, "mul", "sum", x, w))


def test_known_answer_vector(oracle, cuda):
    import json, os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_a6.json")))
    g = dgl.graph((torch.tensor(G["src"]), torch.tensor(G["dst"])), num_nodes=G["n"]).int().to(cuda)
    X = torch.tensor(G["X"], dtype=torch.float32, device=cuda)
    assert n(dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)).tolist() == G["copy_u_sum"]
    assert n(dgl.ops.gspmm(g, "copy_lhs", "mean", X, None)).tolist() == G["copy_u_mean"]
    assert n(dgl.ops.gspmm(g, "copy_lhs", "max", X, None)).tolist() == G["copy_u_max"]
    assert n(dgl.ops.gsddmm(g, "dot", X, X)).reshape(-1).tolist() == G["u_dot_v"]
    z = torch.tensor(G["edge_softmax_logits_div4"], dtype=torch.float32, device=cuda) / 4
    np.testing.assert_allclose(n(dgl.ops.edge_softmax(g, z)), G["edge_softmax"], atol=5e-7)
    assert n(g._graph.csc().indptr).tolist() == G["csc_indptr"]
    assert n(g._graph.csc().indices).tolist() == G["csc_indices"]
    assert n(g._graph.csc().eids).tolist() == G["csc_data"]


@pytest.mark.parametrize("shape", [((4, 16), (4, 1)), ((32,), (32,)), ((100,), (1,))])
def test_hub_rows_u_mul_e_and_copy_e(oracle, cuda, shape, small_hub_threshold):
    ls, rs = shape
    og, g, src, dst = graphs(oracle, 2000, 2000, 100000, seed=6, kind="powerlaw")
    rng = np.random.default_rng(4)
    X = rng.standard_normal((2000,) + ls).astype(np.float32)
    W = rng.standard_normal((100000,) + rs).astype(np.float32)
    want = oracle.gspmm(og, "mul", "sum", X, W)
    got = n(dgl.ops.gspmm(g, "mul", "sum", t(X), t(W)))
    scale = abs_sum_scale_spmm(src, dst, 2000, np.abs(X[src] * W))
    assert_close_sumscaled(got, want, scale, rtol=1e-5, what="hub u_mul_e")
    hub = og.in_degrees() > small_hub_threshold
    assert np.array_equal(got[~hub], want[~hub])
    We = rng.standard_normal((100000, 8)).astype(np.float32)
    for red in ("sum", "max", "min"):
        want = oracle.gspmm(og, "copy_rhs", red, None, We)
        got = n(dgl.ops.gspmm(g, "copy_rhs", red, None, t(We)))
        if red == "sum":
            assert_close_sumscaled(got, want, abs_sum_scale_spmm(src, dst, 2000, np.abs(We)), rtol=1e-5, what="hub copy_e")
        else:
            assert np.array_equal(got, want)


def test_int64_graph_ids_and_noncontiguous_inputs(oracle, cuda):
    """Graphs that were not cast with .int() keep idtype int64 at the API (arg outputs included);
    non-contiguous feature views are accepted."""
    from dgl import sparse as K
    og, g, src, dst = graphs(oracle, 120, 120, 2000, seed=12)
    g64 = g.long()
    assert g64.idtype == torch.int64 and g64.edges()[0].dtype == torch.int64
    X = np.random.default_rng(12).random((120, 24), dtype=np.float32)
    Xt = t(X)
    assert np.array_equal(n(dgl.ops.gspmm(g64, "copy_lhs", "sum", Xt, None)), oracle.gspmm(og, "copy_lhs", "sum", X, None))
    out, (au, _) = K._gspmm(g64._graph, "copy_lhs", "max", Xt, None)
    assert au.dtype == torch.int64
    want, (wu, _) = oracle.gspmm_with_args(og, "copy_lhs", "max", X, None)
    assert np.array_equal(n(au), wu)
    view = t(np.concatenate([X, X], 1))[:, ::2]                      # strided view
    assert not view.is_contiguous()
    got = n(dgl.ops.gspmm(g, "copy_lhs", "sum", view, None))
    assert np.array_equal(got, oracle.gspmm(og, "copy_lhs", "sum", n(view), None))
    assert np.array_equal(n(dgl.ops.gsddmm(g64, "dot", Xt, Xt)), n(dgl.ops.gsddmm(g, "dot", Xt, Xt)))


def test_zero_inf_flag_is_the_upstream_post_pass(oracle, cuda, small_hub_threshold):
    """kernel-level _gspmm keeps -/+inf in rows without in-edges (upstream _CAPI_DGLKernelSpMM contract);
    DGLB_SPMM_ZERO_INF stores what upstream's where(isinf(out), 0, out) would, bit for bit, args untouched."""
    from dgl import sparse as K
    og, g, src, dst = graphs(oracle, 3000, 3000, 30000, seed=77, kind="powerlaw")  # hub rows and empty rows
    X = np.random.default_rng(77).standard_normal((3000, 24)).astype(np.float32)
    X[5, 3], X[6, 2] = np.inf, -np.inf  # genuine infinities in the data are replaced too (as upstream does)
    for red in ("max", "min"):
        raw, (au, _) = K._gspmm(g._graph, "copy_lhs", red, t(X), None)
        fused, (au2, _) = K._gspmm(g._graph, "copy_lhs", red, t(X), None, zero_inf=True)
        want_raw, (wu, _) = oracle.gspmm_with_args(og, "copy_lhs", red, X, None)
        assert np.array_equal(n(raw), want_raw) and np.isinf(n(raw)).any()
        assert np.array_equal(n(fused), np.where(np.isinf(want_raw), np.float32(0), want_raw))
        assert np.array_equal(n(au), wu) and np.array_equal(n(au2), wu)
        assert np.array_equal(n(dgl.ops.gspmm(g, "copy_lhs", red, t(X), None)), n(fused))


@pytest.mark.parametrize("kind", ["uniform", "powerlaw"])
@pytest.mark.parametrize("D", [1, 16, 64, 300])
def test_degree_ordered_row_handout_is_bit_identical(oracle, cuda, kind, D, small_hub_threshold):
    """dglb_hub_t.row_order (rows handed to the lane groups by non-increasing degree, so a warp's rows have equal length)
    changes which group computes a row, never how: gspmm sum / max (+ arg) and gsddmm dot are bit-identical with and
    without it, with and without hub rows, and equal the oracle."""
    from dgl import sparse as K
    nn_, ne = 700, 30000
    og, g, src, dst = graphs(oracle, nn_, nn_, ne, seed=D, kind=kind)
    rng = np.random.default_rng(D)
    X = rng.standard_normal((nn_, D)).astype(np.float32)
    W = rng.standard_normal((ne, 1)).astype(np.float32)
    Xt, Wt = t(X), t(W)
    res = {}
    old = K.ROW_ORDER
    try:
        for mode in ("never", "always"):
            K.ROW_ORDER = mode
            res[mode] = (dgl.ops.gspmm(g, "copy_lhs", "sum", Xt, None), dgl.ops.gspmm(g, "copy_lhs", "max", Xt, None),
                         dgl.ops.gspmm(g, "mul", "sum", Xt, Wt), dgl.ops.gsddmm(g, "dot", Xt, Xt))
    finally:
        K.ROW_ORDER = old
    for a, b in zip(res["never"], res["always"]):
        assert torch.equal(a, b)
    # (hub rows are summed segment by segment here -- small_hub_threshold -- so the oracle comparison is to tolerance)
    assert_close_sumscaled(n(res["always"][0]), oracle.gspmm(og, "copy_lhs", "sum", X, None),
                           abs_sum_scale_spmm(src, dst, nn_, np.abs(X)[src]), rtol=1e-5, what="row-ordered copy_u_sum")
    order = n(g._graph.csc().row_order())
    deg = og.in_degrees()
    assert sorted(order.tolist()) == list(range(nn_)) and (np.diff(deg[order]) <= 0).all()
