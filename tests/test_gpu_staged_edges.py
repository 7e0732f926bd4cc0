"""-m gpu: staged edge order (csrc/edge_stage.cu; include/dglb200.h dglb_edge_stage_plan / dglb_edge_stage).
On graphs whose CSC / CSR carries a non-trivial edge-id permutation, narrow per-edge operands and results go
through one staging pass and the kernels address them with the plan's slot array.  The library switches this on
once the per-edge tensor exceeds 96 MB; the tests lower the cut-off and check that (a) the plan is a valid permutation that is the
identity up to a shuffle inside 32 K-position buckets, slots in edge-id order inside a bucket, (b) every op that
uses it returns exactly what the unstaged path returns, and matches the oracle, forward and backward."""
import numpy as np
import pytest
import torch

import dgl
from dgl import sparse as K
from gpu_util import graphs, n, t

pytestmark = pytest.mark.gpu


@pytest.fixture
def staged():
    old = K.STAGE_MIN_BYTES, K.STAGE_MAX_ROW_FLOATS
    K.STAGE_MIN_BYTES, K.STAGE_MAX_ROW_FLOATS = 1, 8
    yield
    K.STAGE_MIN_BYTES, K.STAGE_MAX_ROW_FLOATS = old


def test_stage_plan_structure(oracle, cuda):
    N, E = 5000, 200000          # 7 buckets of 32 K positions, the last one ragged
    og, g, src, dst = graphs(oracle, N, N, E, seed=41)
    csc = g._graph.csc()
    stage_pos, slot = (x.cpu().numpy() for x in csc.stage_plan())
    eids = csc.eids.cpu().numpy()
    assert np.array_equal(np.sort(stage_pos), np.arange(E))             # a permutation
    assert np.array_equal(slot, stage_pos[eids])
    j = np.arange(E)
    assert np.array_equal(slot >> 15, j >> 15)                          # a slot stays in its position's bucket
    order = np.empty(E, np.int64); order[stage_pos] = np.arange(E)      # staged slot -> edge id
    for b in range((E >> 15) + 1):
        seg = order[b << 15: min(E, (b + 1) << 15)]
        assert (np.diff(seg) > 0).all()                                 # edge-id order inside a bucket
    x = torch.randn(E, 3, device="cuda")
    st = K._stage_move(csc.stage_plan(), x, True)
    assert torch.equal(st[torch.from_numpy(stage_pos).long().cuda()], x)
    assert torch.equal(K._stage_move(csc.stage_plan(), st, False), x)
    # dst-sorted graphs need no plan
    _, g2, _, _ = graphs(oracle, N, N, E, seed=41, order="dst_sorted")
    assert g2._graph.csc().stage_plan() is None


@pytest.mark.parametrize("kind", ["uniform", "powerlaw"])
def test_staged_ops_match_unstaged_and_oracle(oracle, cuda, kind, small_hub_threshold):
    N, E, H, F = 3000, 150000, 4, 16
    og, g, src, dst = graphs(oracle, N, N, E, seed=42, kind=kind)
    rng = np.random.default_rng(42)
    X = rng.standard_normal((N, 64)).astype(np.float32)
    w1 = rng.random((E, 1), dtype=np.float32)
    ft = rng.standard_normal((N, H, F)).astype(np.float32)
    a = rng.random((E, H, 1), dtype=np.float32)
    el = rng.standard_normal((N, H, 1)).astype(np.float32)
    er = rng.standard_normal((N, H, 1)).astype(np.float32)
    logits = rng.standard_normal((E, H, 1)).astype(np.float32)
    gr = rng.standard_normal((E, H, 1)).astype(np.float32)

    def run():
        r = {}
        r["mul_scalar"] = dgl.ops.gspmm(g, "mul", "sum", t(X), t(w1))
        r["mul_head"] = dgl.ops.gspmm(g, "mul", "sum", t(ft), t(a))
        r["copy_e"] = dgl.ops.gspmm(g, "copy_rhs", "sum", None, t(a))
        r["copy_e_max"] = dgl.ops.gspmm(g, "copy_rhs", "max", None, t(a))       # needs real edge ids: never staged
        r["dot"] = dgl.ops.gsddmm(g, "dot", t(X), t(X))
        r["dot_heads"] = dgl.ops.gsddmm(g, "dot", t(ft), t(ft))
        r["add"] = dgl.ops.gsddmm(g, "add", t(el), t(er))
        lt = t(logits).requires_grad_(True)
        sm = dgl.ops.edge_softmax(g, lt)
        sm.backward(t(gr))
        r["softmax"], r["softmax_grad"] = sm.detach(), lt.grad
        wt = t(a).requires_grad_(True)
        ftt = t(ft).requires_grad_(True)
        dgl.ops.gspmm(g, "mul", "sum", ftt, wt).sum().backward()
        r["dW"], r["dft"] = wt.grad, ftt.grad
        return r

    old, K.STAGE_MIN_BYTES = (K.STAGE_MIN_BYTES, K.STAGE_MAX_ROW_FLOATS), 1 << 40
    plain = run()
    K.STAGE_MIN_BYTES, K.STAGE_MAX_ROW_FLOATS = 1, 8      # also stage the (E,4) operands the defaults leave alone
    stagedr = run()
    K.STAGE_MIN_BYTES, K.STAGE_MAX_ROW_FLOATS = old
    assert g._graph.csc()._stage is not None and g._graph.csr()._stage is not None      # both directions were planned
    for k in plain:
        assert torch.equal(plain[k], stagedr[k]), k     # same kernels, same per-row order: bit-identical
    assert np.array_equal(n(stagedr["add"]), oracle.gsddmm(og, "add", el, er))
    np.testing.assert_allclose(n(stagedr["softmax"]), oracle.edge_softmax(og, logits), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(n(stagedr["dot"]), oracle.gsddmm(og, "dot", X, X), rtol=1e-4, atol=1e-4)
    hub = np.bincount(dst, minlength=N) > small_hub_threshold
    want = oracle.gspmm(og, "mul", "sum", X, w1)
    assert np.array_equal(n(stagedr["mul_scalar"])[~hub], want[~hub])
