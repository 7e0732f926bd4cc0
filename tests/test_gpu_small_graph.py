"""-m gpu parity of the batched small-graph pipeline (csrc/small_graph.cu, dgl/batch_store.py, dgl.ops.gcn_norm_relu_sum):
device-side batch construction bit-exact against the oracle's dgl.batch + stable COO->CSR restatement and against the
host path of this package; the fused GCN message kernel bit-exact (forward) against the oracle's written-out UDF +
copy_e-sum and within 1e-5 * sum|terms| (backward); a padded, CUDA-graph-captured training step against the eager,
unpadded one."""
import os
import sys

import numpy as np
import pytest
import torch

import dgl
from conftest import ROOT, assert_close_sumscaled
from gpu_util import graphs, n, t

pytestmark = pytest.mark.gpu


def _molecules(count, seed=0, with_empty=False):
    from dgl.data import synthetic
    out = []
    for i in range(count):
        src, dst, sizes = synthetic.molecule_like_batch(1, seed=seed * 1000 + i, mean_nodes=12.0)
        nn_ = int(sizes[0])
        if with_empty and i % 7 == 3:
            src, dst = src[:0], dst[:0]                      # a member graph without edges
        g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=nn_)
        rng = np.random.default_rng(seed * 1000 + i)
        g.ndata["feat"] = torch.from_numpy(rng.integers(0, 2, size=(nn_, 9)).astype(np.int64))
        g.edata["feat"] = torch.from_numpy(rng.integers(0, 2, size=(len(src), 3)).astype(np.int64))
        out.append(g)
    return out


@pytest.mark.parametrize("batch_size,with_empty", [(1, False), (7, True), (64, True), (300, False)])
def test_device_batch_is_bit_exact(oracle, cuda, batch_size, with_empty):
    mols = _molecules(120, seed=batch_size, with_empty=with_empty)
    labels = torch.arange(len(mols), dtype=torch.float32).view(-1, 1)
    store = dgl.GraphStore(mols, labels, device=cuda)
    rng = np.random.default_rng(batch_size)
    ids = rng.integers(0, len(mols), size=batch_size)            # repeats allowed, any order
    n_pad, e_pad = store.pad_sizes([ids], multiple=32)
    assert store.fits(ids, n_pad, e_pad)
    sb = store.static_batch(batch_size, n_pad, e_pad).load(ids)
    assert not sb.overflowed()
    # oracle: dgl.batch restated + stable counting sort
    og, n_off, e_off = oracle.batch_graphs([(n(m._graph.src), n(m._graph.dst), m.number_of_nodes()) for m in (mols[i] for i in ids)])
    N, E = og.n_dst, og.n_edges
    assert n_pad > N and e_pad >= E
    np.testing.assert_array_equal(n(sb.src)[:E], og.src)
    np.testing.assert_array_equal(n(sb.dst)[:E], og.dst)
    for got_ptr, got_idx, got_eid, want in ((sb.csc_indptr, sb.csc_indices, sb.csc_eids, og.csc),
                                            (sb.csr_indptr, sb.csr_indices, sb.csr_eids, og.csr)):
        np.testing.assert_array_equal(n(got_ptr)[:N + 1], want[0])
        assert (n(got_ptr)[N:] == E).all()                       # padding nodes: empty rows
        np.testing.assert_array_equal(n(got_idx)[:E], want[1])
        np.testing.assert_array_equal(n(got_eid)[:E], want[2])
    np.testing.assert_array_equal(n(sb.out_node_ptr), np.concatenate([n_off, [n_pad]]))
    np.testing.assert_array_equal(n(sb.out_edge_ptr), np.concatenate([e_off, [e_pad]]))
    np.testing.assert_array_equal(n(sb.node_graph)[:N], np.repeat(np.arange(batch_size), np.diff(n_off)))
    assert (n(sb.node_graph)[N:] == batch_size).all()
    # the host path of this package (dgl.batch on CPU graphs -> device -> COO->CSC build) gives the same structure + data
    hb = dgl.batch([mols[i] for i in ids]).to(cuda).int()
    csc = hb._graph.csc()
    np.testing.assert_array_equal(n(csc.indptr), n(sb.csc_indptr)[:N + 1])
    np.testing.assert_array_equal(n(csc.indices), n(sb.csc_indices)[:E])
    eids = n(csc.eids) if csc.eids is not None else np.arange(E)
    np.testing.assert_array_equal(eids, n(sb.csc_eids)[:E])
    np.testing.assert_array_equal(n(hb.ndata["feat"]), n(sb.graph.ndata["feat"])[:N])
    np.testing.assert_array_equal(n(hb.edata["feat"]), n(sb.graph.edata["feat"])[:E])
    np.testing.assert_array_equal(n(sb.labels).ravel(), ids.astype(np.float32))
    np.testing.assert_array_equal(n(sb.graph.in_degrees())[:N], og.in_degrees())
    assert float(sb.n_real_nodes) == N and float(sb.node_mask.sum()) == N
    # a second batch through the same buffers (what a CUDA-graph replay does)
    ids2 = rng.integers(0, len(mols), size=batch_size)
    if store.fits(ids2, n_pad, e_pad):
        sb.load(ids2)
        og2, _, _ = oracle.batch_graphs([(n(m._graph.src), n(m._graph.dst), m.number_of_nodes()) for m in (mols[i] for i in ids2)])
        np.testing.assert_array_equal(n(sb.csc_indptr)[:og2.n_dst + 1], og2.csc[0])
        np.testing.assert_array_equal(n(sb.csc_indices)[:og2.n_edges], og2.csc[1])
        np.testing.assert_array_equal(n(sb.graph.in_degrees())[:og2.n_dst], og2.in_degrees())


def test_device_batch_reports_overflow(cuda):
    mols = _molecules(20, seed=3)
    store = dgl.GraphStore(mols, None, device=cuda)
    ids = np.arange(16)
    assert not store.fits(ids, 8, 8)
    sb = store.static_batch(16, 8, 8).load(ids)
    assert sb.overflowed()                                       # flagged, and nothing was written out of bounds
    assert int(sb.csc_indptr.max()) <= 8 and int(sb.src.max()) < 8


@pytest.mark.parametrize("D", [256, 64, 7, 1])
@pytest.mark.parametrize("kind", ["molecules", "uniform", "powerlaw"])
def test_gcn_norm_relu_sum_matches_written_out_udf(oracle, cuda, D, kind):
    if kind == "molecules":
        mols = _molecules(40, seed=D)
        bg = dgl.batch(mols)
        src, dst, nn_ = n(bg._graph.src), n(bg._graph.dst), bg.number_of_nodes()
        og = oracle.OracleGraph(src, dst, nn_, nn_)
        g = bg.to(cuda).int()
    else:
        nn_, ne = 500, 6000
        og, g, src, dst = graphs(oracle, nn_, nn_, ne, seed=D, kind=kind)
    ne = len(src)
    rng = np.random.default_rng(D)
    x = rng.standard_normal((nn_, D)).astype(np.float32)
    w = rng.standard_normal((ne, D)).astype(np.float32)
    c = ((og.in_degrees() + 1).astype(np.float32) ** np.float32(-0.5)).astype(np.float32)
    want = oracle.gcn_message_sum(og, x, w, c, c)
    xt, wt, ct = t(x).requires_grad_(True), t(w).requires_grad_(True), t(c)
    got = dgl.ops.gcn_norm_relu_sum(g, xt, wt, ct.view(-1, 1))
    np.testing.assert_array_equal(n(got), want)                  # same products, same order of additions: bit-exact
    # the written-out UDF through this package's own update_all (torch ops + copy_e sum kernel) agrees too
    import dgl.function as fn
    lg = g.local_var()
    lg.ndata["c"], lg.ndata["x"], lg.edata["w"] = ct.view(-1, 1), t(x), t(w)
    lg.update_all(lambda e: {"m": e.src["c"] * e.dst["c"] * torch.relu(e.src["x"] + e.data["w"])}, fn.sum("m", "h"))
    np.testing.assert_array_equal(n(lg.ndata["h"]), want)
    # backward
    gout = rng.standard_normal((nn_, D)).astype(np.float32)
    got.backward(t(gout))
    gx, gw = oracle.gcn_message_sum_backward(og, x, w, c, c, gout)
    np.testing.assert_allclose(n(wt.grad), gw, rtol=2e-6, atol=0)   # one product per element
    scale = np.zeros((nn_, D))
    np.add.at(scale, src, np.abs(gw))
    assert_close_sumscaled(n(xt.grad), gx, scale, rtol=1e-5, what="gcn_norm_relu_sum grad_x")


def _load_example_model():
    for p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from examples import small_graph_model
    return small_graph_model


def test_padded_captured_training_step_matches_eager_unpadded(oracle, cuda):
    """The whole iteration of main_dgl_molhiv_gcn.py:95-115 -- batch construction, forward, loss, backward, Adam -- as
    ONE CUDA graph over a padded StaticBatch with the fused message kernel, against the script's own formulation (host
    dgl.batch, message UDF, nn.BatchNorm semantics) run eagerly on the unpadded batch: same losses over several
    iterations with different batches (dropout off; tolerance covers the different summation order of the dense ops)."""
    M = _load_example_model()
    mols = _molecules(96, seed=11)
    labels = torch.from_numpy(np.random.default_rng(0).integers(0, 2, size=(96, 1)).astype(np.float32))
    batches = [np.random.default_rng(i).permutation(96)[:32] for i in range(5)]
    torch.manual_seed(0)
    ref = M.GCN(dim=64, layers=3, dropout=0.0, fused=False).to(cuda)
    fast = M.GCN(dim=64, layers=3, dropout=0.0, fused=True).to(cuda)
    M.copy_parameters(ref, fast)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    opt_fast = torch.optim.Adam(fast.parameters(), lr=1e-3, capturable=True)
    lossf = torch.nn.functional.binary_cross_entropy_with_logits

    want = []
    for ids in batches:
        g = dgl.batch([mols[i] for i in ids]).to(cuda).int().formats("coo")
        opt_ref.zero_grad()
        loss = lossf(ref(g, g.ndata["feat"], g.edata["feat"]), labels[ids].to(cuda))
        loss.backward()
        opt_ref.step()
        want.append(float(loss))

    store = dgl.GraphStore(mols, labels, device=cuda)
    n_pad, e_pad = store.pad_sizes(batches, multiple=64)
    sb = store.static_batch(32, n_pad, e_pad)

    def step():
        sb.refresh()
        g = sb.graph
        pred = fast(g, g.ndata["feat"], g.edata["feat"], sb.node_mask, sb.n_real_nodes)
        loss = lossf(pred[:32], sb.labels)
        loss.backward()
        opt_fast.step()
        return loss

    # capture on a side stream after the warm-up torch asks for; the warm-up must not move the parameters
    state = {k: v.clone() for k, v in fast.state_dict().items()}
    sb.set_ids(batches[0])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            opt_fast.zero_grad(set_to_none=True)
            step()
    torch.cuda.current_stream().wait_stream(side)
    fast.load_state_dict(state)
    opt_fast = torch.optim.Adam(fast.parameters(), lr=1e-3, capturable=True)
    with torch.cuda.stream(side):      # Adam state initialisation outside the capture
        opt_fast.zero_grad(set_to_none=True)
        step()
    torch.cuda.current_stream().wait_stream(side)
    fast.load_state_dict(state)
    for st in opt_fast.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()
    graph = torch.cuda.CUDAGraph()
    opt_fast.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        static_loss = step()
    fast.load_state_dict(state)
    for st in opt_fast.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()
    got = []
    for ids in batches:
        sb.set_ids(ids)
        graph.replay()
        got.append(float(static_loss))
    np.testing.assert_allclose(got, want, rtol=2e-4)
    fsd = fast.state_dict()
    for k, a in ref.state_dict().items():
        if a.dtype.is_floating_point and k in fsd:
            np.testing.assert_allclose(n(fsd[k]), n(a), rtol=0, atol=5e-4, err_msg=k)
    np.testing.assert_allclose(n(fast.atom.weight), n(torch.cat([e.weight for e in ref.atom.embs], 0)), rtol=0, atol=5e-4)


@pytest.mark.parametrize("n_rows,D", [(1, 256), (777, 256), (5000, 64), (300, 7), (0, 32)])
def test_categorical_embedding_sum(cuda, n_rows, D):
    """AtomEncoder arithmetic (sum of 9 per-column embedding lookups, in column order) in one kernel: forward bit-exact
    against the sequential torch loop, backward against torch's embedding backward in float64 (deterministic here)."""
    dims = [119, 4, 12, 12, 10, 6, 6, 2, 2]
    rng = np.random.default_rng(n_rows + D)
    x = torch.from_numpy(np.stack([rng.integers(0, d, size=n_rows) for d in dims], 1).astype(np.int64)).to(cuda)
    table = torch.randn(sum(dims), D, device=cuda, requires_grad=True)
    got = dgl.ops.categorical_embedding_sum(x, table, dims)
    offs = np.concatenate([[0], np.cumsum(dims)])
    want = 0
    for k in range(len(dims)):
        want = want + table.detach()[offs[k] + x[:, k]]
    if n_rows:
        assert torch.equal(got, want)
    gout = torch.randn(n_rows, D, device=cuda)
    got.backward(gout)
    t64 = table.detach().double().requires_grad_(True)
    w64 = 0
    for k in range(len(dims)):
        w64 = w64 + t64[offs[k] + x[:, k]]
    if n_rows:
        w64.backward(gout.double())
        scale = torch.zeros_like(t64)
        for k in range(len(dims)):
            scale.index_add_(0, offs[k] + x[:, k], gout.double().abs())
        assert_close_sumscaled(n(table.grad), n(t64.grad), n(scale), rtol=1e-5, what="categorical_embedding_sum grad")
        g2 = table.grad.clone()
        table.grad = None
        dgl.ops.categorical_embedding_sum(x, table, dims).backward(gout)
        assert torch.equal(table.grad, g2)                       # deterministic
    else:
        assert table.grad is None or float(table.grad.abs().sum()) == 0.0
