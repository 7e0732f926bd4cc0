"""-m gpu parity of the two secondary call sites of SURVEY.md section 8:

  a10  RelGraphConv (main_dgl_proteins_rgcn_for.py:46-60): inside g.local_scope(), once per relation,
       update_all(fn.u_mul_e('feat','weight','m'), fn.mean('m','rel_out')) with (E,1) edge weights that are
       NON-CONTIGUOUS column slices of an (E,R) tensor (:159-161), node widths 1 (layer 1) and 32; forward is
       bit-identical to the oracle (un-split rows keep the CPU kernel's order), backward within 1e-5.
  a11  GCNConv on batched COO-only graphs (main_dgl_molhiv_gcn.py:37-52,101): dgl.batch -> .to(device).int()
       .formats('coo'), in_degrees()+1, a Python UDF message norm*relu(x[src]+w) and builtin fn.sum; and the
       enzymes variant (main_dgl_enzymes_gcn.py:30-39): copy_u sum on the same kind of graph.
"""
import numpy as np
import pytest
import torch

import dgl
import dgl.function as fn
from gpu_util import graphs, n, t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("width", [1, 32])
@pytest.mark.parametrize("kind,order", [("uniform", "shuffled"), ("powerlaw", "shuffled"), ("uniform", "dst_sorted")])
def test_rgcn_u_mul_e_mean_scalar_weights(oracle, cuda, width, kind, order, small_hub_threshold):
    N, E, R = 700, 9000, 8
    og, g, src, dst = graphs(oracle, N, N, E, seed=21, kind=kind, order=order)
    rng = np.random.default_rng(21)
    X = rng.standard_normal((N, width)).astype(np.float32)
    Wall = rng.random((E, R), dtype=np.float32)
    Xt = t(X).requires_grad_(True)
    Wt = t(Wall)
    outs = []
    with g.local_scope():
        g.ndata["feat"] = Xt
        for rel in range(R):
            w = Wt[:, rel:rel + 1]                       # (E,1), stride (R,1): what the script passes
            assert not w.is_contiguous() or R == 1
            g.edata["weight"] = w
            g.update_all(fn.u_mul_e("feat", "weight", "m"), fn.mean("m", "rel_out"))
            outs.append(g.ndata.pop("rel_out"))
    assert "feat" not in g.ndata and "weight" not in g.edata      # local_scope leaves the graph untouched
    hub_rows = (np.bincount(dst, minlength=N) > small_hub_threshold)
    for rel in range(R):
        want = oracle.gspmm(og, "mul", "mean", X, np.ascontiguousarray(Wall[:, rel:rel + 1]))
        got = n(outs[rel])
        assert got.shape == want.shape == (N, width)
        # rows the row-per-group kernel handles are accumulated in CSC order like the CPU kernel: bit-exact
        assert np.array_equal(got[~hub_rows], want[~hub_rows])
        np.testing.assert_allclose(got[hub_rows], want[hub_rows], rtol=1e-5, atol=1e-6)   # segment tree order
    # backward of the sum over relations (what torch.stack(...).sum(0) does upstream of the matmuls)
    gout = rng.standard_normal((R, N, width)).astype(np.float32)
    torch.stack(outs, 0).mul(t(gout)).sum().backward()
    deg = np.maximum(og.in_degrees(), 1).astype(np.float32)
    want_dx = np.zeros_like(X, dtype=np.float64)
    for rel in range(R):
        dZ = (gout[rel] / deg[:, None]).astype(np.float32)
        dX, _ = oracle.gspmm_sum_backward(og, "mul", X, np.ascontiguousarray(Wall[:, rel:rel + 1]), dZ)
        want_dx += dX
    scale = np.zeros_like(want_dx)
    np.add.at(scale, src, np.abs(gout[:, dst, :] / deg[None, dst, None] * Wall.T[:, :, None]).sum(0))
    err = np.abs(n(Xt.grad) - want_dx)
    assert (err <= 1e-5 * scale + 1e-7).all(), float((err / (scale + 1e-30)).max())


def _molecule_batch(batch_size, seed):
    from dgl.data import synthetic
    gs = []
    rng = np.random.default_rng(seed)
    for i in range(batch_size):
        s, d, sizes = synthetic.molecule_like_batch(1, seed=seed * 1000 + i)
        g = dgl.graph((torch.from_numpy(s), torch.from_numpy(d)), num_nodes=int(sizes[0]))
        g.ndata["x"] = torch.from_numpy(rng.standard_normal((int(sizes[0]), 24)).astype(np.float32))
        g.edata["w"] = torch.from_numpy(rng.standard_normal((len(s), 24)).astype(np.float32))
        gs.append(g)
    return gs


@pytest.mark.parametrize("batch_size", [1, 64, 256])
def test_batched_coo_udf_message_sum(oracle, cuda, batch_size):
    """main_dgl_molhiv_gcn.py:37-52 on a batched '.formats(coo)' graph vs the oracle (copy_e sum of the same
    message tensor), forward bit-exact, gradients through the UDF within 1e-5."""
    gs = _molecule_batch(batch_size, seed=5)
    bg_cpu = dgl.batch(gs)
    assert bg_cpu.batch_size == batch_size
    bg = bg_cpu.to(cuda).int().formats("coo")
    assert bg.idtype == torch.int32
    src, dst = (x.numpy() for x in bg_cpu.edges())
    N = bg_cpu.number_of_nodes()
    og = oracle.OracleGraph(src, dst, N, N)
    # the batch keeps each member graph's edges contiguous and shifts ids by the node offsets
    off = 0
    for g_ in gs:
        s_, d_ = g_.edges()
        assert (bg_cpu.edges()[0][off:off + len(s_)] - s_ == bg_cpu.edges()[0][off] - s_[0]).all()
        off += len(s_)
    x = bg.ndata["x"].clone().requires_grad_(True)
    w = bg.edata["w"].clone().requires_grad_(True)
    deg = bg.in_degrees().float().unsqueeze(1) + 1
    assert np.array_equal(n(bg.in_degrees()), og.in_degrees())                     # int32 degrees: bit-exact
    norm = deg.pow(-0.5)

    def message(edges):
        return {"m": edges.src["norm"] * edges.dst["norm"] * torch.relu(edges.src["x"] + edges.data["w"])}

    g2 = bg.local_var()
    g2.ndata["x"], g2.ndata["norm"], g2.edata["w"] = x, norm, w
    g2.update_all(message, fn.sum("m", "h"))
    h = g2.ndata["h"]
    xn, wn, nn_ = n(x), n(w), n(norm)
    msg = (nn_[src] * nn_[dst] * np.maximum(xn[src] + wn, 0)).astype(np.float32)
    # torch computes the message with the same fp32 ops in the same order as numpy does here
    want = oracle.gspmm(og, "copy_rhs", "sum", None, msg)
    assert np.array_equal(n(h), want)
    gout = np.random.default_rng(1).standard_normal(h.shape).astype(np.float32)
    h.backward(t(gout))
    x64 = torch.tensor(xn, dtype=torch.float64, requires_grad=True)
    w64 = torch.tensor(wn, dtype=torch.float64, requires_grad=True)
    s_, d_ = torch.from_numpy(src).long(), torch.from_numpy(dst).long()
    n64 = torch.tensor(nn_, dtype=torch.float64)
    m64 = n64[s_] * n64[d_] * torch.relu(x64[s_] + w64)
    torch.zeros((N, 24), dtype=torch.float64).index_add_(0, d_, m64).backward(torch.tensor(gout, dtype=torch.float64))
    np.testing.assert_allclose(n(x.grad), x64.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(n(w.grad), w64.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_batched_coo_copy_u_sum_and_readout(oracle, cuda):
    """main_dgl_enzymes_gcn.py:30-39 (copy_u sum on a COO-only batch, int64 ids: the train loop there does not
    call .int()) and the AvgPooling readout (main_dgl_molhiv_gcn.py:75,93) vs per-graph means."""
    gs = _molecule_batch(48, seed=9)
    bg_cpu = dgl.batch(gs)
    bg = bg_cpu.to(cuda).formats("coo")
    src, dst = (x.numpy() for x in bg_cpu.edges())
    N = bg_cpu.number_of_nodes()
    og = oracle.OracleGraph(src, dst, N, N)
    g2 = bg.local_var()
    deg = g2.in_degrees().float().unsqueeze(1) + 1
    hx = bg.ndata["x"] * deg.pow(-0.5)
    g2.ndata["h"] = hx
    g2.update_all(fn.copy_u("h", "m"), fn.sum("m", "h"))
    assert np.array_equal(n(g2.ndata["h"]), oracle.gspmm(og, "copy_lhs", "sum", n(hx), None))
    from dgl.nn import AvgPooling
    pooled = n(AvgPooling()(bg, bg.ndata["x"]))
    sizes = np.array([g_.number_of_nodes() for g_ in gs])
    xs = np.split(n(bg.ndata["x"]).astype(np.float64), np.cumsum(sizes)[:-1])
    want = np.stack([a.mean(0) for a in xs])
    np.testing.assert_allclose(pooled, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("R", [2, 4, 8])
@pytest.mark.parametrize("D", [4, 32, 100])
@pytest.mark.parametrize("order", ["shuffled", "dst_sorted"])
def test_rgcn_batched_relations_one_launch(oracle, cuda, R, D, order):
    """All R relations of the RGCN layer in ONE gspmm: lhs (N,1,D) x rhs (E,R,1) -> (N,R,D) (relation-broadcast kernel:
    a neighbour row is gathered once for all relations).  Every relation's slice is bit-identical to the per-relation
    u_mul_e_{sum,mean} the unchanged script computes (main_dgl_proteins_rgcn_for.py:50-53), and to the oracle; the
    gradient w.r.t. the node features equals the sum of the per-relation gradients."""
    from dgl import sparse as K
    N, E = 500, 40000                                      # in-degree ~80: long rows, as on ogbn-proteins
    og, g, src, dst = graphs(oracle, N, N, E, seed=23, order=order)
    old_thr, K.HUB_THRESHOLD = K.HUB_THRESHOLD, 1 << 30    # per-relation reference without split rows: sequential sums
    rng = np.random.default_rng(23)
    X = rng.standard_normal((N, D)).astype(np.float32)
    W = rng.random((E, R), dtype=np.float32)
    for red in ("sum", "mean"):
        xt = t(X).requires_grad_(True)
        launches0 = dgl._capi.launches()
        out = dgl.ops.gspmm(g, "mul", red, xt.unsqueeze(1), t(W).unsqueeze(-1))
        assert dgl._capi.launches() - launches0 == 1
        assert out.shape == (N, R, D)
        want = oracle.gspmm(og, "mul", red, X[:, None, :], W[:, :, None])
        assert np.array_equal(n(out), want)
        gout = rng.standard_normal((N, R, D)).astype(np.float32)
        out.backward(t(gout))
        ref_grad = torch.zeros(N, D, device="cuda")
        for r in range(R):
            xr = t(X).requires_grad_(True)
            o_r = dgl.ops.gspmm(g, "mul", red, xr, t(np.ascontiguousarray(W[:, r:r + 1])))
            assert torch.equal(o_r, out[:, r, :].detach())                       # per-relation result, bit for bit
            o_r.backward(t(np.ascontiguousarray(gout[:, r, :])))
            ref_grad += xr.grad
        np.testing.assert_allclose(n(xt.grad), n(ref_grad), rtol=1e-4, atol=1e-4)
    K.HUB_THRESHOLD = old_thr
