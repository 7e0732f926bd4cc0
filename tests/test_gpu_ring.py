"""-m gpu parity of the bulk-copy ring kernels (csrc/ring.cu: whole-row cp.async.bulk into a per-warp shared-memory
ring, nnz-balanced persistent warps) against the CPU oracle.  The library only takes this path for wide rows on
large graphs; the environment knobs it reads per call are lowered here so small seeded graphs exercise it,
including rows left to the hub kernels, empty rows, the 8-byte-aligned rows of D = 602, the 4-byte-aligned rows of
odd widths and of bf16 D = 602, and the highest row of X (copied by hand: a rounded-up bulk copy would read past
the tensor)."""
import os

import numpy as np
import pytest
import torch

import dgl
from conftest import assert_close_sumscaled
from gpu_util import graphs, n, t

pytestmark = pytest.mark.gpu


@pytest.fixture
def ring_on():
    keys = {"DGLB_RING_MIN_BYTES": "64", "DGLB_RING_MIN_NNZ": "0"}
    old = {k: os.environ.get(k) for k in list(keys) + ["DGLB_RING_STAGES"]}
    os.environ.update(keys)
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("D", [16, 64, 100, 255, 256, 602, 1000, 1433])
@pytest.mark.parametrize("kind", ["uniform", "powerlaw"])
@pytest.mark.parametrize("stages", [0, 2])
def test_ring_copy_u_sum_mean(oracle, cuda, ring_on, small_hub_threshold, D, kind, stages):
    os.environ["DGLB_RING_STAGES"] = str(stages)
    n_src, n_dst, E = 900, 700, 20000
    og, g, src, dst = graphs(oracle, n_src, n_dst, E, seed=31, kind=kind)
    src[:7] = n_src - 1                                   # make sure the highest row of X is gathered
    dst[dst % 9 == 0] = 1                                 # rows 0, 9, 18, ... have no in-edges
    og = oracle.OracleGraph(src, dst, n_src, n_dst)
    g = dgl.create_block((torch.from_numpy(src), torch.from_numpy(dst)), n_src, n_dst).int().to("cuda")
    rng = np.random.default_rng(31)
    X = rng.standard_normal((n_src, D)).astype(np.float32)
    hub = np.bincount(dst, minlength=n_dst) > small_hub_threshold
    for red in ("sum", "mean"):
        got = n(dgl.ops.gspmm(g, "copy_lhs", red, t(X), None))
        want = oracle.gspmm(og, "copy_lhs", red, X, None)
        assert np.array_equal(got[~hub], want[~hub]), (D, kind, red)     # CSR order kept: bit-identical
        scale = np.zeros((n_dst, D)); np.add.at(scale, dst, np.abs(X[src]).astype(np.float64))
        if red == "mean":
            scale /= np.maximum(og.in_degrees(), 1)[:, None]
        assert_close_sumscaled(got, want, scale, 1e-5, "ring gspmm %s" % red)
    assert (og.in_degrees() == 0).any()                  # empty rows were part of the check (-> 0)


@pytest.mark.parametrize("D", [64, 100, 256, 602, 1000])
@pytest.mark.parametrize("kind,order", [("uniform", "shuffled"), ("powerlaw", "shuffled"), ("uniform", "dst_sorted")])
def test_ring_u_dot_v(oracle, cuda, ring_on, small_hub_threshold, D, kind, order):
    N, E = 800, 20000
    og, g, src, dst = graphs(oracle, N, N, E, seed=32, kind=kind, order=order)
    rng = np.random.default_rng(32)
    U = rng.standard_normal((N, D)).astype(np.float32)
    V = rng.standard_normal((N, D)).astype(np.float32)
    got = n(dgl.ops.gsddmm(g, "dot", t(U), t(V)))
    want = oracle.gsddmm(og, "dot", U, V)
    scale = np.abs(U[src].astype(np.float64) * V[dst]).sum(-1, keepdims=True)
    assert_close_sumscaled(got, want, scale, 1e-5, "ring u_dot_v")


@pytest.mark.parametrize("D", [64, 128, 602, 604])
def test_ring_bf16_storage(oracle, cuda, ring_on, D):
    """bf16 storage / fp32 accumulate: |a-b| <= 2^-8 |b| + 1e-5 sum|terms| against fp64 on the bf16-rounded inputs."""
    N, E = 800, 20000
    og, g, src, dst = graphs(oracle, N, N, E, seed=33)
    Xb = torch.randn(N, D, device="cuda").to(torch.bfloat16)
    Vb = torch.randn(N, D, device="cuda").to(torch.bfloat16)
    X64, V64 = Xb.double().cpu().numpy(), Vb.double().cpu().numpy()
    got = dgl.ops.gspmm(g, "copy_lhs", "sum", Xb, None)
    assert got.dtype == torch.bfloat16 and got.is_contiguous()
    want = np.zeros((N, D)); np.add.at(want, dst, X64[src])
    scale = np.zeros((N, D)); np.add.at(scale, dst, np.abs(X64[src]))
    err = np.abs(got.double().cpu().numpy() - want)
    assert (err <= 2.0 ** -8 * np.abs(want) + 1e-5 * scale + 1e-30).all()
    gd = dgl.ops.gsddmm(g, "dot", Xb, Vb).double().cpu().numpy()
    wd = (X64[src] * V64[dst]).sum(-1, keepdims=True)
    sd = np.abs(X64[src] * V64[dst]).sum(-1, keepdims=True)
    assert (np.abs(gd - wd) <= 2.0 ** -8 * np.abs(wd) + 1e-5 * sd + 1e-30).all()


def test_ring_accumulate_flag(oracle, cuda, ring_on):
    from dgl import sparse as K
    N, E, D = 600, 15000, 256
    og, g, src, dst = graphs(oracle, N, N, E, seed=34)
    X = np.random.default_rng(34).random((N, D), dtype=np.float32)
    base = torch.full((N, D), 0.5, device="cuda")
    out = base.clone()
    K._gspmm(g._graph, "copy_lhs", "sum", t(X), None, out=out)
    want = np.float32(0.5) + oracle.gspmm(og, "copy_lhs", "sum", X, None)
    assert np.array_equal(n(out), want.astype(np.float32))
