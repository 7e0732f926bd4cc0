"""CPU tests of the host-side mirror of the reference interface (no kernels are launched)."""
import numpy as np
import pytest
import torch

import dgl
import dgl.function as fn
from dgl import sparse as K


def test_infer_broadcast_shape():
    assert K.infer_broadcast_shape("mul", (4, 8), (4, 1)) == (4, 8)
    assert K.infer_broadcast_shape("add", (8,), (4, 1)) == (4, 8)
    assert K.infer_broadcast_shape("dot", (4, 8), (4, 8)) == (4, 1)
    assert K.infer_broadcast_shape("copy_lhs", (7,), (1,)) == (7,)
    with pytest.raises(dgl.DGLError):
        K.infer_broadcast_shape("add", (3,), (4,))


def test_graph_construction_and_queries():
    g = dgl.graph((torch.tensor([0, 1, 2, 2]), torch.tensor([1, 2, 0, 0])))
    assert g.number_of_nodes() == 3 and g.number_of_edges() == 4
    assert g.number_of_src_nodes() == 3 and g.number_of_dst_nodes() == 3
    assert g.idtype == torch.int64 and g.int().idtype == torch.int32
    assert g.in_degrees().tolist() == [2, 1, 1] and g.out_degrees().tolist() == [1, 1, 2]
    assert isinstance(g, dgl.DGLHeteroGraph) and isinstance(g, dgl.DGLGraph)
    s, d = dgl.graph(g.edges()).edges()  # kernel/utils.py:39 round trip
    assert s.tolist() == [0, 1, 2, 2] and d.tolist() == [1, 2, 0, 0]
    assert "num_nodes=3" in repr(g)
    with pytest.raises(dgl.DGLError):
        dgl.graph((torch.tensor([0, 5]), torch.tensor([1, 2])), num_nodes=3)


def test_self_loops_are_appended_after_existing_edges():
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 2])), num_nodes=3)
    g.edata["w"] = torch.tensor([1.0, 2.0])
    h = dgl.add_self_loop(g)
    assert h.edges()[0].tolist() == [0, 1, 0, 1, 2] and h.edges()[1].tolist() == [1, 2, 0, 1, 2]
    assert h.edata["w"].tolist() == [1.0, 2.0, 0.0, 0.0, 0.0]
    assert dgl.remove_self_loop(h).number_of_edges() == 2


def test_to_bidirected_dedups():
    g = dgl.graph((torch.tensor([0, 1, 1, 0]), torch.tensor([1, 0, 2, 1])), num_nodes=3)
    b = dgl.to_bidirected(g)
    pairs = sorted(zip(b.edges()[0].tolist(), b.edges()[1].tolist()))
    assert pairs == [(0, 1), (1, 0), (1, 2), (2, 1)]


def test_from_networkx_undirected_gives_both_directions():
    import networkx as nx
    nxg = nx.Graph()
    nxg.add_edges_from([(0, 1), (1, 2)])
    g = dgl.from_networkx(nxg)
    assert g.number_of_nodes() == 3 and g.number_of_edges() == 4


def test_frames_and_local_scopes():
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 0])))
    g.ndata["h"] = torch.ones(2, 3)
    assert g.srcdata["h"] is g.dstdata["h"]
    with pytest.raises(dgl.DGLError):
        g.ndata["bad"] = torch.ones(5)
    lv = g.local_var()
    lv.ndata["tmp"] = torch.zeros(2)
    assert "tmp" not in g.ndata and "h" in lv.ndata
    with g.local_scope():
        g.ndata["tmp2"] = torch.zeros(2)
    assert "tmp2" not in g.ndata
    assert g.srcdata.pop("h").shape == (2, 3) and "h" not in g.ndata


def test_formats_restriction():
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 0])))
    h = g.formats(["csr", "csc"])
    assert h.formats()["created"] == ["csc", "csr"]
    assert g.formats("coo").formats()["created"] == ["coo"]
    with pytest.raises(dgl.DGLError):
        g.formats(["ell"])


def test_batch_offsets_and_sizes():
    g1 = dgl.graph((torch.tensor([0]), torch.tensor([1])), num_nodes=2)
    g2 = dgl.graph((torch.tensor([0, 2]), torch.tensor([1, 1])), num_nodes=3)
    g1.ndata["x"] = torch.zeros(2, 1)
    g2.ndata["x"] = torch.ones(3, 1)
    bg = dgl.batch([g1, g2])
    assert bg.number_of_nodes() == 5 and bg.batch_size == 2
    assert bg.edges()[0].tolist() == [0, 2, 4] and bg.edges()[1].tolist() == [1, 3, 3]
    assert bg.batch_num_nodes().tolist() == [2, 3] and bg.ndata["x"].sum().item() == 3


def test_builtin_function_descriptors():
    m = fn.u_mul_e("h", "w", "m")
    assert (m.lhs, m.rhs, m.binary_op, m.name) == ("u", "e", "mul", "u_mul_e")
    assert fn.copy_src("h", "m").name == "copy_u" and fn.copy_u("h", "m").target == "u"
    assert fn.mean("m", "o").name == "mean" and fn.src_mul_edge("a", "b", "c").name == "u_mul_e"


def test_ops_namespace_has_the_generated_shorthands():
    for name in ("gspmm", "gsddmm", "edge_softmax", "copy_u_sum", "copy_e_max", "u_mul_e_sum", "u_add_v",
                 "u_dot_v", "e_div_v", "copy_u_mean"):
        assert hasattr(dgl.ops, name), name


def test_kernels_refuse_cpu_tensors_loudly():
    """No CPU fallback behind the operator API: CPU operands raise instead of computing."""
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 0])))
    with pytest.raises(dgl.DGLError, match="CUDA-only"):
        dgl.ops.gspmm(g, "copy_lhs", "sum", torch.rand(2, 4), None)
    with pytest.raises(dgl.DGLError, match="CUDA-only"):
        dgl.ops.gsddmm(g, "dot", torch.rand(2, 4), torch.rand(2, 4))
    with pytest.raises(dgl.DGLError, match="CUDA-only"):
        dgl.ops.edge_softmax(g, torch.rand(2, 1))


def test_operand_validation_mirrors_upstream_errors():
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 0])))
    with pytest.raises(dgl.DGLError):
        dgl.ops.gspmm(g, "mul", "sum", torch.rand(2, 4), None)      # missing edge operand
    with pytest.raises(dgl.DGLError):
        dgl.ops.gspmm(g, "pow", "sum", torch.rand(2, 4), torch.rand(2, 4))
    with pytest.raises(dgl.DGLError):
        dgl.ops.gspmm(g, "copy_lhs", "prod", torch.rand(2, 4), None)


def test_reduce_grad_sums_broadcast_dims():
    from dgl.backend import _reduce_grad, _need_reduce_last_dim
    g = torch.arange(24.0).reshape(2, 3, 4)
    assert torch.equal(_reduce_grad(g, (2, 3, 1)), g.sum(-1, keepdim=True))
    assert torch.equal(_reduce_grad(g, (2, 4)), g.sum(1))
    assert _reduce_grad(g, (2, 3, 4)) is g
    assert _need_reduce_last_dim(torch.zeros(5, 4, 8), torch.zeros(9, 4, 1))
    assert not _need_reduce_last_dim(torch.zeros(5, 4, 8), torch.zeros(9, 4, 8))


def test_synthetic_generator_is_seeded_and_shaped():
    from dgl.data import synthetic
    a = synthetic.random_edges(100, 80, 1000, seed=1, degree="powerlaw")
    b = synthetic.random_edges(100, 80, 1000, seed=1, degree="powerlaw")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].max() < 100 and a[1].max() < 80
    s, d = synthetic.random_edges(50, 50, 500, seed=2, order="dst_sorted")
    assert np.all(np.diff(d) >= 0)
    n, s, d = synthetic.shaped_edges("cora", self_loops=True)
    assert n == 2708 and len(s) == 10556 + 2708 and np.array_equal(s[-2708:], np.arange(2708))


def test_bench_byte_model_matches_survey_appendix_c():
    """bench.py's algorithmic-byte formulas reproduce SURVEY.md Appendix C (reddit 11.6 M edges)."""
    import importlib.util, os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n, e = 232965, 11606919
    assert abs(bench.spmm_bytes(n, e, 64) / 1e9 - 3.08) < 0.01
    assert abs(bench.spmm_bytes(n, e, 602) / 1e9 - 28.56) < 0.01
    assert abs(bench.sddmm_dot_bytes(n, e, 602, p=0) / 1e9 - 28.6) < 0.1
    assert bench.spmm_bytes(n, e, 602) // e == 2460            # 2 460 B per edge
    assert abs(bench.step_bytes(n, e) / 1e9 - 100.2) < 0.2


def test_masked_batchnorm_equals_batchnorm_on_the_real_rows():
    """examples/small_graph_model.MaskedBatchNorm1d (what lets a padded StaticBatch reproduce the unpadded numbers):
    with a 0/1 row mask it must give, on the real rows, exactly what nn.BatchNorm1d gives when applied to those rows
    alone -- outputs, gradients and running statistics, in training and in eval mode."""
    import torch
    import sys
    import os
    from conftest import ROOT, PKG
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from examples.small_graph_model import MaskedBatchNorm1d
    torch.manual_seed(0)
    n_real, n_pad, d = 37, 48, 5
    x = torch.randn(n_pad, d, dtype=torch.float64)
    x[n_real:] = 1e3                                        # padding rows hold garbage
    mask = torch.zeros(n_pad, 1, dtype=torch.float64)
    mask[:n_real] = 1
    count = torch.tensor(float(n_real), dtype=torch.float64)
    ref = torch.nn.BatchNorm1d(d).double()
    got = MaskedBatchNorm1d(d).double()
    with torch.no_grad():
        for m in (ref, got):
            m.weight.copy_(torch.linspace(0.5, 1.5, d))
            m.bias.copy_(torch.linspace(-1, 1, d))
    for step in range(3):
        xr = (x[:n_real] + step).clone().requires_grad_(True)
        xg = (x + step).clone().requires_grad_(True)
        yr = ref(xr)
        yg = got(xg, mask, count)
        torch.testing.assert_close(yg[:n_real], yr, rtol=1e-10, atol=1e-10)
        w = torch.randn(n_real, d, dtype=torch.float64)
        (yr * w).sum().backward()
        (yg[:n_real] * w).sum().backward()
        torch.testing.assert_close(xg.grad[:n_real], xr.grad, rtol=1e-9, atol=1e-10)
        assert float(xg.grad[n_real:].abs().max()) == 0.0   # padding rows receive no gradient from the real rows' loss
        torch.testing.assert_close(got.running_mean, ref.running_mean, rtol=1e-10, atol=1e-12)
        torch.testing.assert_close(got.running_var, ref.running_var, rtol=1e-10, atol=1e-12)
    assert int(got.num_batches_tracked) == int(ref.num_batches_tracked) == 3
    ref.eval(); got.eval()
    torch.testing.assert_close(got(x, mask, count)[:n_real], ref(x[:n_real]), rtol=1e-10, atol=1e-10)
    torch.testing.assert_close(got(x[:n_real]), ref(x[:n_real]))   # without a mask it IS nn.BatchNorm1d
