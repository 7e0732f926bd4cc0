"""-m gpu: device COO->CSC/CSR is bit-identical to the CPU order oracle (COOToCSR)."""
import numpy as np
import pytest
import torch

import dgl
from gpu_util import graphs, n

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,order", [("uniform", "shuffled"), ("uniform", "dst_sorted"), ("powerlaw", "shuffled")])
@pytest.mark.parametrize("shape", [(50, 50, 400), (1000, 700, 30000), (5, 9, 3)])
def test_csc_csr_bit_exact(oracle, cuda, kind, order, shape):
    ns, nd, ne = shape
    og, g, src, dst = graphs(oracle, ns, nd, ne, seed=ne, kind=kind, order=order)
    for which, (indptr, indices, data) in (("csc", og.csc), ("csr", og.csr)):
        view = getattr(g._graph, which)()
        assert np.array_equal(n(view.indptr), indptr), which
        assert np.array_equal(n(view.indices), indices), which
        if view.eids is None:  # identity permutation is dropped
            assert np.array_equal(data, np.arange(ne)), which
        else:
            assert np.array_equal(n(view.eids), data), which
    if order == "dst_sorted":
        assert g._graph.csc().eids is None
    assert np.array_equal(n(g.in_degrees()), og.in_degrees())
    assert np.array_equal(n(g.out_degrees()), og.out_degrees())


def test_reverse_swaps_formats_zero_copy(oracle, cuda):
    og, g, src, dst = graphs(oracle, 40, 40, 300, seed=1)
    csc = g._graph.csc()
    assert g._graph.reverse().csr() is csc
    ogr = og.reverse()
    assert np.array_equal(n(g._graph.reverse().csc().indices), ogr.csc[1])


def test_empty_graph_and_isolated_nodes(oracle, cuda):
    g = dgl.graph((torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)), num_nodes=4).int().to(cuda)
    assert n(g._graph.csc().indptr).tolist() == [0, 0, 0, 0, 0]
    assert n(g.in_degrees()).tolist() == [0, 0, 0, 0]
    g = dgl.graph((torch.tensor([7]), torch.tensor([2])), num_nodes=10).int().to(cuda)
    assert n(g._graph.csc().indptr).tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1]
    assert n(g._graph.csr().indptr).tolist() == [0] * 8 + [1, 1, 1]


def test_hub_row_list(oracle, cuda):
    og, g, src, dst = graphs(oracle, 2000, 2000, 60000, seed=3, kind="powerlaw")
    deg = og.in_degrees()
    info = g._graph.csc().hubs(100)
    want = np.nonzero(deg > 100)[0]
    assert info.n_hub == len(want) and info.n_hub > 0
    assert n(info.rows).tolist() == want.tolist()
    nseg = -(-deg[want] // 100)
    assert info.n_seg == nseg.sum() and info.seg_len == 100
    assert n(info.seg_ptr).tolist() == [0] + np.cumsum(nseg).tolist()
    assert n(info.seg_hub).tolist() == np.repeat(np.arange(len(want)), nseg).tolist()
    assert g._graph.csc().hubs(10 ** 9) is None
