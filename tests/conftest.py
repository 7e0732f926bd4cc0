import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dgl-0.5-benchmark_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): compiled on first use."""
    from oracle import dgl_ref
    dgl_ref.build()
    return dgl_ref


def make_edges(n_src, n_dst, n_edges, seed, kind="uniform", order="shuffled"):
    from dgl.data import synthetic
    return synthetic.random_edges(n_src, n_dst, n_edges, seed=seed, degree=kind, order=order)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def assert_close_sumscaled(actual, expected, scale, rtol=1e-5, what=""):
    """|a - b| <= rtol * scale elementwise, where `scale` is the sum of |terms| that produced the
    element (abs-sum-scaled relative error, robust to cancellation)."""
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    err = np.abs(actual - expected)
    bound = rtol * np.maximum(scale, 1e-30) + 1e-37
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err / bound), err.shape)
        raise AssertionError("%s: %d elements exceed %g * scale; worst at %s: got %r want %r scale %r"
                             % (what, bad.sum(), rtol, i, actual[i], expected[i], scale[i]))


@pytest.fixture
def small_hub_threshold():
    """Lower the split-row threshold so modest test graphs exercise the hub kernels."""
    from dgl import sparse as K
    old = K.HUB_THRESHOLD
    K.HUB_THRESHOLD = 48
    yield 48
    K.HUB_THRESHOLD = old
