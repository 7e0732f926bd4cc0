"""Seeded synthetic graphs with the node / edge counts and feature widths of the datasets the
reference benchmarks (README.md:19-32) -- the datasets themselves are not available offline.

Two degree models: `uniform` (src, dst i.i.d. uniform; multigraph allowed, as DGL
graphs are multigraphs) and `powerlaw` (destinations drawn with Zipf-like weights, exponent ~2.1,
max in-degree capped near 0.1*N, mimicking reddit's hub nodes); each in two edge orders:
`shuffled` (the CSC edge-id permutation is non-trivial) or `dst_sorted` (identity permutation).
"""
import sys

import numpy as np

_announced = set()


def announce_synthetic(name, what):
    """Loud one-time notice (stderr) whenever a stand-in for a real dataset NAME is instantiated: accuracy /
    rocauc / epoch-time lines an unchanged reference script prints afterwards are for seeded random data of
    that dataset's shape, not for the dataset.  Every stand-in object also carries `.synthetic = True`."""
    if name in _announced:
        return
    _announced.add(name)
    sys.stderr.write("[dgl-b200] WARNING: dataset '%s' is a SEEDED SYNTHETIC stand-in (%s); the real dataset is not "
                     "available offline -- accuracies printed from it are meaningless, only shapes and timings "
                     "carry over.\n" % (name, what))
    sys.stderr.flush()


# name -> (num_nodes, num_directed_edges, feature_width, num_classes)   [README.md:19-25, BASELINE.json]
SHAPES = {
    "cora": (2708, 10556, 1433, 7),
    "citeseer": (3327, 9228, 3703, 6),
    "pubmed": (19717, 88651, 500, 3),
    "reddit": (232965, 11606919, 602, 41),
    "reddit-full": (232965, 114615892, 602, 41),
    "ogbn-arxiv": (169343, 1166243, 128, 40),
    "ogbn-products": (2449029, 61859140, 100, 47),
    "ogbn-products-full": (2449029, 123718280, 100, 47),
    "ogbn-proteins": (132534, 79122504, 8, 112),
}


def random_edges(n_src, n_dst, n_edges, seed=0, degree="uniform", order="shuffled", alpha=2.1,
                 max_frac=0.1):
    """(src, dst) int64 numpy arrays of length n_edges."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_src, size=n_edges, dtype=np.int64)
    if degree == "uniform":
        dst = rng.integers(0, n_dst, size=n_edges, dtype=np.int64)
    elif degree == "powerlaw":
        # in-degree weights w_i ~ rank^(-1/(alpha-1)), capped so the largest hub holds <= max_frac*n_src edges
        ranks = np.arange(1, n_dst + 1, dtype=np.float64)
        w = ranks ** (-1.0 / (alpha - 1.0))
        w /= w.sum()
        cap = max_frac * n_src / max(n_edges, 1)
        for _ in range(8):
            over = w > cap
            if not over.any():
                break
            excess = (w[over] - cap).sum()
            w[over] = cap
            w[~over] += excess * w[~over] / w[~over].sum()
        perm = rng.permutation(n_dst)  # hubs scattered over the id space
        cdf = np.cumsum(w)
        cdf[-1] = 1.0
        dst = perm[np.searchsorted(cdf, rng.random(n_edges), side="right").clip(0, n_dst - 1)]
    else:
        raise ValueError("degree must be 'uniform' or 'powerlaw'")
    if order == "dst_sorted":
        o = np.argsort(dst, kind="stable")
        src, dst = src[o], dst[o]
    elif order != "shuffled":
        raise ValueError("order must be 'shuffled' or 'dst_sorted'")
    return src, dst


def shaped_edges(name, seed=0, degree="uniform", order="shuffled", self_loops=False):
    """Edges of a graph with the node/edge counts of dataset `name`; `self_loops` appends (i, i) for
    every node AFTER the other edges, as dgl.add_self_loop does (GAT scripts)."""
    n, e, _, _ = SHAPES[name]
    src, dst = random_edges(n, n, e, seed=seed, degree=degree, order=order)
    if self_loops:
        loops = np.arange(n, dtype=np.int64)
        src, dst = np.concatenate([src, loops]), np.concatenate([dst, loops])
    return n, src, dst


def molecule_like_batch(batch_size, seed=0, mean_nodes=25.5, extra_edge_frac=0.08):
    """ogbg-molhiv-shaped batch: per graph Poisson(mean_nodes) (>=2) nodes, a random tree plus a few
    ring-closing edges, both directions (~27.5 undirected edges per graph, README.md:31).
    Returns (src, dst, nodes_per_graph) with node ids already offset into the batched graph."""
    rng = np.random.default_rng(seed)
    srcs, dsts, sizes = [], [], []
    off = 0
    for _ in range(batch_size):
        n = max(2, int(rng.poisson(mean_nodes)))
        parent = np.array([rng.integers(0, i) for i in range(1, n)], dtype=np.int64)
        child = np.arange(1, n, dtype=np.int64)
        k = int(round(extra_edge_frac * n))
        a = rng.integers(0, n, size=k)
        b = rng.integers(0, n, size=k)
        keep = a != b
        u = np.concatenate([parent, a[keep]])
        v = np.concatenate([child, b[keep]])
        srcs.append(np.concatenate([u, v]) + off)
        dsts.append(np.concatenate([v, u]) + off)
        sizes.append(n)
        off += n
    return np.concatenate(srcs), np.concatenate(dsts), np.array(sizes, dtype=np.int64)
