"""Synthetic dataset objects with the interface the reference scripts use:

    dgl.data.load_data(args)  -> .features .labels .train_mask .val_mask .test_mask .num_labels .graph
        (main_dgl_citation_sage.py:153-166,190: `.graph` is a networkx graph for the citation sets;
         main_dgl_reddit_sage.py:188: a DGLGraph for reddit)
    dgl.data.RedditDataset()[0]   (kernel/utils.py:52)

Sizes come from dgl.data.synthetic.SHAPES; contents are seeded random (no network, no files).
Set DGLB200_DATA_SCALE=<float <= 1> to shrink node and edge counts for quick functional runs.
"""
import os

import numpy as np
import torch

from . import synthetic

__all__ = ["load_data", "RedditDataset", "CoraGraphDataset", "CiteseerGraphDataset", "PubmedGraphDataset",
           "CitationGraphDataset", "SyntheticNodeDataset", "LegacyTUDataset"]


def _scale():
    return float(os.environ.get("DGLB200_DATA_SCALE", "1"))


class SyntheticNodeDataset:
    """Node-classification dataset of a given SHAPES entry."""

    synthetic = True

    def __init__(self, name, seed=0, degree="uniform", undirected=True, self_loop=False, as_networkx=False):
        n, e, d, c = synthetic.SHAPES[name]
        s = _scale()
        n, e = max(8, int(n * s)), max(8, int(e * s))
        synthetic.announce_synthetic(name, "%d nodes, %d edges, %d features, %d classes, uniform random" % (n, e, d, c))
        self.name, self.num_labels, self.num_classes = name, c, c
        rng = np.random.default_rng(seed)
        if undirected:  # both directions of e/2 random undirected edges, like the citation graphs
            a, b = synthetic.random_edges(n, n, e // 2, seed=seed, degree=degree)
            src, dst = np.concatenate([a, b]), np.concatenate([b, a])
        else:
            src, dst = synthetic.random_edges(n, n, e, seed=seed, degree=degree)
        if self_loop:
            loops = np.arange(n)
            src, dst = np.concatenate([src, loops]), np.concatenate([dst, loops])
        self._src, self._dst, self._n = src, dst, n
        self.features = rng.random((n, d), dtype=np.float32)
        self.labels = rng.integers(0, c, size=n).astype(np.int64)
        perm = rng.permutation(n)
        n_train, n_val = max(1, n // 20), max(1, n // 10)
        self.train_mask = np.zeros(n, bool); self.train_mask[perm[:n_train]] = True
        self.val_mask = np.zeros(n, bool); self.val_mask[perm[n_train:n_train + n_val]] = True
        self.test_mask = np.zeros(n, bool); self.test_mask[perm[n_train + n_val:n_train + 2 * n_val]] = True
        self._as_networkx = as_networkx
        self._graph = None

    @property
    def graph(self):
        if self._graph is None:
            if self._as_networkx:
                import networkx as nx
                g = nx.DiGraph()
                g.add_nodes_from(range(self._n))
                g.add_edges_from(zip(self._src.tolist(), self._dst.tolist()))
                self._graph = g
            else:
                self._graph = self[0]
        return self._graph

    def __getitem__(self, idx):
        from ..heterograph import graph as make_graph
        assert idx == 0
        g = make_graph((torch.from_numpy(self._src), torch.from_numpy(self._dst)), num_nodes=self._n)
        g.ndata["feat"] = torch.from_numpy(self.features)
        g.ndata["label"] = torch.from_numpy(self.labels)
        g.ndata["train_mask"] = torch.from_numpy(self.train_mask)
        g.ndata["val_mask"] = torch.from_numpy(self.val_mask)
        g.ndata["test_mask"] = torch.from_numpy(self.test_mask)
        return g

    def __len__(self):
        return 1


class LegacyTUDataset:
    """dgl.data.LegacyTUDataset('ENZYMES') (main_dgl_enzymes_gcn.py:11,155): 600 small graphs of ~32.6 nodes /
    ~62.1 undirected edges with 18 float node features and 6 classes (README.md:29); dataset[i] -> (graph, label)."""

    synthetic = True
    _SHAPES = {"ENZYMES": (600, 32.63, 18, 6)}

    def __init__(self, name, **kw):
        n, mean_nodes, feat, classes = self._SHAPES[name]
        self.name, self._feat, self.num_labels = name, feat, classes
        self._mean_nodes = mean_nodes
        self._n = max(16, int(n * _scale()))
        synthetic.announce_synthetic(name, "%d molecule-like random graphs of ~%.1f nodes, %d features, %d classes"
                                     % (self._n, mean_nodes, feat, classes))

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        from ..heterograph import graph as make_graph
        i = int(i)
        # ENZYMES has ~1.9 undirected edges per node: a tree plus ~0.9 extra edges per node
        src, dst, sizes = synthetic.molecule_like_batch(1, seed=10_000 + i, mean_nodes=self._mean_nodes, extra_edge_frac=0.9)
        g = make_graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=int(sizes[0]))
        rng = np.random.default_rng(i)
        g.ndata["feat"] = torch.from_numpy(rng.random((int(sizes[0]), self._feat), dtype=np.float32))
        return g, torch.tensor(int(rng.integers(0, self.num_labels)))


def CitationGraphDataset(name, **kw):
    return SyntheticNodeDataset(name, as_networkx=True, **kw)


def CoraGraphDataset(**kw):
    return SyntheticNodeDataset("cora", **kw)


def CiteseerGraphDataset(**kw):
    return SyntheticNodeDataset("citeseer", **kw)


def PubmedGraphDataset(**kw):
    return SyntheticNodeDataset("pubmed", **kw)


def RedditDataset(self_loop=False, **kw):
    return SyntheticNodeDataset("reddit", undirected=False, self_loop=self_loop)


def load_data(args):
    """dgl.data.load_data(args): dispatch on args.dataset like upstream dgl/data/utils.py."""
    name = args.dataset
    if name in ("cora", "citeseer", "pubmed"):
        return CitationGraphDataset(name)
    if name is not None and name.startswith("reddit"):
        return RedditDataset(self_loop=("self-loop" in name))
    raise ValueError("Unknown dataset: {}".format(name))
