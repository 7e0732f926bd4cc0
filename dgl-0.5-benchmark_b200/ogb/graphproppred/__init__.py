import numpy as np
import torch

import dgl
from dgl.data import synthetic

_SHAPES = {"ogbg-molhiv": (41127, 25.5, 1), "ogbg-ppa": (158100, 243.4, 37)}


class _Subset:
    def __init__(self, ds, idx):
        self.ds, self.idx = ds, idx

    def __len__(self):
        return len(self.idx)

    def __getitem__(self, i):
        return self.ds[int(self.idx[i])]


class DglGraphPropPredDataset:
    """dataset[i] -> (graph with ndata['feat'] (n,9) / edata['feat'] (e,3) int64, label[1]);
    dataset[index_tensor] -> subset; molecule-like graphs (tree + a few ring closures)."""

    synthetic = True

    def __init__(self, name, root="dataset", num_graphs=None):
        import os
        self.name = name
        n, self._mean_nodes, self.num_tasks = _SHAPES[name]
        scale = float(os.environ.get("DGLB200_DATA_SCALE", "1"))
        self._n = int(num_graphs or max(16, int(n * scale)))
        synthetic.announce_synthetic(name, "%d molecule-like random graphs of ~%.1f nodes" % (self._n, self._mean_nodes))
        self.eval_metric = "rocauc" if name == "ogbg-molhiv" else "acc"
        self.num_classes = 2 if name == "ogbg-molhiv" else 37

    def __len__(self):
        return self._n

    def get_idx_split(self):
        perm = np.random.default_rng(0).permutation(self._n)
        a, b = int(0.8 * self._n), int(0.9 * self._n)
        return {"train": torch.from_numpy(perm[:a]), "valid": torch.from_numpy(perm[a:b]), "test": torch.from_numpy(perm[b:])}

    def __getitem__(self, i):
        if torch.is_tensor(i) and i.dim() > 0:
            return _Subset(self, i.numpy())
        i = int(i)
        src, dst, sizes = synthetic.molecule_like_batch(1, seed=i, mean_nodes=self._mean_nodes)
        g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=int(sizes[0]))
        rng = np.random.default_rng(i)
        g.ndata["feat"] = torch.from_numpy(rng.integers(0, 2, size=(int(sizes[0]), 9)).astype(np.int64))
        g.edata["feat"] = torch.from_numpy(rng.integers(0, 2, size=(len(src), 3)).astype(np.int64))
        label = torch.tensor([float(rng.integers(0, 2))]) if self.name == "ogbg-molhiv" else torch.tensor([int(rng.integers(0, 37))])
        return g, label


class Evaluator:
    def __init__(self, name):
        self.name = name

    def eval(self, input_dict):
        y_true = torch.as_tensor(input_dict["y_true"]).detach().cpu().numpy()
        y_pred = torch.as_tensor(input_dict["y_pred"]).detach().cpu().numpy()
        if self.name == "ogbg-molhiv":
            from sklearn.metrics import roc_auc_score
            if len(np.unique(y_true)) < 2:
                return {"rocauc": 0.5}
            return {"rocauc": float(roc_auc_score(y_true.reshape(-1), y_pred.reshape(-1)))}
        return {"acc": float((y_true.reshape(-1) == y_pred.reshape(-1)).mean())}
