import numpy as np
import torch

from dgl.data.datasets import SyntheticNodeDataset


class DglNodePropPredDataset:
    """dataset[0] -> (graph, labels[N,1]); .get_idx_split() -> {'train','valid','test'}; .num_classes"""

    synthetic = True

    def __init__(self, name, root="dataset"):
        self.name = name
        key = {"ogbn-arxiv": "ogbn-arxiv", "ogbn-products": "ogbn-products", "ogbn-proteins": "ogbn-proteins"}[name]
        # arxiv is stored directed (scripts call to_bidirected); products is stored symmetric
        self._ds = SyntheticNodeDataset(key, undirected=(name == "ogbn-products"))
        self.num_classes = self._ds.num_classes
        self.num_tasks = 112 if name == "ogbn-proteins" else 1
        self._g = None

    def get_idx_split(self):
        d = self._ds
        return {"train": torch.from_numpy(np.nonzero(d.train_mask)[0]), "valid": torch.from_numpy(np.nonzero(d.val_mask)[0]),
                "test": torch.from_numpy(np.nonzero(d.test_mask)[0])}

    def __getitem__(self, idx):
        assert idx == 0
        if self._g is None:
            g = self._ds[0]
            if self.name == "ogbn-proteins":
                rng = np.random.default_rng(1)
                g.edata["feat"] = torch.from_numpy(rng.random((g.number_of_edges(), 8), dtype=np.float32))
                g.ndata["species"] = torch.zeros(g.number_of_nodes(), 1, dtype=torch.int64)
                labels = torch.from_numpy(rng.integers(0, 2, size=(g.number_of_nodes(), 112)).astype(np.int64))
            else:
                labels = g.ndata["label"].view(-1, 1)
            self._g = (g, labels)
        return self._g

    def __len__(self):
        return 1


class Evaluator:
    def __init__(self, name):
        self.name = name

    def eval(self, input_dict):
        y_true, y_pred = input_dict["y_true"], input_dict["y_pred"]
        y_true = y_true.detach().cpu() if torch.is_tensor(y_true) else torch.as_tensor(y_true)
        y_pred = y_pred.detach().cpu() if torch.is_tensor(y_pred) else torch.as_tensor(y_pred)
        if self.name == "ogbn-proteins":
            from sklearn.metrics import roc_auc_score
            aucs = [roc_auc_score(y_true[:, i], y_pred[:, i]) for i in range(y_true.shape[1])
                    if len(np.unique(y_true[:, i])) == 2]
            return {"rocauc": float(np.mean(aucs)) if aucs else 0.5}
        return {"acc": float((y_true == y_pred).float().mean())}
