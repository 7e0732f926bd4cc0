"""dgl.data.utils -- the one helper the in-scope scripts use (main_dgl_enzymes_gcn.py:161-163)."""


class Subset:
    """Subset of a dataset at the given indices (upstream dgl/data/utils.py::Subset)."""

    def __init__(self, dataset, indices):
        self.dataset, self.indices = dataset, indices

    def __getitem__(self, item):
        return self.dataset[self.indices[item]]

    def __len__(self):
        return len(self.indices)
