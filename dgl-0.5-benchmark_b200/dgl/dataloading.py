"""dgl.dataloading.GraphDataLoader (main_dgl_molhiv_gcn.py:12,163): a torch DataLoader whose
collate function batches (graph, label) samples with dgl.batch."""
import torch
from torch.utils.data import DataLoader

from .batch import batch as _batch
from .heterograph import DGLHeteroGraph


def _collate(samples):
    first = samples[0]
    if isinstance(first, DGLHeteroGraph):
        return _batch(samples)
    if isinstance(first, (tuple, list)):
        cols = list(zip(*samples))
        return [_collate(list(c)) for c in cols]
    if torch.is_tensor(first):
        return torch.stack(samples, 0)
    return torch.utils.data.dataloader.default_collate(samples)


class GraphDataLoader(DataLoader):
    def __init__(self, dataset, collate_fn=None, **kwargs):
        super().__init__(dataset, collate_fn=collate_fn or _collate, **kwargs)
