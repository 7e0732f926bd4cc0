"""AtomEncoder / BondEncoder: sums of per-column embeddings of the categorical atom (9 columns)
and bond (3 columns) features (ogb.graphproppred.mol_encoder)."""
import torch

_ATOM_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
_BOND_DIMS = [5, 6, 2]


class _SumEmbedding(torch.nn.Module):
    def __init__(self, dims, emb_dim):
        super().__init__()
        self.embs = torch.nn.ModuleList(torch.nn.Embedding(d, emb_dim) for d in dims)
        for e in self.embs:
            torch.nn.init.xavier_uniform_(e.weight.data)

    def forward(self, x):
        out = 0
        for i, e in enumerate(self.embs):
            out = out + e(x[:, i])
        return out


class AtomEncoder(_SumEmbedding):
    def __init__(self, emb_dim):
        super().__init__(_ATOM_DIMS, emb_dim)

    @property
    def atom_embedding_list(self):
        return self.embs


class BondEncoder(_SumEmbedding):
    def __init__(self, emb_dim):
        super().__init__(_BOND_DIMS, emb_dim)

    @property
    def bond_embedding_list(self):
        return self.embs
