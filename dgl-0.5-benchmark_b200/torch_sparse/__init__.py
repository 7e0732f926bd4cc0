"""torch_sparse -- import stub.  kernel/utils.py:6 imports SparseTensor at module scope for the PyG
side of the micro-benchmark (kernel/pyg.py, out of scope); the DGL side never touches it."""


class SparseTensor:
    def __init__(self, *args, **kwargs):
        raise ImportError("torch_sparse is not available; only the DGL side of the benchmark is supported")
