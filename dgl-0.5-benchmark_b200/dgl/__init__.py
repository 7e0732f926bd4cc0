"""dgl -- a drop-in stand-in for the slice of DGL v0.6.1 that dglai/dgl-0.5-benchmark exercises,
backed by hand-written sm_100a kernels (lib/libdglb200.so, C-ABI in include/dglb200.h).

Put `dgl-0.5-benchmark_b200/` on sys.path and the reference scripts (`kernel/dgl-new.py`,
`end_to_end/full_graph/**/main_dgl_*.py`) import this package unchanged.  Only the sparse
message-passing hot path is native; there is no CPU fallback for it.
"""
from ._capi import DGLError  # noqa: F401
from .heterograph import DGLGraph, DGLHeteroGraph, graph, create_block  # noqa: F401
from . import function  # noqa: F401
from . import ops  # noqa: F401
from . import utils  # noqa: F401
from .transform import add_self_loop, remove_self_loop, to_bidirected, add_reverse_edges, reverse  # noqa: F401
from .convert import from_networkx, from_scipy, to_networkx  # noqa: F401
from .batch import batch  # noqa: F401
from .batch_store import GraphStore, StaticBatch  # noqa: F401
try:
    from . import nn  # noqa: F401
    from . import data  # noqa: F401
    from . import dataloading  # noqa: F401
except ImportError:  # pragma: no cover  (modules land later in the build)
    pass

__version__ = "0.6.1+b200"
