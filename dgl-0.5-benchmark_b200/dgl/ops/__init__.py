"""dgl.ops -- the operator interface the reference's micro-benchmark calls
(kernel/dgl-new.py:2,20,39): gspmm, gsddmm, edge_softmax and their generated shorthands."""
from .spmm import *  # noqa: F401,F403
from .sddmm import *  # noqa: F401,F403
from .edge_softmax import *  # noqa: F401,F403
from .gat import gat_attention  # noqa: F401
from .gcn import gcn_norm_relu_sum, categorical_embedding_sum  # noqa: F401
