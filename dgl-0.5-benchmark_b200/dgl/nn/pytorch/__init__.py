"""dgl.nn.pytorch: GATConv (main_dgl_*_gat.py), SAGEConv (main_dgl_*_sage_nn.py), GraphConv
(main_dgl_enzymes_gcn_nn.py:29), AvgPooling (main_dgl_molhiv_gcn.py:75)."""
from .conv import GATConv, SAGEConv, GraphConv  # noqa: F401
from .glob import AvgPooling, SumPooling, MaxPooling  # noqa: F401
from ...ops import edge_softmax  # noqa: F401
