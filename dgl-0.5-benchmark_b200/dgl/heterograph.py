"""DGLGraph -- the graph object the reference scripts hold (API surface: SURVEY.md Appendix B).

Mirrors the part of upstream DGL v0.6.1 python/dgl/heterograph.py::DGLHeteroGraph that the in-scope
scripts touch, for graphs with one node type and one edge type (and bipartite "block" graphs with
distinct source / destination node sets): construction from (src, dst), `.int()/.to()/.formats()`,
feature frames (`ndata/srcdata/dstdata/edata`), `update_all`, `apply_edges`, degrees, local scopes.
Message passing is translated into gspmm / gsddmm calls exactly like upstream core.py does.
"""
from collections.abc import MutableMapping
from contextlib import contextmanager

import torch

from . import core
from ._capi import DGLError
from .graph_index import GraphIndex

ALL = None


class Frame(MutableMapping):
    """name -> tensor with a fixed number of rows."""

    def __init__(self, num_rows, data=None):
        self._n = num_rows
        self._d = dict(data) if data else {}

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        if not torch.is_tensor(v):
            raise DGLError("feature data must be a tensor")
        if v.shape[0] != self._n:
            raise DGLError("Expect number of features to match number of nodes/edges (len(u)). "
                           "Got %d and %d instead." % (v.shape[0], self._n))
        self._d[k] = v

    def __delitem__(self, k):
        del self._d[k]

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __repr__(self):
        return repr({k: tuple(v.shape) for k, v in self._d.items()})

    def clone(self):
        return Frame(self._n, self._d)

    def to(self, device):
        return Frame(self._n, {k: v.to(device) for k, v in self._d.items()})


class DGLHeteroGraph:
    def __init__(self, gidx, src_frame=None, dst_frame=None, edge_frame=None, is_block=False):
        self._graph = gidx
        self._is_block = is_block or gidx.n_src != gidx.n_dst
        self._src_frame = src_frame if src_frame is not None else Frame(gidx.n_src)
        if self._is_block:
            self._dst_frame = dst_frame if dst_frame is not None else Frame(gidx.n_dst)
        else:
            self._dst_frame = self._src_frame
        self._edge_frame = edge_frame if edge_frame is not None else Frame(gidx.n_edges)
        self._batch_num_nodes = None
        self._batch_num_edges = None
        self._batch_max_nodes = None   # largest member graph, known on the host (readout: no hub-detection sync)
        self._readout_index = None     # nodes -> member-graph relation used by the readout layers (dgl/nn/pytorch/glob.py)

    # ------------------------------------------------------------------ structure queries
    @property
    def is_block(self):
        return self._is_block

    @property
    def idtype(self):
        return self._graph.idtype

    @property
    def device(self):
        return self._graph.device

    @property
    def ntypes(self):
        return ["_N"]

    @property
    def etypes(self):
        return ["_E"]

    @property
    def canonical_etypes(self):
        return [("_N", "_E", "_N")]

    def number_of_nodes(self, ntype=None):
        if self._is_block:
            return self._graph.n_src + self._graph.n_dst
        return self._graph.n_src

    num_nodes = number_of_nodes

    def number_of_src_nodes(self, ntype=None):
        return self._graph.n_src

    num_src_nodes = number_of_src_nodes

    def number_of_dst_nodes(self, ntype=None):
        return self._graph.n_dst

    num_dst_nodes = number_of_dst_nodes

    def number_of_edges(self, etype=None):
        return self._graph.n_edges

    num_edges = number_of_edges

    def nodes(self):
        return torch.arange(self.number_of_nodes(), dtype=self.idtype, device=self.device)

    def srcnodes(self):
        return torch.arange(self._graph.n_src, dtype=self.idtype, device=self.device)

    def dstnodes(self):
        return torch.arange(self._graph.n_dst, dtype=self.idtype, device=self.device)

    def edges(self, form="uv", order="eid"):
        src, dst = self._graph.src, self._graph.dst
        if form == "uv":
            return src, dst
        eid = torch.arange(self._graph.n_edges, dtype=self.idtype, device=self.device)
        if form == "eid":
            return eid
        if form == "all":
            return src, dst, eid
        raise DGLError('Invalid form: {}. Must be "all", "uv" or "eid".'.format(form))

    all_edges = edges

    def in_degrees(self, v=ALL):
        deg = self._graph.in_degrees().to(self.idtype)
        return deg if v is None else deg[torch.as_tensor(v, device=deg.device).long()]

    def out_degrees(self, u=ALL):
        deg = self._graph.out_degrees().to(self.idtype)
        return deg if u is None else deg[torch.as_tensor(u, device=deg.device).long()]

    # ------------------------------------------------------------------ conversion
    def _with_index(self, gidx, frames=None):
        s, d, e = frames if frames is not None else (self._src_frame, self._dst_frame, self._edge_frame)
        g = DGLHeteroGraph(gidx, s, d if self._is_block else None, e, is_block=self._is_block)
        g._batch_num_nodes, g._batch_num_edges = self._batch_num_nodes, self._batch_num_edges
        g._batch_max_nodes = self._batch_max_nodes
        if gidx is self._graph:
            g._readout_index = self._readout_index
        return g

    def int(self):
        """Cast node / edge ids to int32 (kernel/dgl-new.py:63)."""
        return self._with_index(self._graph.astype(torch.int32))

    def long(self):
        return self._with_index(self._graph.astype(torch.int64))

    def to(self, device, **kwargs):
        device = torch.device(device)
        if device == self.device:
            return self
        frames = (self._src_frame.to(device), self._dst_frame.to(device) if self._is_block else None,
                  self._edge_frame.to(device))
        g = self._with_index(self._graph.to(device), frames)
        if g._batch_num_nodes is not None:
            g._batch_num_nodes = g._batch_num_nodes.to(device)
            g._batch_num_edges = g._batch_num_edges.to(device)
        return g

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def formats(self, formats=None):
        """Query or restrict the allowed sparse formats (main_dgl_product_sage.py:158,
        main_dgl_molhiv_gcn.py:101)."""
        if formats is None:
            allowed = self._graph.formats()
            return {"created": sorted(allowed), "not created": []}
        if isinstance(formats, str):
            formats = [formats]
        bad = set(formats) - {"coo", "csr", "csc"}
        if bad or not formats:
            raise DGLError("formats must be a non-empty subset of coo/csr/csc, got %s" % (formats,))
        return self._with_index(self._graph.restrict_formats(formats))

    def reverse(self, copy_ndata=True, copy_edata=False):
        g = DGLHeteroGraph(self._graph.reverse(), is_block=self._is_block)
        if copy_ndata and not self._is_block:
            g._src_frame = g._dst_frame = self._src_frame.clone()
        if copy_edata:
            g._edge_frame = self._edge_frame.clone()
        return g

    def local_var(self):
        """A graph sharing the structure whose feature frames are private shallow copies
        (main_dgl_citation_sage.py:63)."""
        s = self._src_frame.clone()
        d = self._dst_frame.clone() if self._is_block else None
        return self._with_index(self._graph, (s, d, self._edge_frame.clone()))

    @contextmanager
    def local_scope(self):
        """Feature writes inside the scope are discarded on exit (main_dgl_proteins_rgcn_for.py:47)."""
        saved = (self._src_frame, self._dst_frame, self._edge_frame)
        self._src_frame = saved[0].clone()
        self._dst_frame = saved[1].clone() if self._is_block else self._src_frame
        self._edge_frame = saved[2].clone()
        try:
            yield
        finally:
            self._src_frame, self._dst_frame, self._edge_frame = saved

    # ------------------------------------------------------------------ feature frames
    @property
    def ndata(self):
        if self._is_block:
            raise DGLError("ndata is ambiguous on a block; use srcdata / dstdata")
        return self._src_frame

    @property
    def srcdata(self):
        return self._src_frame

    @property
    def dstdata(self):
        return self._dst_frame

    @property
    def edata(self):
        return self._edge_frame

    # ------------------------------------------------------------------ message passing
    def update_all(self, message_func, reduce_func, apply_node_func=None, etype=None):
        """Send messages along all edges and reduce them at the destinations
        (main_dgl_citation_sage.py:75-77, main_dgl_molhiv_gcn.py:46)."""
        out = core.message_passing(self, message_func, reduce_func, apply_node_func)
        for k, v in out.items():
            self._dst_frame[k] = v

    def apply_edges(self, func, edges=ALL, etype=None):
        """Compute an edge feature from the end points (GATConv: fn.u_add_v; gcmc: fn.u_dot_v)."""
        if edges is not None:
            raise DGLError("apply_edges on an edge subset is not supported")
        if core.is_builtin(func):
            out = core.invoke_gsddmm(self, func)
        else:
            out = core.invoke_edge_udf(self, func)
        for k, v in out.items():
            self._edge_frame[k] = v

    # ------------------------------------------------------------------ batching info
    @property
    def batch_size(self):
        return 1 if self._batch_num_nodes is None else int(self._batch_num_nodes.shape[0])

    def batch_num_nodes(self, ntype=None):
        if self._batch_num_nodes is None:
            return torch.tensor([self.number_of_nodes()], dtype=torch.int64, device=self.device)
        return self._batch_num_nodes

    def batch_num_edges(self, etype=None):
        if self._batch_num_edges is None:
            return torch.tensor([self.number_of_edges()], dtype=torch.int64, device=self.device)
        return self._batch_num_edges

    def __repr__(self):
        if self._is_block:
            return ("Block(num_src_nodes={}, num_dst_nodes={}, num_edges={})"
                    .format(self._graph.n_src, self._graph.n_dst, self._graph.n_edges))
        return ("Graph(num_nodes={}, num_edges={},\n      ndata_schemes={}\n      edata_schemes={})"
                .format(self._graph.n_src, self._graph.n_edges, self._src_frame, self._edge_frame))


DGLGraph = DGLHeteroGraph


def graph(data, num_nodes=None, idtype=None, device=None, **kwargs):
    """dgl.graph((src, dst)) -- edges keep their creation order as edge ids (kernel/utils.py:39)."""
    if isinstance(data, DGLHeteroGraph):
        data = data.edges()
    src, dst = data
    src = torch.as_tensor(src)
    dst = torch.as_tensor(dst)
    if src.shape != dst.shape or src.dim() != 1:
        raise DGLError("src and dst must be 1-D tensors of equal length")
    if idtype is None:
        idtype = torch.int32 if src.dtype == torch.int32 else torch.int64
    src, dst = src.to(idtype), dst.to(idtype)
    if device is not None:
        src, dst = src.to(device), dst.to(device)
    if num_nodes is None:
        num_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
    elif src.numel() and int(max(src.max().item(), dst.max().item())) >= num_nodes:
        raise DGLError("The num_nodes argument must be larger than the max ID in the data")
    return DGLHeteroGraph(GraphIndex(src.contiguous(), dst.contiguous(), num_nodes, num_nodes, idtype))


def create_block(data, num_src_nodes, num_dst_nodes, idtype=None, device=None):
    """Bipartite message-flow block with distinct source / destination node sets."""
    src, dst = data
    src, dst = torch.as_tensor(src), torch.as_tensor(dst)
    if idtype is None:
        idtype = torch.int32 if src.dtype == torch.int32 else torch.int64
    src, dst = src.to(idtype), dst.to(idtype)
    if device is not None:
        src, dst = src.to(device), dst.to(device)
    return DGLHeteroGraph(GraphIndex(src.contiguous(), dst.contiguous(), num_src_nodes, num_dst_nodes, idtype),
                          is_block=True)
