"""Device-side batching of small graphs (BASELINE config 5).

The reference's graph-classification loop (main_dgl_molhiv_gcn.py:95-115 behind GraphDataLoader, :163) builds every
batch on the host (`dgl.batch`: Python concatenation of ~64-256 graphs), copies it to the device and converts COO to
CSC / CSR there -- per iteration.  Here the dataset is uploaded ONCE as a union graph (`GraphStore`); a batch is a list
of member-graph ids, and `StaticBatch.refresh()` materialises COO / CSC / CSR, the node -> member-graph relation of
the readout, features and labels for it with two kernel launches (csrc/small_graph.cu: batch_offsets, batch_gather)
plus one row gather per feature tensor, into buffers of FIXED padded size.  Nothing in refresh() allocates
per-batch-sized memory, synchronises or depends on host-side values, so an entire training step -- batch construction,
forward, backward, optimizer -- can be captured in ONE CUDA graph and replayed with only `graph_ids` changing.

Bit-exactness: the CSC / CSR of the batch equal what `dgl.batch(...).to(dev).int()` + the stable COO -> CSR build of
dglb_coo_to_csr produce (tests/test_gpu_small_graph.py), because member graphs occupy disjoint increasing node ranges.

Padding: nodes beyond the batch's real count are isolated and belong to an extra member graph (row `batch_size` of every
readout; ignore it), edge slots beyond the real count lie outside every indptr range.  `node_mask` / `n_real_nodes` let
a model keep statistics that run over nodes (BatchNorm) exact.
"""
import numpy as np
import torch

from . import _capi
from ._capi import DGLError
from .batch import batch as host_batch
from .graph_index import CSRView, GraphIndex
from .heterograph import DGLHeteroGraph


class GraphStore:
    """All member graphs of a dataset as one device-resident union graph with its CSC / CSR (built once)."""

    def __init__(self, graphs, labels=None, device="cuda"):
        if len(graphs) == 0:
            raise DGLError("GraphStore needs at least one graph")
        dev = torch.device(device)
        self.device = dev
        self.n_nodes_host = np.array([g.number_of_nodes() for g in graphs], dtype=np.int64)
        self.n_edges_host = np.array([g.number_of_edges() for g in graphs], dtype=np.int64)
        if self.n_nodes_host.sum() >= 2 ** 31 or self.n_edges_host.sum() >= 2 ** 31:
            raise DGLError("GraphStore: the union graph must fit int32 ids")
        union = host_batch(list(graphs))                       # host-side, once per dataset
        ne = union.number_of_edges()
        self.max_in_deg = int(torch.bincount(union._graph.dst.long()).max()) if ne else 0     # host-side, once: lets every
        self.max_out_deg = int(torch.bincount(union._graph.src.long()).max()) if ne else 0    # batch skip hub detection
        ug = union.to(dev).int()
        gi = ug._graph
        self.u_src, self.u_dst = gi.coo32()
        self.csc, self.csr = gi.csc(), gi.csr()
        zero = np.zeros(1, dtype=np.int64)
        self.node_ptr = torch.from_numpy(np.concatenate([zero, self.n_nodes_host.cumsum()]).astype(np.int32)).to(dev)
        self.edge_ptr = torch.from_numpy(np.concatenate([zero, self.n_edges_host.cumsum()]).astype(np.int32)).to(dev)
        self.ndata = {k: v.contiguous() for k, v in ug.ndata.items()}
        self.edata = {k: v.contiguous() for k, v in ug.edata.items()}
        self.labels = None if labels is None else torch.as_tensor(labels).to(dev).contiguous()
        self.num_graphs = len(graphs)

    def _store_args(self):
        return [self.node_ptr, self.edge_ptr, self.u_src, self.u_dst, self.csc.indptr, self.csc.indices, self.csc.eids,
                self.csr.indptr, self.csr.indices, self.csr.eids]

    def pad_sizes(self, batches, multiple=64):
        """(n_nodes_pad, n_edges_pad) that fit every batch of `batches` (an iterable of id arrays), rounded up; at least
        one padding node is always present (it anchors the COO entries of padding edge slots)."""
        n = max(int(self.n_nodes_host[np.asarray(b)].sum()) for b in batches)
        e = max(int(self.n_edges_host[np.asarray(b)].sum()) for b in batches)
        up = lambda x: -(-(x + 1) // multiple) * multiple   # noqa: E731
        return up(n), up(e)

    def fits(self, graph_ids, n_nodes_pad, n_edges_pad):
        ids = np.asarray(graph_ids)
        return (int(self.n_nodes_host[ids].sum()) < n_nodes_pad and int(self.n_edges_host[ids].sum()) <= n_edges_pad)

    def static_batch(self, batch_size, n_nodes_pad, n_edges_pad):
        return StaticBatch(self, batch_size, n_nodes_pad, n_edges_pad)


class StaticBatch:
    """Fixed-size device buffers holding one batch of `batch_size` member graphs, and the DGLGraph over them."""

    def __init__(self, store, batch_size, n_nodes_pad, n_edges_pad):
        self.store, self.batch_size = store, int(batch_size)
        self.n_nodes_pad, self.n_edges_pad = int(n_nodes_pad), int(n_edges_pad)
        dev, B, N, E = store.device, self.batch_size, self.n_nodes_pad, self.n_edges_pad
        i32 = dict(dtype=torch.int32, device=dev)
        self.graph_ids = torch.zeros(B, **i32)
        self.out_node_ptr, self.out_edge_ptr = torch.zeros(B + 2, **i32), torch.zeros(B + 2, **i32)
        self.status = torch.zeros(1, **i32)
        self.src, self.dst = torch.zeros(E, **i32), torch.zeros(E, **i32)
        self.csc_indptr, self.csr_indptr = torch.zeros(N + 1, **i32), torch.zeros(N + 1, **i32)
        self.csc_indices, self.csc_eids = torch.zeros(E, **i32), torch.zeros(E, **i32)
        self.csr_indices, self.csr_eids = torch.zeros(E, **i32), torch.zeros(E, **i32)
        self.node_graph, self.node_map, self.edge_map = torch.zeros(N, **i32), torch.zeros(N, **i32), torch.zeros(E, **i32)
        self._batch_args = [self.out_node_ptr, self.out_edge_ptr, self.status, self.src, self.dst, self.csc_indptr,
                            self.csc_indices, self.csc_eids, self.csr_indptr, self.csr_indices, self.csr_eids,
                            self.node_graph, self.node_map, self.edge_map]
        # node-level bookkeeping for models whose statistics run over the nodes
        self.node_mask = torch.zeros(N, 1, dtype=torch.float32, device=dev)      # 1 for real nodes
        self.n_real_nodes = torch.zeros((), dtype=torch.float32, device=dev)
        self._node_iota = torch.arange(N, **i32)
        self.labels = None if store.labels is None else torch.zeros((B,) + tuple(store.labels.shape[1:]),
                                                                     dtype=store.labels.dtype, device=dev)
        # the batched graph over the static buffers
        gi = GraphIndex(self.src, self.dst, N, N, torch.int32)
        self._csc = CSRView(N, N, self.csc_indptr, self.csc_indices, self.csc_eids)
        self._csr = CSRView(N, N, self.csr_indptr, self.csr_indices, self.csr_eids)
        self._csc.max_deg, self._csr.max_deg = store.max_in_deg, store.max_out_deg   # no hub detection, no sync
        gi._c["csc"], gi._c["csr"], gi._c["coo32"] = self._csc, self._csr, (self.src, self.dst)
        gi._c["dst_sorted"] = gi._c["src_sorted"] = False
        gi._c["max_in_deg"], gi._c["max_out_deg"] = store.max_in_deg, store.max_out_deg
        gi._c["padded_edge_slots"] = True   # per-edge gradients: slots outside every CSR row are zero-filled
        self.graph = DGLHeteroGraph(gi)
        for k, v in store.ndata.items():
            self.graph.ndata[k] = torch.zeros((N,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
        for k, v in store.edata.items():
            self.graph.edata[k] = torch.zeros((E,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
        # readout relation "node -> member graph" (dgl/nn/pytorch/glob.py): B real member graphs + the padding graph
        self.counts = torch.ones(B + 1, dtype=torch.int64, device=dev)
        self.counts_f = torch.ones(B + 1, dtype=torch.float32, device=dev)
        mi = GraphIndex(self._node_iota, self.node_graph, N, B + 1, torch.int32)
        m_csc = CSRView(B + 1, N, self.out_node_ptr, self._node_iota, None)
        m_csr = CSRView(N, B + 1, torch.arange(N + 1, **i32), self.node_graph, None)
        # member graphs are at most a few hundred nodes: their rows stay on the row kernels (no hub split, whose
        # detection would need a device sync per batch)
        m_csc.max_deg, m_csr.max_deg = 1, 1
        mi._c["csc"], mi._c["csr"], mi._c["coo32"] = m_csc, m_csr, (self._node_iota, self.node_graph)
        mi._c["dst_sorted"] = mi._c["src_sorted"] = True
        self._m_csc = m_csc
        self.graph._readout_index = (mi, self.counts_f)
        self.graph._batch_num_nodes = self.counts
        self.graph._batch_max_nodes = 1

    def set_ids(self, graph_ids):
        """Copy the member-graph ids of the next batch into the static id buffer (host or device tensor / array)."""
        ids = torch.as_tensor(graph_ids, dtype=torch.int32)
        if ids.numel() != self.batch_size:
            raise DGLError("StaticBatch holds exactly %d graphs, got %d ids" % (self.batch_size, ids.numel()))
        self.graph_ids.copy_(ids, non_blocking=True)

    def refresh(self):
        """Rebuild every buffer from `graph_ids`.  Asynchronous, allocation-free in the batch's size, capturable."""
        st = self.store
        _capi.call(_capi.ops().batch_build, self.graph_ids, st._store_args(), self._batch_args, self.n_nodes_pad,
                   self.n_edges_pad)
        _capi.count_launch(2)
        for view in (self._csc, self._csr, self._m_csc):   # cached degrees / divisors belong to the previous batch
            view._deg = view._deg_f = None
        for k, v in st.ndata.items():
            torch.index_select(v, 0, self.node_map, out=self.graph.ndata[k])
        for k, v in st.edata.items():
            torch.index_select(v, 0, self.edge_map, out=self.graph.edata[k])
        if self.labels is not None:
            torch.index_select(st.labels, 0, self.graph_ids, out=self.labels)
        B = self.batch_size
        torch.sub(self.out_node_ptr[1:B + 2], self.out_node_ptr[:B + 1], out=self.counts)
        self.counts_f.copy_(self.counts.clamp(min=1))
        n_real = self.out_node_ptr[B]
        self.n_real_nodes.copy_(n_real)
        self.node_mask.copy_((self._node_iota < n_real).view(-1, 1))
        return self

    def load(self, graph_ids):
        self.set_ids(graph_ids)
        return self.refresh()

    def overflowed(self):
        """True when the last refreshed batch did not fit the padded sizes (device sync: use it for checks, not in the
        training loop -- GraphStore.fits answers the same from host-side counts)."""
        return bool(self.status.item())
