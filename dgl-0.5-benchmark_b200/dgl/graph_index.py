"""Device-resident graph structure: COO + lazily materialised CSC / CSR with the edge-id permutation.

Plays the role of upstream DGL v0.6.1 src/graph/unit_graph.cc (UnitGraph: COO/CSR/CSC holder with
lazy, cached conversion; `Reverse` swaps CSR and CSC zero-copy) and python/dgl/heterograph_index.py
for the single-node-type / single-edge-type graphs every in-scope script builds
(kernel/utils.py:37-40 `dgl.graph(g.edges())`, main_dgl_citation_sage.py:190-191).

Sparse formats follow SURVEY.md Appendix A.1: CSC = STABLE sort of the edges by destination
(`data` = edge ids, increasing inside a row); CSR = the same by source.  When the COO is already
sorted the permutation is the identity and is dropped (eids=None), which removes a 4-byte
gather per edge from every kernel.
"""
import torch

from . import _capi


class CSRView:
    """indptr / indices / eids (None == identity) over `n_rows` rows, all int32 device tensors,
    plus per-threshold hub-row lists (caller-owned metadata of the C-ABI)."""

    __slots__ = ("n_rows", "n_cols", "indptr", "indices", "eids", "_hubs", "_deg", "_deg_f", "_stage", "max_deg", "_order")

    LOG2_STAGE_BUCKET = 15  # staged edge order: slots are shuffled inside windows of 32 K CSR positions

    def __init__(self, n_rows, n_cols, indptr, indices, eids):
        self.n_rows, self.n_cols = n_rows, n_cols
        self.indptr, self.indices, self.eids = indptr, indices, eids
        self._hubs = {}
        self._deg = None
        self._deg_f = None
        self._stage = None
        self.max_deg = None   # largest row length when it is known from the host side (graphs built from CPU tensors)
        self._order = None

    def row_order(self):
        """Row ids by non-increasing nnz (stable), int32: the degree-ordered hand-out of rows to the row-per-group
        kernels (dglb_hub_t.row_order).  Built once per view, on the device."""
        if self._order is None:
            self._order = torch.argsort(self.degrees(), descending=True, stable=True).to(torch.int32)
        return self._order

    def stage_plan(self):
        """(stage_pos by edge id, slot by CSR position) of the staged edge order (include/dglb200.h,
        dglb_edge_stage_plan), or None when the edge-id permutation is the identity.  Built once, cached."""
        if self.eids is None:
            return None
        if self._stage is None:
            self._stage = tuple(_capi.call(_capi.ops().edge_stage_plan, self.eids, self.LOG2_STAGE_BUCKET))
        return self._stage

    @property
    def nnz(self):
        return self.indices.shape[0]

    def degrees(self):
        if self._deg is None:
            self._deg = _capi.call(_capi.ops().csr_degrees, self.indptr)
        return self._deg

    def mean_divisor(self):
        """float(clamp(deg, 1)) -- the divisor of reducer 'mean' (upstream ops/spmm.py)."""
        if self._deg_f is None:
            self._deg_f = self.degrees().clamp(min=1).to(torch.float32)
        return self._deg_f

    def hubs(self, threshold):
        """HubInfo for rows with nnz > threshold (None when there are none); cached per threshold.
        Hub rows are cut into segments of <= threshold entries (one CTA each in gspmm / gsddmm)."""
        if threshold in self._hubs:
            return self._hubs[threshold]
        info = None
        if self.max_deg is not None and self.max_deg <= threshold:
            pass                  # known from the host side: no hub rows, no kernel, no device sync (batched molecules)
        elif self.nnz > threshold:  # otherwise no row can exceed it: no kernel, no sync (small batched graphs)
            dev = self.indptr.device
            cap = max(1, min(self.n_rows, self.nnz // max(threshold, 1) + 1))
            rows, n_hub_t = _capi.call(_capi.ops().find_hub_rows, self.indptr, int(threshold), cap)
            n_hub = int(n_hub_t.item())  # one-off sync per (graph, threshold)
            assert n_hub <= cap
            if n_hub:
                rows = torch.sort(rows[:n_hub]).values.contiguous()       # deterministic order
                deg = self.degrees()[rows.long()].cpu().numpy().astype("int64")
                seg_len = int(threshold)
                nseg = -(-deg // seg_len)
                seg_ptr = torch.zeros(n_hub + 1, dtype=torch.int32)
                seg_ptr[1:] = torch.from_numpy(nseg.cumsum()).to(torch.int32)
                seg_hub = torch.repeat_interleave(torch.arange(n_hub, dtype=torch.int32), torch.from_numpy(nseg))
                # prefix sum of the non-hub rows' nnz: what the persistent ring kernels balance their warps on
                d = self.degrees()
                light = torch.zeros(self.n_rows + 1, dtype=torch.int32, device=dev)
                light[1:] = torch.cumsum(torch.where(d > threshold, torch.zeros_like(d), d), 0, dtype=torch.int64).to(torch.int32)
                info = HubInfo(rows, seg_ptr.to(dev), seg_hub.to(dev), n_hub, int(nseg.sum()), seg_len, int(threshold), light)
        self._hubs[threshold] = info
        return info


class HubInfo:
    """Caller-owned hub-row metadata of the C-ABI (dglb_hub_t): device arrays + counts."""

    __slots__ = ("rows", "seg_ptr", "seg_hub", "n_hub", "n_seg", "seg_len", "threshold", "light_indptr")

    def __init__(self, rows, seg_ptr, seg_hub, n_hub, n_seg, seg_len, threshold, light_indptr=None):
        self.rows, self.seg_ptr, self.seg_hub = rows, seg_ptr, seg_hub
        self.n_hub, self.n_seg, self.seg_len, self.threshold = n_hub, n_seg, seg_len, threshold
        self.light_indptr = light_indptr

    def pack(self):
        """The hub arguments of the extension's ops: (rows, seg_ptr, seg_hub, light_indptr, row_order slot, [n_hub, n_seg,
        seg_len, threshold]); workspaces are allocated by the op; the row order is filled in by the caller."""
        return (self.rows, self.seg_ptr, self.seg_hub, self.light_indptr, None,
                [self.n_hub, self.n_seg, self.seg_len, self.threshold])

    def struct(self, workspace=None):
        """ctypes dglb_hub_t (keep the returned object alive across the call)."""
        st = _capi.HubStruct()
        st.rows, st.seg_ptr, st.seg_hub = self.rows.data_ptr(), self.seg_ptr.data_ptr(), self.seg_hub.data_ptr()
        st.n_hub, st.n_seg, st.seg_len, st.threshold = self.n_hub, self.n_seg, self.seg_len, self.threshold
        st.workspace = workspace.data_ptr() if workspace is not None else None
        st.workspace_bytes = workspace.numel() * workspace.element_size() if workspace is not None else 0
        st.light_indptr = self.light_indptr.data_ptr() if self.light_indptr is not None else None
        return st


def build_csr(n_rows, n_cols, row, col, row_sorted=None, max_deg=None):
    """Stable sort of (row, col) by row on the device -> CSRView.  `row_sorted` (True/False/None):
    whether `row` is already non-decreasing, i.e. the edge-id permutation is the identity; when it is
    known from the host side (graphs created from CPU tensors) the device check and its sync are skipped."""
    _capi.require_cuda(row, col)
    o = _capi.ops()
    indptr, indices, data = _capi.call(o.coo_to_csr, row.contiguous(), col.contiguous(), int(n_rows))
    if row_sorted is None:
        identity = bool(_capi.call(o.is_identity_perm, data).item())
    else:
        identity = bool(row_sorted)
    view = CSRView(n_rows, n_cols, indptr, indices, None if identity else data)
    view.max_deg = max_deg
    return view


class GraphIndex:
    """One (src type, edge type, dst type) relation: n_src x n_dst, E edges in creation order."""

    def __init__(self, src, dst, n_src, n_dst, idtype=None, _shared=None):
        self.src, self.dst = src, dst  # edge-id order; int32 or int64 tensors (any device)
        self.n_src, self.n_dst = int(n_src), int(n_dst)
        self.idtype = idtype if idtype is not None else src.dtype
        # caches shared with the reversed view
        self._c = _shared if _shared is not None else {"csc": None, "csr": None, "coo32": None,
                                                         "formats": {"coo", "csr", "csc"},
                                                         "dst_sorted": None, "src_sorted": None,
                                                         "max_in_deg": None, "max_out_deg": None}
        self._rev = False
        if _shared is None and src.device.type == "cpu" and src.numel() <= (1 << 22):
            # cheap on the host, saves a device round trip per graph (matters for batched small graphs)
            n = src.numel()
            self._c["dst_sorted"] = bool(n < 2 or bool((dst[1:] >= dst[:-1]).all()))
            self._c["src_sorted"] = bool(n < 2 or bool((src[1:] >= src[:-1]).all()))
            # largest in / out degree: lets the kernels' hub-row detection (a device kernel + a sync per new graph)
            # be skipped for batched small graphs, whose rows are a handful of edges long
            self._c["max_in_deg"] = int(torch.bincount(dst.long(), minlength=1).max()) if n else 0
            self._c["max_out_deg"] = int(torch.bincount(src.long(), minlength=1).max()) if n else 0

    # ---- basic properties
    @property
    def device(self):
        return self.src.device

    @property
    def n_edges(self):
        return self.src.shape[0]

    def reverse(self):
        """Edge-direction reversal; CSC and CSR swap roles zero-copy (upstream UnitGraph::Reverse)."""
        g = GraphIndex(self.dst, self.src, self.n_dst, self.n_src, self.idtype, _shared=self._c)
        g._rev = not self._rev
        return g

    def astype(self, idtype):
        if idtype == self.idtype:
            return self
        g = GraphIndex(self.src.to(idtype), self.dst.to(idtype), self.n_src, self.n_dst, idtype)
        self._carry(g)
        return g

    def to(self, device):
        device = torch.device(device)
        if device == self.device:
            return self
        g = GraphIndex(self.src.to(device), self.dst.to(device), self.n_src, self.n_dst, self.idtype)
        self._carry(g)
        return g

    def restrict_formats(self, formats):
        g = GraphIndex(self.src, self.dst, self.n_src, self.n_dst, self.idtype)
        self._carry(g)
        g._c["formats"] = set(formats)
        return g

    def _carry(self, g):
        """Copy the structure-independent cache entries to a converted copy of this (un-reversed) view."""
        g._c["formats"] = set(self._c["formats"])
        a, b = self._c["dst_sorted"], self._c["src_sorted"]
        g._c["dst_sorted"], g._c["src_sorted"] = (a, b) if not self._rev else (b, a)
        a, b = self._c["max_in_deg"], self._c["max_out_deg"]
        g._c["max_in_deg"], g._c["max_out_deg"] = (a, b) if not self._rev else (b, a)

    def formats(self):
        return set(self._c["formats"])

    # ---- int32 COO for the kernels
    def coo32(self):
        key = "coo32"
        if self._c[key] is None:
            if self.n_src >= 2 ** 31 or self.n_dst >= 2 ** 31 or self.n_edges >= 2 ** 31:
                raise _capi.DGLError("graph too large for int32 ids")
            s = self.src if not self._rev else self.dst
            d = self.dst if not self._rev else self.src
            self._c[key] = (s.to(torch.int32).contiguous(), d.to(torch.int32).contiguous())
        s, d = self._c[key]
        return (s, d) if not self._rev else (d, s)

    # ---- sparse formats (lazy, cached, shared with the reverse view)
    def _fmt(self, which):
        # `which` in the frame of THIS view; translate to the cache's (un-reversed) frame
        key = which if not self._rev else ("csr" if which == "csc" else "csc")
        if self._c[key] is None:
            s, d = self._c["coo32"] if self._c["coo32"] is not None else (None, None)
            if s is None:
                self.coo32()
                s, d = self._c["coo32"]
            n_s = self.n_src if not self._rev else self.n_dst
            n_d = self.n_dst if not self._rev else self.n_src
            if key == "csc":
                self._c[key] = build_csr(n_d, n_s, d, s, self._c["dst_sorted"], self._c["max_in_deg"])
            else:
                self._c[key] = build_csr(n_s, n_d, s, d, self._c["src_sorted"], self._c["max_out_deg"])
        return self._c[key]

    def csc(self):
        """rows = destination nodes, indices = source ids (the matrix SpMM traverses)."""
        return self._fmt("csc")

    def csr(self):
        """rows = source nodes, indices = destination ids."""
        return self._fmt("csr")

    def in_degrees(self):
        if self.device.type == "cuda":
            return self.csc().degrees()
        return torch.bincount(self.dst.long(), minlength=self.n_dst).to(torch.int32)

    def out_degrees(self):
        if self.device.type == "cuda":
            return self.csr().degrees()
        return torch.bincount(self.src.long(), minlength=self.n_src).to(torch.int32)
