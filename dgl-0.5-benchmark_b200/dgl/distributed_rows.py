"""1-D row partition of a full-graph workload across the GPUs of one node (one process per GPU,
torch.distributed over NCCL / NVLink; gloo on CPU for the host-logic tests).

New relative to the reference (single V100, no torch.distributed anywhere -- SURVEY.md 2.1); required
by BASELINE.json's north_star.  Scheme:

  * nodes are split into P contiguous ranges balanced by in-edge count (prefix sum of in-degrees);
    rank r owns the destination rows [lo_r, hi_r), their feature rows, labels and masks;
  * forward aggregation  out[v] = sum_{u->v} X[u]  for v in the rank's range needs X of every source:
    the ranks all-gather their feature rows (each row crosses NVLink once per rank, which beats
    per-edge peer loads by the average in-degree / P on these graphs), then run the SAME gspmm kernel
    on their slice of the CSC -- a block graph with all sources and (hi-lo) local destinations;
  * backward  dX[u] = sum_{u->v} dZ[v]  for u in the rank's range: all-gather dZ, then gspmm on the
    rank's slice of the CSR (edges whose SOURCE is local), so there is no reduce-scatter and no
    atomics and every output row is summed in one place (deterministic);
  * rows are never split between ranks, and a rank's CSC slice keeps the global edge order, so each
    output row is accumulated in exactly the order the single-GPU kernel uses: P-way results are
    bit-identical to 1-GPU results for sums, arg-max and structure.

Layout of gathered operands.  A rank's rows are cut into K chunks of `chunk_rows` rows and every
(chunk, rank) slot is padded to that size, so chunk k of all ranks is ONE equal-sized, in-place
`all_gather_into_tensor` into rows [k*P*chunk_rows, (k+1)*P*chunk_rows) of the buffer the kernels
read (ragged all-gathers fall back to per-rank broadcasts and cost 2x on 8 GPUs; shard-by-shard
point-to-point rounds reached only 160 GB/s -- profiles/r01_notes.md).  The column indices of the
rank's CSC / CSR slices are remapped once, at build time, into that padded id space
((k*P + owner)*chunk_rows + r), so nothing is ever unpacked.

Overlap (K > 1): the CSC slice is also split by source chunk into K column blocks.  All K gathers are
queued on the communication stream up front; the compute stream aggregates block k as soon as gather
k has landed, accumulating into the output (`dglb_gspmm_csr(..., accumulate=1)`), so gather k+1
travels while block k is aggregated.  The block order changes the summation order: this path
matches the exact path to tolerance, not bit for bit.

Peer-to-peer exchange (enable_p2p; the path bench.py and the epoch models use on NVLink boxes).  NCCL's all-gather is
a kernel that occupies SMs and delivers every shard at the same time, so nothing can be aggregated until the whole
collective has landed (8 GPUs, D = 602: 0.93 ms exposed in front of a 0.68 ms kernel).  Here every rank publishes its
rows in a symmetric-memory buffer (torch.distributed._symmetric_memory: the peers' buffers are mapped into this
process over NVLink) and PULLS the other shards with plain device-to-device copies on a side stream -- copy engines,
no SMs -- in ring order: step k fetches the shard of rank (r + k) mod P, so in every step each GPU serves exactly one
reader at full link rate and shard k lands after k/(P-1) of the exchange.  The rank's CSC slice is pre-split by the
ring distance of the source's owner into a few column blocks (`peer_groups`, e.g. [1, 3, 4]: local rows | the next
three peers | the last four); block g is aggregated as soon as its last shard has landed (CUDA events, no spinning
kernels), accumulating into the output, while the later shards are still in flight.  The local block needs no
communication at all.  Block order changes the summation order (tolerance, not bit-identity); exact=True waits for
every shard and runs the single-kernel path instead.

Halo exchange (north_star: "halo and full feature all-gather").  At build time every rank lists, per peer, the sorted
unique rows of that peer which its CSC slice (forward) / CSR slice (backward) actually references.  When a list covers
less than HALO_MAX_FRACTION of the peer's rows, the pull of that shard becomes an indexed row copy
(csrc/row_copy.cu: dst[idx] = src[idx], an SM kernel on the copy stream reading the peer's symmetric buffer over NVLink)
instead of a whole-shard copy-engine transfer: only referenced rows cross the link.  Rows land at the same positions
of the gather buffer, so the blocks, their column ids and the summation order are untouched and results stay
bit-identical to the full exchange; unreferenced rows of the buffer are never written and never read.  On the uniform
random graphs of the benchmark (average degree 25-50 over 8 ranks) every rank references nearly every row and the
exchange stays a full one; graphs with locality (banded / clustered orderings) move a fraction of the bytes.
"""
import numpy as np
import torch
import torch.distributed as dist

from .heterograph import create_block
from . import ops
from . import sparse as K_
from . import _capi


def balanced_row_ranges(in_degrees, world):
    """Contiguous [lo, hi) node ranges with ~equal in-edge counts.  in_degrees: int array (N,)."""
    n = len(in_degrees)
    csum = np.concatenate([[0], np.cumsum(in_degrees, dtype=np.int64)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        b = int(np.searchsorted(csum, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


class RowPartition:
    """Per-rank view of a square graph with N nodes given by its creation-order COO (src, dst)."""

    def __init__(self, n_nodes, world, rank, ranges, chunks, group=None):
        self.n_nodes, self.world, self.rank, self.ranges = n_nodes, world, rank, ranges
        self.lo, self.hi = ranges[rank]
        self.n_local_rows = self.hi - self.lo
        self.group = group
        self.sizes = [hi - lo for lo, hi in ranges]
        self.chunks = chunks
        self.chunk_rows = max(1, -(-max(self.sizes) // chunks))
        self.n_pad = self.chunk_rows * world * chunks
        self.local_graph = None      # all sources (padded ids) -> local destinations (CSC slice)
        self.bwd_graph = None        # all destinations (padded ids) -> local sources (CSR slice, reversed)
        self.chunk_blocks = None     # K > 1: block[k] = the rank's edges whose source lies in chunk k
        self.n_local_edges = 0
        self._fwd_geid = self._bwd_geid = None   # global edge ids of the two blocks, in block order
        self._geid_cache = {}
        self.peer_blocks = None      # [(block graph, last ring step it needs)] forward blocks by owner ring distance
        self.peer_blocks_bwd = None  # the same split of the backward block
        self._p2p = None             # (trailing shape, dtype) -> (symmetric tensor, handle, [peer views], gather buffer)
        self._p2p_group = None
        self._copy_streams = None
        self.peer_group_sizes = None
        self.exact = True            # peer-to-peer autograd path: single-kernel (bit-identical) or blocked aggregation
        self.halo = {"fwd": None, "bwd": None}   # per direction: [int32 device row list per peer, or None = whole shard]
        self.halo_rows = {"fwd": None, "bwd": None}   # per direction: rows referenced per peer (host ints, diagnostics)

    # ------------------------------------------------------------------ construction
    def pad_ids(self, ids):
        """global node id -> row of the padded gather buffer; also returns the chunk id."""
        his = np.array([r[1] for r in self.ranges])
        los = np.array([r[0] for r in self.ranges])
        own = np.searchsorted(his, ids, side="right")
        local = ids - los[own]
        k = local // self.chunk_rows
        return (k * self.world + own) * self.chunk_rows + (local - k * self.chunk_rows), k

    HALO_MAX_FRACTION = 0.5   # pull a peer's shard row by row when fewer than this share of its rows is referenced

    def _halo_lists(self, cols_global, device):
        """Per peer: sorted unique LOCAL row ids (within the peer's range) among `cols_global`, as an int32 device
        tensor when they are few enough for the indexed pull, else None (whole shard); plus the referenced-row counts."""
        his = np.array([r[1] for r in self.ranges])
        los = np.array([r[0] for r in self.ranges])
        uniq = np.unique(cols_global)
        own = np.searchsorted(his, uniq, side="right")
        lists, counts = [], []
        for p in range(self.world):
            rows = (uniq[own == p] - los[p]).astype(np.int32)
            counts.append(int(rows.shape[0]))
            if p != self.rank and rows.shape[0] < self.HALO_MAX_FRACTION * self.sizes[p]:
                lists.append(torch.from_numpy(rows).to(device))
            else:
                lists.append(None)
        return lists, counts

    @staticmethod
    def default_peer_groups(world):
        """Column blocks by ring distance of the source's owner: the local rows, then a few groups of peers --
        enough blocks to start early, few enough that re-reading the partial output stays cheap."""
        if world <= 1:
            return [1]
        if world == 2:
            return [1, 1]
        if world <= 4:
            return [1, 1, world - 2]
        rest = world - 1
        a = rest // 3
        return [1, a, a, rest - 2 * a]

    def ring_step_of(self, ids):
        """ring distance (owner - rank) mod P of the owner of each global node id"""
        his = np.array([r[1] for r in self.ranges])
        own = np.searchsorted(his, ids, side="right")
        return (own - self.rank) % self.world

    def _split_by_ring_step(self, cols_global, s_pad, d_loc, device, groups):
        step = self.ring_step_of(cols_global)
        blocks, lo = [], 0
        for gsz in groups:
            m = (step >= lo) & (step < lo + gsz)
            blocks.append((create_block((torch.from_numpy(s_pad[m]), torch.from_numpy(d_loc[m])), self.n_pad,
                                        self.n_local_rows).int().to(device), lo + gsz - 1))
            lo += gsz
        assert lo == self.world, "peer_groups must sum to the world size"
        return blocks

    @staticmethod
    def build(src, dst, n_nodes, world, rank, device, chunks=1, group=None, peer_groups=None):
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        indeg = np.bincount(dst, minlength=n_nodes)
        part = RowPartition(n_nodes, world, rank, balanced_row_ranges(indeg, world), chunks, group)
        lo, hi = part.lo, part.hi
        sel = (dst >= lo) & (dst < hi)          # keeps the global (edge-id) order of the selected edges
        s_pad, s_chunk = part.pad_ids(src[sel])
        d_loc = dst[sel] - lo
        part.n_local_edges = int(sel.sum())
        part._fwd_geid = np.nonzero(sel)[0].astype(np.int32)
        part.local_graph = create_block((torch.from_numpy(s_pad), torch.from_numpy(d_loc)), part.n_pad, hi - lo).int().to(device)
        selb = (src >= lo) & (src < hi)
        part._bwd_geid = np.nonzero(selb)[0].astype(np.int32)
        # backward: rows = local sources, columns = all destinations; as a block the "sources" are the
        # destination nodes (whose dZ rows are gathered) and the "destinations" the local source nodes
        part.bwd_graph = create_block((torch.from_numpy(part.pad_ids(dst[selb])[0]), torch.from_numpy(src[selb] - lo)),
                                      part.n_pad, hi - lo).int().to(device)
        if chunks == 1:
            part.halo["fwd"], part.halo_rows["fwd"] = part._halo_lists(src[sel], device)
            part.halo["bwd"], part.halo_rows["bwd"] = part._halo_lists(dst[selb], device)
        if peer_groups is not None:
            assert chunks == 1, "peer blocks use the one-slot-per-rank gather layout"
            part.peer_group_sizes = list(peer_groups)
            part.peer_blocks = part._split_by_ring_step(src[sel], s_pad, d_loc, device, peer_groups)
            part.peer_blocks_bwd = part._split_by_ring_step(dst[selb], part.pad_ids(dst[selb])[0], src[selb] - lo, device,
                                                            peer_groups)
        if chunks > 1:
            part.chunk_blocks = []
            for k in range(chunks):
                m = s_chunk == k
                part.chunk_blocks.append(create_block((torch.from_numpy(s_pad[m]), torch.from_numpy(d_loc[m])),
                                                      part.n_pad, hi - lo).int().to(device))
        return part

    # ------------------------------------------------------------------ collectives
    def all_gather_rows(self, x_local, async_op=False):
        """Every rank's rows in the padded layout, (K * P * chunk_rows, ...): one equal-sized in-place
        all_gather_into_tensor per chunk.  async_op=True returns (buffer, [work per chunk])."""
        x_local = x_local.contiguous()
        cr, P = self.chunk_rows, self.world
        out = torch.empty((self.n_pad,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        works = []
        for k in range(self.chunks):
            base = (k * P + self.rank) * cr
            a, b = min(k * cr, x_local.shape[0]), min((k + 1) * cr, x_local.shape[0])
            if b > a:
                out[base: base + (b - a)].copy_(x_local[a:b])
            if P > 1:
                w = dist.all_gather_into_tensor(out[k * P * cr:(k + 1) * P * cr], out[base: base + cr],
                                                group=self.group, async_op=async_op)
                works.append(w)
            else:
                works.append(None)
        return (out, works) if async_op else out

    # ------------------------------------------------------------------ peer-to-peer exchange (symmetric memory)
    def enable_p2p(self, group=None):
        """Switch the exchange to ring-ordered peer pulls through symmetric memory (CUDA + NCCL process group)."""
        assert self.chunks == 1, "the peer-to-peer exchange uses the one-slot-per-rank gather layout"
        self._p2p, self._p2p_group = {}, (group if group is not None else dist.group.WORLD)
        # ONE copy stream: measured at 8 GPUs (profiles/r02_p2p_copy_bench_n8.jsonl) a single in-order stream of pulls
        # moves 70 MB shards at 609 GB/s per rank, two streams 562, four 371
        self._copy_streams = [torch.cuda.Stream()]
        self.halo_bytes_pulled = self.full_bytes_pulled = 0   # bytes requested over NVLink by this rank (host counters)
        return self

    @property
    def p2p(self):
        return self._p2p is not None

    def _p2p_slot(self, x, nth=0):
        """Symmetric buffer, handle and peer views for row tensors shaped like x (nth: n-th operand of that shape in
        one exchange).  The first call per key is a collective (rendezvous): every rank must make it in the same order."""
        key = (tuple(x.shape[1:]), x.dtype, nth)
        if key not in self._p2p:
            import torch.distributed._symmetric_memory as symm_mem
            shape = (self.chunk_rows,) + key[0]
            t = symm_mem.empty(shape, dtype=key[1], device=x.device)
            hdl = symm_mem.rendezvous(t, self._p2p_group)
            views = [hdl.get_buffer(r, shape, key[1]) if r != self.rank else t for r in range(self.world)]
            self._p2p[key] = (t, hdl, views)
        return self._p2p[key]

    def p2p_gather(self, xs, order="operand", bwd=False, halo=True):
        """Start the exchange of a list of local row tensors (order: "operand" = all shards of xs[0], then xs[1], ...;
        "group" = peer group by peer group across the operands).  Returns [(buffer, events)] per operand: the padded
        gather buffer (the local shard is in place in stream order) and events[k], k = 1..P-1, which fire once the
        shard of rank (rank + k) mod P has landed.  Two device-side barriers per call (not per operand): peers have
        finished reading what the symmetric buffers held before / every rank has published its new rows.
        bwd: the operands feed the backward block (rows referenced by the CSR slice); halo=False forces whole shards."""
        P, cr, main = self.world, self.chunk_rows, torch.cuda.current_stream()
        seen, slots, gbufs = {}, [], []
        for x in xs:
            k = (tuple(x.shape[1:]), x.dtype)
            slots.append(self._p2p_slot(x, seen.get(k, 0)))
            seen[k] = seen.get(k, 0) + 1
            # fresh gather buffer (caching allocator): callers may keep it for their backward pass
            gbufs.append(torch.empty((self.n_pad,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device))
        hdl0 = slots[0][1]
        hdl0.barrier(channel=0)
        for x, (t, _, _), gbuf in zip(xs, slots, gbufs):
            x = x.contiguous()
            t[: x.shape[0]].copy_(x)
            gbuf[self.rank * cr: self.rank * cr + x.shape[0]].copy_(x)
        hdl0.barrier(channel=1)
        ready = torch.cuda.Event()
        ready.record(main)
        # pull order: peer group by peer group (the order the blocks are aggregated in), inside a group operand by
        # operand; without peer groups operand by operand
        groups, lo = [], 0
        for gsz in (self.peer_group_sizes or [P]):
            groups.append([k for k in range(lo, min(lo + gsz, P)) if k > 0])   # ring step 0 is the local shard
            lo += gsz
        all_events = [[None] * P for _ in xs]
        halo_lists = self.halo["bwd" if bwd else "fwd"] if halo else None
        for st in self._copy_streams:
            st.wait_event(ready)
        i = 0
        if order == "group":
            plan = [(oi, k) for steps in groups for oi in range(len(xs)) for k in steps]
        else:
            plan = [(oi, k) for oi in range(len(xs)) for steps in groups for k in steps]
        for oi, k in plan:
            views, gbuf = slots[oi][2], gbufs[oi]
            peer = (self.rank + k) % P
            n_peer = self.sizes[peer]
            st = self._copy_streams[i % len(self._copy_streams)]
            i += 1
            idx = halo_lists[peer] if halo_lists is not None else None
            with torch.cuda.stream(st):
                if n_peer and idx is not None:
                    if idx.numel():      # halo: only the rows this rank's slice references, fetched by an SM kernel
                        _capi.call(_capi.ops().copy_rows_indexed, views[peer], gbuf[peer * cr: peer * cr + n_peer], idx)
                        self.halo_bytes_pulled += idx.numel() * gbuf[0].numel() * gbuf.element_size()
                elif n_peer:
                    gbuf[peer * cr: peer * cr + n_peer].copy_(views[peer][:n_peer], non_blocking=True)
                    self.full_bytes_pulled += n_peer * gbuf[0].numel() * gbuf.element_size()
                ev = torch.cuda.Event()
                ev.record(st)
            all_events[oi][k] = ev
        for gbuf in gbufs:
            for st in self._copy_streams:
                gbuf.record_stream(st)
        return [(gbuf, ev) for gbuf, ev in zip(gbufs, all_events)]


    def p2p_all_reduce_flat(self, flat):
        """Sum of a small 1-D float tensor over the ranks through symmetric memory (publish, two device barriers, every
        rank adds the P published vectors in rank order): identical bits on every rank, no NCCL kernel, capturable in a
        CUDA graph next to the exchanges above.  Meant for the dense parameter gradients + loss of a full-graph epoch
        (tens of KB); the first call per length is a collective (rendezvous)."""
        key = ("flat", int(flat.numel()), flat.dtype)
        if key not in self._p2p:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty((int(flat.numel()),), dtype=flat.dtype, device=flat.device)
            hdl = symm_mem.rendezvous(t, self._p2p_group)
            views = [hdl.get_buffer(r, (int(flat.numel()),), flat.dtype) if r != self.rank else t for r in range(self.world)]
            self._p2p[key] = (t, hdl, views)
        t, hdl, views = self._p2p[key]
        hdl.barrier(channel=0)          # every peer has finished reading what the buffer held before
        t.copy_(flat)
        hdl.barrier(channel=1)          # every rank has published
        out = views[0].clone()
        for r in range(1, self.world):
            out += views[r]
        return out

    @staticmethod
    def _wait(events, upto):
        pending = [events[k] for k in range(1, upto + 1) if events[k] is not None]
        if pending:
            cur = torch.cuda.current_stream()
            for ev in pending:
                cur.wait_event(ev)

    def blocked_copy_u_sum(self, buf, events, exact=False, bwd=False):
        """gspmm(copy_lhs, sum) of the rank's rows from a gather in flight: one column block per peer group, each as
        soon as its shards have landed (exact=True: wait for everything, single kernel, bit-identical)."""
        g_all = self.bwd_graph if bwd else self.local_graph
        blocks = self.peer_blocks_bwd if bwd else self.peer_blocks
        small = buf.numel() * buf.element_size() // max(self.world, 1) < self.MIN_PIPELINE_CHUNK_BYTES
        if exact or blocks is None or small:
            self._wait(events, self.world - 1)
            return K_._gspmm(g_all._graph, "copy_lhs", "sum", buf, None)[0]
        out = None
        for blk, last in blocks:
            self._wait(events, last)
            if blk.number_of_edges() == 0 and out is not None:
                continue
            if out is None:
                out = K_._gspmm(blk._graph, "copy_lhs", "sum", buf, None)[0]
            else:
                K_._gspmm(blk._graph, "copy_lhs", "sum", buf, None, out=out)
        return out

    def n_blocks(self):
        return len(self.peer_blocks) if self.peer_blocks is not None else 1

    def _is_small(self, buf):
        return (self.peer_blocks is None
                or buf.numel() * buf.element_size() // max(self.world, 1) < self.MIN_PIPELINE_CHUNK_BYTES)

    def block_copy_u_sum(self, buf, events, gi, out):
        """Block gi of blocked_copy_u_sum (for callers that interleave the blocks of several operands): returns the
        running output.  Small operands are aggregated in one piece at the last block index."""
        if self._is_small(buf):
            if gi != self.n_blocks() - 1:
                return out
            self._wait(events, self.world - 1)
            return K_._gspmm(self.local_graph._graph, "copy_lhs", "sum", buf, None)[0]
        blk, last = self.peer_blocks[gi]
        self._wait(events, last)
        if out is None:
            return K_._gspmm(blk._graph, "copy_lhs", "sum", buf, None)[0]
        if blk.number_of_edges():
            K_._gspmm(blk._graph, "copy_lhs", "sum", buf, None, out=out)
        return out

    def block_u_dot_v(self, buf, events, gi, v_local):
        """Block gi of blocked_u_dot_v; None when the block is folded into the last one (small operands)."""
        if self._is_small(buf):
            if gi != self.n_blocks() - 1:
                return None
            self._wait(events, self.world - 1)
            return K_._gsddmm(self.local_graph._graph, "dot", buf, v_local)
        blk, last = self.peer_blocks[gi]
        self._wait(events, last)
        return K_._gsddmm(blk._graph, "dot", buf, v_local)

    def blocked_u_dot_v(self, buf, events, v_local):
        """gsddmm(dot) for the rank's edges, one peer-group block at a time; returns the per-block (E_g, 1) results
        (block g holds the edges whose source is owned by a rank of group g, in global edge order)."""
        small = buf.numel() * buf.element_size() // max(self.world, 1) < self.MIN_PIPELINE_CHUNK_BYTES
        if self.peer_blocks is None or small:
            self._wait(events, self.world - 1)
            return [K_._gsddmm(self.local_graph._graph, "dot", buf, v_local)]
        outs = []
        for blk, last in self.peer_blocks:
            self._wait(events, last)
            outs.append(K_._gsddmm(blk._graph, "dot", buf, v_local))
        return outs

    def unpad(self, gathered):
        """(N, ...) tensor in global node order from a padded gather buffer (tests / debugging)."""
        cr, P = self.chunk_rows, self.world
        parts = []
        for r, (lo, hi) in enumerate(self.ranges):
            n = hi - lo
            for k in range(self.chunks):
                a, b = min(k * cr, n), min((k + 1) * cr, n)
                if b > a:
                    base = (k * P + r) * cr
                    parts.append(gathered[base: base + (b - a)])
        return torch.cat(parts, 0)

    def global_eids(self, which):
        """Global edge ids of the forward ('fwd') / backward ('bwd') block in the order of that block's
        CSC (int32 device tensor): the dropout counter of the fused GAT kernels, so that the rank that
        applies the mask forward and the rank that replays it backward agree."""
        if which not in self._geid_cache:
            blk = self.local_graph if which == "fwd" else self.bwd_graph
            geid = torch.from_numpy(self._fwd_geid if which == "fwd" else self._bwd_geid).to(blk.device)
            csc = blk._graph.csc()
            self._geid_cache[which] = geid if csc.eids is None else geid[csc.eids.long()].contiguous()
        return self._geid_cache[which]

    def gat_attention(self, ft_local, el_local, er_local, negative_slope=0.2, dropout_p=0.0, seed=0):
        """Row-partitioned fused GAT attention with autograd: all-gather (ft, el) forward; backward =
        local destination pass, all-gather (row_pack, grad_rst), local source pass."""
        from .ops.gat import pad_head_dim
        H = ft_local.shape[1]
        ftp, F = pad_head_dim(ft_local)
        rst = _PartitionedGAT.apply(self, ftp, el_local.reshape(-1, H), er_local.reshape(-1, H),
                                    float(negative_slope), float(dropout_p), int(seed))
        return rst if rst.shape[-1] == F else rst[..., :F]

    # ------------------------------------------------------------------ partitioned ops
    def copy_u_sum(self, x_local, reduce_op="sum"):
        """Row-partitioned gspmm(copy_lhs, sum|mean) with autograd (all-gather fwd, all-gather bwd);
        bit-identical to the single-GPU result."""
        return _PartitionedCopyUSum.apply(self, x_local, reduce_op)

    def u_dot_v(self, u_local, v_local):
        """Row-partitioned gsddmm(dot) for the edges whose destination is local."""
        return ops.gsddmm(self.local_graph, "dot", self.all_gather_rows(u_local), v_local)

    MIN_PIPELINE_CHUNK_BYTES = 12 << 20   # per-rank bytes of one gather chunk below which pipelining is skipped

    def _chunk_bytes(self, buf):
        return buf.numel() * buf.element_size() // (self.chunks * self.world)

    def pipelined_copy_u_sum(self, x_local, gathered=None):
        """gspmm(copy_lhs, sum) over the rank's rows, one source chunk at a time behind its gather
        (no autograd).  Returns (out, gather buffer).  `gathered` = (buffer, works) already in flight."""
        buf, works = gathered if gathered is not None else self.all_gather_rows(x_local, async_op=True)
        if self.chunk_blocks is None or self._chunk_bytes(buf) < self.MIN_PIPELINE_CHUNK_BYTES:
            # small operand: K tiny launches would cost more than the overlap buys -- one kernel
            for w in works:
                if w is not None:
                    w.wait()
            return K_._gspmm(self.local_graph._graph, "copy_lhs", "sum", buf, None)[0], buf
        out = None
        for k, w in enumerate(works):
            if w is not None:
                w.wait()
            blk = self.chunk_blocks[k]._graph
            if out is None:
                out = K_._gspmm(blk, "copy_lhs", "sum", buf, None)[0]
            else:
                K_._gspmm(blk, "copy_lhs", "sum", buf, None, out=out)
        return out, buf

    def pipelined_u_dot_v(self, u_local, v_local, gathered=None):
        """gsddmm(dot) for the rank's edges per source chunk; returns the list of per-chunk (E_k, 1)
        results.  `gathered` = (buffer, works) of the u operand (works may already be complete)."""
        buf, works = gathered if gathered is not None else self.all_gather_rows(u_local, async_op=True)
        if self.chunk_blocks is None or self._chunk_bytes(buf) < self.MIN_PIPELINE_CHUNK_BYTES:
            for w in works:
                if w is not None:
                    w.wait()
            return [K_._gsddmm(self.local_graph._graph, "dot", buf, v_local)]
        outs = []
        for k, w in enumerate(works):
            if w is not None:
                w.wait()
            outs.append(K_._gsddmm(self.chunk_blocks[k]._graph, "dot", buf, v_local))
        return outs


class _PartitionedCopyUSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, part, x_local, reduce_op):
        ctx.part, ctx.reduce_op = part, reduce_op
        with torch.no_grad():
            if part.p2p:
                (buf, ev), = part.p2p_gather([x_local])
                out = part.blocked_copy_u_sum(buf, ev, exact=part.exact)
                if reduce_op == "mean":
                    out = out / part.local_graph._graph.csc().mean_divisor().view(-1, 1)
                return out
            x_full = part.all_gather_rows(x_local)
            out = ops.gspmm(part.local_graph, "copy_lhs", reduce_op, x_full, None)
        return out

    @staticmethod
    def backward(ctx, dz_local):
        part = ctx.part
        with torch.no_grad():
            dz_local = dz_local.contiguous()
            if ctx.reduce_op == "mean":
                deg = part.local_graph.in_degrees().clamp(min=1).to(dz_local.dtype)
                dz_local = dz_local / deg.view(-1, 1)
            if part.p2p:
                (buf, ev), = part.p2p_gather([dz_local], bwd=True)
                return None, part.blocked_copy_u_sum(buf, ev, exact=part.exact, bwd=True), None
            dz_full = part.all_gather_rows(dz_local)
            dx = ops.gspmm(part.bwd_graph, "copy_lhs", "sum", dz_full, None)
        return None, dx, None


class _PartitionedGAT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, part, ft, el, er, slope, dropout_p, seed):
        with torch.no_grad():
            if part.p2p:
                (ft_full, _), (el_full, ev) = part.p2p_gather([ft, el])
                part._wait(ev, part.world - 1)       # the copy stream is in order: el's last shard is the last of both
            else:
                ft_full = part.all_gather_rows(ft)
                el_full = part.all_gather_rows(el)
            rst, row_max, row_sum, _ = K_._gat_fwd(part.local_graph._graph, ft_full, el_full, er.contiguous(), slope,
                                                   dropout_p, seed, eids=part.global_eids("fwd"))
        ctx.part, ctx.args = part, (slope, dropout_p, seed)
        ctx.save_for_backward(ft, el, er, ft_full, el_full, row_max, row_sum)
        return rst

    @staticmethod
    def backward(ctx, grad_rst):
        part = ctx.part
        slope, dropout_p, seed = ctx.args
        ft, el, er, ft_full, el_full, row_max, row_sum = ctx.saved_tensors
        with torch.no_grad():
            grad_rst = grad_rst.contiguous()
            if part.p2p:   # grad_rst is known now: it travels while the destination pass runs
                (grad_full, ev_g), = part.p2p_gather([grad_rst], bwd=True)
            row_pack, grad_er = K_._gat_bwd_dst(part.local_graph._graph, ft_full, el_full, er.contiguous(), row_max,
                                                row_sum, grad_rst, slope, dropout_p, seed, eids=part.global_eids("fwd"))
            if part.p2p:
                (pack_full, ev_p), = part.p2p_gather([row_pack], bwd=True)
                part._wait(ev_g, part.world - 1)
                part._wait(ev_p, part.world - 1)
            else:
                pack_full = part.all_gather_rows(row_pack)
                grad_full = part.all_gather_rows(grad_rst)
            # the backward block's CSC has the LOCAL SOURCE nodes as rows and padded destination ids as columns
            grad_ft, grad_el = K_._gat_bwd_src(part.bwd_graph._graph.csc(), ft.contiguous(), el.contiguous(), pack_full,
                                               grad_full, slope, dropout_p, seed, eids=part.global_eids("bwd"))
        return None, grad_ft, grad_el, grad_er, None, None, None
