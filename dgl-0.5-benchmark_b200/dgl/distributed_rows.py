"""1-D row partition of a full-graph workload across the GPUs of one node (one process per GPU,
torch.distributed over NCCL / NVLink; gloo on CPU for the host-logic tests).

New relative to the reference (single V100, no torch.distributed anywhere -- SURVEY.md 2.1); required
by BASELINE.json's north_star.  Scheme:

  * nodes are split into P contiguous ranges balanced by in-edge count (prefix sum of in-degrees);
    rank r owns the destination rows [lo_r, hi_r), their feature rows, labels and masks;
  * forward aggregation  out[v] = sum_{u->v} X[u]  for v in the rank's range needs X of every source:
    the ranks all-gather their feature rows (each row crosses NVLink once per rank, which beats
    per-edge peer loads by the average in-degree / P on these graphs), then run the SAME gspmm kernel
    on their slice of the CSC -- a block graph with N global sources and (hi-lo) local destinations;
  * backward  dX[u] = sum_{u->v} dZ[v]  for u in the rank's range: all-gather dZ, then gspmm on the
    rank's slice of the CSR (edges whose SOURCE is local), so there is no reduce-scatter and no
    atomics and every output row is summed in one place (deterministic);
  * rows are never split between ranks, and a rank's CSC slice keeps the global edge order, so each
    output row is accumulated in exactly the order the single-GPU kernel uses: P-way results are
    bit-identical to 1-GPU results for sums, arg-max and structure.

Layout of gathered operands: every rank's shard is padded to the largest shard (`max_rows`) so the
collective is ONE equal-sized `all_gather_into_tensor` straight into the buffer the kernels read
(ragged all-gathers fall back to per-rank broadcasts and cost 2x on 8 GPUs, profiles/r01_notes.md);
the column indices of the rank's CSC / CSR slices are remapped once, at build time, into that padded
id space (`owner * max_rows + local id`), so no unpacking copy is ever needed.

Overlap (`ring=True`): the rank's CSC slice is further split by SOURCE OWNER into P column blocks.
The feature shards travel in P-1 point-to-point rounds (round k: receive the shard of rank r+k, send
ours to rank r-k -- every NVLink direction busy, NVSwitch gives all pairs full bandwidth); all rounds
are queued up front on the communication stream, and the compute stream aggregates block (r+k) as
soon as round k has landed, accumulating into the output (`dglb_gspmm_csr(..., accumulate=1)`), so
the transfer of shard k+1 hides behind the aggregation of shard k.  The block order changes the
summation order: results match the exact path to tolerance, not bit for bit.
"""
import numpy as np
import torch
import torch.distributed as dist

from .heterograph import create_block
from . import ops


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

    def __init__(self, n_nodes, world, rank, ranges, fwd_block, bwd_block, fwd_local=None, fwd_remote=None,
                 n_local_edges=0, group=None):
        self.n_nodes, self.world, self.rank, self.ranges = n_nodes, world, rank, ranges
        self.lo, self.hi = ranges[rank]
        self.n_local_rows = self.hi - self.lo
        self.local_graph = fwd_block        # N global sources -> local destinations (CSC slice)
        self.bwd_graph = bwd_block          # N global destinations -> local sources (CSR slice, reversed)
        self.fwd_local, self.fwd_remote = fwd_local, fwd_remote
        self.n_local_edges = n_local_edges
        self.group = group
        self.sizes = [hi - lo for lo, hi in ranges]
        self.shard_blocks = None            # ring mode: block[r] = edges whose source lives on rank r
        self.max_rows = max(self.sizes)
        self.n_pad = self.max_rows * world

    @staticmethod
    def build(src, dst, n_nodes, world, rank, device, overlap=False, group=None, ring=False):
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        indeg = np.bincount(dst, minlength=n_nodes)
        ranges = balanced_row_ranges(indeg, world)
        lo, hi = ranges[rank]
        his = np.array([r[1] for r in ranges])
        los = np.array([r[0] for r in ranges])
        max_rows = int(max(h - l for l, h in ranges))
        n_pad = max_rows * world

        def pad_ids(ids):                       # global node id -> row of the padded gather buffer
            own = np.searchsorted(his, ids, side="right")
            return own * max_rows + (ids - los[own]), own

        sel = (dst >= lo) & (dst < hi)          # keeps the global (edge-id) order of the selected edges
        s_f, d_f = src[sel], dst[sel] - lo
        s_pad, owner = pad_ids(s_f)
        fwd = create_block((torch.from_numpy(s_pad), torch.from_numpy(d_f)), n_pad, hi - lo).int().to(device)
        selb = (src >= lo) & (src < hi)
        # backward: rows = local sources, columns = global destinations; as a block: "sources" are the
        # global dst nodes (whose dZ rows are gathered), "destinations" the local src nodes
        bwd = create_block((torch.from_numpy(pad_ids(dst[selb])[0]), torch.from_numpy(src[selb] - lo)),
                           n_pad, hi - lo).int().to(device)
        fl = fr = None
        if overlap:
            loc = owner == rank
            fl = create_block((torch.from_numpy(s_f[loc] - lo), torch.from_numpy(d_f[loc])), hi - lo, hi - lo).int().to(device)
            fr = create_block((torch.from_numpy(s_pad[~loc]), torch.from_numpy(d_f[~loc])), n_pad, hi - lo).int().to(device)
        part = RowPartition(n_nodes, world, rank, ranges, fwd, bwd, fl, fr, int(sel.sum()), group)
        part.max_rows, part.n_pad = max_rows, n_pad
        if ring:
            part.shard_blocks = []
            for r in range(world):
                m = owner == r
                part.shard_blocks.append(
                    create_block((torch.from_numpy(s_pad[m]), torch.from_numpy(d_f[m])), n_pad, hi - lo).int().to(device))
        return part

    # ------------------------------------------------------------------ collectives
    def all_gather_rows(self, x_local, async_op=False):
        """All ranks' rows in the padded layout: (world * max_rows, ...), rank r's rows starting at
        r * max_rows (rows beyond a shard's size are padding and never referenced by the kernels).
        One equal-sized, in-place all_gather_into_tensor."""
        x_local = x_local.contiguous()
        out = torch.empty((self.n_pad,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        mine = out[self.rank * self.max_rows:(self.rank + 1) * self.max_rows]
        mine[: x_local.shape[0]].copy_(x_local)
        if self.world == 1:
            return (out, None) if async_op else out
        work = dist.all_gather_into_tensor(out, mine, group=self.group, async_op=async_op)
        return (out, work) if async_op else out

    def unpad(self, gathered):
        """(N, ...) tensor in global node order from a padded gather buffer (tests / debugging)."""
        return torch.cat([gathered[r * self.max_rows: r * self.max_rows + (hi - lo)]
                          for r, (lo, hi) in enumerate(self.ranges)], 0)

    def ring_exchange(self, x_local):
        """Start the P-1 point-to-point rounds that fill the (N, ...) buffer with every rank's rows.
        Returns (buffer, [(owner_rank, work or None), ...]) in arrival order, own shard first."""
        x_local = x_local.contiguous()
        buf = torch.empty((self.n_pad,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        buf[self.rank * self.max_rows: self.rank * self.max_rows + x_local.shape[0]].copy_(x_local)
        order = [(self.rank, None)]
        for k in range(1, self.world):
            src_rank = (self.rank + k) % self.world
            dst_rank = (self.rank - k) % self.world
            lo, hi = self.ranges[src_rank]
            ops_ = [dist.P2POp(dist.isend, x_local, dst_rank, group=self.group),
                    dist.P2POp(dist.irecv, buf[src_rank * self.max_rows: src_rank * self.max_rows + (hi - lo)],
                               src_rank, group=self.group)]
            works = dist.batch_isend_irecv(ops_)
            order.append((src_rank, works))
        return buf, order

    def ring_copy_u_sum(self, x_local):
        """gspmm(copy_lhs, sum) over the rank's rows, one source shard at a time (no autograd)."""
        from . import sparse as K
        buf, order = self.ring_exchange(x_local)
        out = None
        for owner, works in order:
            if works is not None:
                for w in works:
                    w.wait()
            blk = self.shard_blocks[owner]._graph
            if out is None:
                out, _ = K._gspmm(blk, "copy_lhs", "sum", buf, None)
            else:
                K._gspmm(blk, "copy_lhs", "sum", buf, None, out=out)
        return out, buf

    def ring_u_dot_v(self, u_local, v_local, gathered=None):
        """gsddmm(dot) for the rank's edges, per source shard; returns the per-shard (E_r, 1) results.
        `gathered` = (buffer, order) from a ring_exchange already in flight for the same operand."""
        from . import sparse as K
        buf, order = gathered if gathered is not None else self.ring_exchange(u_local)
        outs = []
        for owner, works in order:
            if works is not None:
                for w in works:
                    w.wait()
            outs.append(K._gsddmm(self.shard_blocks[owner]._graph, "dot", buf, v_local))
        return outs

    # ------------------------------------------------------------------ partitioned ops
    def copy_u_sum(self, x_local, reduce_op="sum"):
        """Row-partitioned gspmm(copy_lhs, sum|mean) with autograd (all-gather fwd, all-gather bwd)."""
        return _PartitionedCopyUSum.apply(self, x_local, reduce_op)

    def u_dot_v(self, u_local, v_local):
        """Row-partitioned gsddmm(dot) for the edges whose destination is local."""
        u_full = self.all_gather_rows(u_local)
        return ops.gsddmm(self.local_graph, "dot", u_full, v_local)


class _PartitionedCopyUSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, part, x_local, reduce_op):
        ctx.part, ctx.reduce_op = part, reduce_op
        with torch.no_grad():
            if part.fwd_local is not None and part.world > 1:
                x_full, work = part.all_gather_rows(x_local, async_op=True)
                out = ops.gspmm(part.fwd_local, "copy_lhs", "sum", x_local, None)   # overlaps the collective
                if work is not None:
                    work.wait()
                out = out + ops.gspmm(part.fwd_remote, "copy_lhs", "sum", x_full, None)
                if reduce_op == "mean":
                    deg = part.local_graph.in_degrees().clamp(min=1).to(out.dtype)
                    out = out / deg.view(-1, 1)
            else:
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
            dz_full = part.all_gather_rows(dz_local)
            dx = ops.gspmm(part.bwd_graph, "copy_lhs", "sum", dz_full, None)
        return None, dx, None
