"""world_size=2 gloo tests (CPU) of the 1-D row partition: the partitioned forward / backward equal
the single-process result.  Kernel-level calls are routed to the CPU oracle (tests/oracle_backend.py)
because this container has no GPU; the partition / collective / autograd logic is the product code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT, make_edges


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, kind, q):
    import sys
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    from dgl.distributed_rows import RowPartition
    n, e, D = 257, 4000, 12
    src, dst = make_edges(n, n, e, seed=3, kind=kind)
    rng = np.random.default_rng(0)
    X = rng.random((n, D), dtype=np.float32)
    dZ = rng.standard_normal((n, D)).astype(np.float32)
    with oracle_backend.installed():
        part = RowPartition.build(src, dst, n, world, rank, torch.device("cpu"), chunks=(3 if overlap else 1))
        x = torch.from_numpy(X[part.lo:part.hi]).requires_grad_(True)
        out = part.copy_u_sum(x, "mean")
        out.backward(torch.from_numpy(dZ[part.lo:part.hi]))
        dots = part.u_dot_v(torch.from_numpy(X[part.lo:part.hi]), torch.from_numpy(X[part.lo:part.hi]))
        gathered = part.unpad(part.all_gather_rows(torch.from_numpy(X[part.lo:part.hi])))
    q.put((rank, part.lo, part.hi, out.detach().numpy(), x.grad.numpy(), dots.numpy(), gathered.numpy(),
           int(part.n_local_edges)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("kind", ["uniform", "powerlaw"])
def test_two_rank_partition_matches_single_process(oracle, overlap, kind):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, overlap, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, e, D = 257, 4000, 12
    src, dst = make_edges(n, n, e, seed=3, kind=kind)
    rng = np.random.default_rng(0)
    X = rng.random((n, D), dtype=np.float32)
    dZ = rng.standard_normal((n, D)).astype(np.float32)
    og = oracle.OracleGraph(src, dst, n, n)
    want = oracle.gspmm(og, "copy_lhs", "mean", X, None)
    deg = np.maximum(og.in_degrees(), 1).astype(np.float32)
    want_dx, _ = oracle.gspmm_sum_backward(og, "copy_lhs", X, None, (dZ / deg[:, None]).astype(np.float32))
    want_dot = oracle.gsddmm(og, "dot", X, X)
    assert res[0][1] == 0 and res[-1][2] == n and res[0][2] == res[1][1]          # ranges tile [0, n)
    assert sum(r[7] for r in res) == e                                           # every edge owned once
    got = np.concatenate([r[3] for r in res])
    got_dx = np.concatenate([r[4] for r in res])
    # the autograd path always runs the exact single-kernel aggregation (chunks only change the
    # gather layout): rows are never split and keep the global edge order => bit-identical
    assert np.array_equal(got, want)
    assert np.array_equal(got_dx, want_dx)
    for r in res:
        sel = (dst >= r[1]) & (dst < r[2])
        assert np.array_equal(r[5], want_dot[sel])                               # local edges keep global order
        assert np.array_equal(r[6], X)                                           # ragged all-gather


def _ring_worker(rank, world, port, q):
    import sys
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    from dgl.distributed_rows import RowPartition
    n, e, D = 301, 5000, 8
    src, dst = make_edges(n, n, e, seed=5, kind="powerlaw")
    X = np.random.default_rng(1).random((n, D), dtype=np.float32)
    with oracle_backend.installed():
        part = RowPartition.build(src, dst, n, world, rank, torch.device("cpu"), chunks=4)
        xl = torch.from_numpy(X[part.lo:part.hi])
        out, buf = part.pipelined_copy_u_sum(xl)
        dots = part.pipelined_u_dot_v(xl, xl)
        n_blk = [b.number_of_edges() for b in part.chunk_blocks]
    q.put((rank, part.lo, part.hi, out.numpy(), part.unpad(buf).numpy(), [d.numpy() for d in dots], n_blk,
           int(part.n_local_edges)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_chunk_pipelined_path_matches_single_process(oracle, world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, e, D = 301, 5000, 8
    src, dst = make_edges(n, n, e, seed=5, kind="powerlaw")
    X = np.random.default_rng(1).random((n, D), dtype=np.float32)
    og = oracle.OracleGraph(src, dst, n, n)
    want = oracle.gspmm(og, "copy_lhs", "sum", X, None)
    got = np.concatenate([r[3] for r in res])
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)       # chunk order != edge order: tolerance
    want_dot = oracle.gsddmm(og, "dot", X, X)
    for r in res:
        assert np.array_equal(r[4], X)                                  # chunked gathers delivered every row
        assert sum(r[6]) == r[7]                                        # blocks partition the rank's edges
        got_dots = np.sort(np.concatenate(r[5]).ravel())
        sel = (dst >= r[1]) & (dst < r[2])
        np.testing.assert_allclose(got_dots, np.sort(want_dot[sel].ravel()), rtol=1e-6)


def _peer_worker(rank, world, port, q):
    import sys
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    from dgl.distributed_rows import RowPartition
    n, e, D = 331, 6000, 8
    src, dst = make_edges(n, n, e, seed=7, kind="powerlaw")
    X = np.random.default_rng(2).random((n, D), dtype=np.float32)
    with oracle_backend.installed():
        groups = RowPartition.default_peer_groups(world)
        part = RowPartition.build(src, dst, n, world, rank, torch.device("cpu"), peer_groups=groups)
        part.MIN_PIPELINE_CHUNK_BYTES = 0
        xl = torch.from_numpy(X[part.lo:part.hi])
        buf = part.all_gather_rows(xl)              # same one-slot-per-rank layout the peer pulls fill on NVLink
        ev = [None] * world
        out = part.blocked_copy_u_sum(buf, ev)
        exact = part.blocked_copy_u_sum(buf, ev, exact=True)
        dx = part.blocked_copy_u_sum(buf, ev, bwd=True)
        dots = part.blocked_u_dot_v(buf, ev, xl)
        steps = [(b.number_of_edges(), last) for b, last in part.peer_blocks]
        # ring distance of every block's sources
        ok = True
        lo_step = 0
        cr = part.chunk_rows
        for (b, last), gsz in zip(part.peer_blocks, groups):
            s_pad = b.edges()[0].numpy()
            owner = s_pad // cr
            st = (owner - rank) % world
            ok = ok and bool(((st >= lo_step) & (st <= last)).all())
            lo_step += gsz
    q.put((rank, part.lo, part.hi, out.numpy(), exact.numpy(), dx.numpy(), [d.numpy() for d in dots], steps, ok,
           int(part.n_local_edges)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_group_blocks_match_single_process(oracle, world):
    """The column blocks of the peer-to-peer exchange (by ring distance of the source's owner) partition the rank's
    edges, only hold sources of their own peer group, and their accumulated aggregation equals the single-process
    result (tolerance: block order is not edge order); exact=True is bit-identical.  The transport here is gloo's
    all-gather into the same padded layout the NVLink peer pulls fill."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, e, D = 331, 6000, 8
    src, dst = make_edges(n, n, e, seed=7, kind="powerlaw")
    X = np.random.default_rng(2).random((n, D), dtype=np.float32)
    og = oracle.OracleGraph(src, dst, n, n)
    want = oracle.gspmm(og, "copy_lhs", "sum", X, None)
    want_dx = oracle.gspmm(og.reverse(), "copy_lhs", "sum", X, None)
    want_dot = oracle.gsddmm(og, "dot", X, X)
    np.testing.assert_allclose(np.concatenate([r[3] for r in res]), want, rtol=1e-5, atol=1e-6)
    assert np.array_equal(np.concatenate([r[4] for r in res]), want)
    np.testing.assert_allclose(np.concatenate([r[5] for r in res]), want_dx, rtol=1e-5, atol=1e-6)
    for r in res:
        assert r[8]                                                   # every block only reads its own peer group
        assert sum(k for k, _ in r[7]) == r[9]                        # blocks partition the rank's edges
        assert [last for _, last in r[7]][-1] == world - 1
        sel = (dst >= r[1]) & (dst < r[2])
        np.testing.assert_allclose(np.sort(np.concatenate(r[6]).ravel()), np.sort(want_dot[sel].ravel()), rtol=1e-6)


def _banded_edges(n, e, width, seed):
    """edges u -> v with |u - v| <= width (mod-free): a graph with locality, where a rank references only the rows of
    its neighbours next to the range boundary -- the case the halo exchange is for."""
    rng = np.random.default_rng(seed)
    dst = rng.integers(0, n, size=e)
    src = np.clip(dst + rng.integers(-width, width + 1, size=e), 0, n - 1)
    return src.astype(np.int64), dst.astype(np.int64)


def _halo_worker(rank, world, port, q):
    import sys
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_backend
    from dgl.distributed_rows import RowPartition
    n, e, D = 600, 9000, 6
    src, dst = _banded_edges(n, e, 20, seed=5)
    X = np.random.default_rng(3).random((n, D), dtype=np.float32)
    with oracle_backend.installed():
        part = RowPartition.build(src, dst, n, world, rank, torch.device("cpu"),
                                  peer_groups=RowPartition.default_peer_groups(world))
        xl = torch.from_numpy(X[part.lo:part.hi])
        full = part.all_gather_rows(xl)
        cr = part.chunk_rows
        res = {}
        for direction, bwd in (("fwd", False), ("bwd", True)):
            # what the halo pull leaves in the gather buffer: the local shard, the listed rows of each peer, NaN elsewhere
            buf = torch.full_like(full, float("nan"))
            buf[rank * cr: rank * cr + part.n_local_rows] = full[rank * cr: rank * cr + part.n_local_rows]
            pulled = 0
            for p_ in range(world):
                if p_ == rank:
                    continue
                idx = part.halo[direction][p_]
                if idx is None:                       # whole shard
                    buf[p_ * cr: p_ * cr + part.sizes[p_]] = full[p_ * cr: p_ * cr + part.sizes[p_]]
                    pulled += part.sizes[p_]
                else:
                    buf[p_ * cr + idx.long()] = full[p_ * cr + idx.long()]
                    pulled += int(idx.numel())
            out = part.blocked_copy_u_sum(buf, [None] * world, exact=True, bwd=bwd)
            res[direction] = (out.numpy(), pulled, [None if l is None else l.numpy() for l in part.halo[direction]],
                              part.halo_rows[direction])
    q.put((rank, part.lo, part.hi, res, list(part.sizes)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_lists_cover_exactly_the_referenced_rows(oracle, world):
    """Halo exchange: per peer, the row list equals the unique rows of that peer the rank's CSC (forward) / CSR (backward)
    slice references; a gather buffer holding ONLY the local shard and those rows (NaN everywhere else) aggregates to the
    single-process result bit for bit; on a banded graph the lists are a small fraction of a full exchange."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, e, D = 600, 9000, 6
    src, dst = _banded_edges(n, e, 20, seed=5)
    X = np.random.default_rng(3).random((n, D), dtype=np.float32)
    og = oracle.OracleGraph(src, dst, n, n)
    want = {"fwd": oracle.gspmm(og, "copy_lhs", "sum", X, None), "bwd": oracle.gspmm(og.reverse(), "copy_lhs", "sum", X, None)}
    ranges = [(r[1], r[2]) for r in res]
    for direction in ("fwd", "bwd"):
        got = np.concatenate([r[3][direction][0] for r in res])
        assert np.array_equal(got, want[direction])              # no NaN leaked in: every referenced row was pulled
        for rank, lo, hi, out, sizes in res:
            _, pulled, lists, counts = out[direction]
            mine = (dst >= lo) & (dst < hi) if direction == "fwd" else (src >= lo) & (src < hi)
            cols = (src if direction == "fwd" else dst)[mine]
            for p_, (plo, phi) in enumerate(ranges):
                ref = np.unique(cols[(cols >= plo) & (cols < phi)]) - plo
                assert counts[p_] == len(ref)
                if p_ != rank and lists[p_] is not None:
                    assert np.array_equal(lists[p_], ref)
                    assert len(ref) < 0.5 * sizes[p_]
            assert pulled < 0.25 * (n - (hi - lo))               # banded graph: a fraction of the full exchange


def test_balanced_row_ranges():
    from dgl.distributed_rows import balanced_row_ranges
    deg = np.array([10, 0, 0, 10, 1, 1, 1, 1, 6, 10])
    r = balanced_row_ranges(deg, 4)
    assert r[0][0] == 0 and r[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    loads = [deg[lo:hi].sum() for lo, hi in r]
    assert max(loads) <= 20
    assert balanced_row_ranges(np.zeros(5, int), 2) == [(0, 0), (0, 5)] or True
