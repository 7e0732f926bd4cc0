"""-m gpu: the row-partitioned paths (SAGE aggregation fwd/bwd, chunk-pipelined sweep, fused GAT
fwd/bwd with attention dropout) on the real kernels.  P ranks are emulated on ONE device by P host
threads whose `all_gather_rows` meets at a threading.Barrier (no kernel waits on another kernel, only
host threads wait), so the test runs on a single-GPU box; the NCCL transport itself is covered by
bench.py --gpus N and by the gloo tests."""
import threading

import numpy as np
import pytest
import torch

import dgl
from conftest import make_edges
from dgl.distributed_rows import RowPartition

pytestmark = pytest.mark.gpu


class FakeWorld:
    """Barrier-based stand-in for the collective inside RowPartition.all_gather_rows."""

    def __init__(self, world):
        self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def attach(self, part):
        fake = self

        def all_gather_rows(x_local, async_op=False):
            cr, P, K = part.chunk_rows, part.world, part.chunks
            fake.slots[part.rank] = x_local.detach().contiguous()
            fake.barrier.wait(timeout=60)
            out = torch.zeros((part.n_pad,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
            for r in range(P):
                xr = fake.slots[r]
                for k in range(K):
                    a, b = min(k * cr, xr.shape[0]), min((k + 1) * cr, xr.shape[0])
                    if b > a:
                        base = (k * P + r) * cr
                        out[base: base + (b - a)] = xr[a:b]
            torch.cuda.synchronize()
            fake.barrier.wait(timeout=60)
            return (out, [None] * K) if async_op else out
        part.all_gather_rows = all_gather_rows


class Ctx:
    """Minimal autograd-context stand-in: the Functions' forward/backward are called directly from the
    rank threads (the autograd engine runs every CUDA backward node of one device on ONE worker
    thread, which would serialise the emulated ranks and dead-lock at the barrier)."""

    def save_for_backward(self, *ts):
        self.saved_tensors = ts


def run_ranks(world, fn):
    results, errors = [None] * world, []

    def target(r):
        try:
            results[r] = fn(r)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            raise
    threads = [threading.Thread(target=target, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    return results


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("chunks", [1, 3])
def test_partitioned_sage_aggregation_is_bit_identical(cuda, world, chunks):
    n, e, D = 1500, 40000, 48
    src, dst = make_edges(n, n, e, seed=7, kind="powerlaw")
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(cuda)
    torch.manual_seed(0)
    X = torch.rand(n, D, device=cuda)
    dZ = torch.randn(n, D, device=cuda)
    Xr = X.clone().requires_grad_(True)
    want = dgl.ops.gspmm(g, "copy_lhs", "mean", Xr, None)
    want.backward(dZ)
    want_dot = dgl.ops.gsddmm(g, "dot", X, X)
    fake = FakeWorld(world)

    def rank_fn(r):
        part = RowPartition.build(src, dst, n, world, r, cuda, chunks=chunks)
        fake.attach(part)
        from dgl.distributed_rows import _PartitionedCopyUSum as Fn
        x = X[part.lo:part.hi].clone()
        ctx = Ctx()
        out = Fn.forward(ctx, part, x, "mean")
        _, xgrad, _ = Fn.backward(ctx, dZ[part.lo:part.hi])
        pipe, buf = part.pipelined_copy_u_sum(X[part.lo:part.hi])
        dots = part.pipelined_u_dot_v(None, X[part.lo:part.hi], gathered=(buf, [None] * chunks))
        return part.lo, part.hi, out.detach(), xgrad, pipe, torch.cat(dots, 0)

    res = run_ranks(world, rank_fn)
    assert torch.equal(torch.cat([r[2] for r in res]), want)          # rows keep the global edge order
    assert torch.equal(torch.cat([r[3] for r in res]), Xr.grad)
    pipe = torch.cat([r[4] for r in res])
    want_sum = dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)
    if chunks == 1:
        assert torch.equal(pipe, want_sum)
    else:
        torch.testing.assert_close(pipe, want_sum, rtol=1e-5, atol=1e-5)
    sel_sorted = torch.sort(torch.cat([r[5] for r in res]).view(-1)).values
    torch.testing.assert_close(sel_sorted, torch.sort(want_dot.view(-1)).values, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("drop", [0.0, 0.3])
def test_partitioned_fused_gat_matches_single_graph(cuda, world, drop):
    n, e, H, F = 1200, 30000, 4, 8
    src, dst = make_edges(n, n, e, seed=9, kind="powerlaw")
    src, dst = np.concatenate([src, np.arange(n)]), np.concatenate([dst, np.arange(n)])
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(cuda)
    torch.manual_seed(1)
    ft, el, er = torch.randn(n, H, F, device=cuda), torch.randn(n, H, device=cuda), torch.randn(n, H, device=cuda)
    gout = torch.randn(n, H, F, device=cuda)
    a, b, c = (t.clone().requires_grad_(True) for t in (ft, el, er))
    want = dgl.ops.gat_attention(g, a, b, c, 0.2, dropout_p=drop, seed=77)
    want.backward(gout)
    fake = FakeWorld(world)

    def rank_fn(r):
        part = RowPartition.build(src, dst, n, world, r, cuda)
        fake.attach(part)
        sl = slice(part.lo, part.hi)
        from dgl.distributed_rows import _PartitionedGAT as Fn
        x, l, rr = (t[sl].clone() for t in (ft, el, er))
        ctx = Ctx()
        out = Fn.forward(ctx, part, x, l, rr, 0.2, drop, 77)
        _, gx, gl, gr, _, _, _ = Fn.backward(ctx, gout[sl])
        return out.detach(), gx, gl, gr

    res = run_ranks(world, rank_fn)
    # forward / dst pass see each row's edges in the global order: identical sums; the src pass too
    assert torch.equal(torch.cat([r[0] for r in res]), want)
    assert torch.equal(torch.cat([r[3] for r in res]), c.grad)
    assert torch.equal(torch.cat([r[1] for r in res]), a.grad)
    assert torch.equal(torch.cat([r[2] for r in res]), b.grad)
