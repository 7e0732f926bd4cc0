"""RGCN aggregation on an ogbn-proteins-shaped graph (132 534 nodes, 79 122 504 edges, 8 relations = edge-feature channels,
hidden 32): the per-relation loop of main_dgl_proteins_rgcn_for.py:50-53 (one update_all(u_mul_e, mean) per relation) vs
ONE relation-broadcast gspmm over all relations.  Forward and forward + backward (gradient w.r.t. the node features)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
    sys.path.insert(0, _p)
import torch  # noqa: E402

import dgl  # noqa: E402
from dgl.data import synthetic  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda", 0)
    scale = float(os.environ.get("DGLB200_DATA_SCALE", "1"))
    n, e, _, _ = synthetic.SHAPES["ogbn-proteins"]
    e = int(e * scale)
    src, dst = synthetic.random_edges(n, n, e, seed=0)
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(dev)
    R, D = 8, 32
    W = torch.rand(e, R, device=dev)
    cols = [W[:, r:r + 1] for r in range(R)]                   # what the script passes: strided (E,1) views
    x = torch.randn(n, D, device=dev, requires_grad=True)
    gout = torch.randn(n, R, D, device=dev)

    def loop_fwd():
        return [dgl.ops.gspmm(g, "mul", "mean", x, c) for c in cols]

    def batched_fwd():
        return dgl.ops.gspmm(g, "mul", "mean", x.unsqueeze(1), W.unsqueeze(-1))

    def loop_fb():
        x.grad = None
        torch.stack(loop_fwd(), 1).backward(gout)

    def batched_fb():
        x.grad = None
        batched_fwd().backward(gout)

    a = torch.stack(loop_fwd(), 1)
    b = batched_fwd()
    res = {"shape": "ogbn-proteins", "nodes": n, "edges": e, "relations": R, "hidden": D,
           "bit_identical": bool(torch.equal(a, b)),
           "fwd_ms_loop": timeit(loop_fwd), "fwd_ms_batched": timeit(batched_fwd),
           "fwd_bwd_ms_loop": timeit(loop_fb), "fwd_bwd_ms_batched": timeit(batched_fb)}
    gather_bytes = e * (4 + 4 + 4 * D + 4 * R) + n * (4 + 4 * R * D)
    res["fwd_batched_algorithmic_gbs"] = gather_bytes / res["fwd_ms_batched"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
