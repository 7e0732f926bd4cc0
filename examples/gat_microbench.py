"""Kernel-level timing of the GAT attention block: fused forward / backward kernels vs the op-by-op
composition upstream uses (u_add_v, leaky_relu, edge_softmax, u_mul_e_sum and their backward ops).
CUDA events, device-resident inputs, arxiv- / reddit- / products-shaped synthetic graphs."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import dgl  # noqa: E402
from dgl import sparse as K  # noqa: E402
from dgl.data import synthetic  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="ogbn-arxiv")
    ap.add_argument("--edges", type=int, default=0)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--feat", type=int, default=16)
    ap.add_argument("--degree", default="uniform")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--row-hub", default="", help="comma list of hub cut-offs (edges) to sweep for the FUSED kernels only")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, e, _, _ = synthetic.SHAPES[args.shape]
    e = args.edges or e
    src, dst = synthetic.random_edges(n, n, e, seed=0, degree=args.degree)
    src, dst = np.concatenate([src, np.arange(n)]), np.concatenate([dst, np.arange(n)])
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(dev)
    gi = g._graph
    H, F, E = args.heads, args.feat, len(src)
    ft = torch.randn(n, H, F, device=dev)
    el, er = torch.randn(n, H, device=dev), torch.randn(n, H, device=dev)
    dZ = torch.randn(n, H, F, device=dev)
    gi.csc(), gi.csr()
    res = {"shape": args.shape, "nodes": n, "edges": E, "H": H, "F": F, "degree": args.degree}
    for thr in [int(x) for x in args.row_hub.split(",") if x]:
        K.HUB_THRESHOLD = thr
        rst, mx, sm, _ = K._gat_fwd(gi, ft, el, er, 0.2, 0.0, 0)
        print(json.dumps({"row_hub": thr, "H": H, "F": F, "degree": args.degree, "shape": args.shape,
                          "fused_fwd_ms": timeit(lambda: K._gat_fwd(gi, ft, el, er, 0.2, 0.0, 0), args.reps),
                          "fused_bwd_ms": timeit(lambda: K._gat_bwd(gi, ft, el, er, mx, sm, dZ, 0.2, 0.0, 0), args.reps)}),
              flush=True)
    K.HUB_THRESHOLD = None
    rst, mx, sm, _ = K._gat_fwd(gi, ft, el, er, 0.2, 0.0, 0)
    res["fused_fwd_ms"] = timeit(lambda: K._gat_fwd(gi, ft, el, er, 0.2, 0.0, 0), args.reps)
    res["fused_bwd_ms"] = timeit(lambda: K._gat_bwd(gi, ft, el, er, mx, sm, dZ, 0.2, 0.0, 0), args.reps)
    res["fused_fwd_drop_ms"] = timeit(lambda: K._gat_fwd(gi, ft, el, er, 0.2, 0.2, 7), args.reps)
    # unfused pieces
    el3, er3 = el.view(n, H, 1), er.view(n, H, 1)
    t_add = timeit(lambda: K._gsddmm(gi, "add", el3, er3), args.reps)
    escore = torch.nn.functional.leaky_relu(K._gsddmm(gi, "add", el3, er3), 0.2)
    t_lrelu = timeit(lambda: torch.nn.functional.leaky_relu(escore, 0.2), args.reps)
    t_sm = timeit(lambda: K._edge_softmax_fwd(gi, escore), args.reps)
    a = K._edge_softmax_fwd(gi, escore)
    t_mul = timeit(lambda: K._gspmm(gi, "mul", "sum", ft, a), args.reps)
    res.update(unfused_u_add_v_ms=t_add, unfused_lrelu_ms=t_lrelu, unfused_edge_softmax_ms=t_sm, unfused_u_mul_e_sum_ms=t_mul)
    res["unfused_fwd_ms"] = t_add + t_lrelu + t_sm + t_mul
    gr = gi.reverse()
    t_b1 = timeit(lambda: K._gspmm(gr, "mul", "sum", dZ, a), args.reps)          # d ft
    t_b2 = timeit(lambda: K._gsddmm(gi, "dot", ft, dZ), args.reps)               # d a
    da = K._gsddmm(gi, "dot", ft, dZ)
    t_b3 = timeit(lambda: K._edge_softmax_bwd(gi, a, da), args.reps)
    t_b4 = timeit(lambda: K._gspmm(gr, "copy_rhs", "sum", None, da), args.reps)  # d el
    t_b5 = timeit(lambda: K._gspmm(gi, "copy_rhs", "sum", None, da), args.reps)  # d er
    res.update(unfused_bwd_dft_ms=t_b1, unfused_bwd_dot_ms=t_b2, unfused_bwd_softmax_ms=t_b3, unfused_bwd_del_ms=t_b4,
               unfused_bwd_der_ms=t_b5)
    res["unfused_bwd_ms"] = t_b1 + t_b2 + t_b3 + t_b4 + t_b5 + t_lrelu
    # gather-model bytes of the fused forward
    B = 4 * (n + 1) + 4 * E + 4 * H * E + 4 * H * F * E + 4 * H * n + 4 * H * F * n + 8 * H * n
    res["fused_fwd_algorithmic_gbs"] = B / (res["fused_fwd_ms"] * 1e-3) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
