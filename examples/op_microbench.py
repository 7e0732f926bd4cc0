"""Per-op timing (CUDA events) of gspmm / gsddmm on a synthetic graph of a named dataset shape:
algorithmic GB/s (SURVEY.md 8d gather model) and fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import dgl  # noqa: E402
from dgl.data import synthetic  # noqa: E402
from bench import spmm_bytes, sddmm_dot_bytes, measured_peak  # noqa: E402


def timeit(fn, reps=10, warm=3):
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
    ap.add_argument("--shape", default="ogbn-products")
    ap.add_argument("--widths", default="64,100")
    ap.add_argument("--degree", default="uniform")
    ap.add_argument("--order", default="shuffled")
    ap.add_argument("--softmax-heads", default="", help="comma list: also time edge_softmax fwd/bwd with H heads")
    ap.add_argument("--hub-bytes", type=int, default=0, help="override: hub threshold = hub_bytes / (4*D)")
    ap.add_argument("--stage-min-mb", type=int, default=-1, help="override dgl.sparse.STAGE_MIN_BYTES (MB); huge = staging off")
    ap.add_argument("--heads", default="", help="comma list H: also time u_add_v (N,H,1), u_mul_e (N,H,16)x(E,H,1), u_dot_v (N,H,16)")
    ap.add_argument("--softmax-hub", default="", help="comma list of edge_softmax hub cut-offs (edges) to sweep; 0 = default")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, e, _, _ = synthetic.SHAPES[args.shape]
    src, dst = synthetic.random_edges(n, n, e, seed=0, degree=args.degree, order=args.order)
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(dev)
    peak, _ = measured_peak()
    p = 0 if args.order == "dst_sorted" else 1
    from dgl import sparse as K
    if args.stage_min_mb >= 0:
        K.STAGE_MIN_BYTES = args.stage_min_mb << 20
    for H in [int(x) for x in args.heads.split(",") if x]:
        F = 16
        el, er = torch.randn(n, H, 1, device=dev), torch.randn(n, H, 1, device=dev)
        ft, a = torch.randn(n, H, F, device=dev), torch.rand(e, H, 1, device=dev)
        res = {"shape": args.shape, "edges": e, "H": H, "F": F, "order": args.order, "stage_min_mb": args.stage_min_mb}
        for name, fn, B in (
                ("u_add_v(N,H,1)", lambda: dgl.ops.gsddmm(g, "add", el, er), 4 * (n + 1) + 4 * e + 4 * p * e + 8 * H * e + 4 * H * n),
                ("u_mul_e_sum(N,H,F)x(E,H,1)", lambda: dgl.ops.gspmm(g, "mul", "sum", ft, a), spmm_bytes(n, e, H * F) + 4 * H * e + 4 * p * e),
                ("u_dot_v(N,H,F)", lambda: dgl.ops.gsddmm(g, "dot", ft, ft), 4 * (n + 1) + 8 * e * p + 4 * e + 4 * H * F * (e + n) + 4 * H * e)):
            ms = timeit(fn, reps=5)
            res[name] = {"ms": round(ms, 4), "gbs": round(B / ms / 1e6), "frac": round(B / ms / 1e6 / peak, 3)}
        print(json.dumps(res), flush=True)
        del el, er, ft, a
    for D in [int(x) for x in args.widths.split(",") if x]:
        K.HUB_THRESHOLD = max(32, args.hub_bytes // (4 * D)) if args.hub_bytes else None
        X, V = torch.rand(n, D, device=dev), torch.rand(n, D, device=dev)
        W = torch.rand(e, 1, device=dev)
        res = {"shape": args.shape, "nodes": n, "edges": e, "D": D, "degree": args.degree, "order": args.order,
               "hub_threshold": K.HUB_THRESHOLD}
        for name, fn, B in (
                ("copy_u_sum", lambda: dgl.ops.gspmm(g, "copy_lhs", "sum", X, None), spmm_bytes(n, e, D)),
                ("copy_u_mean", lambda: dgl.ops.gspmm(g, "copy_lhs", "mean", X, None), spmm_bytes(n, e, D)),
                ("copy_u_max", lambda: dgl.ops.gspmm(g, "copy_lhs", "max", X, None), spmm_bytes(n, e, D) + 8 * D * n),
                ("u_mul_e_sum(E,1)", lambda: dgl.ops.gspmm(g, "mul", "sum", X, W), spmm_bytes(n, e, D) + 4 * e + 4 * p * e),
                ("u_dot_v", lambda: dgl.ops.gsddmm(g, "dot", X, V), sddmm_dot_bytes(n, e, D, p=p))):
            ms = timeit(fn)
            res[name] = {"ms": round(ms, 4), "gbs": round(B / ms / 1e6), "frac": round(B / ms / 1e6 / peak, 3)}
        print(json.dumps(res), flush=True)
    if args.softmax_heads:
        from dgl import sparse as K2
        for H, thr in [(int(x), int(y)) for x in args.softmax_heads.split(",") for y in (args.softmax_hub or "0").split(",")]:
            K2.HUB_THRESHOLD = thr or None
            z = torch.randn(e, H, device=dev)
            a = K2._edge_softmax_fwd(g._graph, z)
            gr = torch.randn(e, H, device=dev)
            Bf = 4 * (n + 1) + 4 * p * e + 2 * 4 * H * e
            Bb = 4 * (n + 1) + 4 * p * e + 3 * 4 * H * e
            tf = timeit(lambda: K2._edge_softmax_fwd(g._graph, z))
            tb = timeit(lambda: K2._edge_softmax_bwd(g._graph, a, gr))
            print(json.dumps({"shape": args.shape, "edges": e, "H": H, "order": args.order, "degree": args.degree, "hub": thr,
                              "edge_softmax_fwd": {"ms": round(tf, 4), "gbs": round(Bf / tf / 1e6), "frac": round(Bf / tf / 1e6 / peak, 3)},
                              "edge_softmax_bwd": {"ms": round(tb, 4), "gbs": round(Bb / tb / 1e6), "frac": round(Bb / tb / 1e6 / peak, 3)}}),
                  flush=True)


if __name__ == "__main__":
    main()
