"""A/B of the bulk-copy ring kernels (csrc/ring.cu) against the register-staged row kernels on one graph shape:
checks that both paths agree (gspmm: bit for bit; u_dot_v: 1e-5 sum-scaled) and prints ms / algorithmic GB/s.

    python examples/ring_tune.py --shape reddit --widths 256,602 --stages 0,3,4,5,6 --smem 100,110
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
    sys.path.insert(0, _p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dgl  # noqa: E402
from dgl.data import synthetic  # noqa: E402
from bench_extras import spmm_bytes, sddmm_dot_bytes  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--degree", default="uniform")
    ap.add_argument("--widths", default="256,602")
    ap.add_argument("--stages", default="0")
    ap.add_argument("--smem", default="100")
    ap.add_argument("--dtypes", default="f32")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--hub", type=int, default=0, help="hub-row cut-off override (0 = library default)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    if args.hub:
        from dgl import sparse as K
        K.HUB_THRESHOLD = args.hub
    n, src, dst = synthetic.shaped_edges(args.shape, degree=args.degree)
    E = len(src)
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(dev)
    peak = 6454.6
    for dt in args.dtypes.split(","):
        tdt = torch.float32 if dt == "f32" else torch.bfloat16
        s = 4 if dt == "f32" else 2
        for D in [int(x) for x in args.widths.split(",")]:
            X = torch.rand(n, D, device=dev).to(tdt)
            V = torch.rand(n, D, device=dev).to(tdt)
            os.environ["DGLB_RING_MIN_BYTES"] = str(1 << 30)          # ring off
            with torch.no_grad():
                ref_o = dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)
                ref_d = dgl.ops.gsddmm(g, "dot", X, V)
                t_o = timeit(lambda: dgl.ops.gspmm(g, "copy_lhs", "sum", X, None), args.reps)
                t_d = timeit(lambda: dgl.ops.gsddmm(g, "dot", X, V), args.reps)
            print(json.dumps({"D": D, "dtype": dt, "path": "rows", "spmm_ms": t_o, "spmm_frac": spmm_bytes(n, E, D, s) / t_o / 1e6 / peak,
                              "dot_ms": t_d, "dot_frac": sddmm_dot_bytes(n, E, D, s) / t_d / 1e6 / peak}), flush=True)
            os.environ["DGLB_RING_MIN_BYTES"] = "64"
            for smem in [int(x) for x in args.smem.split(",")]:
                os.environ["DGLB_RING_SMEM"] = str(smem * 1024)
                for S in [int(x) for x in args.stages.split(",")]:
                    os.environ["DGLB_RING_STAGES"] = str(S)
                    with torch.no_grad():
                        try:
                            o = dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)
                            d = dgl.ops.gsddmm(g, "dot", X, V)
                            torch.cuda.synchronize()
                        except Exception as ex:  # noqa: BLE001
                            print(json.dumps({"D": D, "S": S, "smem_kb": smem, "error": str(ex)[:200]}), flush=True)
                            continue
                        same = bool(torch.equal(o, ref_o)) if dt == "f32" else bool(torch.allclose(o.float(), ref_o.float(), rtol=2 ** -7))
                        derr = float(((d.float() - ref_d.float()).abs() / ref_d.float().abs().clamp(min=1e-30)).max())
                        t_o = timeit(lambda: dgl.ops.gspmm(g, "copy_lhs", "sum", X, None), args.reps)
                        t_d = timeit(lambda: dgl.ops.gsddmm(g, "dot", X, V), args.reps)
                    print(json.dumps({"D": D, "dtype": dt, "path": "ring", "S": S, "smem_kb": smem, "spmm_equal": same,
                                      "dot_max_rel": derr, "spmm_ms": t_o, "spmm_frac": spmm_bytes(n, E, D, s) / t_o / 1e6 / peak,
                                      "dot_ms": t_d, "dot_frac": sddmm_dot_bytes(n, E, D, s) / t_d / 1e6 / peak}), flush=True)
            del X, V


if __name__ == "__main__":
    main()
