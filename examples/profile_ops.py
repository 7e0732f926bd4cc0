"""One launch of each secondary op (for `ncu -k regex:...`): copy_u_sum (the yardstick), copy_u_max,
u_mul_e_sum with (E,1) weights, edge_softmax fwd/bwd, fused GAT fwd/bwd -- on a synthetic graph of a named shape."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dgl-0.5-benchmark_b200"))
import dgl  # noqa: E402
from dgl import sparse as K  # noqa: E402
from dgl.data import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="ogbn-products")
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--degree", default="uniform")
    ap.add_argument("--order", default="shuffled")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, e, _, _ = synthetic.SHAPES[args.shape]
    src, dst = synthetic.random_edges(n, n, e, seed=0, degree=args.degree, order=args.order)
    g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(dev)
    D = args.width
    X = torch.rand(n, D, device=dev)
    W = torch.rand(e, 1, device=dev)
    z = torch.randn(e, args.heads, device=dev)
    gr = torch.randn(e, args.heads, device=dev)
    dgl.ops.gspmm(g, "copy_lhs", "sum", X, None)
    dgl.ops.gspmm(g, "copy_lhs", "max", X, None)
    dgl.ops.gspmm(g, "mul", "sum", X, W)
    a = K._edge_softmax_fwd(g._graph, z)
    K._edge_softmax_bwd(g._graph, a, gr)
    # fused GAT attention (H heads x 16), forward and both backward passes
    H = args.heads
    ft = torch.randn(n, H, 16, device=dev)
    el, er = torch.randn(n, H, device=dev), torch.randn(n, H, device=dev)
    dZ = torch.randn(n, H, 16, device=dev)
    rst, mx, sm, _ = K._gat_fwd(g._graph, ft, el, er, 0.2, 0.0, 0)
    K._gat_bwd(g._graph, ft, el, er, mx, sm, dZ, 0.2, 0.0, 0)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
