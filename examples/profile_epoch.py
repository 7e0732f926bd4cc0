"""Kernel-level time table of one training epoch (torch.profiler, CUDA activities) for a config of epoch_bench.py:
which kernels the epoch's device time goes to (ours vs cuBLAS vs torch elementwise / indexing).

    python examples/profile_epoch.py --config products_sage --out profiles/r02_epoch_products_sage.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
    sys.path.insert(0, _p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import dgl  # noqa: E402
import epoch_bench  # noqa: E402
from examples.full_graph import GAT, GraphSAGE, synthetic_task  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="products_sage")
    ap.add_argument("--out", default="")
    ap.add_argument("--epochs", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    shape, kind, kw, _ = epoch_bench.CONFIGS[args.config]
    (n, src, dst), feats, labels, train_idx, n_classes = synthetic_task(shape, dev, self_loops=(kind == "gat"), edges=kw.get("edges"))
    graph = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n).int().to(dev)
    feats, labels, train_idx = feats.to(dev), labels.to(dev), train_idx.to(dev)
    torch.manual_seed(0)
    if kind == "sage":
        model = GraphSAGE(feats.shape[1], kw["hidden"], n_classes, kw["layers"], kw["aggr"], kw["dropout"]).to(dev)
    else:
        model = GAT(feats.shape[1], kw["hidden"], n_classes, kw["heads"], kw["dropout"], kw["dropout"]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=kw["lr"], weight_decay=kw["wd"])

    def step():
        model.train()
        opt.zero_grad()
        out = model(graph, feats)
        loss = (F.cross_entropy if kind == "sage" else F.nll_loss)(out[train_idx], labels[train_idx])
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.epochs):
            step()
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in ev)
    lines = ["# %s: device time per epoch %.3f ms over %d kernels/memcpys per epoch (torch.profiler, %d epochs)"
             % (args.config, total / args.epochs / 1e3, sum(e.count for e in ev) // args.epochs, args.epochs),
             "# share  ms/epoch  calls/epoch  kernel"]
    for e in ev[:40]:
        lines.append("%5.1f%%  %8.3f  %5d  %s" % (100 * e.device_time_total / total, e.device_time_total / args.epochs / 1e3,
                                                  e.count // args.epochs, e.key[:150]))
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
