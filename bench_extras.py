"""Extra blocks of bench.py's JSON line (kept out of bench.py's timed `value` region):

  parity_block      the GPU results of the benchmark's own ops at the benchmark's own size compared with
                    the full-size CPU oracle on the same seeded inputs (bit-exactness / abs-sum-scaled error /
                    arg-max equality) -- the one place besides cpu_baseline where bench.py executes oracle/;
  secondary_kernels algorithmic-byte roofline fractions of the ops the headline sweep does not time
                    (u_add_v, copy_u_max, u_mul_e, edge_softmax fwd/bwd, fused GAT fwd/bwd), each timed by CUDA
                    events around the public API call on the launching stream;
  epochs_block      full-graph SAGE / GAT epoch times on the arxiv and products shapes (BASELINE.json
                    configs[2] and configs[3]; second half of `metric`), 1 GPU or row-partitioned.

Algorithmic bytes follow SURVEY.md section 8(d) (gather model, fp32, ids 4 B, p = 1 when the CSC carries an
edge-id permutation); DESIGN.md section 5 restates them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "dgl-0.5-benchmark_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


# ------------------------------------------------------------------ algorithmic bytes
def spmm_bytes(n_dst, n_edges, D, s=4):
    return 4 * (n_dst + 1) + 4 * n_edges + s * D * n_edges + s * D * n_dst


def spmm_max_bytes(n_dst, n_edges, D):
    return spmm_bytes(n_dst, n_edges, D) + 4 * D * n_dst            # + the one arg output upstream records


def u_mul_e_bytes(n_dst, n_edges, D, De, p):
    return spmm_bytes(n_dst, n_edges, D) + 4 * De * n_edges + 4 * p * n_edges


def sddmm_dot_bytes(n_dst, n_edges, D, s=4, p=1):
    return 4 * (n_dst + 1) + 4 * n_edges + 4 * p * n_edges + s * D * n_edges + s * D * n_dst + s * n_edges


def u_add_v_bytes(n_dst, n_edges, D, p):
    return 4 * (n_dst + 1) + 4 * n_edges + 4 * p * n_edges + 4 * D * n_edges + 4 * D * n_dst + 4 * D * n_edges


def edge_softmax_bytes(n_dst, n_edges, H, p, bwd=False):
    return 4 * (n_dst + 1) + 4 * p * n_edges + (3 if bwd else 2) * 4 * H * n_edges


def gat_fwd_bytes(n_dst, n_edges, H, F):
    return (4 * (n_dst + 1) + 4 * n_edges + 4 * H * n_edges + 4 * H * F * n_edges
            + 4 * H * n_dst + 4 * H * F * n_dst + 2 * 4 * H * n_dst)


def gat_bwd_bytes(n_dst, n_src, n_edges, H, F):
    dst_pass = (4 * (n_dst + 1) + 4 * n_edges + n_edges * (4 * H * F + 4 * H)
                + n_dst * (4 * H * F + 4 * H + 8 * H + 16 * H + 4 * H))
    src_pass = (4 * (n_src + 1) + 4 * n_edges + n_edges * (4 * H * F + 16 * H)
                + n_src * (4 * H * F + 4 * H + 4 * H * F + 4 * H))
    return dst_pass + src_pass


# ------------------------------------------------------------------ secondary kernels
def _time(fn, reps=5, warm=2):
    """ms per call, CUDA events around `reps` back-to-back calls (th_op_time semantics, kernel/utils.py:18-34).  A
    sub-0.3 ms op can be bound by the Python dispatch of the call rather than by its kernel (the GPU idles between
    launches), so such ops are ALSO timed as one CUDA-graph replay of the same `reps` calls and the smaller figure is
    reported: the roofline fraction is a statement about the kernel, not about the interpreter."""
    import torch
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    if ms < 0.3:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(reps):
                    fn()
            graph.replay()
            torch.cuda.synchronize()
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ms = min(ms, a.elapsed_time(b) / reps)
            del graph
        except Exception:   # an op that cannot be captured keeps its eager figure
            torch.cuda.synchronize()
    return ms


def secondary_kernels(src, dst, n_nodes, dev, peak, reps=5):
    """Times the ops outside the headline sweep on the reddit-shaped graph, in the benchmark's shuffled edge
    order and (per-edge-tensor ops) in dst-sorted order, where the edge-id permutation is the identity."""
    import torch
    import dgl
    from dgl import backend as B
    E = len(src)
    N = n_nodes
    out = []

    def add(op, shape, order, ms, nbytes, **kw):
        d = {"op": op, "shape": shape, "edge_order": order, "ms": ms, "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9}
        d["frac_of_peak"] = d["algorithmic_gbs"] / peak
        d.update(kw)
        out.append(d)

    o = np.argsort(dst, kind="stable")
    graphs = {"shuffled": (src, dst), "dst_sorted": (src[o], dst[o])}
    for order, (s_, d_) in graphs.items():
        p = 1 if order == "shuffled" else 0
        g = dgl.graph((torch.from_numpy(s_), torch.from_numpy(d_)), num_nodes=N).int().to(dev)
        gi = g._graph
        with torch.no_grad():
            if order == "shuffled":
                for D in (64, 602):
                    X = torch.rand(N, D, device=dev)
                    add("gspmm_copy_u_max", "D=%d" % D, order,
                        _time(lambda: dgl.ops.gspmm(g, "copy_lhs", "max", X, None), reps), spmm_max_bytes(N, E, D))
                    del X
            for D in (64,):
                X = torch.rand(N, D, device=dev)
                V = torch.rand(N, D, device=dev)
                add("gsddmm_u_add_v", "D=%d" % D, order,
                    _time(lambda: dgl.ops.gsddmm(g, "add", X, V), reps), u_add_v_bytes(N, E, D, p))
                W1 = torch.rand(E, 1, device=dev)
                add("gspmm_u_mul_e_sum", "D=%d x (E,1)" % D, order,
                    _time(lambda: dgl.ops.gspmm(g, "mul", "sum", X, W1), reps), u_mul_e_bytes(N, E, D, 1, p))
                del X, V, W1
            for H, F in ((1, 16), (4, 16)):
                el = torch.randn(N, H, 1, device=dev)
                er = torch.randn(N, H, 1, device=dev)
                add("gsddmm_u_add_v", "(N,%d,1)" % H, order,
                    _time(lambda: dgl.ops.gsddmm(g, "add", el, er), reps), u_add_v_bytes(N, E, H, p))
                ft = torch.randn(N, H, F, device=dev)
                a = torch.rand(E, H, 1, device=dev)
                add("gspmm_u_mul_e_sum", "(N,%d,%d) x (E,%d,1)" % (H, F, H), order,
                    _time(lambda: dgl.ops.gspmm(g, "mul", "sum", ft, a), reps), u_mul_e_bytes(N, E, H * F, H, p))
                del a
                logits = torch.randn(E, H, 1, device=dev)
                sm = dgl.ops.edge_softmax(g, logits)
                add("edge_softmax_fwd", "H=%d" % H, order,
                    _time(lambda: dgl.ops.edge_softmax(g, logits), reps), edge_softmax_bytes(N, E, H, p))
                gr = torch.randn(E, H, 1, device=dev)
                from dgl import sparse as K
                add("edge_softmax_bwd", "H=%d" % H, order,
                    _time(lambda: K._edge_softmax_bwd(gi, sm, gr), reps), edge_softmax_bytes(N, E, H, p, bwd=True))
                del logits, sm, gr
                if order == "shuffled":       # the fused kernels never touch a per-edge tensor: order-independent
                    el2, er2 = el.view(N, H), er.view(N, H)
                    rst, rmax, rsum, _ = K._gat_fwd(gi, ft, el2, er2, 0.2, 0.0, 0)
                    add("gat_fused_fwd", "(H,F)=(%d,%d)" % (H, F), order,
                        _time(lambda: K._gat_fwd(gi, ft, el2, er2, 0.2, 0.0, 0), reps), gat_fwd_bytes(N, E, H, F))
                    dz = torch.randn(N, H, F, device=dev)
                    gi.csr()
                    add("gat_fused_bwd", "(H,F)=(%d,%d)" % (H, F), order,
                        _time(lambda: K._gat_bwd(gi, ft, el2, er2, rmax, rsum, dz, 0.2, 0.0, 0), reps),
                        gat_bwd_bytes(N, N, E, H, F))
                    del rst, rmax, rsum, dz
                del el, er, ft
        del g, gi
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------ parity at the benchmark's own size
def parity_block(src, dst, n_nodes, g, dev, widths=(64, 602)):
    """GPU vs full-size CPU oracle on the same seeded host inputs, through the public API.
    gspmm copy_u_sum and copy_u_max (+ arg_u), gsddmm u_dot_v.  Scaled error = |gpu - cpu| / sum|terms| (inputs are
    the micro-benchmark's non-negative U[0,1) features, so sum|terms| is the result itself)."""
    import torch
    import dgl
    from dgl import sparse as K
    from oracle import dgl_ref as R
    R.build()
    R.set_num_threads(os.cpu_count() or 1)
    og = R.OracleGraph(src, dst, n_nodes, n_nodes)
    csc = g._graph.csc()
    structure_equal = bool(np.array_equal(csc.indptr.cpu().numpy(), og.csc[0])
                           and np.array_equal(csc.indices.cpu().numpy(), og.csc[1])
                           and (csc.eids is None or np.array_equal(csc.eids.cpu().numpy(), og.csc[2])))
    checks, worst, arg_equal, bit_exact_spmm = [], 0.0, True, True
    rng = np.random.default_rng(1234)
    for D in widths:
        X = rng.random((n_nodes, D), dtype=np.float32)
        V = rng.random((n_nodes, D), dtype=np.float32)
        Xt, Vt = torch.from_numpy(X).to(dev), torch.from_numpy(V).to(dev)
        with torch.no_grad():
            got = dgl.ops.gspmm(g, "copy_lhs", "sum", Xt, None).cpu().numpy()
        want = R.gspmm(og, "copy_lhs", "sum", X, None)
        err = float(np.max(np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)))
        exact = bool(np.array_equal(got, want))
        bit_exact_spmm &= exact
        worst = max(worst, err)
        checks.append({"op": "gspmm_copy_u_sum", "D": D, "max_err_scaled": err, "bit_exact": exact})
        del got, want
        with torch.no_grad():
            got = dgl.ops.gsddmm(g, "dot", Xt, Vt).cpu().numpy()
        want = R.gsddmm(og, "dot", X, V)
        err = float(np.max(np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)))
        worst = max(worst, err)
        checks.append({"op": "gsddmm_u_dot_v", "D": D, "max_err_scaled": err, "bit_exact": bool(np.array_equal(got, want))})
        del got, want
        if D == widths[0]:
            with torch.no_grad():
                gmax, (gau, _) = K._gspmm(g._graph, "copy_lhs", "max", Xt, None)
            wmax, (wau, _) = R.gspmm_with_args(og, "copy_lhs", "max", X, None)
            eq = bool(np.array_equal(gau.cpu().numpy(), wau)) and bool(np.array_equal(gmax.cpu().numpy(), wmax))
            arg_equal &= eq
            checks.append({"op": "gspmm_copy_u_max", "D": D, "values_bit_exact": bool(np.array_equal(gmax.cpu().numpy(), wmax)),
                           "argmax_equal": bool(np.array_equal(gau.cpu().numpy(), wau))})
            del gmax, gau, wmax, wau
        del Xt, Vt, X, V
    return {"against": "oracle/dgl_cpu_oracle.c (C restatement of DGL v0.6.1 SpMMSumCsr / SpMMCmpCsr / SDDMMCoo; parity "
                       "unpinned: no reference-held vectors exist), full reddit-shaped graph, all rows / all edges",
            "max_err_scaled": worst, "tolerance": 1e-5, "within_tolerance": bool(worst <= 1e-5),
            "gspmm_sum_bit_exact": bit_exact_spmm, "argmax_equal": arg_equal, "csc_structure_equal": structure_equal,
            "checks": checks}


# ------------------------------------------------------------------ epochs (BASELINE configs[2], configs[3])
def epochs_block(rank, world, dev, epochs=9, configs=None):
    """seconds/epoch of full-graph training on synthetic graphs of the named shapes (3 warm-up epochs skipped,
    device synchronised, max over ranks).  N = 1: arxiv GAT, products SAGE, products GAT; N > 1: the two
    products configs, row-partitioned."""
    import torch
    import epoch_bench
    if configs is None:
        configs = (["arxiv_gat"] if world == 1 else []) + ["products_sage", "products_gat"]
    res = {}
    for name in configs:
        r = epoch_bench.run_config(name, epochs, rank, world, dev, "uniform")
        res[name] = {k: r[k] for k in ("epoch_s", "epoch_s_min", "epochs_timed", "nodes", "edges", "n_gpus",
                                       "sparse_launches_per_epoch", "v100_dgl_epoch_s_published", "loss_first",
                                       "loss_last", "step_launch")}
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ batched small graphs (BASELINE configs[4])
def small_graph_block(dev, batch_size=64, iters=40, n_graphs=512):
    """ms per training iteration of the molhiv-shaped GCN (main_dgl_molhiv_gcn.py:20-115; emb 256, 5 layers) at batch 64 on
    synthetic molecule-like graphs: (a) the script's formulation -- host dgl.batch + H2D + COO->CSC per iteration, Python
    message UDF, eager launches; (b) the whole iteration (device-side batch construction, fused message / encoder kernels,
    loss, backward, Adam) replayed as ONE CUDA graph, `batch_size` graph ids copied from pinned host memory per step."""
    import time
    import torch
    import torch.nn.functional as F
    import dgl
    from examples.molhiv_bench import captured_runner
    from examples.small_graph_model import GCN
    from ogb.graphproppred import DglGraphPropPredDataset
    ds = DglGraphPropPredDataset("ogbg-molhiv", num_graphs=n_graphs)
    samples = [ds[i] for i in range(n_graphs)]
    nb = n_graphs // batch_size
    model = GCN().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def eager(i):
        lo = (i % nb) * batch_size
        g = dgl.batch([s[0] for s in samples[lo:lo + batch_size]]).to(dev).int().formats("coo")
        y = torch.stack([s[1] for s in samples[lo:lo + batch_size]]).to(dev)
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(model(g, g.ndata["feat"], g.edata["feat"]), y)
        loss.backward()
        opt.step()

    for i in range(5):
        eager(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        eager(i)
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) / iters * 1e3
    store = dgl.GraphStore([s[0] for s in samples], torch.stack([s[1] for s in samples]), device=dev)
    id_batches = [np.arange(j * batch_size, (j + 1) * batch_size) for j in range(nb)]
    pinned = [torch.from_numpy(b.astype(np.int32)).pin_memory() for b in id_batches]
    run, sb = captured_runner(GCN(fused=True).to(dev), store, id_batches, batch_size)
    for i in range(5):
        run(pinned[i % nb])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        loss = run(pinned[i % nb])
    torch.cuda.synchronize()
    graph_ms = (time.perf_counter() - t0) / iters * 1e3
    iters_per_epoch = -(-32901 // batch_size)
    return {"config": "molhiv-shaped GCN (emb 256, 5 layers), batch %d, synthetic molecule-like graphs" % batch_size,
            "nodes_per_batch": int(store.n_nodes_host[:batch_size].sum()), "edges_per_batch": int(store.n_edges_host[:batch_size].sum()),
            "ms_per_iter_reference_formulation_eager": eager_ms,
            "ms_per_iter_cuda_graph_with_device_batching": graph_ms,
            "epoch_s_cuda_graph_with_device_batching": graph_ms * iters_per_epoch / 1e3,
            "epoch_s_reference_formulation_eager": eager_ms * iters_per_epoch / 1e3,
            "v100_dgl_epoch_s_published": 15.089, "final_loss": float(loss.detach()),
            "padded_nodes": sb.n_nodes_pad, "padded_edges": sb.n_edges_pad}
