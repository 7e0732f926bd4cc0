"""-m gpu: the UNCHANGED reference scripts (staged byte for byte under baseline/_ref by
baseline/stage_reference.py; /root/reference does not exist on the GPU box) run on cuda:0 through the
drop-in `dgl` package, i.e. through the C-ABI into the sm_100a kernels -- there is no other backend on a
CUDA tensor (dgl/sparse.py raises on CPU tensors, dgl/_capi.py raises when the library is missing).

Covers north_star's "kernel/dgl-new.py and the end_to_end/full_graph call sites are unchanged":
kernel/dgl-new.py:10-46,56-70, main_dgl_citation_sage.py, main_dgl_arxiv_gat.py, main_dgl_product_sage.py,
main_dgl_molhiv_gcn.py, main_dgl_proteins_rgcn_for.py (+ the citation GAT and the nn.SAGEConv variants).
Datasets are the seeded synthetic stand-ins of each dataset's shape, scaled down by DGLB200_DATA_SCALE so
the whole file runs in about a minute; profiles/ holds the full-size runs.
"""
import os
import re
import subprocess
import sys

import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref")
NC = "end_to_end/full_graph/node_classification/"
GC = "end_to_end/full_graph/graph_classification/"


@pytest.fixture(scope="module")
def staged(cuda):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import stage_reference
    finally:
        sys.path.pop(0)
    if not os.path.isdir(REF):
        if stage_reference.stage() is None:
            pytest.fail("baseline/_ref is missing: run `python baseline/stage_reference.py` (or "
                        "__graft_entry__.build()) where /root/reference exists, before shipping to the GPU box")
    assert stage_reference.verify(), "baseline/_ref differs from the manifest written when it was staged"
    return REF


def run_script(rel, argv, scale, timeout=600):
    env = dict(os.environ, DGLB200_DATA_SCALE=str(scale), PYTHONUNBUFFERED="1")
    cmd = [sys.executable, os.path.join(PKG, "run_reference.py"), os.path.join(REF, rel)] + argv
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0, "%s failed:\n%s\n%s" % (rel, r.stdout[-3000:], r.stderr[-3000:])
    m = re.search(r"C-ABI compute calls: (\d+) \(library (.*)\)", r.stderr)
    assert m, r.stderr[-2000:]
    assert "libdglb200.so via " in m.group(2) and m.group(2).endswith("libdglb200_torch.so"), m.group(2)
    return r.stdout, int(m.group(1))


def test_kernel_microbench_unchanged_on_gpu(staged):
    """kernel/dgl-new.py -g 0: gspmm(copy_lhs,sum) and gsddmm(add) for hidden 1..128 on reddit / arxiv /
    proteins shaped graphs, timed by the script's own th_op_time (CUDA events)."""
    out, calls = run_script("kernel/dgl-new.py", ["-g", "0"], scale=0.05)
    assert "OOM" not in out, out                      # the script's bare except prints OOM on ANY error
    assert out.count("SPMM") == 3 and out.count("SDDMM") == 3, out
    assert out.count("avg time") == 3 * 2 * 8, out    # 3 graphs x (spmm, sddmm) x 8 hidden sizes
    times = [float(x) for x in re.findall(r"avg time: ([0-9.eE+-]+)", out)]
    assert all(0 < t < 1e3 for t in times), times
    assert calls >= 3 * 2 * 8 * 10, calls             # 10 reps each, every one through the C-ABI


@pytest.mark.parametrize("binary,reduce_op,sddmm", [("mul", "sum", "dot"), ("copy_lhs", "max", "mul"),
                                                   ("add", "mean", "sub"), ("copy_rhs", "min", "div")])
def test_kernel_microbench_other_ops_on_gpu(staged, binary, reduce_op, sddmm):
    out, calls = run_script("kernel/dgl-new.py", ["-g", "0", "--spmm-binary", binary, "--spmm-reduce", reduce_op,
                                                  "--sddmm-binary", sddmm], scale=0.01)
    assert "OOM" not in out and out.count("avg time") == 48, out
    assert calls >= 480


def test_citation_sage_unchanged_on_gpu(staged):
    out, calls = run_script(NC + "main_dgl_citation_sage.py", ["--dataset", "cora", "--epochs", "12", "--runs", "1", "--eval"], 1)
    assert "Training time/epoch" in out and "Final Test" in out, out[-2000:]
    assert calls >= 12 * 3                            # 2 forward + 1 backward aggregation per epoch


def test_citation_gat_unchanged_on_gpu(staged):
    out, calls = run_script(NC + "main_dgl_citation_gat.py", ["--dataset", "cora", "--epochs", "8", "--runs", "1"], 1)
    assert "Training time/epoch" in out, out[-2000:]
    assert calls >= 8 * 3


def test_arxiv_gat_unchanged_on_gpu(staged):
    out, calls = run_script(NC + "main_dgl_arxiv_gat.py", ["--epochs", "8", "--runs", "1", "--eval"], 0.25)
    assert "Training time/epoch" in out and "Test:" in out, out[-2000:]
    assert calls >= 8 * 3 * 3                         # fused forward + two backward passes per layer


def test_product_sage_unchanged_on_gpu(staged):
    out, calls = run_script(NC + "main_dgl_product_sage.py", ["--epochs", "6", "--runs", "1"], 0.05)
    assert "Training time/epoch" in out, out[-2000:]
    assert calls >= 6 * 5


def test_reddit_sage_nn_unchanged_on_gpu(staged):
    out, calls = run_script(NC + "main_dgl_reddit_sage_nn.py", ["--dataset", "reddit", "--epochs", "6", "--runs", "1"], 0.05)
    assert "Training time/epoch" in out, out[-2000:]
    assert calls >= 6 * 3


def test_proteins_rgcn_unchanged_on_gpu(staged):
    """main_dgl_proteins_rgcn_for.py:46-60: one update_all(u_mul_e, mean) with (E,1) weights per relation
    (8 relations) per layer (3 layers), forward and backward."""
    out, calls = run_script(NC + "main_dgl_proteins_rgcn_for.py", ["--epochs", "5", "--runs", "1", "--eval"], 0.02)
    assert "Training time/epoch" in out and "Test:" in out, out[-2000:]
    assert calls >= 5 * 8 * 3


def test_molhiv_gcn_unchanged_on_gpu(staged):
    """main_dgl_molhiv_gcn.py:95-115: batched COO-only graphs, UDF message + builtin sum."""
    out, calls = run_script(GC + "main_dgl_molhiv_gcn.py", ["--epochs", "3", "--runs", "1", "--num_workers", "0", "--eval"], 0.02)
    assert "Training time/epoch" in out and "Valid:" in out, out[-2000:]
    assert calls >= 3 * 5


def test_enzymes_gcn_unchanged_on_gpu(staged):
    out, calls = run_script(GC + "main_dgl_enzymes_gcn.py", ["--epochs", "3", "--runs", "1", "--num_workers", "0"], 1)
    assert "Training time/epoch" in out, out[-2000:]
    assert calls >= 3
