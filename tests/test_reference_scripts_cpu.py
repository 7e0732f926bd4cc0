"""The UNCHANGED reference scripts run against the dgl / ogb stand-ins (API-surface check of the
drop-in boundary).  This container has no GPU, so the kernel-level calls are
routed to the CPU oracle by tests/oracle_backend.py (test infrastructure); on the GPU box
/root/reference does not exist and these tests skip."""
import contextlib
import io
import os
import runpy
import sys

import pytest

from conftest import PKG
import oracle_backend

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _run(script, argv, scale):
    script = os.path.join(REF, script)
    old_argv, old_path, old_env = sys.argv, list(sys.path), os.environ.get("DGLB200_DATA_SCALE")
    old_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")}
    os.environ["DGLB200_DATA_SCALE"] = str(scale)
    sys.argv = [script] + argv
    sys.path.insert(0, os.path.dirname(script))
    sys.path.insert(0, PKG)
    buf = io.StringIO()
    try:
        with oracle_backend.installed(), contextlib.redirect_stdout(buf):
            runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
        sys.modules.pop("utils", None)
        sys.modules.update(old_mods)
        if old_env is None:
            os.environ.pop("DGLB200_DATA_SCALE", None)
        else:
            os.environ["DGLB200_DATA_SCALE"] = old_env
    return buf.getvalue()


def test_kernel_microbench_script_runs_unchanged():
    out = _run("kernel/dgl-new.py", ["-g", "-1"], scale=0.002)
    assert out.count("hidden size: 128, avg time") == 3, out   # reddit, arxiv, proteins
    assert "OOM" not in out, out                               # the script's bare except prints OOM on ANY error


def test_kernel_microbench_other_ops():
    out = _run("kernel/dgl-new.py", ["-g", "-1", "--spmm-binary", "mul", "--spmm-reduce", "max"], scale=0.001)
    assert "OOM" not in out and out.count("avg time") == 24, out


def test_citation_sage_script_runs_unchanged():
    out = _run("end_to_end/full_graph/node_classification/main_dgl_citation_sage.py",
               ["--dataset", "cora", "--epochs", "6", "--runs", "1", "--eval"], scale=1)
    assert "Training time/epoch" in out and "Final Test" in out, out[-2000:]


def test_citation_gat_script_runs_unchanged():
    out = _run("end_to_end/full_graph/node_classification/main_dgl_citation_gat.py",
               ["--dataset", "cora", "--epochs", "5", "--runs", "1"], scale=1)
    assert "Training time/epoch" in out, out[-2000:]


def test_arxiv_gat_script_runs_unchanged():
    out = _run("end_to_end/full_graph/node_classification/main_dgl_arxiv_gat.py",
               ["--epochs", "4", "--runs", "1", "--eval"], scale=0.01)
    assert "Training time/epoch" in out and "Test:" in out, out[-2000:]


def test_product_sage_script_runs_unchanged():
    out = _run("end_to_end/full_graph/node_classification/main_dgl_product_sage.py",
               ["--epochs", "4", "--runs", "1"], scale=0.001)
    assert "Training time/epoch" in out, out[-2000:]


def test_reddit_sage_nn_script_runs_unchanged():
    out = _run("end_to_end/full_graph/node_classification/main_dgl_reddit_sage_nn.py",
               ["--dataset", "reddit", "--epochs", "5", "--runs", "1"], scale=0.002)
    assert "Training time/epoch" in out, out[-2000:]


def test_molhiv_gcn_script_runs_unchanged():
    out = _run("end_to_end/full_graph/graph_classification/main_dgl_molhiv_gcn.py",
               ["--epochs", "3", "--runs", "1", "--num_workers", "0", "--eval"], scale=0.005)
    assert "Training time/epoch" in out and "Valid:" in out, out[-2000:]


def test_enzymes_gcn_script_runs_unchanged():
    out = _run("end_to_end/full_graph/graph_classification/main_dgl_enzymes_gcn.py",
               ["--epochs", "3", "--runs", "1", "--eval"], scale=0.2)
    assert "Training time/epoch" in out and "Valid:" in out, out[-2000:]


def test_proteins_rgcn_script_runs_unchanged():
    # main_dgl_proteins_rgcn_for.py:46-60: update_all(u_mul_e('feat','weight'), mean) once per relation
    out = _run("end_to_end/full_graph/node_classification/main_dgl_proteins_rgcn_for.py",
               ["--epochs", "5", "--runs", "1", "--eval"], scale=0.0005)
    assert "Training time/epoch" in out and "Test:" in out, out[-2000:]
