"""Regression guard (CPU-only, cuobjdump): the hot gather kernels must issue a whole batch of
neighbour-row loads before the first floating-point consumer.  Twice in round 1 a harmless-looking
source change made the compiler interleave each load with its FMA (1-3 loads in flight instead of 8),
costing 1.5-2.6x at identical DRAM traffic (profiles/r01_notes.md); this test catches that in SASS."""
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

KERNELS = {  # mangled-name fragment -> minimum run of vector loads with no FP instruction in between
    "spmm_rows_kernelILi2ELi4ELi4ELi0ELi0EfLb0EEE": 6,        # gspmm copy_u_sum, D=602 (LDG.64)
    "spmm_rows_kernelILi4ELi1ELi4ELi0ELi0EfLb0EEE": 6,        # gspmm copy_u_sum, D=64..128 (LDG.128)
    "spmm_rows_kernelILi4ELi2ELi4ELi0ELi0EfLb0EEE": 6,        # D=256
    "sddmm_dot_kernelILi2ELi4ELi5ELb0ELb0EfEE": 8,        # gsddmm u_dot_v, D=602
    "sddmm_dot_kernelILi4ELi1ELi4ELb1ELb0EfEE": 6,        # D=64
    "sddmm_dot_kernelILi4ELi2ELi5ELb1ELb0EfEE": 6,        # D=256
    "spmm_rows_kernelILi8ELi4ELi4ELi0ELi0E13__nv_bfloat16Lb0E": 6,  # bf16 storage, D=608
    # u_mul_e_sum with (E,1) weights, D=64: the weight shuffles are hoisted in front of the gathers; with them in
    # the consume phase ptxas issued the batch as 3 + 5 gathers
    "spmm_rows_kernelILi4ELi1ELi2ELi0ELi3EfLb0EEE": 7,
}


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("dglb_build", os.path.join(ROOT, "dgl-0.5-benchmark_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = mod.build()
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    return funcs


@pytest.mark.parametrize("fragment,min_run", sorted(KERNELS.items()))
def test_gathers_are_issued_as_a_batch(sass, fragment, min_run):
    names = [n for n in sass if fragment in n]
    assert names, "kernel %s not found in libdglb200.so" % fragment
    best = run = 0
    for line in sass[names[0]]:
        if re.search(r"\bLDG\.E\.(64|128)", line):
            run += 1
            best = max(best, run)
        elif re.search(r"\b(FFMA|FADD|FMUL)\b", line):
            run = 0
    assert best >= min_run, "%s: longest run of vector gathers before an FP consumer is %d (< %d)" % (fragment, best, min_run)


# Residency guard: 256-thread CTAs need <= 80 registers per thread for 3 CTAs per SM (48 for 5, 32 for 8).  Measured on the
# products graph: copy_u_sum 2.72 ms at 70 registers vs 4.78 ms at 84; copy_u_max 4.04 ms at 98 vs 2.88 ms
# at 80; fused GAT 1.2-1.35x from 2 -> 3 CTAs per SM (profiles/r01_notes.md sections 9-10).
REG_BUDGET = {
    "spmm_rows_kernelILi4ELi1ELi4ELi0ELi0EfLb0EEE": 80,   # copy_u_sum D=64/128
    "spmm_rows_kernelILi4ELi2ELi4ELi0ELi0EfLb0EEE": 80,   # D=256
    "spmm_rows_kernelILi2ELi4ELi4ELi0ELi0EfLb0EEE": 80,   # D=602
    "spmm_rows_kernelILi4ELi1ELi4ELi1ELi0EfLb0EEE": 80,   # copy_u_max D=64
    "spmm_rows_kernelILi2ELi4ELi4ELi1ELi0EfLb0EEE": 80,   # copy_u_max D=602
    "spmm_rows_kernelILi4ELi1ELi2ELi0ELi3EfLb0EEE": 80,   # u_mul_e_sum with (E,1) weights, D=64
    "sddmm_dot_kernelILi2ELi4ELi5ELb0ELb0EfEE": 80,   # u_dot_v D=602
    "gat_fwd_kernelILi4ELi1ELi1ELi4ELi0EEE": 80,      # fused GAT forward, UT = 4, ordinary rows
    "gat_fwd_kernelILi4ELi1ELi4ELi4ELi2EEE": 80,      # ... hub-row segments, 4 heads
    "gat_bwd_kernelILi4ELi1ELi1ELi4ELb0ELi0EEE": 80,  # fused GAT backward (dst pass)
    "gat_bwd_kernelILi4ELi1ELi1ELi4ELb1ELi0EEE": 80,  # fused GAT backward (src pass)
    "gat_bwd_kernelILi4ELi1ELi4ELi4ELb1ELi2EEE": 80,  # ... hub-row segments, 4 heads
    "edge_softmax_rows_kernelILb0ELi8ELb0EEE": 32,    # edge_softmax fwd, dst-sorted graph, R = 8: 8 CTAs/SM
    "edge_softmax_rows_kernelILb1ELi8ELb0EEE": 32,    # ... bwd
    "edge_softmax_rows_kernelILb0ELi16ELb0EEE": 48,   # R = 16: 5 CTAs/SM
    "edge_softmax_rows_kernelILb1ELi16ELb0EEE": 48,
}


@pytest.fixture(scope="module")
def res_usage(sass):
    import os
    lib = os.path.join(ROOT, "dgl-0.5-benchmark_b200", "lib", "libdglb200.so")
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    regs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    return regs


@pytest.mark.parametrize("fragment,budget", sorted(REG_BUDGET.items()))
def test_hot_kernels_keep_three_ctas_per_sm(res_usage, fragment, budget):
    names = [n for n in res_usage if fragment in n]
    assert names, "kernel %s not found in libdglb200.so" % fragment
    assert res_usage[names[0]] <= budget, "%s uses %d registers (> %d: only 2 CTAs per SM)" % (
        fragment, res_usage[names[0]], budget)
