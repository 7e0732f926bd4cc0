"""The C-ABI library loads and exports every symbol include/dglb200.h declares, and the ctypes
signatures in dgl/_capi.py have the same arity as the header (no compute calls: CPU-only test)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_functions():
    text = open(os.path.join(ROOT, "include", "dglb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = {}
    for m in re.finditer(r"\b(?:int|size_t|int32_t|const char\*)\s+(dglb_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        funcs[name] = n
    return funcs


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("dglb_build", os.path.join(ROOT, "dgl-0.5-benchmark_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_header_declares_the_expected_entry_points():
    funcs = _header_functions()
    for name in ("dglb_coo_to_csr", "dglb_gspmm_csr", "dglb_gsddmm_csr", "dglb_gsddmm_coo", "dglb_edge_softmax_fwd",
                 "dglb_edge_softmax_bwd", "dglb_gat_fused_fwd", "dglb_gat_fused_bwd_dst", "dglb_gat_fused_bwd_src"):
        assert name in funcs


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _header_functions():
        assert hasattr(lib, name), "libdglb200.so does not export %s" % name
    assert lib.dglb_abi_version() == 3


def test_ctypes_signatures_match_header_arity(built_lib):
    from dgl import _capi
    funcs = _header_functions()
    assert set(_capi.exported_symbols()) == set(funcs)
    for name, (_, argtypes) in _capi._SIGNATURES.items():
        assert len(argtypes) == funcs[name], (name, len(argtypes), funcs[name])
    _capi.lib()  # binds every symbol with its argtypes


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    from dgl import _capi
    l = _capi.lib()
    shp = _capi.shape_arr((4,))
    # unknown op -> DGLB_E_INVALID before any CUDA call
    rc = l.dglb_gspmm_csr(99, 0, 0, 1, 1, 0, None, None, None, None, None, 1, shp, shp, None, None, None, None,
                          0, None, None)
    assert rc == -1 and b"unknown op" in l.dglb_last_error()
    bad = _capi.shape_arr((3,))
    rc = l.dglb_gsddmm_coo(0, 0, 0, 2, 1, 1, 1, None, None, ctypes.c_void_p(8), ctypes.c_void_p(8), 1, shp, bad,
                           ctypes.c_void_p(8), None)
    assert rc == -1 and b"broadcast" in l.dglb_last_error()
    assert l.dglb_default_hub_threshold(602) == 256
    assert l.dglb_default_hub_threshold(16) == 64


def test_host_side_helpers_of_the_hub_paths(built_lib):
    """cut-offs and workspace sizes are pure host functions (no GPU): the values DESIGN.md / the notes quote."""
    from dgl import _capi
    l = _capi.lib()
    assert [l.dglb_default_row_hub_threshold(w) for w in (16, 64, 160, 602)] == [128] * 4     # fused GAT (notes §11)
    assert [l.dglb_default_softmax_hub_threshold(h) for h in (1, 3, 4, 8, 32, 40)] == [1024, 256, 256, 128, 64, 64]
    # edge_softmax: (n_seg + n_hub) x heads padded to a power of two x {max|acc, sum} floats
    assert l.dglb_edge_softmax_workspace_bytes(10, 3, 3) == (10 + 3) * 4 * 2 * 4
    # fused GAT: per segment one partial feature row (H*F) + 4 floats per head
    assert l.dglb_gat_hub_workspace_bytes(7, 4, 16) == 7 * (64 + 16) * 4
    assert l.dglb_hub_workspace_bytes(5, 100, 0) == 5 * 100 * 4 and l.dglb_hub_workspace_bytes(5, 100, 1) == 5 * 100 * 12
    # unknown gspmm flag bits are rejected (after the pointer checks, so give it non-null dummies)
    shp = _capi.shape_arr((4,))
    d = ctypes.c_void_p(16)
    rc = l.dglb_gspmm_csr(4, 0, 0, 1, 1, 1, d, d, None, d, None, 1, shp, shp, d, None, None, None, 8, None, None)
    assert rc == -1 and b"flag" in l.dglb_last_error()
    rc = l.dglb_gspmm_csr(4, 1, 0, 1, 1, 1, d, d, None, d, None, 1, shp, shp, d, None, None, None, 1, None, None)
    assert rc == -1 and b"accumulate" in l.dglb_last_error()


def test_torch_extension_loads_and_fails_loudly_on_cpu_tensors():
    """The PyTorch C++ extension (csrc_torch/ops.cpp, TORCH_LIBRARY "dglb200") is the product's only door into the
    kernels: it must load, expose every op the Python layer calls, agree with the C library on the ABI version, and
    refuse CPU tensors with an error (there is no CPU fallback to fall into)."""
    import torch
    from dgl import _capi
    o = _capi.ops()
    assert o.abi_version() == 3
    for name in ("coo_to_csr", "csr_degrees", "is_identity_perm", "find_hub_rows", "edge_stage_plan", "edge_stage",
                 "gspmm", "gsddmm_csr", "gsddmm_coo", "edge_softmax_fwd", "edge_softmax_bwd", "gat_fwd", "gat_bwd_dst",
                 "gat_bwd_src", "default_hub_threshold", "gcn_msg_sum_fwd", "gcn_msg_sum_bwd", "batch_build"):
        assert hasattr(o, name), name
    assert o.default_hub_threshold(0, 602) == 256 and o.default_hub_threshold(1, 64) == 128
    with pytest.raises(_capi.DGLError, match="CUDA-only"):
        _capi.call(o.csr_degrees, torch.zeros(4, dtype=torch.int32))
    ip = torch.tensor([0, 1, 2], dtype=torch.int32)
    with pytest.raises(_capi.DGLError, match="CUDA-only"):
        _capi.call(o.gspmm, ip, ip[:2], None, 2, 4, 0, torch.ones(2, 4), None, [4], [4], [4], None, None, 0, *_capi.NO_HUB)


def test_small_graph_entry_points_validate_arguments_without_a_gpu(built_lib):
    """dglb_batch_gather / dglb_gcn_msg_sum_* reject inconsistent argument sets before any CUDA call."""
    from dgl import _capi
    l = _capi.lib()
    assert l.dglb_batch_gather(None, None) == -1 and b"io is null" in l.dglb_last_error()
    io = _capi.BatchIO()
    io.n_sel, io.n_nodes_pad, io.n_edges_pad = 2, 8, 8
    for k in ("graph_ids", "node_ptr", "edge_ptr", "out_node_ptr", "out_edge_ptr", "csc_indptr"):
        setattr(io, k, 8)
    assert l.dglb_batch_gather(ctypes.byref(io), None) == -1 and b"union CSC" in l.dglb_last_error()
    assert l.dglb_gcn_msg_sum_fwd(4, 4, 3, 8, ctypes.c_void_p(8), None, None, None, None, None, ctypes.c_void_p(8),
                                  ctypes.c_void_p(8), None) == -1
    assert b"null operand" in l.dglb_last_error()
    with pytest.raises(_capi.DGLError, match="CUDA-only"):
        import torch
        _capi.call(_capi.ops().batch_build, torch.zeros(2, dtype=torch.int32), [None] * 10, [None] * 14, 4, 4)
