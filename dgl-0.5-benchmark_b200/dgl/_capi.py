"""The door from Python into the sm_100a kernels.

Product path: `ops()` loads lib/libdglb200_torch.so -- the PyTorch C++ extension (TORCH_LIBRARY "dglb200",
csrc_torch/ops.cpp) that allocates outputs, takes torch's current stream under a device guard and calls the C-ABI of
include/dglb200.h in lib/libdglb200.so -- and `call()` runs one of its ops, turning the extension's RuntimeError into
DGLError.  There is NO CPU or PyTorch fallback behind it: if a library is missing, or an op is asked to run on
tensors that are not on a CUDA device, the call raises.

`lib()` is a plain ctypes binding of the same C-ABI; the product does not use it any more -- it is what
tests/test_capi_abi.py checks the header, the exports and the argument validation through.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdglb200.so")

OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "copy_lhs": 4, "copy_rhs": 5, "dot": 6}
REDUCERS = {"sum": 0, "max": 1, "min": 2}
TARGETS = {"u": 0, "e": 1, "v": 2}
F32 = 0
BF16 = 1

TORCH_LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdglb200_torch.so")

_lib = None
_ops = None
_launches = 0  # number of C-ABI compute calls issued (bench.py reports kernel launches from this)


class DGLError(RuntimeError):
    """Mirrors dgl.DGLError (upstream dgl._ffi.base.DGLError)."""


_vp, _i64, _i32, _int, _f32, _u64 = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int,
                                     ctypes.c_float, ctypes.c_uint64)
_shape_t = ctypes.POINTER(ctypes.c_int64)


class HubStruct(ctypes.Structure):
    """dglb_hub_t of include/dglb200.h."""
    _fields_ = [("rows", ctypes.c_void_p), ("seg_ptr", ctypes.c_void_p), ("seg_hub", ctypes.c_void_p),
                ("n_hub", ctypes.c_int32), ("n_seg", ctypes.c_int32), ("seg_len", ctypes.c_int32),
                ("threshold", ctypes.c_int32), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
                ("light_indptr", ctypes.c_void_p), ("row_order", ctypes.c_void_p)]


_hub_t = ctypes.POINTER(HubStruct)


class BatchIO(ctypes.Structure):
    """dglb_batch_io_t of include/dglb200.h."""
    _fields_ = ([("n_sel", ctypes.c_int32), ("n_nodes_pad", ctypes.c_int32), ("n_edges_pad", ctypes.c_int32)] +
                [(k, ctypes.c_void_p) for k in (
                    "graph_ids", "node_ptr", "edge_ptr", "out_node_ptr", "out_edge_ptr", "u_src", "u_dst",
                    "u_csc_indptr", "u_csc_indices", "u_csc_eids", "u_csr_indptr", "u_csr_indices", "u_csr_eids",
                    "src", "dst", "csc_indptr", "csc_indices", "csc_eids", "csr_indptr", "csr_indices", "csr_eids",
                    "node_graph", "node_map", "edge_map")])

_SIGNATURES = {
    "dglb_abi_version": (_int, []),
    "dglb_last_error": (ctypes.c_char_p, []),
    "dglb_device_info": (_int, [ctypes.POINTER(_int)] * 3 + [ctypes.POINTER(_i64)]),
    "dglb_set_device": (_int, [_int]),
    "dglb_coo_to_csr_workspace_bytes": (ctypes.c_size_t, [_i64, _i64]),
    "dglb_coo_to_csr": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "dglb_csr_degrees": (_int, [_i64, _vp, _vp, _vp]),
    "dglb_is_identity_perm": (_int, [_i64, _vp, _vp, _vp]),
    "dglb_csr_find_hub_rows": (_int, [_i64, _vp, _i32, _vp, _i64, _vp, _vp]),
    "dglb_edge_stage_plan_workspace_bytes": (ctypes.c_size_t, [_i64, _int]),
    "dglb_edge_stage_plan": (_int, [_i64, _vp, _int, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "dglb_edge_stage": (_int, [_int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "dglb_default_hub_threshold": (_i32, [_i64]),
    "dglb_default_row_hub_threshold": (_i32, [_i64]),
    "dglb_default_softmax_hub_threshold": (_i32, [_i64]),
    "dglb_edge_softmax_workspace_bytes": (ctypes.c_size_t, [_i64, _i64, _i64]),
    "dglb_gat_hub_workspace_bytes": (ctypes.c_size_t, [_i64, _i64, _i64]),
    "dglb_hub_workspace_bytes": (ctypes.c_size_t, [_i64, _i64, _int]),
    "dglb_gspmm_csr": (_int, [_int, _int, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _shape_t,
                              _shape_t, _vp, _vp, _vp, _vp, _int, _hub_t, _vp]),
    "dglb_gsddmm_csr": (_int, [_int, _int, _int, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int,
                               _shape_t, _shape_t, _vp, _hub_t, _vp]),
    "dglb_gsddmm_coo": (_int, [_int, _int, _int, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _shape_t,
                               _shape_t, _vp, _vp]),
    "dglb_edge_softmax_fwd": (_int, [_int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _hub_t, _vp]),
    "dglb_edge_softmax_bwd": (_int, [_int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _hub_t, _vp]),
    "dglb_gat_fused_fwd": (_int, [_int, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _u64, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _hub_t, _vp]),
    "dglb_gat_fused_bwd_dst": (_int, [_int, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _u64, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _hub_t, _vp]),
    "dglb_gat_fused_bwd_src": (_int, [_int, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _u64, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _hub_t, _vp]),
    "dglb_gcn_msg_sum_fwd": (_int, [_i64, _i64, _i64, _i64] + [_vp] * 9),
    "dglb_gcn_msg_sum_bwd": (_int, [_i64, _i64, _i64, _i64] + [_vp] * 11),
    "dglb_cat_embed_sum_fwd": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "dglb_cat_embed_sum_bwd_workspace_bytes": (ctypes.c_size_t, [_i64, _i64, _i64]),
    "dglb_cat_embed_sum_bwd": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "dglb_batch_offsets": (_int, [_i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "dglb_batch_gather": (_int, [ctypes.POINTER(BatchIO), _vp]),
    "dglb_copy_rows_indexed": (_int, [_i64, _vp, _i64, _vp, _vp, _vp]),
}


def exported_symbols():
    """Every entry point include/dglb200.h declares (tests check the .so exports them all)."""
    return sorted(_SIGNATURES)


def lib():
    """Load lib/libdglb200.so (built by dgl-0.5-benchmark_b200/build.py).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DGLError(
                "native library %s not found: build it with `python dgl-0.5-benchmark_b200/build.py` "
                "(there is no CPU / PyTorch fallback for the sparse kernels)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.dglb_abi_version() != 3:
            raise DGLError("libdglb200.so ABI version mismatch")
        _lib = l
    return _lib


_restore_device = None  # device index to switch back to after the C-ABI call that follows an enter()


def ops():
    """torch.ops.dglb200 (loads lib/libdglb200_torch.so, which pulls in lib/libdglb200.so).  Fails loudly if absent."""
    global _ops
    if _ops is None:
        for path in (LIB_PATH, TORCH_LIB_PATH):
            if not os.path.exists(path):
                raise DGLError(
                    "native library %s not found: build it with `python dgl-0.5-benchmark_b200/build.py` "
                    "(there is no CPU / PyTorch fallback for the sparse kernels)" % path)
        torch.ops.load_library(TORCH_LIB_PATH)
        o = torch.ops.dglb200
        if o.abi_version() != 3:
            raise DGLError("libdglb200.so ABI version mismatch")
        _ops = o
    return _ops


def call(fn, *args):
    """Run an op of the extension; its RuntimeError (bad argument, CUDA error, non-CUDA tensor) becomes DGLError."""
    try:
        return fn(*args)
    except RuntimeError as ex:
        raise DGLError(str(ex).split("\n")[0]) from None


NO_HUB = (None, None, None, None, None, [])


def check(rc, what):
    global _restore_device
    if _restore_device is not None:
        # enter() switched the thread's current CUDA device for an op on another GPU's tensors: switch back, so
        # later torch allocations / current_stream() calls of a multi-GPU process stay on the caller's device
        prev, _restore_device = _restore_device, None
        lib().dglb_set_device(prev)
    if rc != 0:
        msg = lib().dglb_last_error().decode("utf-8", "replace")
        raise DGLError("%s failed (status %d): %s" % (what, rc, msg))


def launches():
    return _launches


def count_launch(n=1):
    global _launches
    _launches += n


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def shape_arr(shape):
    return (ctypes.c_int64 * len(shape))(*shape)


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DGLError("dgl-b200 sparse kernels are CUDA-only (sm_100a); got a tensor on %s. "
                           "There is no CPU fallback." % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise DGLError("expected all operands on one device, got %s and %s" % (dev, t.device))
    return dev


def enter(dev):
    """Make `dev` current for the C-ABI call that follows (every such call is followed by check(), which
    restores the caller's device) and return the handle of torch's current stream on `dev`."""
    global _restore_device
    cur = torch.cuda.current_device()
    idx = dev.index if dev.index is not None else cur
    if idx != cur:
        rc = lib().dglb_set_device(idx)
        if rc != 0:
            raise DGLError("dglb_set_device(%d) failed: %s" % (idx, lib().dglb_last_error().decode("utf-8", "replace")))
        _restore_device = cur
    return torch.cuda.current_stream(idx).cuda_stream
