"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the CPU oracle (oracle/dgl_cpu_oracle.c).

PARITY UNPINNED (see the header of dgl_cpu_oracle.c): DGL v0.6.1 is an un-vendored pip dependency
of the reference (docker/build.dockerfile:14, README.md:6) and is absent from this image, and the
reference holds no tests or golden vectors for this path.  This module restates the *Python-level*
semantics of dmlc/dgl@0.6.1 on top of the C kernels:

    python/dgl/ops/spmm.py::gspmm            -> gspmm            (mean = sum / clamp(in_deg,1); +-inf -> 0)
    python/dgl/ops/sddmm.py::gsddmm          -> gsddmm
    python/dgl/backend/pytorch/sparse.py     -> sub -> add(-rhs), div -> mul(1/rhs); GSpMM/GSDDMM/EdgeSoftmax
                                                backward formulas (gspmm_backward, gsddmm_backward, ...)
    python/dgl/sparse.py::infer_broadcast_shape + include/dgl/bcast.h::CalcBcastOff -> calc_bcast
    python/dgl/nn/pytorch/conv/gatconv.py    -> gat_forward  (written-out twin in the reference:
                                                end_to_end/full_graph/node_classification/main_pyg_arxiv_gat.py:98-111)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (dgl-0.5-benchmark_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "copy_lhs": 4, "copy_rhs": 5, "dot": 6}
REDUCERS = {"sum": 0, "max": 1, "min": 2}
TARGETS = {"u": 0, "e": 1, "v": 2}


def build(force=False):
    """Compile the C oracle (gcc, OpenMP, no FMA contraction) into oracle/_build/liboracle.so."""
    src = os.path.join(_HERE, "dgl_cpu_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    cmd = ["gcc", "-O3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-std=c11",
           "-o", _LIB_PATH, src, "-lm"]
    subprocess.check_call(cmd)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def coo_to_csr(n_rows, row, col):
    """Stable counting sort of edges by `row` -> (indptr, indices, data=edge ids)."""
    row, col = _i32(row), _i32(col)
    nnz = row.shape[0]
    indptr = np.empty(n_rows + 1, np.int32)
    indices = np.empty(nnz, np.int32)
    data = np.empty(nnz, np.int32)
    lib().oracle_coo_to_csr(ctypes.c_int64(n_rows), ctypes.c_int64(nnz), _p(row), _p(col),
                            _p(indptr), _p(indices), _p(data))
    return indptr, indices, data


class OracleGraph:
    """Multigraph given by its creation-order COO; CSC (by dst) and CSR (by src) built lazily
    with the stable counting sort."""

    def __init__(self, src, dst, n_src=None, n_dst=None):
        self.src, self.dst = _i32(src), _i32(dst)
        if n_src is None and n_dst is None:
            n = int(max(self.src.max(initial=-1), self.dst.max(initial=-1)) + 1)
            n_src = n_dst = n
        self.n_src, self.n_dst = int(n_src), int(n_dst)
        self.n_edges = int(self.src.shape[0])
        self._csc = self._csr = None

    @property
    def csc(self):
        if self._csc is None:
            self._csc = coo_to_csr(self.n_dst, self.dst, self.src)
        return self._csc

    @property
    def csr(self):
        if self._csr is None:
            self._csr = coo_to_csr(self.n_src, self.src, self.dst)
        return self._csr

    def reverse(self):
        g = OracleGraph(self.dst, self.src, self.n_dst, self.n_src)
        g._csc, g._csr = self._csr, self._csc
        return g

    def in_degrees(self):
        return np.diff(self.csc[0]).astype(np.int32)

    def out_degrees(self):
        return np.diff(self.csr[0]).astype(np.int32)


# ----------------------------------------------------------------------------- broadcasting
def calc_bcast(op, lhs_shape, rhs_shape):
    """Restates CalcBcastOff (dmlc/dgl@0.6.1 include/dgl/bcast.h + src/array/kernel.cc): the
    per-node / per-edge feature shapes are broadcast numpy-style, iterating the axes from back to
    front; a size-1 axis contributes offset 0.  For `dot` the last axis is the reduction axis and
    offsets count units of reduce_size.  Returns dict(use_bcast, lhs_len, rhs_len, out_len,
    reduce_size, lhs_off, rhs_off, out_shape)."""
    lhs_shape, rhs_shape = list(lhs_shape), list(rhs_shape)
    nd = max(len(lhs_shape), len(rhs_shape))
    lhs_shape = [1] * (nd - len(lhs_shape)) + lhs_shape
    rhs_shape = [1] * (nd - len(rhs_shape)) + rhs_shape
    reduce_size = 1
    if op == "dot":
        if lhs_shape[-1] != rhs_shape[-1]:
            raise ValueError("dot: last dims differ")
        reduce_size = lhs_shape[-1]
    out_shape = []
    for a, b in zip(lhs_shape, rhs_shape):
        if a != b and a != 1 and b != 1:
            raise ValueError("cannot broadcast %s with %s" % (lhs_shape, rhs_shape))
        out_shape.append(max(a, b))
    lhs_len = int(np.prod(lhs_shape))
    rhs_len = int(np.prod(rhs_shape))
    if op == "dot":
        out_shape[-1] = 1
        cmp_l, cmp_r = lhs_shape[:-1] + [1], rhs_shape[:-1] + [1]
    else:
        cmp_l, cmp_r = lhs_shape, rhs_shape
    out_len = int(np.prod(out_shape))
    use_bcast = cmp_l != cmp_r
    lhs_off = np.broadcast_to(np.arange(int(np.prod(cmp_l))).reshape(cmp_l), out_shape).reshape(-1)
    rhs_off = np.broadcast_to(np.arange(int(np.prod(cmp_r))).reshape(cmp_r), out_shape).reshape(-1)
    return dict(use_bcast=use_bcast, lhs_len=lhs_len, rhs_len=rhs_len, out_len=out_len,
                reduce_size=reduce_size, lhs_off=lhs_off.astype(np.int64),
                rhs_off=rhs_off.astype(np.int64), out_shape=tuple(out_shape))


# ----------------------------------------------------------------------------- kernel level
def _gspmm(g, op, reduce, X, W, both_args=False):
    """python/dgl/sparse.py::_gspmm on the CSC of g.  Returns (out, (arg_u, arg_e)).
    both_args=True also records the arg of the unused operand (the C kernel tracks both)."""
    indptr, indices, eids = g.csc
    X, W = _f32(X), _f32(W)
    use_lhs, use_rhs = op != "copy_rhs", op != "copy_lhs"
    expand_l = expand_r = False
    if use_lhs and X.ndim == 1:
        X, expand_l = X[:, None], True
    if use_rhs and W.ndim == 1:
        W, expand_r = W[:, None], True
    lshape = X.shape[1:] if use_lhs else (W.shape[1:])
    rshape = W.shape[1:] if use_rhs else (X.shape[1:])
    bc = calc_bcast(op, lshape, rshape)
    out = np.zeros((g.n_dst,) + bc["out_shape"], np.float32)
    arg_u = arg_e = None
    if reduce != "sum":
        arg_u = np.zeros(out.shape, np.int32) if (use_lhs or both_args) else None
        arg_e = np.zeros(out.shape, np.int32) if (use_rhs or both_args) else None
    if g.n_edges > 0:
        lo = bc["lhs_off"] if bc["use_bcast"] else None
        ro = bc["rhs_off"] if bc["use_bcast"] else None
        lib().oracle_spmm_csr(
            OPS[op], REDUCERS[reduce], ctypes.c_int64(g.n_dst), _p(indptr), _p(indices), _p(eids),
            _p(X) if use_lhs else None, _p(W) if use_rhs else None,
            ctypes.c_int64(bc["lhs_len"]), ctypes.c_int64(bc["rhs_len"]),
            ctypes.c_int64(bc["out_len"]), _p(lo), _p(ro), _p(out), _p(arg_u), _p(arg_e))
    if (expand_l or not use_lhs) and (expand_r or not use_rhs):  # scalar features: squeeze back
        out = out.reshape(out.shape[:-1])
        arg_u = None if arg_u is None else arg_u.reshape(arg_u.shape[:-1])
        arg_e = None if arg_e is None else arg_e.reshape(arg_e.shape[:-1])
    return out, (arg_u, arg_e)


def _gsddmm(g, op, L, R, lhs_target="u", rhs_target="v"):
    """python/dgl/sparse.py::_gsddmm on the COO of g; output in edge-id order."""
    L, R = _f32(L), _f32(R)
    use_lhs, use_rhs = op != "copy_rhs", op != "copy_lhs"
    expand_l = expand_r = False
    if use_lhs and L.ndim == 1:
        L, expand_l = L[:, None], True
    if use_rhs and R.ndim == 1:
        R, expand_r = R[:, None], True
    lshape = L.shape[1:] if use_lhs else R.shape[1:]
    rshape = R.shape[1:] if use_rhs else L.shape[1:]
    bc = calc_bcast(op, lshape, rshape)
    out = np.zeros((g.n_edges,) + bc["out_shape"], np.float32)
    if g.n_edges > 0:
        lo = bc["lhs_off"] if bc["use_bcast"] else None
        ro = bc["rhs_off"] if bc["use_bcast"] else None
        lib().oracle_sddmm_coo(
            OPS[op], TARGETS[lhs_target], TARGETS[rhs_target], ctypes.c_int64(g.n_edges),
            _p(g.src), _p(g.dst), _p(L) if use_lhs else None, _p(R) if use_rhs else None,
            ctypes.c_int64(bc["lhs_len"]), ctypes.c_int64(bc["rhs_len"]),
            ctypes.c_int64(bc["out_len"]), ctypes.c_int64(bc["reduce_size"]), _p(lo), _p(ro), _p(out))
    if (expand_l or not use_lhs) and (expand_r or not use_rhs):
        out = out.reshape(out.shape[:-1])
    return out


# ----------------------------------------------------------------------------- ops level
def _reshape_lhs_rhs(l, r):
    """python/dgl/ops/spmm.py::reshape_lhs_rhs: pad the shorter feature shape with 1s after dim 0."""
    ls, rs = l.shape, r.shape
    if len(ls) != len(rs):
        m = max(len(ls), len(rs))
        l = l.reshape((ls[0],) + (1,) * (m - len(ls)) + ls[1:])
        r = r.reshape((rs[0],) + (1,) * (m - len(rs)) + rs[1:])
    return l, r


def gspmm(g, op, reduce_op, lhs_data, rhs_data):
    """dgl.ops.gspmm (kernel/dgl-new.py:20)."""
    lhs_data, rhs_data = _f32(lhs_data), _f32(rhs_data)
    if op not in ("copy_lhs", "copy_rhs"):
        lhs_data, rhs_data = _reshape_lhs_rhs(lhs_data, rhs_data)
    if op == "sub":
        op, rhs_data = "add", -rhs_data
    if op == "div":
        op, rhs_data = "mul", (np.float32(1.0) / rhs_data).astype(np.float32)
    out, _ = _gspmm(g, op, "sum" if reduce_op == "mean" else reduce_op, lhs_data, rhs_data)
    if reduce_op == "mean":
        deg = np.maximum(g.in_degrees(), 1).astype(np.float32)
        out = (out / deg.reshape((-1,) + (1,) * (out.ndim - 1))).astype(np.float32)
    if reduce_op in ("max", "min"):
        out = np.where(np.isinf(out), np.float32(0), out)
    return out


def gspmm_with_args(g, op, reduce_op, lhs_data, rhs_data, both_args=False):
    return _gspmm(g, op, reduce_op, lhs_data, rhs_data, both_args=both_args)


def gsddmm(g, op, lhs_data, rhs_data, lhs_target="u", rhs_target="v"):
    """dgl.ops.gsddmm (kernel/dgl-new.py:39)."""
    lhs_data, rhs_data = _f32(lhs_data), _f32(rhs_data)
    if op not in ("copy_lhs", "copy_rhs"):
        lhs_data, rhs_data = _reshape_lhs_rhs(lhs_data, rhs_data)
    if op == "sub":
        op, rhs_data = "add", -rhs_data
    if op == "div":
        op, rhs_data = "mul", (np.float32(1.0) / rhs_data).astype(np.float32)
    return _gsddmm(g, op, lhs_data, rhs_data, lhs_target, rhs_target)


def edge_softmax(g, logits):
    """EdgeSoftmax.forward (dmlc/dgl@0.6.1 python/dgl/backend/pytorch/sparse.py), norm_by='dst'."""
    score = _f32(logits)
    score_max, _ = _gspmm(g, "copy_rhs", "max", None, score)
    # upstream calls the kernel-level _gsddmm(gidx, 'sub' / 'div', ...): true subtract / divide
    score = np.exp(_gsddmm(g, "sub", score, score_max, "e", "v")).astype(np.float32)
    score_sum, _ = _gspmm(g, "copy_rhs", "sum", None, score)
    return _gsddmm(g, "div", score, score_sum, "e", "v")


def edge_softmax_backward(g, out, grad_out):
    """EdgeSoftmax.backward: sds = out*grad; accum = copy_rhs-sum(sds); grad = sds - out*accum[dst]."""
    out, grad_out = _f32(out), _f32(grad_out)
    sds = (out * grad_out).astype(np.float32)
    accum, _ = _gspmm(g, "copy_rhs", "sum", None, sds)
    return (sds - _gsddmm(g, "mul", out, accum, "e", "v")).astype(np.float32)


# ----------------------------------------------------------------------------- autograd restatements
def _reduce_grad(grad, shape):
    """backend/pytorch/sparse.py::_reduce_grad: sum broadcast dims so grad matches `shape`."""
    grad_shape = grad.shape[1:]
    in_shape = tuple(shape[1:])
    if in_shape == grad_shape:
        return grad
    nd = len(grad_shape)
    in_shape = (1,) * (nd - len(in_shape)) + in_shape
    axes = tuple(i + 1 for i in range(nd) if in_shape[i] != grad_shape[i])
    grad = grad.sum(axis=axes, keepdims=True, dtype=np.float32)
    return grad.reshape((-1,) + tuple(shape[1:]))


def gspmm_sum_backward(g, op, X, W, dZ):
    """GSpMM.backward for reducer sum.  Returns (dX, dW)."""
    dZ = _f32(dZ)
    gr = g.reverse()
    dX = dW = None
    if op != "copy_rhs":
        if op == "mul":
            dX = _reduce_grad(_gspmm(gr, "mul", "sum", dZ, _bshape(W, dZ))[0], X.shape)
        elif op in ("add", "copy_lhs"):
            dX = _reduce_grad(_gspmm(gr, "copy_lhs", "sum", dZ, None)[0], X.shape)
    if op != "copy_lhs":
        if op == "mul":
            Xb = _bshape(X, dZ)
            if W.shape[-1] == 1 and Xb.shape[-1] > 1 and W.shape[1:-1] == Xb.shape[1:-1]:
                dW = _gsddmm(g, "dot", Xb, dZ)
            else:
                dW = _reduce_grad(_gsddmm(g, "mul", Xb, dZ), W.shape)
        elif op in ("add", "copy_rhs"):
            dW = _reduce_grad(_gsddmm(g, "copy_rhs", None, dZ, "u", "v"), W.shape)
    return dX, dW


def _bshape(a, like):
    a = _f32(a)
    if a.ndim < like.ndim:
        a = a.reshape((a.shape[0],) + (1,) * (like.ndim - a.ndim) + a.shape[1:])
    return a


# ----------------------------------------------------------------------------- GATConv forward
def gat_forward(g, ft, attn_l, attn_r, negative_slope=0.2):
    """GATConv.forward attention part (upstream gatconv.py; twin main_pyg_arxiv_gat.py:98-111).
    ft (N,H,F); returns rst (N_dst,H,F), a (E,H,1), el, er."""
    ft = _f32(ft)
    el = (ft * _f32(attn_l)).sum(-1, keepdims=True, dtype=np.float32)
    er = (ft[: g.n_dst] * _f32(attn_r)).sum(-1, keepdims=True, dtype=np.float32)
    e = _gsddmm(g, "add", el, er, "u", "v")
    e = np.where(e > 0, e, e * np.float32(negative_slope)).astype(np.float32)
    a = edge_softmax(g, e)
    rst, _ = _gspmm(g, "mul", "sum", ft, a)
    return rst, a, el, er


# ----------------------------------------------------------------------------- batched small graphs
def gcn_message_sum(g, x, w, c_src, c_dst):
    """The message UDF + reducer of the graph-classification GCN layer, written out as in the reference:
    end_to_end/full_graph/graph_classification/main_dgl_molhiv_gcn.py:50-52 (norm = c[src] * c[dst];
    m = norm * relu(x[src] + w)) followed by update_all(..., fn.sum('m', 'h')) (:46) = gspmm(copy_rhs, sum) in CSC order.
    Elementwise float32 with one rounding per operation, like the torch ops the UDF runs."""
    x, w = _f32(x), _f32(w)
    c_src, c_dst = _f32(c_src).reshape(-1), _f32(c_dst).reshape(-1)
    norm = (c_src[g.src] * c_dst[g.dst]).astype(np.float32)[:, None]
    m = (norm * np.maximum(x[g.src] + w, np.float32(0))).astype(np.float32)
    return _gspmm(g, "copy_rhs", "sum", None, m)[0]


def gcn_message_sum_backward(g, x, w, c_src, c_dst, grad_out):
    """Gradients of gcn_message_sum w.r.t. x and w in float64 (the reference gets them from torch autograd through
    the same ops: d m = grad_out[dst]; d relu = d m * norm where x[src] + w > 0; d x = scatter-add over src)."""
    x64, w64 = np.asarray(x, np.float64), np.asarray(w, np.float64)
    norm = (np.asarray(c_src, np.float64).reshape(-1)[g.src] * np.asarray(c_dst, np.float64).reshape(-1)[g.dst])[:, None]
    pre = _f32(x)[g.src] + _f32(w)                      # the mask is decided in float32, like the forward
    gw = np.where(pre > 0, np.asarray(grad_out, np.float64)[g.dst] * norm, 0.0)
    gx = np.zeros_like(x64)
    np.add.at(gx, g.src, gw)
    return gx, gw


def batch_graphs(members):
    """dgl.batch restated (upstream python/dgl/batch.py::batch; call site main_dgl_molhiv_gcn.py:163 through
    GraphDataLoader): members = [(src, dst, n_nodes), ...]; node ids of member i are shifted by the node counts of the
    members before it, edges keep member order.  Returns (OracleGraph, node_offsets, edge_offsets)."""
    n_off = np.concatenate([[0], np.cumsum([m[2] for m in members])]).astype(np.int64)
    e_off = np.concatenate([[0], np.cumsum([len(m[0]) for m in members])]).astype(np.int64)
    src = np.concatenate([np.asarray(m[0], np.int64) + o for m, o in zip(members, n_off)]) if members else np.zeros(0)
    dst = np.concatenate([np.asarray(m[1], np.int64) + o for m, o in zip(members, n_off)]) if members else np.zeros(0)
    return OracleGraph(src, dst, int(n_off[-1]), int(n_off[-1])), n_off, e_off
