"""Kernel-level sparse ops: shape inference, output allocation and the C-ABI calls.

Mirrors upstream DGL v0.6.1 python/dgl/sparse.py (`_gspmm`, `_gsddmm`, `infer_broadcast_shape`):
same argument meaning (`op`, `reduce_op`, operands may be None, 1-D operands are treated as
(n,1) and squeezed back), same error behaviour (DGLError on bad broadcast / missing operand),
same "skip the kernel when the graph has no edges" rule.  Outputs are allocated by torch;
everything else happens in lib/libdglb200.so on the current CUDA stream.
"""
import ctypes
import os

import torch

from . import _capi
from ._capi import DGLError

_TARGET = _capi.TARGETS

# Rows with more nnz than this are handled by the split-row ("hub") kernels.  None = the library's
# width-dependent default (dglb_default_hub_threshold); tests lower it to exercise the hub path.
HUB_THRESHOLD = None


def _hub_threshold(width, row_cta=False):
    """Hub cut-off for a feature width: segmented path (gspmm / gsddmm) or, with row_cta=True, the
    one-CTA-per-row path of edge_softmax and the fused GAT kernels."""
    if HUB_THRESHOLD is not None:
        return int(HUB_THRESHOLD)
    return _capi.ops().default_hub_threshold(1 if row_cta else 0, int(width))


# Degree-ordered hand-out of rows to the row-per-group kernels of gspmm / gsddmm (dglb_hub_t.row_order): "auto" = on
# graphs that have hub rows (skewed degrees: the 4-32 rows a warp walks together differ wildly in length otherwise),
# "always", "never".  Results are bit-identical either way.
ROW_ORDER = os.environ.get("DGLB_ROW_ORDER", "auto")


def _hub_pack(info, view=None):
    """(hub arguments of an extension op, extra launches the hub path adds)."""
    order = None
    if view is not None and (ROW_ORDER == "always" or (ROW_ORDER == "auto" and info is not None)):
        order = view.row_order()
    if info is None:
        return ((None, None, None, None, order, []) if order is not None else _capi.NO_HUB), 0
    p = info.pack()
    return (p[0], p[1], p[2], p[3], order, p[5]), 1


# Staged edge order (include/dglb200.h, dglb_edge_stage_plan): on graphs whose CSC / CSR carries a non-trivial edge-id
# permutation, narrow per-edge tensors are moved into "staged" order by one cheap pass and the kernels address them
# through the plan's `slot` array, so a 4..64-byte per-edge access no longer costs a 128-byte DRAM line.
# It pays once the per-edge tensor no longer fits the 126 MB L2 (reddit shape, E = 11.6 M: (E,1) weights = 46 MB are
# L2-resident and the extra pass costs more than it saves; (E,4) scores = 186 MB and everything on the products
# shape win 1.3-1.6x -- profiles/r02_notes.md).  Tests lower the cut-off.
STAGE_MIN_BYTES = 96 << 20
# Per-edge rows wider than this are addressed directly: measured on the products shape (profiles/r02_notes.md), (E,1)
# operands / results gain 1.3-2.5x, (E,4) ones lose 5-8 % in gspmm / gsddmm (a 16-byte access already uses half a
# sector pair) but still gain 1.2x in edge_softmax.  A second pass that brings the tensor all the way into CSR-position
# order (so the kernels run with eids = NULL) was measured too and loses to the single staging pass everywhere.
NARROW_COO_MAX_FLOATS = int(os.environ.get("DGLB_NARROW_COO", "8"))   # 0 disables (A/B measurements)
STAGE_MAX_ROW_FLOATS = 2
STAGE_MAX_ROW_FLOATS_SOFTMAX = 8


def _stage_plan(view, row_floats, max_floats=None):
    if max_floats is None:
        max_floats = STAGE_MAX_ROW_FLOATS
    if (view.eids is None or row_floats < 1 or row_floats > max_floats
            or view.nnz * row_floats * 4 < STAGE_MIN_BYTES):
        return None
    return view.stage_plan()


def _stage_move(plan, t, to_staged):
    """Edge-id order -> staged order (to_staged) or back, for a contiguous per-edge tensor (E, ...)."""
    _capi.count_launch(1)
    return _capi.call(_capi.ops().edge_stage, plan[0], t, bool(to_staged))


def infer_broadcast_shape(op, shp1, shp2):
    """Feature shape of op(lhs, rhs) under numpy-style broadcasting of the per-node / per-edge
    feature shapes (leading node/edge dim excluded).  `dot` reduces the last dim to 1."""
    pad1, pad2 = tuple(shp1), tuple(shp2)
    if op == "copy_lhs":
        return pad1
    if op == "copy_rhs":
        return pad2
    if len(pad1) != len(pad2):
        n = max(len(pad1), len(pad2))
        pad1 = (1,) * (n - len(pad1)) + pad1
        pad2 = (1,) * (n - len(pad2)) + pad2
    for d1, d2 in zip(pad1, pad2):
        if d1 != d2 and d1 != 1 and d2 != 1:
            raise DGLError("Feature shapes {} and {} are not valid for broadcasting.".format(shp1, shp2))
    rst = tuple(max(d1, d2) for d1, d2 in zip(pad1, pad2))
    return rst[:-1] + (1,) if op == "dot" else rst


def _shapes_for_abi(op, lhs, rhs):
    """Right-aligned trailing shapes (>= 1 dim) handed to the C-ABI."""
    ls = tuple(lhs.shape[1:]) if lhs is not None else None
    rs = tuple(rhs.shape[1:]) if rhs is not None else None
    if ls is None:
        ls = rs
    if rs is None:
        rs = ls
    n = max(len(ls), len(rs), 1)
    ls = (1,) * (n - len(ls)) + ls
    rs = (1,) * (n - len(rs)) + rs
    if n > 5:
        raise DGLError("feature tensors with more than 5 trailing dims are not supported")
    return n, list(ls), list(rs)


def _check_float32(*ts):
    for t in ts:
        if t is not None and t.dtype != torch.float32:
            raise DGLError("dgl-b200 kernels compute in float32; got %s" % t.dtype)


def _dtype_code(what, bf16_ok, *ts):
    """C-ABI dtype of the operands: float32 everywhere; bfloat16 STORAGE (fp32 accumulate, one rounding
    at the store) for gspmm copy_u sum/mean and gsddmm u_dot_v."""
    dts = {t.dtype for t in ts if t is not None}
    if dts == {torch.float32}:
        return _capi.F32
    if dts == {torch.bfloat16}:
        if not bf16_ok:
            raise DGLError("%s: bfloat16 storage is implemented for gspmm copy_u sum/mean and gsddmm u_dot_v only" % what)
        return _capi.BF16
    raise DGLError("%s: operands must be all float32 (or all bfloat16 where supported); got %s" % (what, sorted(map(str, dts))))


def _gspmm(gidx, op, reduce_op, u, e, row_scale=None, out=None, zero_inf=False):
    """out[v] = reduce_{(s->v)} op(u[s], e[eid]).  Returns (out, (arg_u, arg_e)).

    `gidx` is a GraphIndex; the kernel walks its CSC.  reduce_op in {sum, max, min}; `row_scale`
    (float32, n_dst) fuses the mean divide.  arg_u / arg_e (graph idtype) only for max / min.
    `out` (reducer sum only): accumulate into this existing tensor instead of allocating a result.
    `zero_inf` (max / min): store 0 instead of -/+inf (rows without in-edges), i.e. upstream's
    where(isinf(out), 0, out) post-pass folded into the kernel's store.
    """
    use_u = op != "copy_rhs"
    use_e = op != "copy_lhs"
    if use_u and u is None:
        raise DGLError("gspmm: op %s needs node data" % op)
    if use_e and e is None:
        raise DGLError("gspmm: op %s needs edge data" % op)
    if use_u and use_e and u.dtype != e.dtype:
        raise DGLError("The node features' data type {} doesn't match edge features' data type {}, "
                       "please convert them to the same type.".format(u.dtype, e.dtype))
    if op not in ("add", "sub", "mul", "div", "copy_lhs", "copy_rhs"):
        raise DGLError("gspmm: unknown op %s" % op)
    if reduce_op not in ("sum", "max", "min"):
        raise DGLError("gspmm: unknown reducer %s" % reduce_op)
    u = u if use_u else None
    e = e if use_e else None
    dtype = _dtype_code("gspmm", op == "copy_lhs" and reduce_op == "sum", u, e)
    dev = _capi.require_cuda(u, e, gidx.src)
    expand_u = expand_e = False
    if use_u:
        if u.shape[0] != gidx.n_src:
            raise DGLError("gspmm: expect %d source-node rows, got %d" % (gidx.n_src, u.shape[0]))
        if u.dim() == 1:
            u, expand_u = u.unsqueeze(-1), True
        u = u.contiguous()
    if use_e:
        if e.shape[0] != gidx.n_edges:
            raise DGLError("gspmm: expect %d edge rows, got %d" % (gidx.n_edges, e.shape[0]))
        if e.dim() == 1:
            e, expand_e = e.unsqueeze(-1), True
        e = e.contiguous()
    feat_shape = infer_broadcast_shape(op, u.shape[1:] if use_u else (1,), e.shape[1:] if use_e else (1,))
    ref = u if use_u else e
    out_shape = (gidx.n_dst,) + tuple(feat_shape)
    use_cmp = reduce_op in ("max", "min")
    if dtype == _capi.BF16 and out is None:
        # bf16 rows whose width is not a multiple of 8 (e.g. 602) would be gathered with 4-byte loads:
        # zero-pad the row once (one cheap pass), run with 128-bit loads, slice the padding off
        width = 1
        for s_ in feat_shape:
            width *= s_
        # (wide rows on large graphs go through the bulk-copy ring kernel, which takes 4-byte-aligned rows as they are)
        ring = width * 2 >= 1024 and width % 2 == 0 and gidx.n_edges >= (1 << 18)
        if width >= 32 and width % 8 and not ring:
            padded = torch.nn.functional.pad(u.reshape(u.shape[0], width), (0, 8 - width % 8))
            res, _ = _gspmm(gidx, op, reduce_op, padded, None, row_scale)
            return res[:, :width].contiguous().reshape(out_shape), (None, None)
    arg_u = arg_e = None
    if out is not None:
        if reduce_op != "sum" or tuple(out.shape) != out_shape or not out.is_contiguous() or out.dtype != ref.dtype:
            raise DGLError("gspmm: `out` must be a contiguous %s tensor and the reducer must be sum" % (out_shape,))
    if gidx.n_edges == 0 or gidx.n_dst == 0:
        v = out if out is not None else torch.zeros(out_shape, dtype=ref.dtype, device=dev)
        if use_cmp:
            arg_u = torch.zeros(out_shape, dtype=gidx.idtype, device=dev) if use_u else None
            arg_e = torch.zeros(out_shape, dtype=gidx.idtype, device=dev) if use_e else None
    else:
        csc = gidx.csc()
        out_len = 1
        for s_ in feat_shape:
            out_len *= s_
        # relation broadcast (N,1,D) x (E,R,1) -> (N,R,D), R in {2,4,8}: the batched RGCN kernel gathers a neighbour row
        # once for all R relations; it has no split-row path, so no hub rows are handed to it
        rel = (op == "mul" and not use_cmp and out is None and dtype == _capi.F32 and u.dim() == 3 and e.dim() == 3
               and u.shape[1] == 1 and e.shape[2] == 1 and e.shape[1] in (2, 4, 8) and 1 < u.shape[2] <= 128)
        hub, hub_launches = (_capi.NO_HUB, 0) if rel else _hub_pack(csc.hubs(_hub_threshold(out_len)), csc)
        if hub_launches:
            hub_launches = 2          # segment kernel + combine kernel
        ndim, ls, rs = _shapes_for_abi(op, u, e)
        eids = csc.eids
        if use_e and not use_cmp and dtype == _capi.F32:   # max / min record eids as arg_e: they need the real ids
            plan = _stage_plan(csc, e.numel() // max(e.shape[0], 1))
            if plan is not None:
                e, eids = _stage_move(plan, e, True), plan[1]
        flags = (1 if out is not None else 0) | (2 if (zero_inf and use_cmp) else 0)
        v, arg_u, arg_e = _capi.call(_capi.ops().gspmm, csc.indptr, csc.indices, eids, csc.n_cols, _capi.OPS[op],
                                     _capi.REDUCERS[reduce_op], u, e, list(feat_shape), ls, rs, row_scale, out, flags, *hub)
        _capi.count_launch(1 + hub_launches)
        arg_u = arg_u if (use_cmp and use_u) else None
        arg_e = arg_e if (use_cmp and use_e) else None
        if use_cmp and gidx.idtype != torch.int32:
            arg_u = arg_u.to(gidx.idtype) if arg_u is not None else None
            arg_e = arg_e.to(gidx.idtype) if arg_e is not None else None
    if (expand_u or not use_u) and (expand_e or not use_e):
        v = v.squeeze(-1)
        if arg_u is not None:
            arg_u = arg_u.squeeze(-1)
        if arg_e is not None:
            arg_e = arg_e.squeeze(-1)
    return v, (arg_u, arg_e)


def _gsddmm(gidx, op, lhs, rhs, lhs_target="u", rhs_target="v"):
    """out[eid] = op(lhs[sel(lhs_target, eid)], rhs[sel(rhs_target, eid)]), edge-id order."""
    if op not in ("add", "sub", "mul", "div", "dot", "copy_lhs", "copy_rhs"):
        raise DGLError("gsddmm: unknown op %s" % op)
    if lhs_target not in _TARGET or rhs_target not in _TARGET:
        raise DGLError("gsddmm: targets must be one of u, e, v")
    use_lhs = op != "copy_rhs"
    use_rhs = op != "copy_lhs"
    if use_lhs and lhs is None:
        raise DGLError("gsddmm: op %s needs lhs data" % op)
    if use_rhs and rhs is None:
        raise DGLError("gsddmm: op %s needs rhs data" % op)
    if use_lhs and use_rhs and lhs.dtype != rhs.dtype:
        raise DGLError("The operands data type don't match: {} and {}, please convert them to the same "
                       "type.".format(lhs.dtype, rhs.dtype))
    lhs = lhs if use_lhs else None
    rhs = rhs if use_rhs else None
    dtype = _dtype_code("gsddmm", op == "dot" and lhs_target == "u" and rhs_target == "v" and "csc" in gidx.formats(),
                        lhs, rhs)
    dev = _capi.require_cuda(lhs, rhs, gidx.src)
    n_of = {"u": gidx.n_src, "e": gidx.n_edges, "v": gidx.n_dst}
    expand_lhs = expand_rhs = False
    if use_lhs:
        if lhs.shape[0] != n_of[lhs_target]:
            raise DGLError("gsddmm: lhs has %d rows, target '%s' has %d" % (lhs.shape[0], lhs_target, n_of[lhs_target]))
        if lhs.dim() == 1:
            lhs, expand_lhs = lhs.unsqueeze(-1), True
        lhs = lhs.contiguous()
    if use_rhs:
        if rhs.shape[0] != n_of[rhs_target]:
            raise DGLError("gsddmm: rhs has %d rows, target '%s' has %d" % (rhs.shape[0], rhs_target, n_of[rhs_target]))
        if rhs.dim() == 1:
            rhs, expand_rhs = rhs.unsqueeze(-1), True
        rhs = rhs.contiguous()
    feat_shape = infer_broadcast_shape(op, lhs.shape[1:] if use_lhs else (1,), rhs.shape[1:] if use_rhs else (1,))
    ref = lhs if use_lhs else rhs
    ring = lhs is not None and lhs.dim() == 2 and lhs.shape[1] * 2 >= 1024 and lhs.shape[1] % 2 == 0 and gidx.n_edges >= (1 << 18)
    if dtype == _capi.BF16 and lhs.dim() == 2 and lhs.shape[1] >= 32 and lhs.shape[1] % 8 and not ring:
        pad = 8 - lhs.shape[1] % 8  # zero columns do not change the dot product; 128-bit loads instead of 32-bit
        return _gsddmm(gidx, op, torch.nn.functional.pad(lhs, (0, pad)), torch.nn.functional.pad(rhs, (0, pad)),
                       lhs_target, rhs_target)
    out = None
    out_numel = gidx.n_edges
    for s_ in feat_shape:
        out_numel *= s_
    if gidx.n_edges > 0 and out_numel > 0:
        o = _capi.ops()
        ndim, ls, rs = _shapes_for_abi(op, lhs, rhs)
        lt, rt = _TARGET[lhs_target], _TARGET[rhs_target]
        fmts = gidx.formats()
        use_csr = ("csc" in fmts) and ((lhs_target == "u" and rhs_target == "v") or "coo" not in fmts)
        # rows of <= NARROW_COO_MAX_FLOATS floats with the same shape on both sides (attention logits, (N,1) scores):
        # one thread per edge over the COO -- coalesced ids and results in edge-id order, no permutation, L2-resident
        # gathers (csrc/sddmm.cu sddmm_coo_narrow_kernel); the destination-major kernel is for wide rows
        if use_csr and "coo" in fmts and dtype == _capi.F32 and NARROW_COO_MAX_FLOATS > 0:
            same = (not (use_lhs and use_rhs)) or tuple(lhs.shape[1:]) == tuple(rhs.shape[1:])
            row_floats = 1
            for s_ in ref.shape[1:]:
                row_floats *= s_
            if same and row_floats <= NARROW_COO_MAX_FLOATS and (op != "dot" or ref.dim() == 2):
                use_csr = False
        if use_csr:
            csc = gidx.csc()
            width = 1
            for s_ in ref.shape[1:]:
                width *= s_
            hub, hub_launches = _hub_pack(csc.hubs(_hub_threshold(width)), csc)
            # narrow results (u_dot_v, u_add_v on (N,H,1) scores) on a shuffled graph: the kernel writes them in
            # staged order (stores stay inside a 32 K-slot window), one pass then puts them in edge-id order
            plan = None
            if lhs_target == "u" and rhs_target == "v" and dtype == _capi.F32:
                plan = _stage_plan(csc, out_numel // gidx.n_edges)
            out = _capi.call(o.gsddmm_csr, csc.indptr, csc.indices, plan[1] if plan is not None else csc.eids, csc.n_cols,
                             _capi.OPS[op], lt, rt, lhs, rhs, list(feat_shape), ls, rs, *hub)
            _capi.count_launch(1 + hub_launches)
            if plan is not None:
                out = _stage_move(plan, out, False)
        else:
            if dtype != _capi.F32:
                raise DGLError("gsddmm_coo: only f32 is implemented")
            s32, d32 = gidx.coo32()
            out = _capi.call(o.gsddmm_coo, s32, d32, gidx.n_src, gidx.n_dst, _capi.OPS[op], lt, rt, lhs, rhs,
                             list(feat_shape), ls, rs)
            _capi.count_launch(1)
    if out is None:
        out = torch.empty((gidx.n_edges,) + tuple(feat_shape), dtype=ref.dtype, device=dev)
    if (expand_lhs or not use_lhs) and (expand_rhs or not use_rhs):
        out = out.squeeze(-1)
    return out


def _softmax_hub(csc, heads):
    """Hub-row segments for edge_softmax: (hub arguments without light_indptr, extra launches)."""
    thr = int(HUB_THRESHOLD) if HUB_THRESHOLD is not None else _capi.ops().default_hub_threshold(2, int(heads))
    info = csc.hubs(thr)
    if info is None or heads > 32:
        return (None, None, None, []), 0
    p = info.pack()
    return (p[0], p[1], p[2], p[5]), 3


def _edge_softmax_fwd(gidx, logits):
    """softmax over each destination's in-edges; logits (E, ...) in edge-id order."""
    _check_float32(logits)
    _capi.require_cuda(logits, gidx.src)
    if logits.shape[0] != gidx.n_edges:
        raise DGLError("edge_softmax: expect %d edge rows, got %d" % (gidx.n_edges, logits.shape[0]))
    logits = logits.contiguous()
    if gidx.n_edges == 0:
        return torch.empty_like(logits)
    heads = logits.numel() // gidx.n_edges
    csc = gidx.csc()
    hub, hub_launches = _softmax_hub(csc, heads)
    plan = _stage_plan(csc, heads, STAGE_MAX_ROW_FLOATS_SOFTMAX)
    eids = csc.eids
    if plan is not None:
        logits, eids = _stage_move(plan, logits, True), plan[1]
    out = _capi.call(_capi.ops().edge_softmax_fwd, csc.indptr, eids, logits, heads, *hub)
    _capi.count_launch(1 + hub_launches)
    if plan is not None:
        out = _stage_move(plan, out, False)
    return out


def _edge_softmax_bwd(gidx, out, grad_out):
    _check_float32(out, grad_out)
    _capi.require_cuda(out, grad_out, gidx.src)
    out = out.contiguous()
    grad_out = grad_out.contiguous()
    if gidx.n_edges == 0:
        return torch.empty_like(out)
    heads = out.numel() // gidx.n_edges
    csc = gidx.csc()
    hub, hub_launches = _softmax_hub(csc, heads)
    plan = _stage_plan(csc, heads, STAGE_MAX_ROW_FLOATS_SOFTMAX)
    eids = csc.eids
    if plan is not None:
        out, grad_out, eids = _stage_move(plan, out, True), _stage_move(plan, grad_out, True), plan[1]
    grad = _capi.call(_capi.ops().edge_softmax_bwd, csc.indptr, eids, out, grad_out, heads, *hub)
    _capi.count_launch(1 + hub_launches)
    if plan is not None:
        grad = _stage_move(plan, grad, False)
    return grad


GAT_HUB_SEGMENTS = True  # False: one CTA per hub row (the pre-segmentation path, kept for comparison)


def _gat_hub(view, H, F, launches):
    """Hub rows of a CSRView for the fused GAT kernels: (hub arguments without light_indptr, extra launches)."""
    info = view.hubs(_hub_threshold(H * F, row_cta=True))
    if info is None:
        return (None, None, None, []), 0
    p = info.pack()
    return (p[0], p[1], p[2], p[5]), (launches if GAT_HUB_SEGMENTS else 1)


def _gat_fwd(gidx, ft, el, er, slope, dropout_p, seed, want_scores=False, eids=None):
    """Fused GAT attention forward.  ft (n_src,H,F), el (n_src,H), er (n_dst,H) ->
    rst (n_dst,H,F), row_max, row_sum (n_dst,H) [, scores (E,H)].  `eids` (int32, CSC order) overrides
    the edge ids that key the dropout mask (row-partitioned graphs pass GLOBAL edge ids so the
    forward block and the backward block of different ranks regenerate the same mask)."""
    _check_float32(ft, el, er)
    _capi.require_cuda(ft, el, er, gidx.src)
    ft, el, er = ft.contiguous(), el.contiguous(), er.contiguous()
    H, F = ft.shape[1], ft.shape[2]
    csc = gidx.csc()
    hub, hub_launches = _gat_hub(csc, H, F, 4)
    rst, row_max, row_sum, scores = _capi.call(
        _capi.ops().gat_fwd, csc.indptr, csc.indices, eids if eids is not None else csc.eids, ft, el, er, float(slope),
        float(dropout_p), int(seed), bool(want_scores), bool(GAT_HUB_SEGMENTS), *hub)
    if gidx.n_dst:
        _capi.count_launch(1 + hub_launches)
    return rst, row_max, row_sum, (scores if want_scores else None)


def _gat_bwd_dst(gidx, ft, el, er, row_max, row_sum, grad_rst, slope, dropout_p, seed, eids=None):
    """Destination pass of the fused GAT backward on gidx's CSC -> (row_pack (n_dst,H,4), grad_er)."""
    _capi.require_cuda(ft, el, er, grad_rst, gidx.src)
    grad_rst = grad_rst.contiguous()
    H, F = ft.shape[1], ft.shape[2]
    csc = gidx.csc()
    hub, hub_launches = _gat_hub(csc, H, F, 2)
    row_pack, grad_er = _capi.call(
        _capi.ops().gat_bwd_dst, csc.indptr, csc.indices, eids if eids is not None else csc.eids, ft, el, er, row_max,
        row_sum, grad_rst, float(slope), float(dropout_p), int(seed), bool(GAT_HUB_SEGMENTS), *hub)
    if gidx.n_dst:
        _capi.count_launch(1 + hub_launches)
    return row_pack, grad_er


def _gat_bwd_src(csr, ft, el, row_pack, grad_rst, slope, dropout_p, seed, eids=None):
    """Source pass of the fused GAT backward on a CSRView whose rows are the source nodes and whose
    indices are destination ids -> (grad_ft (n_src,H,F), grad_el (n_src,H)).  row_pack / grad_rst are
    indexed by destination id."""
    _capi.require_cuda(ft, el, row_pack, grad_rst)
    grad_rst = grad_rst.contiguous()
    H, F = ft.shape[1], ft.shape[2]
    hub, hub_launches = _gat_hub(csr, H, F, 3)
    grad_ft, grad_el = _capi.call(
        _capi.ops().gat_bwd_src, csr.indptr, csr.indices, eids if eids is not None else csr.eids, csr.n_cols, ft, el,
        row_pack, grad_rst, float(slope), float(dropout_p), int(seed), bool(GAT_HUB_SEGMENTS), *hub)
    if csr.n_rows:
        _capi.count_launch(1 + hub_launches)
    return grad_ft, grad_el


def _gat_bwd(gidx, ft, el, er, row_max, row_sum, grad_rst, slope, dropout_p, seed):
    """Fused GAT attention backward -> (grad_ft, grad_el, grad_er)."""
    row_pack, grad_er = _gat_bwd_dst(gidx, ft, el, er, row_max, row_sum, grad_rst, slope, dropout_p, seed)
    grad_ft, grad_el = _gat_bwd_src(gidx.csr(), ft, el, row_pack, grad_rst, slope, dropout_p, seed)
    return grad_ft, grad_el, grad_er


# ------------------------------------------------------------------ batched small graphs (csrc/small_graph.cu)
def _gcn_msg_sum_fwd(gidx, x, w, c_src, c_dst):
    """out[v] = sum_{e=(u->v)} (c_src[u] c_dst[v]) relu(x[u] + w[e]) over the CSC, one launch."""
    _check_float32(x, w, c_src, c_dst)
    _capi.require_cuda(x, w, c_src, c_dst, gidx.src)
    if w.shape[0] != gidx.n_edges or x.shape[0] != gidx.n_src:
        raise DGLError("gcn_norm_relu_sum: expect x with %d rows and w with %d rows, got %d and %d"
                       % (gidx.n_src, gidx.n_edges, x.shape[0], w.shape[0]))
    csc = gidx.csc()
    out = _capi.call(_capi.ops().gcn_msg_sum_fwd, csc.indptr, csc.indices, csc.eids, x.contiguous(), w.contiguous(),
                     c_src.contiguous().view(-1), c_dst.contiguous().view(-1))
    _capi.count_launch(1)
    return out


def _gcn_msg_sum_bwd(gidx, x, w, c_src, c_dst, grad_out):
    """(grad_x, grad_w) over the CSR, one launch; grad_w is zero-filled first when the graph has edge slots outside
    every CSR row (fixed-size padded batches: indptr[-1] < number of edge slots)."""
    csr = gidx.csr()
    padded = bool(gidx._c.get("padded_edge_slots", False))
    gx, gw = _capi.call(_capi.ops().gcn_msg_sum_bwd, csr.indptr, csr.indices, csr.eids, x.contiguous(), w.contiguous(),
                        c_src.contiguous().view(-1), c_dst.contiguous().view(-1), grad_out.contiguous(), padded)
    _capi.count_launch(1)
    return gx, gw
