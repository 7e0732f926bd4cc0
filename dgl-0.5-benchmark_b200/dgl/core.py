"""Translation of update_all / apply_edges into gspmm / gsddmm calls.

Mirrors upstream DGL v0.6.1 python/dgl/core.py (message_passing, invoke_gspmm, invoke_gsddmm,
invoke_edge_udf): a (built-in message, built-in reduce) pair is ONE gspmm; a user-defined message
function materialises its (E, ...) messages with torch gathers and is then reduced by a copy_e
gspmm (main_dgl_molhiv_gcn.py:46,50-52); a lone built-in message function is ONE gsddmm.
"""
from . import function as fn
from . import ops
from ._capi import DGLError


def is_builtin(func):
    return isinstance(func, fn.BuiltinFunction)


class EdgeBatch:
    """What a message UDF receives: edges.src[...] / edges.dst[...] / edges.data[...]."""

    def __init__(self, graph):
        self._g = graph
        self._src = self._dst = None

    class _Gathered:
        def __init__(self, frame, index):
            self._frame, self._index, self._cache = frame, index, {}

        def __getitem__(self, key):
            if key not in self._cache:
                self._cache[key] = self._frame[key].index_select(0, self._index)
            return self._cache[key]

        def __contains__(self, key):
            return key in self._frame

    @property
    def src(self):
        if self._src is None:
            self._src = EdgeBatch._Gathered(self._g.srcdata, self._g._graph.src.long())
        return self._src

    @property
    def dst(self):
        if self._dst is None:
            self._dst = EdgeBatch._Gathered(self._g.dstdata, self._g._graph.dst.long())
        return self._dst

    @property
    def data(self):
        return self._g.edata

    def edges(self):
        return self._g.edges(form="all")

    def batch_size(self):
        return self._g.number_of_edges()

    def __len__(self):
        return self._g.number_of_edges()


def _frame_of(graph, target):
    return {"u": graph.srcdata, "v": graph.dstdata, "e": graph.edata}[target]


def invoke_edge_udf(graph, func):
    out = func(EdgeBatch(graph))
    if not isinstance(out, dict):
        raise DGLError("a message / edge UDF must return a dict of tensors")
    return out


def invoke_gsddmm(graph, func):
    """apply_edges with a built-in: one gsddmm."""
    if isinstance(func, fn.BinaryMessageFunction):
        x = _frame_of(graph, func.lhs)[func.lhs_field]
        y = _frame_of(graph, func.rhs)[func.rhs_field]
        z = ops.gsddmm(graph, func.binary_op, x, y, lhs_target=func.lhs, rhs_target=func.rhs)
    else:
        x = _frame_of(graph, func.target)[func.in_field]
        if func.target == "e":
            z = x
        elif func.target == "u":
            z = ops.gsddmm(graph, "copy_lhs", x, None)
        else:
            z = ops.gsddmm(graph, "copy_rhs", None, x)
    return {func.out_field: z}


def invoke_gspmm(graph, mfunc, rfunc, edata=None):
    """update_all with built-in message and reduce functions: one gspmm."""
    if mfunc.out_field != rfunc.msg_field:
        raise DGLError('Cannot find message field "{}" (the message function writes "{}").'
                       .format(rfunc.msg_field, mfunc.out_field))
    if isinstance(mfunc, fn.BinaryMessageFunction):
        x = _frame_of(graph, mfunc.lhs)[mfunc.lhs_field]
        y = _frame_of(graph, mfunc.rhs)[mfunc.rhs_field]
        if mfunc.binary_op == "dot":
            raise DGLError("dot is not a valid message op for update_all")
        if (mfunc.lhs, mfunc.rhs) == ("u", "e"):
            z = ops.gspmm(graph, mfunc.binary_op, rfunc.name, x, y)
        elif (mfunc.lhs, mfunc.rhs) == ("e", "u") and mfunc.binary_op in ("add", "mul"):
            z = ops.gspmm(graph, mfunc.binary_op, rfunc.name, y, x)
        else:
            # messages that involve the destination end point: materialise them per edge, then reduce
            m = ops.gsddmm(graph, mfunc.binary_op, x, y, lhs_target=mfunc.lhs, rhs_target=mfunc.rhs)
            z = ops.gspmm(graph, "copy_rhs", rfunc.name, None, m)
    else:
        x = _frame_of(graph, mfunc.target)[mfunc.in_field]
        if mfunc.target == "u":
            z = ops.gspmm(graph, "copy_lhs", rfunc.name, x, None)
        elif mfunc.target == "e":
            z = ops.gspmm(graph, "copy_rhs", rfunc.name, None, x)
        else:
            raise DGLError("copy_v is not a valid message function for update_all")
    return {rfunc.out_field: z}


def message_passing(graph, mfunc, rfunc, afunc=None):
    if not is_builtin(rfunc):
        raise DGLError("user-defined reduce functions are not supported; use fn.sum/mean/max/min")
    if is_builtin(mfunc):
        out = invoke_gspmm(graph, mfunc, rfunc)
    else:
        msgs = invoke_edge_udf(graph, mfunc)
        if rfunc.msg_field not in msgs:
            raise DGLError('Cannot find message field "{}".'.format(rfunc.msg_field))
        out = {rfunc.out_field: ops.gspmm(graph, "copy_rhs", rfunc.name, None, msgs[rfunc.msg_field])}
    if afunc is not None:
        raise DGLError("apply_node_func is not supported; apply it to dstdata after update_all")
    return out
