"""dgl.function -- built-in message and reduce functions (upstream python/dgl/function/).

The reference scripts use fn.copy_src / fn.copy_u / fn.u_mul_e with fn.sum / fn.mean
(main_dgl_citation_sage.py:75-77, main_dgl_proteins_rgcn_for.py:52) and, through the layers,
fn.u_add_v / fn.u_dot_v / fn.max.  Built-ins are plain descriptors; dgl.core turns a
(message, reduce) pair into one gspmm call and a lone message built-in into one gsddmm call.
"""
import sys

__all__ = ["copy_u", "copy_src", "copy_e", "copy_edge", "sum", "mean", "max", "min",
           "BuiltinFunction", "BinaryMessageFunction", "CopyMessageFunction", "SimpleReduceFunction"]


class BuiltinFunction:
    """Marker base class."""


class BinaryMessageFunction(BuiltinFunction):
    def __init__(self, binary_op, lhs, rhs, lhs_field, rhs_field, out_field):
        self.binary_op, self.lhs, self.rhs = binary_op, lhs, rhs
        self.lhs_field, self.rhs_field, self.out_field = lhs_field, rhs_field, out_field

    @property
    def name(self):
        return "{}_{}_{}".format(self.lhs, self.binary_op, self.rhs)


class CopyMessageFunction(BuiltinFunction):
    def __init__(self, target, in_field, out_field):
        self.target, self.in_field, self.out_field = target, in_field, out_field

    @property
    def name(self):
        return "copy_{}".format(self.target)


class SimpleReduceFunction(BuiltinFunction):
    def __init__(self, name, msg_field, out_field):
        self.name, self.msg_field, self.out_field = name, msg_field, out_field


def copy_u(u, out):
    return CopyMessageFunction("u", u, out)


def copy_e(e, out):
    return CopyMessageFunction("e", e, out)


def copy_src(src, out):
    """Alias of copy_u (the spelling main_dgl_citation_sage.py:75 uses)."""
    return copy_u(src, out)


def copy_edge(edge, out):
    return copy_e(edge, out)


def sum(msg, out):  # noqa: A001  (the upstream name)
    return SimpleReduceFunction("sum", msg, out)


def mean(msg, out):
    return SimpleReduceFunction("mean", msg, out)


def max(msg, out):  # noqa: A001
    return SimpleReduceFunction("max", msg, out)


def min(msg, out):  # noqa: A001
    return SimpleReduceFunction("min", msg, out)


def _gen_binary():
    mod = sys.modules[__name__]
    for lhs in ("u", "v", "e"):
        for rhs in ("u", "v", "e"):
            if lhs == rhs:
                continue
            for op in ("add", "sub", "mul", "div", "dot"):
                name = "{}_{}_{}".format(lhs, op, rhs)

                def f(lhs_field, rhs_field, out, _op=op, _l=lhs, _r=rhs):
                    return BinaryMessageFunction(_op, _l, _r, lhs_field, rhs_field, out)
                f.__name__ = name
                setattr(mod, name, f)
                __all__.append(name)
    # src/dst/edge spellings of the 0.4-era API
    for op in ("add", "sub", "mul", "div", "dot"):
        for (l, ln), (r, rn) in ((("u", "src"), ("e", "edge")), (("e", "edge"), ("u", "src")),
                                 (("u", "src"), ("v", "dst")), (("v", "dst"), ("u", "src")),
                                 (("v", "dst"), ("e", "edge")), (("e", "edge"), ("v", "dst"))):
            setattr(mod, "{}_{}_{}".format(ln, op, rn), getattr(mod, "{}_{}_{}".format(l, op, r)))


_gen_binary()
