"""dgl.utils -- the one helper the in-scope scripts import (main_dgl_citation_sage.py:16,29)."""
from collections.abc import Mapping


def expand_as_pair(input_, g=None):
    """(src_input, dst_input) from a single value or a pair; for a block given one tensor, the
    destination part is its first number_of_dst_nodes rows (upstream dgl/utils/internal.py)."""
    if isinstance(input_, tuple):
        return input_
    if g is not None and getattr(g, "is_block", False):
        if isinstance(input_, Mapping):
            return input_, {k: v[: g.number_of_dst_nodes()] for k, v in input_.items()}
        return input_, input_[: g.number_of_dst_nodes()]
    return input_, input_
