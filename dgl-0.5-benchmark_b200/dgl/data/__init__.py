"""dgl.data -- synthetic stand-ins for the datasets the reference scripts load (the real datasets
are not available offline): same node / edge counts, feature widths, class counts and split API."""
from . import synthetic  # noqa: F401
from . import utils  # noqa: F401
try:
    from .datasets import *  # noqa: F401,F403
except ImportError:  # pragma: no cover
    pass
