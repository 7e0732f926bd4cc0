"""dgl.nn -- the layers the reference scripts import (PyTorch backend only)."""
from .pytorch import *  # noqa: F401,F403
from . import pytorch  # noqa: F401
