"""Run an UNCHANGED reference script against this package:

    python dgl-0.5-benchmark_b200/run_reference.py baseline/_ref/kernel/dgl-new.py -g 0
    python dgl-0.5-benchmark_b200/run_reference.py baseline/_ref/end_to_end/full_graph/node_classification/main_dgl_citation_sage.py --dataset cora

Puts this directory first on sys.path (so `import dgl`, `import ogb`, `import torch_sparse` resolve to
the stand-ins here), then the script's own directory (so its `from utils import ...` keeps working),
and executes the script as __main__.  On exit it reports on stderr how many calls crossed the C-ABI
into lib/libdglb200.so, so a caller can tell that the sm_100a kernels (and nothing else) did the work.
"""
import atexit
import os
import runpy
import sys


def _report():
    mod = sys.modules.get("dgl._capi")
    if mod is not None:
        sys.stderr.write("[dgl-b200] C-ABI compute calls: %d (library %s)\n"
                         % (mod.launches(), ("%s via %s" % (mod.LIB_PATH, mod.TORCH_LIB_PATH)) if mod._ops is not None
                            else "NOT LOADED"))
        sys.stderr.flush()


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    sys.argv = [script] + sys.argv[2:]
    sys.path.insert(0, os.path.dirname(script))
    sys.path.insert(0, here)
    atexit.register(_report)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
