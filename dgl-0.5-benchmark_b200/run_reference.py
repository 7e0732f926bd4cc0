"""Run an UNCHANGED reference script against this package:

    python dgl-0.5-benchmark_b200/run_reference.py /path/to/kernel/dgl-new.py -g 0
    python dgl-0.5-benchmark_b200/run_reference.py /path/to/main_dgl_citation_sage.py --dataset cora

Puts this directory first on sys.path (so `import dgl`, `import ogb`, `import torch_sparse` resolve to
the stand-ins here), then the script's own directory (so its `from utils import ...` keeps working),
and executes the script as __main__.
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    sys.argv = [script] + sys.argv[2:]
    sys.path.insert(0, os.path.dirname(script))
    sys.path.insert(0, here)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
