"""Rough MLP check on SASS: for every kernel in a .so, the longest run of vector/scalar global loads
issued with no floating-point consumer (FADD/FFMA/FMUL/FMNMX/FSETP) in between.  A run of 1-2 in a
gather loop means the compiler serialised load->use pairs (seen once in sddmm_dot_kernel)."""
import re, subprocess, sys, collections
so = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None; best = collections.OrderedDict(); run = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); best[name] = 0; run = 0; continue
    if name is None: continue
    if re.search(r"\bLDG\.E(\.(64|128))?\.CONSTANT", line) or re.search(r"\bLDG\.E\.(64|128)\b", line):
        run += 1; best[name] = max(best[name], run)
    elif re.search(r"\b(FADD|FFMA|FMUL|FMNMX|FSETP)\b", line):
        run = 0
import subprocess as sp
for k, v in best.items():
    d = sp.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    if pat in d:
        print("%3d  %s" % (v, d[:120]))
