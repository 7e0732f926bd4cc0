"""Summarise an ncu --page source --csv dump: per kernel, top source lines by stall samples."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = list(csv.reader(open(path)))
kern = None; hdr = None; cur_file = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": kern = r[1]; agg.setdefault(kern, []); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            int(d["# Samples"] or 0); int(d["Instructions Executed"] or 0)
        except ValueError:
            continue
        agg[kern].append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), cur_file.split("/")[-1], r[0], r[1][:110], d.get("stall_long_sb","0"), d.get("stall_short_sb","0"), d.get("stall_wait","0")))
for k, v in agg.items():
    tot = sum(x[0] for x in v) or 1; ti = sum(x[1] for x in v)
    print("==", k[:100], "samples", tot, "inst", ti)
    for s, i, f, ln, src, lsb, ssb, w in sorted(v, reverse=True)[:topn]:
        print("  %5.1f%% inst %5.1f%% %s:%s  long_sb=%s short_sb=%s wait=%s | %s" % (100.0*s/tot, 100.0*i/max(ti,1), f, ln, lsb, ssb, w, src.strip()))
