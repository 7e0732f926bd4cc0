set -x
timeout 400 python -m pytest tests/test_gpu_sddmm.py tests/test_gpu_autograd.py tests/test_gpu_softmax_gat.py tests/test_gpu_staged_edges.py tests/test_gpu_training.py -x -q 2>&1 | tail -4
rm -f gpurun_out/r02_narrow_coo_ab.jsonl
for m in 0 8; do
DGLB_NARROW_COO=$m timeout 200 python examples/op_microbench.py --shape reddit --widths "" --heads 1,4,8 2>/dev/null | sed "s/^/{\"narrow_coo\": $m} /" >> gpurun_out/r02_narrow_coo_ab.jsonl
DGLB_NARROW_COO=$m timeout 200 python examples/op_microbench.py --shape ogbn-products --widths "" --heads 1,4 2>/dev/null | sed "s/^/{\"narrow_coo\": $m} /" >> gpurun_out/r02_narrow_coo_ab.jsonl
done
python - <<EOF
import json
for l in open("gpurun_out/r02_narrow_coo_ab.jsonl"):
    i=l.index("} ")+2; w=json.loads(l[:i]); d=json.loads(l[i:])
    print(w["narrow_coo"], d["shape"], d["H"], {k:(v["ms"], v["frac"]) for k,v in d.items() if isinstance(v,dict)})
EOF
