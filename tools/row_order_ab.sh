set -x
timeout 300 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_sddmm.py -x -q 2>&1 | tail -3
rm -f gpurun_out/r02_row_order_ab.jsonl
for mode in never always; do
for deg in powerlaw uniform; do
DGLB_ROW_ORDER=$mode timeout 200 python examples/op_microbench.py --shape reddit --degree $deg --widths 64,128,256 2>/dev/null | sed "s/^/{\"row_order\": \"$mode\"} /" >> gpurun_out/r02_row_order_ab.jsonl
done; done
python - <<EOF
import json
for l in open("gpurun_out/r02_row_order_ab.jsonl"):
    i=l.index("} ")+2; w=json.loads(l[:i]); d=json.loads(l[i:])
    print(w["row_order"], d["degree"], d["D"], {k:v["ms"] for k,v in d.items() if isinstance(v,dict)})
EOF
