set -x
for w in 0 1; do
DGLB_ESM_WINDOW=$w timeout 300 ncu --set full --clock-control none --import-source on -k regex:"edge_softmax" -c 1 -f -o gpurun_out/r02_esm_fwd_products_h4_w$w python examples/op_microbench.py --shape ogbn-products --order dst_sorted --widths "" --softmax-heads 4 > gpurun_out/ncu_esm_w$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
