# window-kernel edge_softmax: parity tests, then A/B against the row kernel (DGLB_ESM_WINDOW=0) on dst-sorted graphs
set -x
timeout 300 python -m pytest tests/test_gpu_softmax_gat.py -x -q -k "edge_softmax" 2>&1 | tail -5 > gpurun_out/esm_window_tests.txt; cat gpurun_out/esm_window_tests.txt
for shape in reddit ogbn-products; do
for w in 0 1; do
DGLB_ESM_WINDOW=$w timeout 200 python examples/op_microbench.py --shape $shape --order dst_sorted --widths "" --softmax-heads 1,2,4,8 2>/dev/null | sed "s/^/{\"window\": $w} /" >> gpurun_out/esm_window_ab.jsonl
done; done
cat gpurun_out/esm_window_ab.jsonl
