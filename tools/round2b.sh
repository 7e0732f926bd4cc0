set -x
timeout 300 python -m pytest tests/test_gpu_small_graph.py tests/test_gpu_rgcn_batched.py -x -q 2>&1 | tail -25 > gpurun_out/small_graph_tests.txt; cat gpurun_out/small_graph_tests.txt
timeout 400 python examples/molhiv_bench.py > gpurun_out/r02_molhiv_v2.jsonl 2> gpurun_out/molhiv_v2.err; tail -5 gpurun_out/molhiv_v2.err; cat gpurun_out/r02_molhiv_v2.jsonl
rm -f gpurun_out/esm_ming.jsonl
for mg in 8 4 2; do
DGLB_ESM_MIN_G=$mg timeout 200 python examples/op_microbench.py --shape ogbn-products --order dst_sorted --widths "" --softmax-heads 1,2,4 2>/dev/null | sed "s/^/{\"min_g\": $mg} /" >> gpurun_out/esm_ming.jsonl
DGLB_ESM_MIN_G=$mg timeout 200 python examples/op_microbench.py --shape ogbn-arxiv --order dst_sorted --widths "" --softmax-heads 1,4 2>/dev/null | sed "s/^/{\"min_g\": $mg} /" >> gpurun_out/esm_ming.jsonl
done
cat gpurun_out/esm_ming.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum.per_cycle_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"spmm|sddmm|softmax|gat_|ring" -c 400 --csv --log-file gpurun_out/r02_issue_util_extras.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-epochs > gpurun_out/ncuD.log 2>&1
tail -c 300 gpurun_out/ncuD.log
