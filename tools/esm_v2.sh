# rewritten register path of edge_softmax: parity, timings (dst-sorted and shuffled), issue-slot utilisation of every kernel of the bench extras
set -x
timeout 300 python -m pytest tests/test_gpu_softmax_gat.py -x -q 2>&1 | tail -5 > gpurun_out/esm_v2_tests.txt; cat gpurun_out/esm_v2_tests.txt
rm -f gpurun_out/esm_v2.jsonl
for shape in reddit ogbn-products; do for order in dst_sorted shuffled; do
timeout 200 python examples/op_microbench.py --shape $shape --order $order --widths "" --softmax-heads 1,2,4,8 2>/dev/null >> gpurun_out/esm_v2.jsonl
done; done
cat gpurun_out/esm_v2.jsonl
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-epochs > gpurun_out/plainD.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum.per_cycle_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"dglb" -c 400 --csv --log-file gpurun_out/r02_issue_util_extras.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-epochs > gpurun_out/ncuD.log 2>&1
tail -3 gpurun_out/ncuD.log
