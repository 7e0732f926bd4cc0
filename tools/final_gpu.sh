# final single-GPU evidence run (one gpurun call): tests, bench, ncu launch list + full capture of the dominant kernel,
# DRAM traffic of the staged per-edge ops on the products shape
set -x
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_gputests.txt; cat gpurun_out/final_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_ref.json 2>/dev/null; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/plainA.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dglb|ring_kernel" -c 200 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncuA.log 2>&1
python examples/ring_tune.py --shape reddit --widths 602 --stages 0 --smem 0 --reps 2 > gpurun_out/plainB.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ring_kernel" -c 2 -o gpurun_out/r02_ring_d602_final python examples/ring_tune.py --shape reddit --widths 602 --stages 0 --smem 0 --reps 2 > gpurun_out/ncuB.log 2>&1
python examples/op_microbench.py --shape ogbn-products --widths 64 --heads 1 --softmax-heads 1,4 > gpurun_out/plainC.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"edge_stage|edge_softmax|spmm_rows_kernel|sddmm_" -c 120 --csv --log-file gpurun_out/r02_staged_products_traffic.csv python examples/op_microbench.py --shape ogbn-products --widths 64 --heads 1 --softmax-heads 1,4 > gpurun_out/ncuC.log 2>&1
python examples/molhiv_bench.py 2>/dev/null | tail -3 > gpurun_out/final_molhiv.jsonl
python epoch_bench.py --configs cora_sage,arxiv_gat,arxiv_sage,reddit_sage,reddit_gat,products_sage,products_gat 2>/dev/null > gpurun_out/final_epochs_n1.jsonl
ls -la gpurun_out | tail -20
