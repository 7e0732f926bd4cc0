# final single-GPU evidence run (one gpurun call): smoke, tests, bench (+ reference arm), ncu launch list, full captures
set -x
python __graft_entry__.py smoke 2>&1 | tail -2
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final2_gputests.txt; cat gpurun_out/final2_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final2_bench_ref.json 2>/dev/null; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/plainA.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|sddmm|ring" -c 200 --csv --log-file gpurun_out/r02_launches_final2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncuA.log 2>&1
python examples/ring_tune.py --shape reddit --widths 602 --stages 0 --smem 0 --reps 2 > gpurun_out/plainB.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ring_kernel" -c 1 -f -o gpurun_out/r02_ring_d602_final2 python examples/ring_tune.py --shape reddit --widths 602 --stages 0 --smem 0 --reps 2 > gpurun_out/ncuB.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"edge_softmax" -c 1 -f -o gpurun_out/r02_esm_fwd_products_h4_after python examples/op_microbench.py --shape ogbn-products --order dst_sorted --widths "" --softmax-heads 4 > gpurun_out/ncuE.log 2>&1
ls -la gpurun_out | tail -8
