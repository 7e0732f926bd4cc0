run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 "$@" > gpurun_out/bench_n8_$tag.json 2> gpurun_out/bench_n8_$tag.err; echo $tag rc=$?; tail -c 300 gpurun_out/bench_n8_$tag.err | grep -v NCCL; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n8_$tag.json"))
    print("$tag", round(d["ms_per_step"],3), round(d["value"]), d["config"].get("step_launch"), "e2e", round(d["e2e"]["ms_per_step"],2), [(k["op"][:6],k["D"],round(k["ms"],3)) for k in d["kernels"]])
    if "epochs" in d: print({k:(round(v["epoch_s"]*1e3,2)) for k,v in d["epochs"].items()})
except Exception as ex: print("parse failed", ex)
PY
}
run wf_default --no-extras
run wf_1_2_2_2_1 --no-extras --peer-groups 1,2,2,2,1
