"""Peer-copy bandwidth through symmetric memory, every rank at once in a ring pattern (what the row-partitioned
exchange does): pull (local <- peer) vs push (peer <- local), 1 or 2 copy streams, a few shard sizes.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 examples/p2p_copy_bench.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    for mb in (8, 70):
        n = mb * (1 << 20) // 4
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        views = [hdl.get_buffer(r, (n,), torch.float32) if r != rank else t for r in range(world)]
        local = torch.randn(world, n, device=dev)
        streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]
        for mode in ("pull", "push"):
            for ns in (1, 2, 4):
                def run():
                    main_s = torch.cuda.current_stream()
                    ready = torch.cuda.Event(); ready.record(main_s)
                    for s in streams[:ns]:
                        s.wait_event(ready)
                    for k in range(1, world):
                        s = streams[(k - 1) % ns]
                        with torch.cuda.stream(s):
                            if mode == "pull":
                                local[k].copy_(views[(rank + k) % world], non_blocking=True)
                            else:
                                views[(rank - k) % world].copy_(local[0], non_blocking=True)
                    for s in streams[:ns]:
                        main_s.wait_stream(s)
                hdl.barrier(channel=0)
                run()
                hdl.barrier(channel=0)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    run()
                b.record()
                torch.cuda.synchronize()
                hdl.barrier(channel=0)
                ms = a.elapsed_time(b) / 5
                tt = torch.tensor([ms], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(json.dumps({"world": world, "shard_mb": mb, "mode": mode, "streams": ns, "ms": round(tt.item(), 4),
                                      "gbs_per_rank": round((world - 1) * n * 4 / tt.item() / 1e6, 1)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
