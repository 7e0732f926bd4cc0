"""NCCL bandwidth probe on the box (torchrun): all_gather_into_tensor and ring-style batch_isend_irecv
for the shard sizes bench.py moves.  Prints GB/s received per rank."""
import os
import sys
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for total_mb in (60, 240, 561):
    n = total_mb * 1024 * 1024 // 4 // world
    x = torch.rand(n, device=dev)
    out = torch.empty(n * world, device=dev)
    t = timeit(lambda: dist.all_gather_into_tensor(out, x))
    recv = n * 4 * (world - 1)
    msg = "all_gather total %4d MB: %.3f ms, %.0f GB/s recv per rank" % (total_mb, t, recv / t / 1e6)

    def ring():
        works = []
        for k in range(1, world):
            s, d = (rank + k) % world, (rank - k) % world
            works += dist.batch_isend_irecv([dist.P2POp(dist.isend, x, d), dist.P2POp(dist.irecv, out[s * n:(s + 1) * n], s)])
        for w in works:
            w.wait()
    t2 = timeit(ring)
    msg += " | p2p rounds: %.3f ms, %.0f GB/s" % (t2, recv / t2 / 1e6)

    def allp2p():
        ops = []
        for k in range(1, world):
            s, d = (rank + k) % world, (rank - k) % world
            ops += [dist.P2POp(dist.isend, x, d), dist.P2POp(dist.irecv, out[s * n:(s + 1) * n], s)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    t3 = timeit(allp2p)
    msg += " | one p2p batch: %.3f ms, %.0f GB/s" % (t3, recv / t3 / 1e6)
    if rank == 0:
        print(msg, flush=True)
dist.destroy_process_group()
