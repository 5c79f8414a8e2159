#!/usr/bin/env python
"""tools/d2h_bench.py -- pinned device->host copy bandwidth per rank, alone and with every rank copying at once.

    python tools/d2h_bench.py                                             # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
           tools/d2h_bench.py [--bind]

What bounds bench.py's `e2e` at N > 1 is the copy of the observations into pinned host memory (e2e
efficiency 0.31 at N=8 with 482 MB/step/GPU).  This measures the ceiling directly: each rank copies `--mb` MiB
from its GPU into a pinned buffer `--iters` times (CUDA events), first one rank at a time (the others idle), then
all ranks together after a barrier.  `--bind` pins each process to its GPU's NUMA node before allocating.
Prints one JSON line (rank 0): GB/s per rank alone, per rank concurrently, and the aggregate."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--bind", action="store_true")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    binding = None
    if a.bind:
        import bench
        binding = bench.bind_to_gpu_numa_node(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = a.mb << 20
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    dst = torch.empty(n, dtype=torch.uint8).pin_memory()
    back = torch.empty(n, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return n * a.iters / (e0.elapsed_time(e1) / 1e3) / 1e9

    d2h = lambda: dst.copy_(src, non_blocking=True)
    h2d = lambda: back.copy_(dst, non_blocking=True)
    alone_d2h = alone_h2d = 0.0
    for r in range(world):          # one rank at a time
        barrier()
        if r == rank:
            alone_d2h, alone_h2d = timed(d2h), timed(h2d)
    barrier()
    together_d2h = timed(d2h)       # every rank at once
    barrier()
    together_h2d = timed(h2d)
    barrier()
    vals = torch.tensor([alone_d2h, together_d2h, alone_h2d, together_h2d], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.empty_like(vals) for _ in range(world)]
        dist.all_gather(out, vals)
    else:
        out = [vals]
    if rank == 0:
        rows = [o.tolist() for o in out]
        print(json.dumps({
            "what": f"pinned D2H / H2D of {a.mb} MiB x {a.iters}, GB/s", "world": world, "bound": binding,
            "host_cpus": os.cpu_count(),
            "d2h_alone_gbs": [round(r[0], 1) for r in rows], "d2h_concurrent_gbs": [round(r[1], 1) for r in rows],
            "d2h_concurrent_aggregate_gbs": round(sum(r[1] for r in rows), 1),
            "h2d_alone_gbs": [round(r[2], 1) for r in rows], "h2d_concurrent_gbs": [round(r[3], 1) for r in rows],
            "h2d_concurrent_aggregate_gbs": round(sum(r[3] for r in rows), 1)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
