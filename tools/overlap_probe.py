"""Does splitting the environments of one GPU over K streams pay?  K BatchedEnv objects of E/K envs, each stepping
on its own CUDA stream, against one BatchedEnv of E envs on one stream (same total work).  The composite kernel is
issue-bound, the kernels in front of it (physics, LBS, preprocess, emit, sort) are latency / memory-bound: on separate
streams the front end of one group runs under the compositing of another.
    python tools/overlap_probe.py [--envs 256] [--groups 1 2 4] [--steps 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=256)
ap.add_argument("--groups", type=int, nargs="+", default=[1, 2, 4])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--fast", action="store_true")
ap.add_argument("--split", action="store_true", help="one high-priority stream for everything up to the sort, one low-priority stream for the compositing kernels")
a = ap.parse_args()
dev = torch.device("cuda", 0)
for K in a.groups:
    E = a.envs // K
    envs = [BatchedEnv(EnvBatchConfig(E=E, env_offset=k * E, success_start_frame=0, fast_composite=a.fast), dev) for k in range(K)]
    streams = [torch.cuda.Stream(dev) for _ in range(K)]
    t = lambda x: torch.tensor(np.ascontiguousarray(x), device=dev)
    feed = [[(tuple(t(c) for c in e.make_commands(f)), t(e.make_link_poses(f))) for f in range(a.warmup + a.steps)] for e in envs]
    main = torch.cuda.current_stream(dev)

    hi, lo = torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev, priority=0)

    def run(f0, n):
        if a.split:
            hi.wait_stream(main); lo.wait_stream(main)
            for f in range(f0, f0 + n):
                for k, e in enumerate(envs):
                    with torch.cuda.stream(hi):
                        e.step(command=feed[k][f][0], link_pose=feed[k][f][1], composite_stream=lo)
            main.wait_stream(hi); main.wait_stream(lo)
            return
        for s in streams:
            s.wait_stream(main)
        for f in range(f0, f0 + n):
            for k, e in enumerate(envs):
                with torch.cuda.stream(streams[k]):
                    e.step(command=feed[k][f][0], link_pose=feed[k][f][1])
        for s in streams:
            main.wait_stream(s)

    run(0, a.warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    run(a.warmup, a.steps)
    e1.record(main)
    torch.cuda.synchronize()
    for e in envs:
        e.check()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"split={a.split} groups={K} envs/group={E}: {ms:.3f} ms/step, {a.envs / ms * 1e3:.0f} env.step/s, "
          f"checksum {sum(float(e.color.double().sum()) for e in envs):.3f}", flush=True)
    del envs, feed
    torch.cuda.empty_cache()
