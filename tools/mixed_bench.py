"""BASELINE configs[4] on one rank: the mixed workload -- rope : sloth : T-block environments in the ratio 2 : 1 : 1
(2048 envs = 1024 + 512 + 512 over 8 GPUs = 128 + 64 + 64 per GPU), 512x512 render of 200k Gaussians per env,
10 substeps per step.  Every rank gets the same scene-type ratio (shard.interleave_scene_types); a step advances all
three batches back to back on one stream.  Launch like bench.py (plain python for one GPU, torchrun for N):

    python tools/mixed_bench.py [--envs 256] [--steps 10] [--warmup 3] [--res 512 512] [--gaussians 200000]

Prints one JSON line (rank 0): whole-job env.step/s (max over ranks of the CUDA-event time), per-scene ms, the
metrics all-gather {steps, seconds, checksum(x), checksum(rgb), episodes_succeeded, frames_passed} per rank."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402  (ClockSampler, dist_env)
from real2sim_eval_b200 import shard   # noqa: E402
from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256, help="environments per GPU (split 2:1:1)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", type=int, nargs=2, default=[512, 512])
    ap.add_argument("--gaussians", type=int, default=200_000)
    ap.add_argument("--substeps", type=int, default=10)
    a = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = bench.dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    total = world * a.envs
    counts = {"rope": total // 2, "sloth": total // 4, "tblock": total - total // 2 - total // 4}
    mine = shard.interleave_scene_types(counts, world)[rank]
    W, H = a.res
    envs, offset = {}, rank * a.envs
    for name in ("rope", "sloth", "tblock"):
        E = mine.count(name)
        if E == 0:
            continue
        # rope and sloth are handled by the gripper, the T-block by the pusher rod (25,312 triangles)
        cfg = EnvBatchConfig(scene=name, E=E, W=W, H=H, n_substeps=a.substeps, P=a.gaussians, env_offset=offset,
                             gripper=name != "tblock", pusher=name == "tblock", success_start_frame=0)
        envs[name] = BatchedEnv(cfg, dev)
        offset += E
    n_frames = a.warmup + a.steps
    t = lambda x: torch.tensor(np.ascontiguousarray(x), device=dev)
    feed = {n: [(tuple(None if c is None else t(c) for c in e.make_commands(f)), t(e.make_link_poses(f)))
                for f in range(n_frames)] for n, e in envs.items()}

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = bench.ClockSampler(local)
    sampler.start()
    import time
    time.sleep(1.0)
    sampler.mark()
    for f in range(a.warmup):
        for n, e in envs.items():
            e.step(command=feed[n][f][0], link_pose=feed[n][f][1])
    for n, e in envs.items():
        e.check()               # nothing silently dropped: instance-list overflow, candidate-row overflow
    barrier()
    ev = {n: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
          for n in envs}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        f = a.warmup + k
        for n, e in envs.items():
            ev[n][k][0].record()
            e.step(command=feed[n][f][0], link_pose=feed[n][f][1])
            ev[n][k][1].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ms_max = shard.max_over_ranks(ms, dev)
    clocks = sampler.stop()
    for e in envs.values():
        e.check()
    per_scene = {n: round(float(np.mean([x.elapsed_time(y) for x, y in ev[n]])), 3) for n in envs}
    cx = sum(float(e.phys.x.double().sum()) for e in envs.values())
    crgb = sum(float(e.color.double().sum()) for e in envs.values())
    succ = sum(int(e.success.result()[0].sum()) for e in envs.values())
    hits = sum(int(e.success.result()[1].sum()) for e in envs.values())
    gathered = shard.gather_metrics([a.steps, ms / 1e3, cx, crgb, succ, hits], dev)
    line = {"metric": bench.METRIC, "value": total * a.steps / (ms_max / 1e3), "unit": bench.UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"mixed rope+sloth+T ({counts}) x {W}x{H}, {a.gaussians} Gaussians/env, "
                                   f"{a.substeps} substeps/step (BASELINE configs[4] at {total} envs)",
                       "envs_per_gpu": {n: e.cfg.E for n, e in envs.items()}, "parallelism": f"env-shard x{world}"},
            "ms_per_step_by_scene": per_scene, "clocks": clocks,
            "comm": {"collective": "one all-gather (NCCL) of 6 float64 per rank after the timed region",
                     "payload_bytes_per_rank": 48, "data_path_collectives": 0},
            "metrics_allgather": {"per_rank": gathered, "fields": ["steps", "seconds", "checksum_x", "checksum_rgb",
                                                                   "episodes_succeeded", "frames_passed"]}}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
