#!/usr/bin/env python
"""bench.py -- env.step()/sec (physics + render) for parallel PhysTwin environments.

    python bench.py --gpus 1 --steps 5 --warmup 3              # this repository's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): rope PhysTwin (synthetic, N=2048 particles / S=32889
springs), 256 parallel envs per GPU, 10 substeps per step, one 512x512 RGB-D render per env of
200,000 Gaussians (10% bound to the particles).  A step is one pass of the hot path over all
envs: end-effector command -> per-substep finger tables + grasp hysteresis (device) -> per-frame
collision-graph rebuild -> 10 substeps -> LBS of the object Gaussians -> robot link re-posing -> render.
Multi-GPU: envs shard statically, one process per GPU, no data-path collective; one NCCL
all-gather of {steps, seconds, checksum(x), checksum(rgb)} at the end (weak scaling).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` is the same loop driven from HOST buffers: per step the end-effector commands, the
robot's link poses and the camera matrices are copied from pinned host memory, and particle states + rendered RGB-D are
copied back to pinned host memory, all inside the timed region.

`--impl reference` times the reference pipeline on the same box: the CPU restatement of
sim/physics/spring_mass_warp.py (oracle/physics_ref.c -- the reference's Warp path cannot run:
warp-lang is not installable offline) on all host cores, plus the UNMODIFIED reference CUDA
rasterizer (oracle/_ref, built from /root/reference) called once per env with its own blocking
num_rendered read-back, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "env_steps_per_sec"
UNIT = "env.step/s"
STAGES = ("preprocess", "scan", "emit", "tile_sort", "composite")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="rope")
    ap.add_argument("--envs", type=int, default=256, help="environments per GPU")
    ap.add_argument("--res", type=int, nargs=2, default=[512, 512], metavar=("W", "H"))
    ap.add_argument("--cameras", type=int, default=1)
    ap.add_argument("--substeps", type=int, default=10)
    ap.add_argument("--gaussians", type=int, default=200_000)
    ap.add_argument("--ref-envs", type=int, default=0, help="reference arm: envs per step (0 = host cores, <= 64)")
    ap.add_argument("--cpu-sample-envs", type=int, default=0, help="cpu_baseline sample size (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fast", action="store_true", help="skip the fast-composite variant measured beside the headline")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines, self.t0 = gpu_index, None, [], 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "25"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Samples arriving from now on count as 'under load' (call right before the measured span)."""
        self.t0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            if t < self.t0:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local: int):
    """Pin this process to the cores of its GPU's NUMA node BEFORE any pinned host memory is allocated, so the
    D2H targets live on the node the GPU's PCIe root port hangs off.  Returns a description for the JSON line.
    (On a single-node host -- the 8-GPU box of round 1 reports NUMA 0 / CPUs 0-31 for every GPU -- this is a no-op.)"""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        prop = torch.cuda.get_device_properties(local)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["pci"] = bdf
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if node < 0 or len(nodes) <= 1:
            info["numa_node"] = node
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(numa_node=node, cpus=len(allowed), bound=True)
    except Exception as ex:   # no sysfs / no permission: report and carry on unbound
        info["error"] = str(ex)[:80]
    return info


def config_label(args):
    """Which BASELINE.json config the arguments describe (configs[1] is the headline the metric is quoted on)."""
    W, H = args.res
    key = (args.scene, args.envs, W, H, args.cameras, args.substeps)
    if key == ("rope", 256, 512, 512, 1, 10):
        return "BASELINE configs[1]"
    if key == ("sloth", 64, 640, 480, 2, 10):
        return "BASELINE configs[2]"
    if args.scene == "tblock" and (W, H) == (256, 256):
        return "BASELINE configs[3] shape with rendering"
    return "custom shape"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    from real2sim_eval_b200 import _lib
    from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig
    from real2sim_eval_b200.shard import shard_envs

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    binding = bind_to_gpu_numa_node(local)      # before the first pinned allocation
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    W, H = args.res
    cfg = EnvBatchConfig(scene=args.scene, E=args.envs, W=W, H=H, cameras=args.cameras, n_substeps=args.substeps,
                         P=args.gaussians, env_offset=shard_envs(world * args.envs, world, rank).start,
                         success_start_frame=0)
    env = BatchedEnv(cfg, dev)
    lib = _lib.load()
    E, ns = cfg.E, cfg.n_substeps
    n_frames = args.warmup + args.steps + (0 if args.no_e2e else 2 * (args.warmup + args.steps)) + 2
    # host numpy, made before any timing: end-effector commands (19 floats per env) + the robot's FK link poses of every frame
    acts = [env.make_commands(f) + (env.make_link_poses(f),) for f in range(n_frames)]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    acts_pinned = [tuple(pin(a) for a in act) for act in acts]
    dev_bufs = tuple(torch.empty_like(t, device=dev) for t in acts_pinned[0])

    def upload(i):
        for d, h in zip(dev_bufs, acts_pinned[i]):
            d.copy_(h, non_blocking=True)
        return dev_bufs

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident loop: motion tables already on the device
    motions_dev = [tuple(t.to(dev) for t in act) for act in acts_pinned[:args.warmup + args.steps]]
    sampler = ClockSampler(local)   # samples clocks / throttle reasons from the warm-up through the e2e loop
    sampler.start()
    time.sleep(1.0)                 # let nvidia-smi start streaming before the measured span
    sampler.mark()
    for i in range(args.warmup):
        env.step(command=motions_dev[i][:5], link_pose=motions_dev[i][5])
    total, overflow = env.raster.status()
    if overflow:
        raise RuntimeError(f"instance capacity exceeded ({total} > {env.max_instances})")
    lib.r2s_raster_set_profile(1)
    stage_ms = np.zeros(len(STAGES))
    phys_ms = 0.0
    barrier()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pe1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pe2 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pe3 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pem = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pes = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    peg = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    import ctypes
    prof = (ctypes.c_float * 5)()
    ev0.record()
    for k in range(args.steps):
        m = motions_dev[args.warmup + k]
        # (same body as env.step, with events around the physics launch for the per-kernel report)
        peg[k].record()
        if env.phys.self_collision:
            env.phys.update_collision_graph()
        pem[k].record()
        env.eef.forward(*m[:5])
        env.x_prev4.copy_(env.phys.x4)
        pe0[k].record()
        env.phys.step()
        pe1[k].record()
        env.success.update(env.phys.x4)
        pes[k].record()
        env.lbs.forward(env.x_prev4, env.phys.x4, env.means3D)
        pe2[k].record()
        env.links.forward(m[5], env.means3D, env.rotations)
        pe3[k].record()
        env.raster.forward(env.means3D, env.opacities, viewmatrix=env.view, projmatrix=env.proj, campos=env.campos,
                           bg=env.bg, W=W, H=H, tanfovx=env.tanfovx, tanfovy=env.tanfovy, shs=env.shs,
                           scales=env.scales, rotations=env.rotations, sh_degree=0, z_threshold=0.05,
                           views_per_scene=cfg.cameras, max_instances=env.max_instances, out_color=env.color,
                           out_depth=env.depth, want_radii=False)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    _lib.check(lib.r2s_raster_get_profile(prof), "get_profile")   # stages of the LAST timed step
    stage_ms = np.array(list(prof), dtype=np.float64)
    phys_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(pe0, pe1)]))
    succ_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(pe1, pes)]))
    lbs_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(pes, pe2)]))
    links_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(pe2, pe3)]))
    eef_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(pem, pe0)]))   # incl. the x_prev copy
    grid_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(peg, pem)]))
    lib.r2s_raster_set_profile(0)
    total, overflow = env.raster.status()
    if overflow:
        raise RuntimeError(f"instance capacity exceeded in the timed region ({total} > {env.max_instances})")
    env.check()                                 # sticky overflow counter + candidate-row overflow, outside the timed region
    from real2sim_eval_b200 import shard
    ms_serial = shard.max_over_ranks(ms_total, dev)     # single stream, per-kernel events: the kernel accounting pass

    # ---- the same loop, pipelined: everything up to the sort on a high-priority stream, the compositing kernel on a
    # second stream (r2s_raster_args.composite_stream), so the latency / memory-bound front end of step k+1 runs under
    # the issue-bound compositing of step k.  Same kernels, same results.  This is the headline `value`; the serial loop
    # above is the kernel-accounting pass (per-kernel CUDA events need a single stream).
    hi, lo = torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev, priority=0)
    main_stream = torch.cuda.current_stream(dev)

    def pipelined(n, first):
        hi.wait_stream(main_stream); lo.wait_stream(main_stream)
        with torch.cuda.stream(hi):
            for k in range(n):
                m = motions_dev[first + k]
                env.step(command=m[:5], link_pose=m[5], composite_stream=lo)
        main_stream.wait_stream(hi); main_stream.wait_stream(lo)

    pipelined(args.warmup, 0)
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    q0.record()
    pipelined(args.steps, args.warmup)
    q1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_max = shard.max_over_ranks(q0.elapsed_time(q1), dev)
    value = world * E * args.steps / (ms_max / 1e3)     # the headline: the same K steps, pipelined over two streams
    serial = {"value": world * E * args.steps / (ms_serial / 1e3), "unit": UNIT, "ms_per_step": ms_serial / args.steps,
              "what": "the same K steps on ONE stream with CUDA events around every kernel: the pass the per-kernel times "
                      "and roofline figures come from (events inside the two-stream region would time co-running kernels)"}
    env.check()

    # ---- the same loop with the fast compositing variant (ex2.approx; 1e-4 relative contract, not bit-identical):
    # reported beside the headline, never as the headline
    fast_variant = None
    if not args.no_fast:
        env.cfg.fast_composite = True
        lib.r2s_raster_set_profile(1)
        for i in range(args.warmup):
            env.step(command=motions_dev[i][:5], link_pose=motions_dev[i][5])
        barrier()
        _lib.check(lib.r2s_raster_get_profile(prof), "get_profile")     # composite time of a serial step
        lib.r2s_raster_set_profile(0)
        pipelined(args.warmup, 0)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        pipelined(args.steps, args.warmup)
        f1.record()
        barrier()
        fms = shard.max_over_ranks(f0.elapsed_time(f1), dev)
        fast_variant = {"value": world * E * args.steps / (fms / 1e3), "unit": UNIT, "ms_per_step": fms / args.steps,
                        "composite_ms": round(float(prof[4]), 4),
                        "what": "composite_mode=FAST: log2(e) folded into the staged conic + ex2.approx instead of the IEEE "
                                "expf the bit-identical path keeps; held to 1e-4 relative vs the live reference in "
                                "tests/test_gpu_raster.py::test_fast_composite_variant_within_contract"}
        env.cfg.fast_composite = False
        env.check()

    # ---- end-to-end loop: host buffers in, host buffers out, copies inside the timed region.
    # Device->host copies of step k (observations + particle state) run on a side stream while step k+1
    # computes into the other output buffer (double buffering); every byte is still moved and waited for
    # inside the timed region.  Observation format: "u8" = RGB as [B,H,W,3] uint8 (what the reference's
    # evaluation loop holds on the host, experiments/eval_policy.py:248) + float32 depth; "f32" = the
    # float32 CHW colour image + depth (the device-side tensors, 16 B/pixel).  The headline `e2e` is "u8";
    # the float variant is reported next to it.
    e2e = None
    if not args.no_e2e:
        N = env.base.N
        import ctypes
        main = hi                                    # the step runs on the high-priority stream, its compositing on `lo`
        cs = torch.cuda.Stream(dev)
        phys_lib = env.phys.lib
        h2d = sum(t.numel() * 4 for t in acts_pinned[0]) + (env.view_h.numel() + env.proj_h.numel() + env.campos_h.numel()) * 4

        def run_e2e(fmt, base_i, with_depth):
            hostbuf, devbuf = [], []
            for slot in range(2):
                hb = dict(x=torch.empty((E, N, 3)).pin_memory(), v=torch.empty((E, N, 3)).pin_memory())
                if with_depth:
                    hb["depth"] = torch.empty(env.depth.shape).pin_memory()
                color = env.color if slot == 0 else torch.empty_like(env.color)
                depth = env.depth if slot == 0 else torch.empty_like(env.depth)
                if fmt == "u8":
                    hb["rgb8"] = torch.empty((env.B, H, W, 3), dtype=torch.uint8).pin_memory()
                    devbuf.append((color, depth, torch.empty((env.B, H, W, 3), dtype=torch.uint8, device=dev)))
                else:
                    hb["color"] = torch.empty(env.color.shape).pin_memory()
                    devbuf.append((color, depth))
                hostbuf.append(hb)
            devstate = [(torch.empty((E, N, 3), device=dev), torch.empty((E, N, 3), device=dev)) for _ in range(2)]
            d2h = sum(t.numel() * t.element_size() for t in hostbuf[0].values())
            computed = [torch.cuda.Event() for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]
            for ev in copied:
                ev.record(cs)

            def e2e_step(i, slot):
              with torch.cuda.stream(main):
                m = upload(i)                                    # H2D: end-effector commands + link poses (pinned -> device)
                env.view.copy_(env.view_h, non_blocking=True)    # H2D: cameras
                env.proj.copy_(env.proj_h, non_blocking=True)
                env.campos.copy_(env.campos_h, non_blocking=True)
                main.wait_event(copied[slot])                    # the buffers of step i-2 have left the device
                env.step(command=m[:5], out=devbuf[slot], link_pose=m[5], composite_stream=lo)
                xs, vs = devstate[slot]
                _lib.check(phys_lib.r2s_phys_get_state(env.phys.h, ctypes.c_void_p(xs.data_ptr()),
                                                       ctypes.c_void_p(vs.data_ptr()),
                                                       ctypes.c_void_p(main.cuda_stream)), "get_state")
                lo.wait_stream(main)                             # images (lo) and state (main) of this step are complete
                computed[slot].record(lo)
                with torch.cuda.stream(cs):                      # D2H on the side stream
                    cs.wait_event(computed[slot])
                    hb = hostbuf[slot]
                    hb["x"].copy_(xs, non_blocking=True)
                    hb["v"].copy_(vs, non_blocking=True)
                    if with_depth:
                        hb["depth"].copy_(devbuf[slot][1], non_blocking=True)
                    if fmt == "u8":
                        hb["rgb8"].copy_(devbuf[slot][2], non_blocking=True)
                    else:
                        hb["color"].copy_(devbuf[slot][0], non_blocking=True)
                    copied[slot].record(cs)

            for i in range(args.warmup):
                e2e_step(base_i + i, i % 2)
            cs.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main.wait_stream(main_stream); lo.wait_stream(main_stream)
            e0.record(main)
            for i in range(args.steps):
                e2e_step(base_i + args.warmup + i, i % 2)
            main.wait_stream(lo)
            main.wait_stream(cs)                                 # all copies have landed
            e1.record(main)
            cs.synchronize()
            barrier()
            ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)
            last = hostbuf[(args.steps - 1) % 2]
            out = {"value": world * E * args.steps / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": ms / args.steps,
                   "observations": (("rgb uint8 [B,H,W,3] (reference host format, eval_policy.py:248)" if fmt == "u8" else
                                     "colour f32 [B,3,H,W]") + (" + depth f32 [B,1,H,W]" if with_depth else "")
                                    + " + particle x,v f32 [E,N,3]"),
                   "host_buffers": {k: list(v.shape) for k, v in hostbuf[0].items()},
                   "overlap": "D2H of step k on a side stream under the compute of step k+1 (double-buffered outputs); "
                              "the compositing kernel on its own stream (composite_stream) under the next step's front end"}
            if fmt == "u8":
                out["checksum_rgb8_host"] = int(last["rgb8"].long().sum())
            else:
                out["checksum_rgb_host"] = float(last["color"].double().sum())
            return out

        # headline: what the reference's evaluation loop moves to the host every step -- the RGB image
        # (experiments/eval_policy.py:139-157, 248 copy RGB only; depth stays on the device) + the particle state it
        # pickles (:209-213).  The variants that also bring the f32 depth / the f32 colour image are reported beside it.
        base_i = args.warmup + args.steps
        e2e = run_e2e("u8", base_i, with_depth=False)
        e2e["with_depth_variant"] = run_e2e("u8", base_i, with_depth=True)
        e2e["float_image_variant"] = run_e2e("f32", base_i + args.warmup + args.steps, with_depth=True)
        e2e["host_binding"] = binding
        env.check()

    clocks = sampler.stop()
    # ---- metrics all-gather (the only collective)
    cx = float(env.phys.x.double().sum())
    crgb = float(env.color.double().sum())
    succ, hits = env.success.result()
    gathered = shard.gather_metrics([args.steps, ms_total / 1e3, cx, crgb, float(succ.sum()), float(hits.sum())], dev)

    # ---- roofline of the dominant kernel (algorithmic bytes: SURVEY.md §8d)
    pk, pk_kind = peaks()
    B, P, T = env.B, cfg.P, (W // 16) * (H // 16)
    R = total
    alg = {
        "phys_frame": E * ns * (52 * env.base.N + 16 * env.base.S),
        "lbs": E * env.base.N * (32 + 36) + E * env.n_obj * 24 + env.n_obj * env.K * 8 + env.base.N * cfg.k_rel * 4,
        "links": E * env.n_robot * 28 + env.n_robot * 32 + E * env.links.L * 64,
        "preprocess": B * P * (44 + 12 * 1) + B * P * 40,
        "emit": R * 12,
        "tile_sort": R * 24,
        "composite": R * 40 + B * W * H * 16,
    }
    alg["eef"] = E * (19 * 4 + ns * (env.eef.V + 1) * 12 + 36 + env.eef.V * 24)   # command in, tables out (+ 2 table rows)
    alg["success"] = E * (env.base.N * 16 + (env.base.S * 8 if cfg.scene == "rope" else 0) + 24)
    times = {"eef": eef_ms, "phys_frame": phys_ms, "success": succ_ms, "lbs": lbs_ms, "links": links_ms, "preprocess": stage_ms[0], "scan": stage_ms[1], "emit": stage_ms[2],
             "tile_sort": stage_ms[3], "composite": stage_ms[4]}
    times["collision_graph"] = grid_ms
    dom = max(alg, key=lambda k: times[k])   # dominant kernel among those with an algorithmic-byte model
    ach = alg[dom] / (times[dom] / 1e3) / 1e9
    # per-kernel DRAM traffic: ncu dram__bytes_read + write per launch at this workload (profiles/traffic.json, one
    # `ncu --set full` capture per round; only valid for the default workload it was captured on)
    tmap = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and config_label(args) == "BASELINE configs[1]":
        try:
            tmap = json.load(open(tpath))
        except Exception:
            tmap = {}
    # emit / sort work on (Gaussian, super-tile) instances, 0.177x the reference pipeline's (Gaussian, tile) count their
    # algorithmic-byte model is written in: their "achieved" would exceed the HBM peak, so they are reported against
    # their measured traffic only
    over_model = {"emit", "tile_sort"}
    kernels = {}
    for k, v in times.items():
        ent = {"ms": round(float(v), 4),
               "alg_gbs": round(alg[k] / (v / 1e3) / 1e9, 1) if k in alg and k not in over_model and v > 0 else None}
        if isinstance(tmap.get(k), (int, float)) and v > 0:
            ent["dram_bytes"] = int(tmap[k])
            ent["dram_gbs"] = round(tmap[k] / (v / 1e3) / 1e9, 1)
        kernels[k] = ent
    traffic = tmap.get(dom) if isinstance(tmap.get(dom), (int, float)) else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "peak_source": pk_kind,
                "unit": "GB/s", "frac": round(ach / pk["hbm_gbs"], 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": int(alg[dom]), "kernels": kernels}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "ours",
        "config": {"workload": f"{cfg.scene} PhysTwin x {E} envs/GPU, {ns} substeps/step, {cfg.cameras} x {W}x{H} "
                               f"render of {P} Gaussians per env ({config_label(args)})",
                   "envs_per_gpu": E, "global_envs": world * E, "particles": env.base.N, "springs": env.base.S,
                   "substeps": ns, "resolution": [W, H], "cameras": cfg.cameras, "gaussians_per_env": P,
                   "gaussian_rows": {"object_lbs": env.n_obj, "robot_links": env.n_robot, "static": P - env.n_obj - env.n_robot},
                   "instances_per_step": int(R), "instances_per_gaussian": round(R / (B * P), 3),
                   "super_tile_instances_per_step": int(env.raster.intermediates()["super_offset"][-1].item()),
                   "mean_tile_list": round(R / (B * T), 1), "parallelism": f"env-shard x{world}",
                   "streams": "2 per GPU: everything up to the sort on one (high priority), the compositing kernel on the "
                              "other (r2s_raster_args.composite_stream); steps enqueued back to back, no host sync",
                   "l2_policy": "inputs larger than L2 (2.9 GB of Gaussians + 1.07 GB of images per step)",
                   "outputs": "colour f32 CHW + depth f32 (+ uint8 HWC in e2e); the `radii` output of the reference API "
                              "(205 MB per step, ~0.05 ms) is not requested in the loop (want_radii=False)"},
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "fast_composite_variant": fast_variant, "serial_kernel_accounting": serial,
        "metrics_allgather": {"per_rank": gathered, "fields": ["steps", "seconds", "checksum_x", "checksum_rgb", "episodes_succeeded", "frames_passed"]},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # timed on rank 0 at N=1 only
        line["cpu_baseline"] = cb = cpu_baseline(args, env)
        if cb.get("reference_rasterizer_us_per_view"):
            ours_us = float(sum(stage_ms)) / env.B * 1e3
            line["raster_vs_reference"] = {
                "ours_us_per_view": round(ours_us, 2), "reference_us_per_view": round(cb["reference_rasterizer_us_per_view"], 2),
                "ratio": round(cb["reference_rasterizer_us_per_view"] / ours_us, 2),
                "what": f"five raster kernels of the last timed step / {env.B} views (CUDA events) vs the unmodified reference "
                        "CUDA rasterizer (oracle/_ref) called view by view on the same device tensors with its blocking "
                        "num_rendered read-back (wall clock, 3 passes over 16 envs); images bit-identical "
                        "(tests/test_gpu_raster.py::test_bench_configurations_bit_identical_to_reference)"}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _oracle_envs(scene, n, n_substeps, seed, offset=0, mesh=None):
    import r2s_testutil as util
    from real2sim_eval_b200 import synth
    envs = []
    for e in range(n):
        s = synth.pose_scene(scene, seed + offset + e)
        envs.append(util.oracle_from_scene(s, n_substeps, mesh=mesh))
    return envs


def cpu_baseline(args, env=None):
    """oracle/ on the host cores: physics (all cores, one env per thread) + CPU rasterizer oracle
    on a bounded sample of the same workload."""
    import r2s_testutil as util
    from oracle import physics_ref, raster_ref
    from real2sim_eval_b200 import synth
    from concurrent.futures import ThreadPoolExecutor

    cores = os.cpu_count() or 1
    physics_ref.set_threads(cores)
    n = args.cpu_sample_envs or min(cores, 32)
    W, H = args.res
    scene = {"rope": synth.make_rope, "sloth": synth.make_sloth, "tblock": synth.load_tblock}[args.scene]()
    g = synth.make_gripper(center=(float(scene.x[:, 0].mean()), float(scene.x[:, 1].mean()), 0.004), gap=0.03)
    mesh = util.gripper_mesh_dict(g)
    envs = _oracle_envs(scene, n, args.substeps, 1234, mesh=mesh)
    gss = [synth.make_gaussians(1234 + e, args.gaussians, scene.x, n_object=0) for e in range(min(n, 4))]
    cams = [synth.make_camera(W, H, "side", jitter_seed=e) for e in range(n)]

    def render(e):
        gs = gss[e % len(gss)]
        c = cams[e]
        raster_ref.rasterize(gs.means3D, gs.opacities, viewmatrix=c.view, projmatrix=c.proj, campos=c.campos,
                             bg=np.zeros(3, np.float32), W=W, H=H, tanfovx=c.tanfovx, tanfovy=c.tanfovy, shs=gs.shs,
                             scales=gs.scales, rotations=gs.rotations, z_threshold=0.05)

    t0 = time.perf_counter()
    for o in envs:
        o.update_collision_graph()
    physics_ref.step_batch(envs)                       # OpenMP: one env per thread
    t_phys = time.perf_counter() - t0
    t1 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(cores, n)) as ex:   # ctypes releases the GIL
        list(ex.map(render, range(n)))
    t_rend = time.perf_counter() - t1
    dt = t_phys + t_rend
    ref_raster_us = None
    if env is not None:   # the UNMODIFIED reference CUDA rasterizer (oracle/_ref) on this run's own scenes, view by view
        import ref_raster
        if ref_raster.available():
            dev = env.device
            k = min(16, env.cfg.E)
            color = torch.empty((3, H, W), device=dev)
            depth = torch.empty((1, H, W), device=dev)
            radii = torch.empty(env.cfg.P, dtype=torch.int32, device=dev)
            views = [(e, c) for e in range(k) for c in range(env.cfg.cameras)]

            def ref_pass():
                for e, c in views:
                    b = e * env.cfg.cameras + c
                    t = dict(means3D=env.means3D[e], scales=env.scales[e], rotations=env.rotations[e],
                             opacities=env.opacities[e], shs=env.shs[e])
                    ref_raster.forward_torch(t, env.view[b], env.proj[b], env.campos[b], env.bg, W, H,
                                             env.cams[b].tanfovx, env.cams[b].tanfovy, 0, 0.05, color, depth, radii)
                torch.cuda.synchronize(dev)

            ref_pass()
            t2 = time.perf_counter()
            for _ in range(3):
                ref_pass()
            ref_raster_us = (time.perf_counter() - t2) / (3 * len(views)) * 1e6
    return {"value": n / dt, "unit": UNIT, "cores": min(cores, n), "kind": "port",
            "reference_rasterizer_us_per_view": ref_raster_us,
            "sample": f"{n} envs x 1 step ({args.substeps} substeps + one {W}x{H} render of {args.gaussians} Gaussians) "
                      f"with oracle/physics_ref.c + oracle/raster_ref.c, one env per host thread",
            "physics_s": round(t_phys, 4), "render_s": round(t_rend, 4),
            "physics_env_substeps_per_s": n * args.substeps / t_phys}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    """CPU physics (oracle port of the Warp path, all host threads) + the unmodified reference CUDA
    rasterizer called once per env, as the reference's env loop would (one render + one blocking
    read-back per call).  Bounded sample: ref_envs environments per step."""
    rank, world, local = dist_env()
    import r2s_testutil as util
    import ref_raster
    from oracle import physics_ref
    from real2sim_eval_b200 import synth

    if not ref_raster.available():
        return {"impl": "reference", "unavailable": "oracle/_ref/libref_raster.so missing (build needs /root/reference)"} \
            if rank == 0 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # N > 1: one reference process per GPU, as the reference's own multi-GPU driver runs it
    # (experiments/eval_policy_parallel.py:266-279: multiprocessing.Pool, episode i -> GPU i % n): every rank steps its
    # share of the sample with its share of the host threads and its own GPU for the rasterizer; gloo carries the
    # barrier and the per-rank times.
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    cores_total = os.cpu_count() or 1
    cores = max(1, cores_total // world)
    omp_threads = physics_ref.set_threads(cores)     # torchrun exports OMP_NUM_THREADS=1: ask for this rank's share
    n_total = args.ref_envs or min(cores_total, 64)
    n = max(1, n_total // world)
    W, H = args.res
    P = args.gaussians
    scene = {"rope": synth.make_rope, "sloth": synth.make_sloth, "tblock": synth.load_tblock}[args.scene]()
    g = synth.make_gripper(center=(float(scene.x[:, 0].mean()), float(scene.x[:, 1].mean()), 0.004), gap=0.03)
    envs = _oracle_envs(scene, n, args.substeps, 1234, offset=rank * n, mesh=util.gripper_mesh_dict(g))
    tables = synth.gripper_motion(g, args.substeps, 5e-5, eef_vel=(0.02, 0.0, -0.01))
    for o in envs:
        o.set_mesh_interactive(*tables)
    # Gaussians / cameras: same generator as our arm's BatchedEnv (device RNG), one set per env
    gen = torch.Generator(device=dev)
    scenes_t, cams = [], []
    lo, hi = torch.tensor([-0.1, -0.6, 0.0], device=dev), torch.tensor([1.1, 0.6, 0.6], device=dev)
    for e in range(rank * n, rank * n + n):
        gen.manual_seed(1234 + 7 * e + 1)
        means = lo + (hi - lo) * torch.rand((P, 3), device=dev, generator=gen)
        scales = torch.exp(np.log(0.006) + 0.5 * torch.randn((P, 3), device=dev, generator=gen))
        q = torch.randn((P, 4), device=dev, generator=gen)
        rots = q / q.norm(dim=1, keepdim=True)
        opac = torch.sigmoid(1.5 + 1.5 * torch.randn((P, 1), device=dev, generator=gen))
        shs = (torch.rand((P, 1, 3), device=dev, generator=gen) - 0.5) / 0.28209479177387814
        scenes_t.append(dict(means3D=means.contiguous(), scales=scales.contiguous(), rotations=rots.contiguous(),
                             opacities=opac.contiguous(), shs=shs.contiguous()))
        c = synth.make_camera(W, H, "side", jitter_seed=1234 + 13 * e)
        cams.append((c, torch.tensor(c.view, device=dev), torch.tensor(c.proj, device=dev),
                     torch.tensor(c.campos, device=dev)))
    bg = torch.zeros(3, device=dev)
    color = torch.empty((3, H, W), device=dev)
    depth = torch.empty((1, H, W), device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)

    def step():
        for o in envs:
            o.update_collision_graph()
        physics_ref.step_batch(envs)
        rendered = 0
        for e in range(n):
            c, v, pr, cp = cams[e]
            rendered += ref_raster.forward_torch(scenes_t[e], v, pr, cp, bg, W, H, c.tanfovx, c.tanfovy, 0, 0.05,
                                                 color, depth, radii)
        torch.cuda.synchronize(dev)
        return rendered

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(1.0)
    sampler.mark()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rendered = step()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    if dist is not None:                      # whole job = all ranks' envs over the slowest rank's time
        tt = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    value = world * n * args.steps / dt
    # split for the record (one more untimed step)
    t1 = time.perf_counter()
    for o in envs:
        o.update_collision_graph()
    physics_ref.step_batch(envs)
    t_phys = time.perf_counter() - t1
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{args.scene} PhysTwin, bounded sample of {world * n} envs per step (of {world * args.envs}), "
                               f"{args.substeps} substeps/step, one {W}x{H} render of {P} Gaussians per env",
                   "envs_per_step": world * n, "reference_processes": world, "host_threads_per_process": cores,
                   "substeps": args.substeps, "resolution": [W, H], "gaussians_per_env": P,
                   "instances_last_step": int(rendered)},
        "clocks": clocks, "gpu_launches": 0,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores * world, "kind": "port+reference",
                         "sample": f"{world * n} envs per step over {world} process(es): physics = oracle/physics_ref.c (CPU "
                                   f"restatement of the Warp kernels, pinned to the reference's kernel source; warp-lang "
                                   f"not installable) on {cores} host threads per process, render = unmodified "
                                   f"reference CUDA rasterizer (oracle/_ref) once per env with its blocking read-back",
                         "physics_s_per_step": round(t_phys, 4), "omp_threads": omp_threads},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    args = parse()
    # Exactly ONE line may reach stdout.  Libraries (NCCL prints its version banner with printf) write to
    # fd 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to
    # a saved copy of the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        line = run_reference(args)
    else:
        line = run_ours(args)
    sys.stdout.flush()
    if line is not None:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    os.close(real_stdout)


if __name__ == "__main__":
    main()
