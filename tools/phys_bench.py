"""Physics-only timing of the frame kernel (CUDA events), for kernel iteration:
    python tools/phys_bench.py [--scene rope] [--envs 256] [--substeps 10] [--no-mesh] [--threads 1024]
Prints ms per frame, env-substeps/s and algorithmic GB/s (52N + 16S bytes per env-substep)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from real2sim_eval_b200 import synth
from real2sim_eval_b200.physics import BatchedSpringMass

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="rope")
ap.add_argument("--envs", type=int, default=256)
ap.add_argument("--substeps", type=int, default=10)
ap.add_argument("--no-mesh", action="store_true")
ap.add_argument("--no-self", action="store_true")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--precise", action="store_true")
ap.add_argument("--pusher", action="store_true", help="25,312-triangle rigid rod instead of the two-finger gripper")
ap.add_argument("--shared-pose", action="store_true", help="all envs share one pose (one rest-length table)")
a = ap.parse_args()
base = {"rope": synth.make_rope, "sloth": synth.make_sloth, "tblock": synth.load_tblock}[a.scene]()
E = a.envs
poses = [base] * E if a.shared_pose else [synth.pose_scene(base, 1234 + e) for e in range(E)]
p = dict(base.params)
p["self_collision"] = not a.no_self
if a.pusher:
    p["use_pusher"], p["collide_eef_fric"] = True, 0.2          # phystwin.py:305-306
rest = base.rest if a.shared_pose else np.stack([q.rest for q in poses])
s = BatchedSpringMass(E, base.springs, rest, num_particles=base.N,
                      n_substeps=a.substeps, log_spring_Y=base.log_Y, masses=base.mass, threads=a.threads, precise=a.precise, **p)
s.set_state(base.x if a.shared_pose else np.stack([q.x for q in poses]), None if a.shared_pose else np.stack([q.v for q in poses]))
if s.self_collision:
    s.create_resting_case()
if a.pusher:
    g = synth.make_pusher(center=(float(base.x[:, 0].min()) - 0.0375 + 0.0006, 0.0, 0.004), n_circ=112, n_len=112)
    s.set_mesh(g.verts, g.faces, g.mesh_map, g.face_map, len(g.verts))
    t = synth.rigid_motion_tables(g, a.substeps, p["dt"], vel=(0.05, 0.0, 0.0), omega=(0.0, 0.0, 0.2))
    s.set_mesh_motion(*[torch.tensor(x).cuda() for x in t])
elif not a.no_mesh:
    c = base.x.mean(0)
    g = synth.make_gripper(center=(float(c[0]), float(c[1]), 0.004), gap=0.03)
    s.set_mesh(g.verts, g.faces, g.mesh_map, g.face_map, len(g.verts))
    t = synth.gripper_motion(g, a.substeps, p["dt"], eef_vel=(0.02, 0.0, -0.01))
    s.set_mesh_motion(*[torch.tensor(x).cuda() for x in t])
for _ in range(3):
    if s.self_collision:
        s.update_collision_graph()
    s.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tg = 0.0
e0.record()
for _ in range(a.iters):
    s.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
if s.self_collision:
    g0.record()
    for _ in range(a.iters):
        s.update_collision_graph()
    g1.record()
    torch.cuda.synchronize()
    tg = g0.elapsed_time(g1) / a.iters
alg = E * a.substeps * s.algorithmic_bytes_per_env_substep()
print(f"{a.scene} pusher={a.pusher} precise={a.precise} E={E} substeps={a.substeps} mesh={not a.no_mesh} threads={a.threads or 1024} smem={s.smem_bytes}B: "
      f"frame {ms:.3f} ms, {E * a.substeps / ms * 1e3:.3e} env-substeps/s, alg {alg / ms / 1e6:.1f} GB/s; "
      f"collision graph {tg:.3f} ms")
