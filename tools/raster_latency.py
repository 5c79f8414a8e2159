"""Drop-in use case: ONE 848x480 view (the reference's camera size, cfg/env/xarm_gripper.yaml:21-35) of P Gaussians
through `GaussianRasterizer(...)(...)`, ours vs the unmodified reference CUDA rasterizer (oracle/_ref), same inputs,
CUDA events around N back-to-back calls.  Also B views in one enqueue (BatchedRasterizer) for the same scene.
    python tools/raster_latency.py [--gaussians 200000] [--iters 50]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from real2sim_eval_b200 import synth
from real2sim_eval_b200.rasterizer import BatchedRasterizer, GaussianRasterizationSettings, GaussianRasterizer

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=200_000)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--res", type=int, nargs=2, default=[848, 480])
a = ap.parse_args()
W, H = a.res
rope = synth.make_rope()
g = synth.make_gaussians(1234, a.gaussians, rope.x, n_object=0)
cam = synth.make_camera(W, H, "side")
t = lambda x: torch.tensor(np.ascontiguousarray(x)).cuda()
G = dict(means3D=t(g.means3D), scales=t(g.scales), rotations=t(g.rotations), opacities=t(g.opacities), shs=t(g.shs))
rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3).cuda(), 1.0, t(cam.view), t(cam.proj), 0,
                                   t(cam.campos), False, 0.05)
rast = GaussianRasterizer(rs)


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


ours = lambda: rast(means3D=G["means3D"], means2D=None, opacities=G["opacities"], shs=G["shs"], scales=G["scales"],
                    rotations=G["rotations"])
us_ours = timeit(ours, a.iters)
color, radii, depth = ours()
line = f"{W}x{H}, P={a.gaussians}: drop-in GaussianRasterizer call {us_ours:.1f} us/view"
import ref_raster
if ref_raster.available():
    oc = torch.empty((3, H, W), device="cuda"); od = torch.empty((1, H, W), device="cuda")
    orad = torch.empty(a.gaussians, dtype=torch.int32, device="cuda")
    ref = lambda: ref_raster.forward_torch(G, t(cam.view), t(cam.proj), t(cam.campos), torch.zeros(3).cuda(), W, H, cam.tanfovx,
                                           cam.tanfovy, 0, 0.05, oc, od, orad)
    v, p_, c_, bg = t(cam.view), t(cam.proj), t(cam.campos), torch.zeros(3).cuda()
    ref = lambda: ref_raster.forward_torch(G, v, p_, c_, bg, W, H, cam.tanfovx, cam.tanfovy, 0, 0.05, oc, od, orad)
    us_ref = timeit(ref, a.iters)
    n = ref()
    torch.cuda.synchronize()
    line += (f"; reference CUDA rasterizer {us_ref:.1f} us/view (speed-up {us_ref / us_ours:.2f}x); "
             f"bit-identical colour: {bool(torch.equal(color, oc))}, depth: {bool(torch.equal(depth, od))}, num_rendered {n}")
B = 16
br = BatchedRasterizer("cuda")
rep = lambda x: x[None].expand(B, *x.shape).contiguous()
bk = dict(viewmatrix=rep(t(cam.view)), projmatrix=rep(t(cam.proj)), campos=rep(t(cam.campos)), bg=torch.zeros(3).cuda(), W=W, H=H,
          tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, shs=rep(G["shs"]), scales=rep(G["scales"]), rotations=rep(G["rotations"]),
          max_instances=12 * B * a.gaussians, want_radii=False)
mb, ob = rep(G["means3D"]), rep(G["opacities"])
us_b = timeit(lambda: br.forward(mb, ob, **bk), max(5, a.iters // 5)) / B
print(line + f"; batched (B={B}, no host sync) {us_b:.1f} us/view")
