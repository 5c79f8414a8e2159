import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import test_gpu_raster as T
import ref_raster
from real2sim_eval_b200 import synth
W, H = 640, 480
g = T._bench_scene(0)
cam = synth.make_camera(W, H, "wrist", jitter_seed=0)
r, color, radii, depth, total, ov = T._run_cuda(g, cam, max_instances=8 * 200000)
rc, rr, rd, n = ref_raster.forward(g, cam)
pl, rg = ref_raster.lists(n, W, H)
print("total", total, n, "radii equal", np.array_equal(radii, rr), "color equal", np.array_equal(color, rc), (color != rc).sum(), "depth eq", np.array_equal(depth, rd))
lists = T._tile_lists(r, 0)
it = r.intermediates()
dep = it["depths"][0].cpu().numpy()
nbad = 0
for t, ids in enumerate(lists):
    want = pl[rg[t, 0]:rg[t, 1]]
    if not np.array_equal(ids, want):
        nbad += 1
        if nbad <= 3:
            print("tile", t, len(ids), len(want))
            if len(ids) == len(want):
                d = np.nonzero(ids != want)[0]
                print(" first diffs at", d[:10], "ours", ids[d[:6]], "ref", want[d[:6]])
                print(" depths ours", dep[ids[d[:6]]].view(np.uint32), "ref", dep[want[d[:6]]].view(np.uint32))
            else:
                so, sw = set(ids.tolist()), set(want.tolist())
                print(" only ours", list(so - sw)[:10], "only ref", list(sw - so)[:10])
print("bad tiles", nbad, "of", len(lists))
