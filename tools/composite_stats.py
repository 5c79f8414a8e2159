"""CPU analysis of the compositing workload of one bench view (no GPU): from the oracle's per-tile lists,
count what a 16x16-pixel tile kernel has to evaluate -- list entries walked per pixel until the reference's
stop rule, entries a 16x2 warp strip can skip, live lanes per visited entry -- and, per candidate shape of the
pixel block a warp owns (--shapes), how many power evaluations and blend evaluations a pixel pays when a warp
skips an entry only if none of its pixels is live: the count behind composite_kernel's 8x8 warp blocks (16x4
strips: 70.2 / 64.5 per pixel, 8x8 blocks: 68.3 / 60.4, really blended: 51.1).  Test/analysis tooling: uses
oracle/ (allowed for tools that are neither product nor bench legs)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import raster_ref          # noqa: E402
from real2sim_eval_b200 import synth   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=200_000)
    ap.add_argument("--W", type=int, default=512)
    ap.add_argument("--H", type=int, default=512)
    ap.add_argument("--tiles", type=int, default=128, help="number of tiles sampled")
    ap.add_argument("--shapes", default="16x4,8x8,16x2,8x4,16x16", help="warp pixel blocks (width x height) to compare")
    a = ap.parse_args()
    shapes = {k: (int(k.split("x")[1]), int(k.split("x")[0])) for k in a.shapes.split(",") if k}
    blocks = lambda m, bh, bw: m.reshape(16 // bh, bh, 16 // bw, bw).transpose(0, 2, 1, 3).reshape(-1, bh * bw)
    shape_tot = {k: [0, 0] for k in shapes}   # pixel-evaluations of power, of exp + blend
    g = synth.make_gaussians(1234, a.P)
    cam = synth.make_camera(a.W, a.H, "side", jitter_seed=1)
    col, rad, dep, aux = raster_ref.rasterize(g.means3D, g.opacities, viewmatrix=cam.view, projmatrix=cam.proj,
                                              campos=cam.campos, bg=np.zeros(3, np.float32), W=a.W, H=a.H,
                                              tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, shs=g.shs, scales=g.scales,
                                              rotations=g.rotations, z_threshold=cam.z_threshold, aux=True)
    R = aux["num_rendered"]
    gx, gy = (a.W + 15) // 16, (a.H + 15) // 16
    print(f"R={R} R/P={R / a.P:.2f} mean list={R / (gx * gy):.0f} visible={int((rad > 0).sum())}")
    rng = np.random.default_rng(0)
    tiles = rng.choice(gx * gy, min(a.tiles, gx * gy), replace=False)
    xy, co = aux["means2D"], aux["conic_opacity"]
    tot = dict(list=0, walked_tile=0, walked_warp=0, warp_visit=0, warp_live=0, lane_live=0, lane_blend=0,
               pix_walk=0, kept=0)
    for t in tiles:
        s, e = aux["ranges"][t]
        ids = aux["point_list"][s:e]
        tx, ty = t % gx, t // gx
        px = (tx * 16 + np.arange(16))[None, :].repeat(16, 0).astype(np.float32)
        py = (ty * 16 + np.arange(16))[:, None].repeat(16, 1).astype(np.float32)
        T = np.ones((16, 16), np.float32)
        done = np.zeros((16, 16), bool)
        tot["list"] += len(ids)
        # tile-level cull as in composite_kernel (bound on the minimum of q over the tile) -- approximate by
        # the exact per-pixel test: entry kept if any pixel has alpha >= 1/255
        for j, i in enumerate(ids):
            if done.all():
                break
            dx, dy = xy[i, 0] - px, xy[i, 1] - py
            A, B, Cc, o = co[i]
            power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
            alpha = np.minimum(0.99, o * np.exp(power))
            vis = (power <= 0) & (alpha >= 1.0 / 255.0)
            if not vis.any():
                continue
            tot["kept"] += 1
            live = vis & ~done
            for k, (bh, bw) in shapes.items():         # a warp evaluates power unless all its pixels are done,
                shape_tot[k][0] += int((~blocks(done, bh, bw).all(1)).sum()) * bh * bw
                shape_tot[k][1] += int(blocks(live, bh, bw).any(1).sum()) * bh * bw   # and blends if any is live
            wl = live.reshape(8, 32).any(1)            # 16x2 strips = rows (2k, 2k+1)
            wd = done.reshape(8, 32).all(1)
            tot["warp_visit"] += int((~wd).sum())
            tot["warp_live"] += int(wl.sum())
            tot["lane_live"] += int(live.sum())
            tot["pix_walk"] += int((~done).sum())
            test_T = T * (1 - alpha)
            stop = live & (test_T < 1e-4)
            done |= stop
            ok = live & ~stop
            tot["lane_blend"] += int(ok.sum())
            T = np.where(ok, test_T, T)
        tot["walked_tile"] += j + 1
    n = len(tiles)
    print(f"per tile (n={n}): list {tot['list'] / n:.0f}, walked until tile done {tot['walked_tile'] / n:.0f}, "
          f"kept (visible on tile) {tot['kept'] / n:.0f}")
    print(f"  warp visits of kept entries {tot['warp_visit'] / n:.0f} (of {8 * tot['kept'] / n:.0f}), "
          f"warp-live {tot['warp_live'] / n:.0f}")
    print(f"  lane-live {tot['lane_live'] / n:.0f}, lane-blend {tot['lane_blend'] / n:.0f}, "
          f"not-done pixel visits {tot['pix_walk'] / n:.0f}")
    print(f"  lanes live per live warp {tot['lane_live'] / max(1, tot['warp_live']):.1f} / 32")
    npx = 256.0 * n
    print(f"per pixel: live {tot['lane_live'] / npx:.1f}, blended {tot['lane_blend'] / npx:.1f}; by warp pixel block:")
    for k, (pw, bl) in shape_tot.items():
        print(f"  {k:6s} power evaluations {pw / npx:5.1f}  blend evaluations {bl / npx:5.1f}")


if __name__ == "__main__":
    main()
