"""Generate rasterizer golden vectors from the UNMODIFIED reference CUDA rasterizer.

Must run on a GPU box (the reference path is CUDA-only), with oracle/_ref/libref_raster.so
built beforehand in the container that has /root/reference (make -C oracle):

    gpurun -- 'python tests/golden/make_raster_golden.py gpurun_out/golden'

then copy gpurun_out/golden/raster_*.npz into tests/golden/.  Each file holds the seeded
inputs (float32, as fed to the reference), the camera in the reference's convention
(sim/utils/gs/transform_utils.py:7-31) and the reference's outputs: color (3,H,W),
depth (1,H,W), radii (P,), num_rendered.  Sizes are kept small (fixtures are committed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))           # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # repo root

import r2s_testutil as _util  # noqa: E402
import ref_raster  # noqa: E402

CASES = [  # name, W, H, P, seed, sh degree, bg, gaussian scale
    ("a_64x64_deg0", 64, 64, 1500, 101, 0, (0.0, 0.0, 0.0), 0.02),
    ("b_128x96_deg0_bg", 128, 96, 4000, 102, 0, (0.1, 0.2, 0.3), 0.02),
    ("c_96x96_deg3", 96, 96, 1200, 103, 3, (0.0, 0.0, 0.0), 0.03),
    ("d_200x120_small", 200, 120, 3000, 104, 1, (1.0, 1.0, 1.0), 0.008),
]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, W, H, P, seed, deg, bg, scale in CASES:
        g = _util.small_gaussians(seed, P, scale=scale, sh_coeffs=(deg + 1) ** 2)
        cam = _util.make_test_camera(W, H)
        color, radii, depth, n = ref_raster.forward(g, cam, sh_degree=deg, bg=bg)
        np.savez_compressed(os.path.join(out_dir, f"raster_{name}.npz"), **g, W=W, H=H, tanfovx=cam.tanfovx,
                            tanfovy=cam.tanfovy, view=cam.view, proj=cam.proj, campos=cam.campos,
                            z_threshold=cam.z_threshold, sh_degree=deg, bg=np.asarray(bg, np.float32),
                            color=color.astype(np.float32), depth=depth.astype(np.float32),
                            radii=radii.astype(np.int32), num_rendered=n)
        print(name, "num_rendered", n, "visible", int((radii > 0).sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE))
