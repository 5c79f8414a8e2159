"""Generate LBS golden vectors from the reference's OWN interpolate_motions (torch on CPU).

Run in the build container (needs /root/reference):  python tests/golden/make_lbs_golden.py
Writes tests/golden/lbs_*.npz {bones, motions, relations, xyz, weights, weights_indices, out}.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lbs_ref  # noqa: E402
from real2sim_eval_b200 import synth  # noqa: E402


def case(name, bones, motions, n_pts, seed):
    rng = np.random.default_rng(seed)
    rel = lbs_ref.knn_relations(bones, 8)
    src = bones[rng.integers(0, len(bones), n_pts)] + rng.normal(0, 0.003, (n_pts, 3))
    w, wi = lbs_ref.knn_weights(bones, src, 16)
    ref = lbs_ref.load_reference()
    t = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt)
    out, _, _ = ref(bones=t(bones), motions=t(motions), relations=rel, xyz=t(src), weights=t(w),
                    weights_indices=t(wi, torch.int64), quat=None, device="cpu")
    np.savez_compressed(os.path.join(HERE, f"lbs_{name}.npz"), bones=bones.astype(np.float32),
                        motions=motions.astype(np.float32), relations=rel, xyz=src.astype(np.float32), weights=w,
                        weights_indices=wi, out=out.numpy().astype(np.float32))
    print(name, out.shape, float(np.abs(out.numpy() - src).max()))


def main():
    rope = synth.make_rope(n=600)
    rng = np.random.default_rng(3)
    # (a) smooth bending + stretch of a rope
    x = rope.x.astype(np.float64)
    bend = np.stack([0.02 * np.sin(6 * x[:, 0]), 0.01 * x[:, 0] ** 2, 0.03 * np.sin(3 * x[:, 0]) ** 2], 1)
    case("a_rope_bend", rope.x, (bend + rng.normal(0, 2e-4, x.shape)).astype(np.float32), 2500, 11)
    # (b) rigid rotation + translation of a blob
    sl = synth.make_sloth(n=500)
    ang = 0.4
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    c = sl.x.mean(0)
    new = (sl.x - c) @ Rz.T + c + np.array([0.01, -0.02, 0.005])
    case("b_blob_rigid", sl.x, (new - sl.x).astype(np.float32), 2000, 12)
    # (c) a mirror-like local deformation (det F < 0 for some bones: the reflection fix is exercised)
    m = sl.x.copy()
    m[:, 0] = 2 * c[0] - m[:, 0]
    case("c_blob_mirrored", sl.x, (m - sl.x).astype(np.float32), 1500, 13)


if __name__ == "__main__":
    main()
