"""Generates tests/golden/metrics_*.npz by running the REFERENCE's own success tests
(/root/reference/experiments/utils/calculate_success_T.py:17-29 is_pusht_success,
calculate_success_rope.py:77-170 count_xz_plane_intersections / is_rope_success; unmodified, numpy) on
synthetic particle states around their decision thresholds.  Run in the build container (needs
/root/reference):   python tests/golden/make_metrics_golden.py
  metrics_pusht.npz : x (K,N,3) states of the shipped T-block (rigidly displaced target + noise), target,
                      ref_pass (K,), ref_mse (K,)
  metrics_rope.npz  : x (K,N,3) rope states threaded through / beside the routing box, springs,
                      ref_pass (K,), ref_counts (K,2)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import metrics_ref               # noqa: E402
from real2sim_eval_b200 import synth         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rope_states(rope, rng, K):
    """The synthetic rope (1 m along x) turned to run along y through the routing box at (0.62, 0.05),
    shifted sideways / lifted / bent so that some states route through both faces and some do not."""
    lo, hi = metrics_ref.rope_box()
    out = []
    for k in range(K):
        ang = np.pi / 2 + rng.normal(0, 0.15 if k % 3 else 0.6)
        c, s = np.cos(ang), np.sin(ang)
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        p = (rope.x.astype(np.float64) - [0.5, 0.0, 0.0]) @ R.T
        side = [0.0, 0.004, 0.012, 0.03][k % 4] * (1 if k % 2 else -1)
        lift = [0.0, 0.01, 0.025][k % 3]
        p += [0.62 + side, 0.05 + rng.normal(0, 0.05), lift]
        p[:, 2] += 0.02 * np.sin(6 * p[:, 1]) * (k % 5 == 0)
        out.append(p.astype(np.float32))
    return np.stack(out)


def main():
    rng = np.random.default_rng(7)
    # ---- push-T
    mod = metrics_ref.load_reference("pusht")
    assert mod is not None, "needs /root/reference"
    target = np.load(os.path.join(HERE, "tblock.npz"))["x"]
    xs = []
    for k, d in enumerate([0.0, 0.01, 0.03, 0.04, 0.0440, 0.0446, 0.0448, 0.0450, 0.05, 0.08, 0.2]):
        th = rng.uniform(0, 2 * np.pi)
        x = target + np.array([d * np.cos(th), d * np.sin(th), 0.0], np.float32)
        if k % 2:
            x = x + rng.normal(0, 0.002, x.shape)
        xs.append(x.astype(np.float32))
    xs = np.stack(xs)
    ref_pass = np.array([metrics_ref.reference_frame_test(mod, "pusht", x, target=target) for x in xs])
    ref_mse = np.array([((x - target) ** 2).sum(1).mean() for x in xs], np.float32)
    assert ref_pass.any() and not ref_pass.all()
    np.savez_compressed(os.path.join(HERE, "metrics_pusht.npz"), x=xs, target=target, ref_pass=ref_pass, ref_mse=ref_mse)
    print("pusht", ref_pass.astype(int).tolist(), np.round(ref_mse, 5).tolist())
    # ---- rope
    mod = metrics_ref.load_reference("rope")
    rope = synth.make_rope()
    xs = rope_states(rope, rng, 16)
    lo, hi = metrics_ref.rope_box()
    ref_pass, ref_counts = [], []
    for x in xs:
        ref_pass.append(metrics_ref.reference_frame_test(mod, "rope", x, springs=rope.springs))
        r = mod.count_xz_plane_intersections(x, rope.springs, (lo, hi))
        ref_counts.append((r["y_min_count"], r["y_max_count"]))
    ref_pass, ref_counts = np.array(ref_pass), np.array(ref_counts, np.int32)
    assert ref_pass.any() and not ref_pass.all()
    np.savez_compressed(os.path.join(HERE, "metrics_rope.npz"), x=xs, springs=rope.springs.astype(np.int32),
                        ref_pass=ref_pass, ref_counts=ref_counts)
    print("rope", ref_pass.astype(int).tolist(), ref_counts.tolist())


if __name__ == "__main__":
    main()
