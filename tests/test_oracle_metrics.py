"""The success-metric oracle (oracle/metrics_ref.py) against the golden vectors made from the reference's own
calculate_success_T.py / calculate_success_rope.py (tests/golden/make_metrics_golden.py), against the live
reference scripts when /root/reference is mounted, and known answers."""
import os

import numpy as np
import pytest

from oracle import metrics_ref
from real2sim_eval_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def test_pusht_matches_reference_golden():
    d = np.load(os.path.join(G, "metrics_pusht.npz"))
    assert d["ref_pass"].any() and not d["ref_pass"].all()
    for k, x in enumerate(d["x"]):
        ok, mse, _ = metrics_ref.frame_test("pusht", x, target=d["target"])
        assert ok == bool(d["ref_pass"][k]) and np.float32(mse) == d["ref_mse"][k]


def test_rope_counts_match_reference_golden_exactly():
    d = np.load(os.path.join(G, "metrics_rope.npz"))
    assert d["ref_pass"].any() and not d["ref_pass"].all()
    for k, x in enumerate(d["x"]):
        ok, c0, c1 = metrics_ref.frame_test("rope", x, springs=d["springs"])
        assert (int(c0), int(c1)) == tuple(d["ref_counts"][k]) and ok == bool(d["ref_pass"][k])


@pytest.mark.skipif(not os.path.exists(metrics_ref.REF_DIR), reason="/root/reference not mounted")
def test_matches_live_reference_scripts():
    rope = synth.make_rope()
    mod = metrics_ref.load_reference("rope")
    sloth_mod = metrics_ref.load_reference("sloth")   # same plane test, open3d stubbed
    rng = np.random.default_rng(3)
    lo, hi = metrics_ref.rope_box()
    for k in range(6):
        R = synth._rot_from_rotvec([0, 0, np.pi / 2 + rng.normal(0, 0.2)])
        x = ((rope.x - [0.5, 0, 0]) @ R.T + [0.62 + rng.normal(0, 0.01), 0.05, 0.0]).astype(np.float32)
        want = mod.count_xz_plane_intersections(x, rope.springs, (lo, hi))
        assert metrics_ref.rope_counts(x, rope.springs) == (want["y_min_count"], want["y_max_count"])
        assert sloth_mod.count_xz_plane_intersections(x, rope.springs, (lo, hi)) == want
        assert metrics_ref.frame_test("rope", x, springs=rope.springs)[0] == \
            metrics_ref.reference_frame_test(mod, "rope", x, springs=rope.springs)
    # segments lying IN a face plane: the coplanar branch (an endpoint inside the rectangle counts)
    xd = np.array([[0.61, lo[1], 0.01], [0.63, lo[1], 0.02], [0.9, lo[1], 0.01], [0.95, lo[1], 0.01]])
    springs = np.array([[0, 1], [2, 3], [1, 2]])
    want = mod.count_xz_plane_intersections(xd, springs, (lo, hi))
    a = metrics_ref.plane_hits(xd[springs[:, 0]], xd[springs[:, 1]], lo[1], lo[0], hi[0], lo[2], hi[2])
    assert int(a.sum()) == want["y_min_count"] == 2


def test_known_answers():
    # push-T: a rigid shift by d gives mse = d^2
    t = np.load(os.path.join(G, "tblock.npz"))["x"]
    ok, mse, _ = metrics_ref.frame_test("pusht", t + np.float32([0.03, 0, 0]), target=t)
    assert ok and abs(mse - 9e-4) < 1e-7
    assert not metrics_ref.frame_test("pusht", t + np.float32([0.05, 0, 0]), target=t)[0]
    # OBB: axis-aligned unit cube at the origin, rotated box
    pts = np.array([[0.4, 0.4, 0.4], [0.6, 0, 0], [0, -0.5, 0], [0, 0, 0.51]])
    assert metrics_ref.obb_count(pts, [0, 0, 0], np.eye(3), [1, 1, 1]) == 2
    Rz = synth._rot_from_rotvec([0, 0, np.pi / 4])
    assert metrics_ref.obb_count([[0.6, 0.6, 0]], [0, 0, 0], Rz, [2, 0.2, 1]) == 1     # along the rotated x axis
    assert metrics_ref.obb_count([[0.6, -0.6, 0]], [0, 0, 0], Rz, [2, 0.2, 1]) == 0
    # episode rule: frames before the start frame do not count; success latches at 30 hits
    passed = [True] * 10 + [False] * 5 + [True] * 40
    out = metrics_ref.episode_rule(passed, start_frame=12, need=30)
    assert out[11] == (0, False) and out[15 + 28] == (29, False) and out[15 + 29] == (30, True) and out[-1][1]


def test_restated_obb_rule_agrees_with_a_convex_hull_test():
    """Open3D is not installable here, so its point-in-oriented-box rule (the sloth success test) is restated (parity
    unpinned).  An independent geometric test -- membership in the convex hull of the box's eight corners, by
    scipy's Delaunay triangulation -- selects the same points wherever a point is not within 1e-9 of a face."""
    from scipy.spatial import Delaunay
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(12)
    R = Rotation.random(random_state=5).as_matrix()
    center, extent = np.array([0.31, -0.12, 0.07]), np.array([0.20, 0.11, 0.06])
    x = center + rng.uniform(-0.16, 0.16, (20000, 3))
    corners = center + (np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * extent / 2) @ R.T
    hull = Delaunay(corners)
    inside_hull = hull.find_simplex(x) >= 0
    local = np.abs((x - center) @ R) - extent / 2
    clear = np.abs(local).min(1) > 1e-9                       # not on a face, where the two tests may round differently
    n = metrics_ref.obb_count(x[clear], center, R, extent)
    assert n == int(inside_hull[clear].sum()) and 200 < n < 19000
