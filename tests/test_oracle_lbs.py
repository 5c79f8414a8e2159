"""CPU tests of the LBS oracle (oracle/lbs_ref.py) against golden vectors generated from the reference's own
`interpolate_motions` (tests/golden/make_lbs_golden.py) and, when /root/reference is mounted, against the
live function."""
import glob
import os

import numpy as np
import pytest

from oracle import lbs_ref

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lbs_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_restatement_matches_reference_golden(path):
    d = np.load(path)
    out = lbs_ref.interpolate_motions(d["bones"], d["motions"], d["relations"], d["xyz"], d["weights"],
                                      d["weights_indices"])
    assert np.abs(out - d["out"]).max() <= 1e-6


def test_golden_fixtures_present():
    assert len(GOLD) >= 3


def test_rigid_motion_is_reproduced_exactly_and_reflections_are_removed():
    rng = np.random.default_rng(1)
    bones = rng.uniform(0, 0.1, (300, 3)).astype(np.float32)
    rel = lbs_ref.knn_relations(bones, 8)
    pts = rng.uniform(0, 0.1, (400, 3)).astype(np.float32)
    w, wi = lbs_ref.knn_weights(bones, pts, 16)
    ang = 0.7
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    new = (bones - 0.05) @ Rz.T + 0.05 + np.array([0.01, 0.0, -0.02])
    out = lbs_ref.interpolate_motions(bones, (new - bones).astype(np.float32), rel, pts, w, wi)
    assert np.abs(out - ((pts - 0.05) @ Rz.T + 0.05 + np.array([0.01, 0.0, -0.02]))).max() < 1e-6
    R, ok = lbs_ref.bone_rotations(bones, (new - bones).astype(np.float32), rel)
    assert ok and np.allclose(R, Rz[None], atol=1e-5)
    mirrored = bones.copy()
    mirrored[:, 0] = 0.1 - mirrored[:, 0]
    R, ok = lbs_ref.bone_rotations(bones, mirrored - bones, rel)
    assert ok and np.allclose(np.linalg.det(R), 1.0, atol=1e-4), "det F < 0: the reflection is removed (Kabsch)"


def test_one_rank_deficient_bone_gives_identity_everywhere():
    """transform_utils.py:159-167: R is built for the rank >= 2 bones only; when a bone is missing the
    assignment into bone_transforms raises and the except branch sets the identity for ALL bones."""
    rng = np.random.default_rng(2)
    bones = rng.uniform(0, 0.1, (60, 3)).astype(np.float32)
    bones[:10] = np.stack([np.linspace(0.5, 0.6, 10), np.full(10, 0.5), np.full(10, 0.5)], 1)  # a collinear cluster
    rel = lbs_ref.knn_relations(bones, 8)
    motions = (0.1 * bones[:, [1, 2, 0]] - 0.01).astype(np.float32)                              # non-rigid
    R, ok = lbs_ref.bone_rotations(bones, motions, rel)
    assert not ok and np.array_equal(R, np.tile(np.eye(3, dtype=np.float32), (60, 1, 1)))
    ref = lbs_ref.load_reference()
    if ref is not None:
        import torch
        pts = rng.uniform(0, 0.1, (50, 3)).astype(np.float32)
        w, wi = lbs_ref.knn_weights(bones, pts, 16)
        t = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt)
        out, _, _ = ref(bones=t(bones), motions=t(motions), relations=rel, xyz=t(pts), weights=t(w),
                        weights_indices=t(wi, torch.int64), quat=None, device="cpu")
        mine = lbs_ref.interpolate_motions(bones, motions, rel, pts, w, wi)
        assert np.abs(out.numpy() - mine).max() < 1e-6


@pytest.mark.skipif(lbs_ref.load_reference() is None, reason="/root/reference not mounted")
def test_restatement_matches_live_reference_function():
    import torch
    ref = lbs_ref.load_reference()
    rng = np.random.default_rng(7)
    for trial in range(3):
        bones = rng.uniform(0, 0.2, (250, 3)).astype(np.float32)
        rel = lbs_ref.knn_relations(bones, 8)
        pts = (bones[rng.integers(0, 250, 600)] + rng.normal(0, 0.004, (600, 3))).astype(np.float32)
        w, wi = lbs_ref.knn_weights(bones, pts, 16)
        motions = (0.05 * np.sin(20 * bones[:, [2, 0, 1]]) + rng.normal(0, 1e-3, bones.shape)).astype(np.float32)
        t = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt)
        out, _, _ = ref(bones=t(bones), motions=t(motions), relations=rel, xyz=t(pts), weights=t(w),
                        weights_indices=t(wi, torch.int64), quat=None, device="cpu")
        mine = lbs_ref.interpolate_motions(bones, motions, rel, pts, w, wi)
        assert np.abs(out.numpy() - mine).max() < 2e-6
