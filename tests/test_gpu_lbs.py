"""GPU parity of the LBS kernels (through the C ABI) against the golden vectors made from the reference's own
interpolate_motions and against the CPU restatement.  Tolerance 1e-5 m absolute on the transformed positions:
the kernel fits each bone rotation from an fp32 Jacobi eigen-decomposition of F^T F, the reference from a LAPACK
SVD of F; offsets (xyz - bone) are centimetres, so rotation errors of ~1e-5 rad stay below 1e-6 m."""
import glob
import os

import numpy as np
import pytest

from oracle import lbs_ref

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lbs_*.npz")))
TOL = 1e-5


def _cuda(bones, motions, rel, xyz, w, wi):
    import torch
    from real2sim_eval_b200.lbs import interpolate_motions
    t = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device="cuda")
    out, rot, wts = interpolate_motions(bones=t(bones), motions=t(motions), relations=rel, xyz=t(xyz), weights=t(w),
                                        weights_indices=t(wi, torch.int64), quat=None, device="cuda")
    assert rot is None
    return out.cpu().numpy()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_cuda_matches_reference_golden(path):
    d = np.load(path)
    out = _cuda(d["bones"], d["motions"], d["relations"], d["xyz"], d["weights"], d["weights_indices"])
    assert np.abs(out - d["out"]).max() <= TOL


@pytest.mark.parametrize("slots", [False, True], ids=["bone-order", "morton-slots"])
def test_batched_envs_and_rank_deficient_env(slots):
    """(slots: the layout hint of r2s_lbs.h -- bone transforms stored along a Morton curve, each Gaussian's bones
    listed by ascending slot; same sums in another order, and the identity fallback still reads the bones by index.)
    E = 3 environments sharing relations / weights: env 1 has a collinear bone cluster AFTER..BEFORE the frame, so
    the reference rule gives it the identity rotation for every bone; envs 0 and 2 are regular."""
    import torch
    from real2sim_eval_b200.lbs import BatchedLBS
    rng = np.random.default_rng(4)
    n, P, n_obj = 400, 900, 700
    base = rng.uniform(0, 0.1, (n, 3)).astype(np.float32)
    rel = lbs_ref.knn_relations(base, 8)
    pts = (base[rng.integers(0, n, n_obj)] + rng.normal(0, 0.003, (n_obj, 3))).astype(np.float32)
    w, wi = lbs_ref.knn_weights(base, pts, 16)
    bones = np.stack([base, base.copy(), base + np.float32(0.5)])
    k = rel[0][:8]                                   # collapse bone 0's neighbourhood onto a line in env 1
    bones[1][k] = bones[1][0] + np.outer(np.arange(1, 9), [0.001, 0.0, 0.0]).astype(np.float32)
    motions = np.stack([0.02 * np.sin(30 * b[:, [1, 2, 0]]) for b in bones]).astype(np.float32)
    means = np.zeros((3, P, 3), np.float32)
    means[:, :n_obj] = pts[None] + (bones - base[None]).mean(1, keepdims=True)
    means[:, n_obj:] = 7.0
    pad = lambda a: torch.tensor(np.concatenate([a, np.zeros_like(a[..., :1])], -1)).cuda().contiguous()
    lbs = BatchedLBS(3, n, P, n_obj, rel, w, wi, bone_positions=base if slots else None)
    assert (lbs.bone_slot is not None) == slots
    m = torch.tensor(means).cuda()
    lbs.forward(pad(bones), pad(bones + motions), m)
    flags = lbs.rank_flags.cpu().numpy().tolist()
    assert flags == [1, 0, 1]
    out = m.cpu().numpy()
    assert (out[:, n_obj:] == 7.0).all(), "rows >= n_obj are untouched"
    for e in range(3):
        want = lbs_ref.interpolate_motions(bones[e], motions[e], rel, means[e, :n_obj], w, wi)
        assert np.abs(out[e, :n_obj] - want).max() <= TOL, e


def test_exactly_collinear_chain_is_rank_one():
    """a thin chain whose bone neighbourhoods are exactly collinear along a GENERIC direction, stretched
    along itself: every F is rank 1 to rounding (s2 / s1 ~ 3e-8; numpy and torch both report rank 1 for all bones), so
    the reference ends with the identity for every bone.  Singular values taken from the eigenvalues of F^T F carry
    ~sqrt(eps) * s1 of noise and called these rank 2; |F v2| does not."""
    import torch
    from real2sim_eval_b200.lbs import BatchedLBS
    n, n_obj = 300, 500
    d = np.array([1.0, 2.0, 3.0]) / np.sqrt(14.0)
    base = (np.array([0.3, -0.2, 0.1]) + np.outer(np.arange(n) * 0.004, d)).astype(np.float32)
    rel = lbs_ref.knn_relations(base, 8)
    motions = np.outer(0.0005 * np.arange(n), d).astype(np.float32)
    _, ok = lbs_ref.bone_rotations(base, motions, rel)
    assert not ok, "the reference's rank test (numpy SVD) finds rank-1 bones"
    rng = np.random.default_rng(9)
    pts = (base[rng.integers(0, n, n_obj)] + rng.normal(0, 0.002, (n_obj, 3))).astype(np.float32)
    w, wi = lbs_ref.knn_weights(base, pts, 16)
    pad = lambda a: torch.tensor(np.concatenate([a, np.zeros_like(a[..., :1])], -1)[None]).cuda().contiguous()
    lbs = BatchedLBS(1, n, n_obj, n_obj, rel, w, wi)
    m = torch.tensor(pts[None]).cuda()
    lbs.forward(pad(base), pad(base + motions), m)
    assert lbs.rank_flags.cpu().numpy().tolist() == [0]
    want = lbs_ref.interpolate_motions(base, motions, rel, pts, w, wi)      # identity rotations
    assert np.abs(m[0].cpu().numpy() - want).max() <= TOL


def test_dropin_rejects_the_quat_path():
    import torch
    from real2sim_eval_b200.lbs import interpolate_motions
    z = torch.zeros((4, 3), device="cuda")
    with pytest.raises(NotImplementedError):
        interpolate_motions(z, z, np.zeros((4, 2), np.int64), z, quat=torch.zeros((4, 4), device="cuda"),
                            weights=torch.ones((4, 1), device="cuda"), weights_indices=torch.zeros((4, 1), dtype=torch.int64))
