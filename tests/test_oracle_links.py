"""The links oracle (oracle/links_ref.py) against the golden vectors made from the reference's own
RobotPcSampler.transform_gs_torch (tests/golden/make_links_golden.py), against the live reference when
/root/reference is mounted, and known answers."""
import glob
import os

import numpy as np
import pytest

from oracle import links_ref
from real2sim_eval_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "links_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_golden(path):
    d = np.load(path)
    p, q = links_ref.transform_gs(d["points"], d["quats"], d["link_id"], d["link_pose"], d["base_pose"], d["link_offset"])
    # torch's matmul / inverse and numpy's differ by float32 rounding only
    assert np.abs(p - d["out_points"]).max() < 2e-6
    assert np.abs(q - d["out_quats"]).max() < 2e-6
    unmoved = d["link_id"] < 0
    assert unmoved.any() and np.array_equal(p[unmoved], d["points"][unmoved])


def test_golden_set_is_present():
    assert len(GOLD) == 3


@pytest.mark.skipif(not os.path.exists(links_ref.REF_FILE), reason="/root/reference not mounted")
def test_oracle_matches_live_reference():
    mod = links_ref.load_reference()
    scan = synth.make_robot_scan(900, 77)
    pose = synth.robot_link_poses(scan, 78, 0.8)
    rp, rq = links_ref.reference_transform_gs(mod, scan.points, scan.quats, scan.total_mask, list(synth.XARM_LINK_IDS),
                                              scan.link_names, pose, scan.base_pose, scan.link_offset)
    p, q = links_ref.transform_gs(scan.points, scan.quats, scan.link_id, pose, scan.base_pose, scan.link_offset)
    assert np.abs(p - rp).max() < 2e-6 and np.abs(q - rq).max() < 2e-6
    # the reference's in-tree quaternion product is the restated one, bit for bit
    import torch
    a = np.random.default_rng(0).normal(size=(50, 4)).astype(np.float32)
    b = np.random.default_rng(1).normal(size=(50, 4)).astype(np.float32)
    assert np.array_equal(mod.quat_mult_torch(torch.tensor(a), torch.tensor(b)).numpy(), links_ref.quat_mult(a, b))


def test_known_answers():
    # rest pose -> identity transform: points unchanged (to rounding), quaternions just normalised
    scan = synth.make_robot_scan(300, 5)
    p, q = links_ref.transform_gs(scan.points, scan.quats, scan.link_id, scan.base_pose, scan.base_pose, scan.link_offset)
    assert np.abs(p - scan.points).max() < 1e-6
    qn = links_ref.normalize(scan.quats)
    assert np.abs(np.abs((q * qn).sum(-1)) - 1).max() < 1e-5
    # quarter turn about z of a single link, no offset: (1,0,0) -> (0,1,0), identity quat -> (c,0,0,s)
    Rz = np.eye(4); Rz[:2, :2] = [[0, -1], [1, 0]]
    p, q = links_ref.transform_gs([[1, 0, 0]], [[2, 0, 0, 0]], [0], Rz[None], np.eye(4)[None], np.eye(4)[None])
    assert np.allclose(p, [[0, 1, 0]], atol=1e-7)
    assert np.allclose(q, [[np.sqrt(0.5), 0, 0, np.sqrt(0.5)]], atol=1e-6)
    # the matrix -> quaternion conversion reproduces the rotation in all four branches
    rng = np.random.default_rng(3)
    for axis in ([0.1, 0.2, 0.3], [3.1, 0, 0], [0, 3.1, 0], [0, 0, 3.1]):
        R = synth._rot_from_rotvec(axis).astype(np.float32)
        w, x, y, z = links_ref.rotation_matrix_to_quaternion(R)
        R2 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(R - R2).max() < 1e-5


def test_restated_kornia_quaternion_agrees_with_scipy():
    """kornia is not installable here, so `rotation_matrix_to_quaternion` is restated from its published four-branch
    algorithm (parity unpinned).  An independent implementation narrows what is unpinned to kornia's branch choice and
    its eps: scipy's Rotation.from_matrix gives the same rotation (as (x, y, z, w), up to the overall sign) on random
    rotations that exercise all four branches, to 2e-6."""
    from scipy.spatial.transform import Rotation
    rot = Rotation.random(4000, random_state=7)
    R = rot.as_matrix().astype(np.float32)
    q = links_ref.rotation_matrix_to_quaternion(R)                       # (w, x, y, z)
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    d = np.stack([R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]], 1)
    assert (tr > 0).any() and all(((tr <= 0) & (d.argmax(1) == k)).any() for k in range(3)), "all four branches hit"
    s = Rotation.from_matrix(R.astype(np.float64)).as_quat()[:, [3, 0, 1, 2]]
    sign = np.sign((q * s).sum(1, keepdims=True))
    assert np.abs(q * sign - s).max() < 2e-6
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-6
