"""GPU parity of r2s_links_forward (through the C ABI) against oracle/links_ref.py and the golden vectors made
from the reference's transform_gs_torch.  Tolerance: 2e-6 absolute on positions (metres) and quaternion
components -- fp32 rounding of two 4x4 products and a 3x3 apply; the arithmetic is otherwise the reference's."""
import glob
import os

import numpy as np
import pytest

from real2sim_eval_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "links_*.npz")))
TOL = 2e-6


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference_golden(path):
    from real2sim_eval_b200.links import transform_gs
    d = np.load(path)
    p, q = transform_gs(d["points"], d["quats"], d["total_mask"], list(synth.XARM_LINK_IDS), d["link_pose"],
                        d["base_pose"], d["link_offset"])
    p, q = p.cpu().numpy(), q.cpu().numpy()
    assert np.abs(p - d["out_points"]).max() < TOL
    assert np.abs(q - d["out_quats"]).max() < TOL
    unmoved = d["link_id"] < 0
    assert np.array_equal(p[unmoved], d["points"][unmoved])


def test_batched_envs_match_oracle_and_leave_other_rows_alone():
    import torch
    from oracle import links_ref
    from real2sim_eval_b200.links import BatchedLinkTransform
    E, P, first = 5, 3000, 700
    scan = synth.make_robot_scan(1800, 21)
    lt = BatchedLinkTransform(E, P, first, scan.link_id, scan.points, scan.quats, scan.link_offset, scan.base_pose)
    poses = np.stack([synth.robot_link_poses(scan, 300 + e, 0.2 + 0.3 * e) for e in range(E)])
    means = torch.full((E, P, 3), 7.0, device="cuda")
    rots = torch.full((E, P, 4), 9.0, device="cuda")
    lt.forward(torch.tensor(poses, dtype=torch.float32, device="cuda").contiguous(), means, rots)
    m, r = means.cpu().numpy(), rots.cpu().numpy()
    for e in range(E):
        p, q = links_ref.transform_gs(scan.points, scan.quats, scan.link_id, poses[e], scan.base_pose, scan.link_offset)
        assert np.abs(m[e, first:first + 1800] - p).max() < TOL
        assert np.abs(r[e, first:first + 1800] - q).max() < TOL
    assert (m[:, :first] == 7.0).all() and (m[:, first + 1800:] == 7.0).all()
    assert (r[:, :first] == 9.0).all() and (r[:, first + 1800:] == 9.0).all()


def test_half_turn_branches_and_argument_errors():
    import torch
    from oracle import links_ref
    from real2sim_eval_b200 import _lib
    from real2sim_eval_b200.links import BatchedLinkTransform
    # rotations by ~pi about x, y, z take the three non-trace branches of the matrix -> quaternion conversion
    L = 3
    base = np.stack([np.eye(4)] * L)
    pose = np.stack([synth._rigid(ax, (0.1, 0.2, 0.3)) for ax in ([3.1, 0, 0], [0, 3.1, 0], [0, 0, 3.1])])
    rng = np.random.default_rng(4)
    pts = rng.normal(size=(90, 3)).astype(np.float32)
    qs = rng.normal(size=(90, 4)).astype(np.float32)
    ids = (np.arange(90) % 4 - 1).astype(np.int32)         # -1, 0, 1, 2
    lt = BatchedLinkTransform(1, 90, 0, ids, pts, qs, base, base)
    means = torch.empty((1, 90, 3), device="cuda")
    rots = torch.empty((1, 90, 4), device="cuda")
    lt.forward(torch.tensor(pose[None], dtype=torch.float32, device="cuda").contiguous(), means, rots)
    p, q = links_ref.transform_gs(pts, qs, ids, pose, base, base)
    assert np.abs(means[0].cpu().numpy() - p).max() < TOL and np.abs(rots[0].cpu().numpy() - q).max() < TOL
    with pytest.raises(ValueError):
        BatchedLinkTransform(1, 90, 0, np.full(90, 5, np.int32), pts, qs, base, base)      # slot outside the table
    with pytest.raises(ValueError):
        BatchedLinkTransform(1, 50, 0, ids, pts, qs, base, base)                            # rows do not fit
    a = _lib.LinksArgs()
    a.E, a.L, a.P, a.n_robot = 1, 0, 10, 0
    lib = _lib.load()
    assert lib.r2s_links_forward(a, None) != 0 and b"bad sizes" in lib.r2s_last_error()
