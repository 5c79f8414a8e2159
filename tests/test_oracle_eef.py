"""The end-effector oracle (oracle/eef_ref.py) against the golden vectors made from the reference's own
SpringMassDynamicsModule.step (tests/golden/make_eef_golden.py), against the live reference when
/root/reference is mounted, against scipy's interp1d, and known answers."""
import glob
import os

import numpy as np
import pytest

from oracle import eef_ref
from real2sim_eval_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eef_*.npz")))
KEYS = ("interp_pts", "interp_center", "dyn_vel", "dyn_omega")


def close(a, b):
    """1e-6 m on positions; velocities (a few m/s while the fingers close) to float32 rounding."""
    return bool((np.abs(a - b) <= 1e-6 + 2e-6 * np.abs(b)).all())


def replay(d, step_fn):
    """Runs step_fn frame by frame over a golden sequence, carrying (current_openness, grasped)."""
    cur, grasped, out = None, False, []
    pusher = bool(d["use_pusher"])
    faces = None if pusher else eef_ref.force_faces(d["mesh_map"])
    for f in range(len(d["eef_xyz"])):
        o = step_fn(d, f, cur, grasped, faces)
        cur, grasped = o["current_openness"], o["grasped"]
        out.append(o)
    return out


def oracle_step(d, f, cur, grasped, faces):
    return eef_ref.eef_step(d["table"], d["init_eef_xyz"], d["eef_xyz"][f], d["eef_vel"][f], d["eef_rot"][f],
                            d["eef_rot_vel"][f], d["openness_cmd"][f], dt=float(d["dt"]), n_substeps=int(d["n_substeps"]),
                            current_openness=cur, grasped=grasped, forces=d["forces"][f], faces=faces,
                            threshold=float(d["threshold"]), use_pusher=bool(d["use_pusher"]))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_golden(path):
    d = np.load(path)
    for f, o in enumerate(replay(d, oracle_step)):
        assert o["current_openness"] == d["ref_current_openness"][f], f      # the hysteresis is exact (float64)
        assert o["grasped"] == bool(d["ref_grasped"][f]), f
        for k in KEYS:   # torch's bmm / mean and numpy's differ by float32 rounding only
            assert close(o[k], d["ref_" + k][f]), (f, k)


def test_golden_set_is_present_and_covers_the_grasp_states():
    assert len(GOLD) == 3
    d = np.load(GOLD[0])
    g, o = d["ref_grasped"], d["ref_current_openness"]
    assert g.any() and not g.all()
    held = [f for f in range(1, len(g)) if g[f] and o[f] == o[f - 1] and d["openness_cmd"][f] < o[f]]
    slow = [f for f in range(1, len(g)) if g[f] and abs((o[f - 1] - o[f]) - 0.05) < 1e-9]
    assert held and slow, "the sequence must hold a grasp and close by 0.05 while grasped"


def test_interpolation_is_scipys():
    table = synth.gripper_opening_table((0.5, 0.0, 0.03))
    func = eef_ref.make_eef_pts_func(table)
    rng = np.random.default_rng(0)
    for x in list(rng.uniform(0, 1, 40)) + [0.0, 1.0, 0.5, 0.37, 0.99, 0.01, 0.3, 0.30000000000000004]:
        a, b = func(x), eef_ref.interp_table(table, x)
        assert a.dtype == np.float64 and np.array_equal(a, b), x


@pytest.mark.skipif(not os.path.exists(eef_ref.REF_FILE), reason="/root/reference not mounted")
def test_oracle_matches_live_reference():
    mod = eef_ref.load_reference()
    center = (0.45, 0.05, 0.02)
    g = synth.make_gripper(center)
    table = synth.gripper_opening_table(center)
    S, dt = 7, 5e-5
    ref = eef_ref.ReferenceModule(mod, dt=dt, n_substeps=S, threshold=3e4, use_pusher=False, mesh_map=g.mesh_map,
                                  n_faces=len(g.faces))
    func, faces = eef_ref.make_eef_pts_func(table), eef_ref.force_faces(g.mesh_map)
    rng = np.random.default_rng(5)
    cur, grasped = None, False
    for f in range(6):
        xyz = (np.asarray(center) + rng.normal(size=3) * 0.01).astype(np.float32)
        vel, rvel = rng.uniform(-0.2, 0.2, 3), rng.normal(size=3) * (0.02 if f % 2 else 3.0)
        rot = synth._rot_from_rotvec(rng.normal(size=3)) @ synth.EEF_ROT_DOWN
        cmd = np.float32(0.9 - 0.15 * f)
        forces = (rng.normal(size=(len(g.faces), 3)) * (5e4 if f in (2, 3) else 50)).astype(np.float32)
        r = ref.step(func, center, xyz, vel, rot, rvel, cmd, forces)
        o = eef_ref.eef_step(table, center, xyz, vel, rot, rvel, cmd, dt=dt, n_substeps=S, current_openness=cur,
                             grasped=grasped, forces=forces, faces=faces)
        cur, grasped = o["current_openness"], o["grasped"]
        assert (cur, grasped) == (r["current_openness"], r["grasped"])
        for k in KEYS:
            assert close(o[k], r[k]), (f, k)


def test_known_answers():
    # pure translation, gripper pointing down, opening unchanged: every vertex moves by vel * t
    center = (0.5, 0.0, 0.03)
    table = synth.gripper_opening_table(center)
    o = eef_ref.eef_step(table, center, center, [0.3, 0.0, -0.6], synth.EEF_ROT_DOWN, [0, 0, 0], 0.5, dt=5e-5, n_substeps=4)
    rest = eef_ref.interp_table(table, 0.5)
    for s in range(4):
        want = rest + np.array([0.3, 0.0, -0.6]) * 5e-5 * (s + 1)
        assert np.abs(o["interp_pts"][s] - want).max() < 1e-6
    assert np.allclose(o["dyn_vel"], [[0.15, 0, -0.3]] * 2, atol=1e-7) and np.allclose(o["dyn_omega"], 0)
    # closing from 0.5 to 0.4 over the frame: fingers approach symmetrically, closing velocity +-y
    o2 = eef_ref.eef_step(table, center, center, [0, 0, 0], synth.EEF_ROT_DOWN, [0, 0, 0], 0.4, dt=5e-5, n_substeps=4,
                          current_openness=0.5)
    end = eef_ref.interp_table(table, 0.4)
    assert np.abs(o2["interp_pts"][-1] - end).max() < 1e-6
    assert o2["dyn_vel"][0, 1] > 0 > o2["dyn_vel"][1, 1] and abs(o2["dyn_vel"][0, 1] + o2["dyn_vel"][1, 1]) < 1e-4
    # both branches of the axis-angle conversion are rotations about the right axis
    for aa in ([0.3, -0.2, 0.5], [2e-4, -3e-4, 1e-4]):
        R = eef_ref.axis_angle_to_rotation_matrix([aa])[0]
        assert np.abs(R - synth._rot_from_rotvec(aa)).max() < 2e-6
    # hysteresis: closing command, both fingers loaded -> opening held and grasp set; small forces -> released
    faces = [0, 1, 2, 3, 4, 5]
    big, small = np.full((6, 3), 2e4, np.float32), np.full((6, 3), 1.0, np.float32)
    assert eef_ref.hysteresis(0.2, 0.5, False, big, faces, 3e4) == (0.5, 0.5, 0.5, True)
    assert eef_ref.hysteresis(0.2, 0.5, True, small, faces, 3e4)[2:] == (float(np.float32(0.2)), False)
    mid = np.full((6, 3), 100.0, np.float32)   # |sum| ~ 520: neither small nor large
    assert eef_ref.hysteresis(0.2, 0.5, True, mid, faces, 3e4) == (0.45, 0.5, 0.45, True)


def test_restated_kornia_axis_angle_agrees_with_scipy():
    """kornia is not installable here, so `axis_angle_to_rotation_matrix` is restated from its published algorithm
    (parity unpinned).  scipy's Rotation.from_rotvec is an independent implementation of the same map: the restated
    function equals it to 3e-6 over rotation angles from 1e-5 rad (the first-order branch, theta^2 <= 1e-6) to pi --
    what is left unpinned is kornia's `+ 1e-6` in the axis normalisation and its branch threshold, both restated."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    axis = rng.normal(size=(3000, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    theta = np.concatenate([10 ** rng.uniform(-5, -3, 1000), rng.uniform(1e-3, 0.05, 1000), rng.uniform(0.05, np.pi, 1000)])
    aa = (axis * theta[:, None]).astype(np.float32)
    R = eef_ref.axis_angle_to_rotation_matrix(aa)
    want = Rotation.from_rotvec(aa.astype(np.float64)).as_matrix()
    assert np.abs(R - want).max() < 3e-6
    assert (theta ** 2 <= 1e-6).sum() > 100 and (theta ** 2 > 1e-6).sum() > 100, "both branches exercised"
