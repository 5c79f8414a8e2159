"""GPU parity of r2s_eef_forward (through the C ABI) against the golden vectors made from the reference's own
SpringMassDynamicsModule.step and against oracle/eef_ref.py.  The grasp state (current_openness in float64,
grasped) must be exact; tables within 1e-6 m, velocities to float32 rounding (2e-6 relative)."""
import glob
import os

import numpy as np
import pytest

from real2sim_eval_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eef_*.npz")))
KEYS = ("interp_pts", "interp_center", "dyn_vel", "dyn_omega")


def close(a, b):
    return bool((np.abs(a - b) <= 1e-6 + 2e-6 * np.abs(b)).all())


def dev(a, shape=None):
    import torch
    a = np.asarray(a, np.float32)
    return torch.tensor(a.reshape(shape) if shape else a, device="cuda").contiguous()


def outputs(m, e=0):
    return dict(interp_pts=m.interp_pts[e].cpu().numpy(), interp_center=m.interp_center[e].cpu().numpy(),
                dyn_vel=m.dyn_vel[e].cpu().numpy(), dyn_omega=m.dyn_omega[e].cpu().numpy())


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference_golden(path):
    from real2sim_eval_b200.eef import BatchedEefMotion
    d = np.load(path)
    pusher = bool(d["use_pusher"])
    m = BatchedEefMotion(1, d["table"], d["init_eef_xyz"], dt=float(d["dt"]), n_substeps=int(d["n_substeps"]),
                         grasp_force_threshold=float(d["threshold"]), use_pusher=pusher, mesh_map=d["mesh_map"])
    for f in range(len(d["eef_xyz"])):
        m.forward(dev(d["eef_xyz"][f], (1, 3)), dev(d["eef_vel"][f], (1, 3)), dev(d["eef_rot"][f], (1, 3, 3)),
                  dev(d["eef_rot_vel"][f], (1, 3)), None if pusher else dev(d["openness_cmd"][f], (1,)),
                  collision_forces=dev(d["forces"][f])[None].contiguous())
        assert float(m.current_openness[0]) == float(d["ref_current_openness"][f]), f
        assert bool(m.grasped[0]) == bool(d["ref_grasped"][f]), f
        got = outputs(m)
        for k in KEYS:
            assert close(got[k], d["ref_" + k][f]), (f, k, np.abs(got[k] - d["ref_" + k][f]).max())


def test_batched_envs_in_place_into_the_physics_handle_match_the_oracle():
    """E environments with different commands; the tables are written into the physics handle's own buffers and
    the forces read from its collision_forces; the substeps then run on them."""
    import torch
    import r2s_testutil as util
    from oracle import eef_ref
    from real2sim_eval_b200.eef import BatchedEefMotion
    E, S = 6, 10
    sc = synth.make_rope(v_scale=0.01)
    center = (0.5, 0.0, 0.03)
    g = synth.make_gripper(center)
    table = synth.gripper_opening_table(center)
    phys = util.cuda_from_scenes([sc] * E, S, per_env_rest=False)
    phys.set_mesh(**util.gripper_mesh_dict(g))
    m = BatchedEefMotion(E, table, center, dt=sc.params["dt"], n_substeps=S, mesh_map=g.mesh_map, phys=phys)
    views = phys.motion_tables(per_env=True)
    rng = np.random.default_rng(11)
    faces = eef_ref.force_faces(g.mesh_map)
    cur, grasped = [None] * E, [False] * E
    for frame in range(4):
        xyz = (np.asarray(center) + rng.normal(size=(E, 3)) * 0.01).astype(np.float32)
        vel = rng.uniform(-0.2, 0.2, (E, 3)).astype(np.float32)
        rvel = (rng.normal(size=(E, 3)) * np.where(np.arange(E) % 2, 2.0, 0.01)[:, None]).astype(np.float32)
        rot = np.stack([synth._rot_from_rotvec(rng.normal(size=3) * 0.4) @ synth.EEF_ROT_DOWN for _ in range(E)]).astype(np.float32)
        cmd = np.clip(0.9 - 0.2 * frame - 0.05 * np.arange(E), 0, 1).astype(np.float32)
        forces = (rng.normal(size=(E, len(g.faces), 3)) * np.where(np.arange(E) < 3, 5e4, 20.0)[:, None, None]).astype(np.float32)
        phys.collision_forces.copy_(torch.tensor(forces, device="cuda"))
        m.forward(dev(xyz), dev(vel), dev(rot), dev(rvel), dev(cmd))
        for e in range(E):
            o = eef_ref.eef_step(table, center, xyz[e], vel[e], rot[e], rvel[e], cmd[e], dt=sc.params["dt"], n_substeps=S,
                                 current_openness=cur[e], grasped=grasped[e], forces=forces[e], faces=faces)
            cur[e], grasped[e] = o["current_openness"], o["grasped"]
            assert float(m.current_openness[e]) == cur[e] and bool(m.grasped[e]) == grasped[e], (frame, e)
            got = dict(zip(KEYS, (v[e].cpu().numpy() for v in views)))
            for k in KEYS:
                assert close(got[k], o[k]), (frame, e, k)
    assert any(grasped) and not all(grasped)
    # the substeps run on the tables just written (same result as uploading them through set_mesh_motion)
    x0, v0 = phys.get_state()
    phys.step()
    xa, _ = phys.get_state()
    phys.set_state(x0, v0)
    phys.set_mesh_motion(*[v.clone() for v in views])
    phys.step()
    xb, _ = phys.get_state()
    assert torch.equal(xa, xb)


def test_pusher_rows_and_argument_checks():
    import torch
    from real2sim_eval_b200 import _lib
    from real2sim_eval_b200.eef import BatchedEefMotion
    pusher = synth.make_pusher((0.4, 0.0, 0.05), n_circ=8, n_len=3)
    table = np.repeat(pusher.verts[None], 101, 0)
    m = BatchedEefMotion(2, table, (0.4, 0.0, 0.05), dt=5e-5, n_substeps=5, use_pusher=True)
    z = torch.zeros((2, 3), device="cuda")
    rot = torch.tensor(np.stack([synth.EEF_ROT_DOWN] * 2), device="cuda")
    m.forward(torch.tensor([[0.4, 0.0, 0.05]] * 2, device="cuda"), z, rot, z)
    assert m.dyn_vel.shape == (2, 1, 3) and torch.all(m.current_openness == 1.0)
    # zero velocity at the initial pose: every substep row is the sampled mesh
    assert np.abs(m.interp_pts[0, -1].cpu().numpy() - pusher.verts).max() < 1e-6
    with pytest.raises(ValueError):
        m.forward(z[:1], z, rot, z)
    with pytest.raises(_lib.R2SError, match="no CPU path"):
        BatchedEefMotion(1, table, (0, 0, 0), dt=5e-5, n_substeps=5, use_pusher=True, device="cpu")
