"""Generates tests/golden/eef_*.npz by running the REFERENCE's own SpringMassDynamicsModule.step
(/root/reference/sim/physics/phystwin.py:362-510, unmodified, torch on CPU) over short action sequences
with a fake simulator that records the arguments of set_mesh_interactive (oracle/eef_ref.py explains the
stubs; kornia's axis_angle_to_rotation_matrix is the one restated piece).  Run in the build container
(needs /root/reference):   python tests/golden/make_eef_golden.py
Each file holds, per frame: the inputs (eef pose / velocities / commanded opening / finger forces of the
previous frame) and the reference's outputs (vertex table, centres, velocities, current_openness, grasped)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import eef_ref                   # noqa: E402
from real2sim_eval_b200 import synth         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sequence(seed, frames, S, use_pusher, rot_scale):
    rng = np.random.default_rng(seed)
    center = (0.5, 0.0, 0.03)
    if use_pusher:
        pusher = synth.make_pusher(center, n_circ=10, n_len=4)
        table = np.repeat(pusher.verts[None], 101, 0).astype(np.float32)   # robot_pc_transformations.py:222-225
        mesh_map, n_faces = pusher.mesh_map, len(pusher.faces)
    else:
        g = synth.make_gripper(center)
        table = synth.gripper_opening_table(center)
        mesh_map, n_faces = g.mesh_map, len(g.faces)
    init = np.asarray(center, np.float32)
    xyz = init.copy()
    cmds = np.clip(0.95 - 0.09 * np.arange(frames), 0.0, 1.0)
    cmds[frames * 3 // 4:] = 0.8                              # re-open at the end
    ins = dict(eef_xyz=[], eef_vel=[], eef_rot=[], eef_rot_vel=[], openness_cmd=[], forces=[])
    for f in range(frames):
        vel = rng.uniform(-0.1, 0.1, 3).astype(np.float32)
        rot = (synth._rot_from_rotvec(rng.normal(size=3) * 0.3) @ synth.EEF_ROT_DOWN).astype(np.float32)
        rvel = (rng.normal(size=3) * (rot_scale if f % 2 else 0.01)).astype(np.float32)   # both kornia branches
        forces = np.zeros((n_faces, 3), np.float32)
        if frames // 3 <= f < frames * 2 // 3:
            forces[:] = rng.normal(size=forces.shape) * 4e4   # both fingers loaded: grasp establishes / holds
        elif f == frames * 2 // 3:
            forces[:] = rng.normal(size=forces.shape) * 300   # neither large nor small: grasped closing by 0.05
        elif f == frames * 2 // 3 + 1:
            forces[:] = rng.normal(size=forces.shape) * 10    # both small: release
        for k, v in zip(ins, (xyz.copy(), vel, rot, rvel, np.float32(cmds[f]), forces)):
            ins[k].append(v)
        xyz = (xyz + vel / np.float32(30)).astype(np.float32)
    return table, init, mesh_map, n_faces, {k: np.stack(v) for k, v in ins.items()}


def main():
    mod = eef_ref.load_reference()
    assert mod is not None, "needs /root/reference"
    cases = [("eef_a_gripper_s10", dict(seed=1, frames=14, S=10, use_pusher=False, rot_scale=2.0)),
             ("eef_b_gripper_s40", dict(seed=2, frames=10, S=40, use_pusher=False, rot_scale=0.5)),
             ("eef_c_pusher_s10", dict(seed=3, frames=5, S=10, use_pusher=True, rot_scale=1.0))]
    dt = 5e-5
    for name, c in cases:
        table, init, mesh_map, n_faces, ins = sequence(**c)
        ref = eef_ref.ReferenceModule(mod, dt=dt, n_substeps=c["S"], threshold=3e4, use_pusher=c["use_pusher"],
                                      mesh_map=mesh_map, n_faces=n_faces)
        func = eef_ref.make_eef_pts_func(table)
        outs = dict(interp_pts=[], interp_center=[], dyn_vel=[], dyn_omega=[], current_openness=[], grasped=[])
        for f in range(c["frames"]):
            r = ref.step(func, init, ins["eef_xyz"][f], ins["eef_vel"][f], ins["eef_rot"][f], ins["eef_rot_vel"][f],
                         ins["openness_cmd"][f], ins["forces"][f])
            for k in outs:
                outs[k].append(r[k])
        outs = {"ref_" + k: np.stack(v) for k, v in outs.items()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), table=table, init_eef_xyz=init, mesh_map=mesh_map,
                            dt=np.float64(dt), n_substeps=np.int32(c["S"]), use_pusher=np.int32(c["use_pusher"]),
                            threshold=np.float32(3e4), **ins, **outs)
        print(name, "frames", c["frames"], "openness", np.round(outs["ref_current_openness"], 3).tolist(),
              "grasped", outs["ref_grasped"].astype(int).tolist())


if __name__ == "__main__":
    main()
