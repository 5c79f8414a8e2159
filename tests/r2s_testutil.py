"""Shared builders for the parity tests: the same seeded inputs are fed to the
CPU oracle (oracle/) and to the CUDA path (through the C ABI)."""
import numpy as np

from real2sim_eval_b200 import synth


def oracle_from_scene(scene, n_substeps, mesh=None, **over):
    from oracle.physics_ref import SpringMassOracle
    p = dict(scene.params)
    p.pop("reverse_z", None)
    p.update(over)
    return SpringMassOracle(scene.x, scene.v, scene.springs, scene.rest, scene.log_Y, scene.mass,
                            n_substeps=n_substeps, mesh=mesh, **p)


def cuda_from_scenes(scenes, n_substeps, device="cuda", per_env_rest=True, **over):
    """BatchedSpringMass over a list of posed clones of one scene."""
    import torch
    from real2sim_eval_b200.physics import BatchedSpringMass
    s0 = scenes[0]
    p = dict(s0.params)
    p.update(over)
    rest = np.stack([s.rest for s in scenes]) if per_env_rest and len(scenes) > 1 else s0.rest
    sys = BatchedSpringMass(len(scenes), s0.springs, rest, num_particles=s0.N, n_substeps=n_substeps,
                            log_spring_Y=s0.log_Y, masses=s0.mass, device=device, **p)
    sys.set_state(np.stack([s.x for s in scenes]), np.stack([s.v for s in scenes]))
    if sys.self_collision:
        sys.create_resting_case()  # the reference constructor does this from the initial state (SMW:714-721)
    return sys


def gripper_mesh_dict(g):
    return dict(verts=g.verts, faces=g.faces, mesh_map=g.mesh_map, face_map=g.face_map, n_dyn_verts=len(g.verts))


def small_gaussians(seed, P, box=((-0.3, -0.3, 0.0), (0.3, 0.3, 0.4)), scale=0.02, sh_coeffs=1):
    """Random Gaussians in front of make_test_camera()."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(box[0], box[1], (P, 3)).astype(np.float32)
    scales = np.exp(rng.normal(np.log(scale), 0.6, (P, 3))).astype(np.float32)
    q = rng.normal(size=(P, 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    opa = (1 / (1 + np.exp(-rng.normal(1.0, 1.5, (P, 1))))).astype(np.float32)
    shs = rng.normal(0, 0.8, (P, sh_coeffs, 3)).astype(np.float32)
    return dict(means3D=means, scales=scales, rotations=q, opacities=opa, shs=shs)


def make_test_camera(W, H, eye=(0.9, 0.05, 0.5), target=(0.0, 0.0, 0.15), fov_deg=50.0):
    """Look-at camera in the reference's convention (camera looks down +z, y down)."""
    eye, target = np.asarray(eye, float), np.asarray(target, float)
    f = target - eye
    f /= np.linalg.norm(f)
    up = np.array([0.0, 0.0, 1.0])
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = r, d, f, eye
    fx = 0.5 * W / np.tan(np.radians(fov_deg) / 2)
    k = np.array([[fx, 0, W / 2.0 - 0.3], [0, fx, H / 2.0 + 0.2], [0, 0, 1]])
    return synth.setup_camera(W, H, k, np.linalg.inv(c2w))


# ---------------------------------------------------------------------------- physics cases (tests/phys_cases.py)
def load_phys_golden(name):
    """tests/golden/phys_<name>.npz (outputs of the reference's own spring_mass_warp.py, see
    tests/golden/make_physics_golden.py) + the case rebuilt from the seeded builders; the input digest must agree."""
    import os
    import phys_cases
    case = phys_cases.CASES[name]()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"phys_{name}.npz"))
    assert str(g["input_sha"]) == phys_cases.input_digest(case), f"{name}: seeded inputs drifted from the golden's"
    return case, g


def oracle_from_case(case, **kw):
    import phys_cases
    o = oracle_from_scene(case["scene"], case["n_substeps"], mesh=phys_cases.merged_mesh(case),
                          use_pusher=case["use_pusher"], **case["over"], **kw)
    if case["reset"] is not None:
        o.x[:], o.v[:] = case["reset"]
    return o


def cuda_from_case(case, E=1, **kw):
    import phys_cases
    c = cuda_from_scenes([case["scene"]] * E, case["n_substeps"], per_env_rest=False, use_pusher=case["use_pusher"],
                         **case["over"], **kw)
    m = phys_cases.merged_mesh(case)
    if m is not None:
        c.set_mesh(**m)
    if case["reset"] is not None:
        x, v = case["reset"]
        c.set_state(np.repeat(x[None], E, 0), np.repeat(v[None], E, 0))
    return c


def golden_coll_rows(g, k):
    """Candidate rows of frame k as a list of arrays."""
    num = g[f"f{k}_coll_num"]
    return num, np.split(g[f"f{k}_coll_flat"], np.cumsum(num)[:-1]) if len(num) else []
