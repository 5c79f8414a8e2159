"""Seeded physics scenarios shared by tests/golden/make_physics_golden.py (which runs the reference's own
spring_mass_warp.py on them, see oracle/warp_exec.py) and by the parity tests (which run the C oracle and the
CUDA path on the same inputs and compare with the committed outputs).

A case is a dict:
  scene      synth.Scene (x, v, springs, rest, log_Y, mass, params)
  n_substeps substeps per frame
  frames     list of per-frame mesh motion tables (or [None] * n) -- one entry per frame stepped
  meshes     None or dict(dynamic=[(verts, faces), ...], static=[(verts, faces), ...])
  use_pusher bool
  reset      None, or (x, v) to install with set_init_state AFTER construction (the constructor's resting
             pairs come from scene.x; the run starts from `reset`)
  over       parameter overrides on top of scene.params
"""
import hashlib

import numpy as np

from real2sim_eval_b200 import synth

DT = 5e-5


def _two_ropes(dz, vz):
    half = synth.make_rope(n=1024)
    n = half.N
    x = np.concatenate([half.x, half.x + np.array([0.0, 0.0, dz], np.float32)], 0).astype(np.float32)
    v = np.zeros_like(x)
    v[n:, 2] = vz
    springs = np.concatenate([half.springs, half.springs + n], 0)
    cat = lambda a: np.concatenate([a, a], 0)
    return synth.Scene("two_ropes", x, v, springs, cat(half.rest), cat(half.log_Y), cat(half.mass), dict(half.params))


def _gripper_meshes(g):
    half, fh = len(g.verts) // 2, len(g.faces) // 2
    return [(g.verts[:half], g.faces[:fh]), (g.verts[half:], g.faces[fh:] - half)]


def _case(scene, n_substeps, frames=None, meshes=None, use_pusher=False, reset=None, over=None, n_frames=1):
    return dict(scene=scene, n_substeps=n_substeps, frames=frames if frames is not None else [None] * n_frames,
                meshes=meshes, use_pusher=use_pusher, reset=reset, over=over or {})


def rope_s10():
    """BASELINE configs[1] physics: rope-synth, 10 substeps per frame, two frames."""
    return _case(synth.make_rope(v_scale=0.05), 10, n_frames=2)


def tblock_s100():
    """The shipped T-block graph, 100 substeps."""
    return _case(synth.load_tblock(v_scale=0.05), 100)


def two_ropes_collide():
    """Self-collision: resting pairs from a far-apart pose, then the upper copy falls onto the lower one."""
    close = _two_ropes(0.0125, -0.8)
    return _case(_two_ropes(0.1, 0.0), 20, reset=(close.x, close.v))


def chain_ground():
    """Free fall onto the ground with time of impact; no self-collision."""
    sc = synth.make_chain(n=8, z=0.0008)
    sc.v[:] = [0.0, 0.0, -0.5]
    return _case(sc, 40, over=dict(self_collision=False))


def chain_reverse_z():
    """reverse_z (gravity and ground flipped), non-default restitution / friction, per-spring stiffness partly
    above spring_Y_max and partly below spring_Y_min."""
    sc = synth.make_chain(n=16, z=-0.0006)
    sc.v[:] = [0.1, 0.0, 0.4]
    sc.log_Y = np.log(np.linspace(5e3, 2e5, sc.S).astype(np.float32))
    return _case(sc, 30, over=dict(self_collision=False, reverse_z=True, collide_elas=0.8, collide_fric=0.1,
                                   spring_Y_min=1e4))


def _gripper(gap, ns=30, n_frames=1, vz=-0.3):
    sc = synth.make_rope(v_scale=0.0)
    g = synth.make_gripper(center=(0.5, 0.0, 0.004), gap=gap)
    frames = []
    gg = g
    for _ in range(n_frames):
        t = synth.gripper_motion(gg, ns, DT, eef_vel=(0.0, 0.0, vz), close_speed=0.6, omega=(0.0, 0.0, 0.4))
        frames.append(t)
        gg = synth.Gripper(t[0][-1].copy(), g.faces, g.mesh_map, g.face_map)   # next frame starts where this ended
    return _case(sc, ns, frames=frames, meshes=dict(dynamic=_gripper_meshes(g), static=[]))


def gripper_graze():
    """Fingers (gap 22 mm) close on the rope from outside while descending and rotating; two frames."""
    return _gripper(0.022, n_frames=2)


def gripper_inside():
    """Gap 8 mm: rope particles start INSIDE the finger volume (sign -1 branch)."""
    return _gripper(0.008)


def static_and_gripper():
    """Two dynamic fingers + one static obstacle (mesh_map -1, margin 1 mm, projection without re-query)."""
    sc = synth.make_rope(v_scale=0.0)
    sc.v[:, 2] = -0.6
    ns = 40
    g = synth.make_gripper(center=(0.5, 0.0, 0.004), gap=0.022)
    bv, bf = synth.make_finger_mesh(length=0.06, half_w=0.02, half_t=0.0045)
    box = (bv[:, [2, 0, 1]] + np.array([0.17, 0.0, 0.0045], np.float32)).astype(np.float32)
    t = synth.gripper_motion(g, ns, DT, eef_vel=(0.0, 0.0, -0.2), close_speed=0.5)
    return _case(sc, ns, frames=[t], meshes=dict(dynamic=_gripper_meshes(g), static=[(box, bf)]))


def pusher_tblock():
    """use_pusher=True: an 816-triangle rod pushes the real T-block (eef friction 0.2, phystwin.py:305-306)."""
    ns = 40
    sc = synth.load_tblock()
    g = synth.make_pusher(center=(0.32 - 0.0375 + 0.0006, 0.0, 0.004))
    t = synth.rigid_motion_tables(g, ns, DT, vel=(0.4, 0.02, 0.0), omega=(0.0, 0.0, 1.5))
    return _case(sc, ns, frames=[t], meshes=dict(dynamic=[(g.verts, g.faces)], static=[]), use_pusher=True,
                 over=dict(collide_eef_fric=0.2))


def pusher_static_tblock():
    """The rod pushes the T-block against a STATIC obstacle standing at its far side (mesh_map -1 after the tool's 0):
    tool and static faces in one merged mesh, as the reference keeps them in one BVH (SMW:636-676)."""
    c = pusher_tblock()
    sc = c["scene"]
    wv, wf = synth.make_finger_mesh(length=0.06, half_w=0.02, half_t=0.03)
    wall = (wv + np.array([float(sc.x[:, 0].max()) + 0.02 - 0.0005, 0.0, 0.0], np.float32)).astype(np.float32)
    c["meshes"] = dict(dynamic=c["meshes"]["dynamic"], static=[(wall, wf)])
    return c


CASES = dict(pusher_static_tblock=pusher_static_tblock, rope_s10=rope_s10, tblock_s100=tblock_s100, two_ropes_collide=two_ropes_collide,
             chain_ground=chain_ground, chain_reverse_z=chain_reverse_z, gripper_graze=gripper_graze,
             gripper_inside=gripper_inside, static_and_gripper=static_and_gripper, pusher_tblock=pusher_tblock)


def params_of(case):
    p = dict(case["scene"].params)
    p.update(case["over"])
    return p


def merged_mesh(case):
    """The merged mesh arrays in the order of spring_mass_warp.py:626-689 (dynamic first, then static with
    mesh_map -1, -2, ...; face_map = arange), as the dict the oracle / BatchedSpringMass.set_mesh take."""
    m = case["meshes"]
    if m is None:
        return None
    verts, faces, mesh_map, off = [], [], [], 0
    for k, (v, f) in enumerate(m["dynamic"]):
        verts.append(v); faces.append(f + off); mesh_map.append(np.full(len(f), k)); off += len(v)
    n_dyn = off
    for k, (v, f) in enumerate(m["static"]):
        verts.append(v); faces.append(f + off); mesh_map.append(np.full(len(f), -1 - k)); off += len(v)
    faces = np.concatenate(faces).astype(np.int32)
    return dict(verts=np.concatenate(verts).astype(np.float32), faces=faces,
                mesh_map=np.concatenate(mesh_map).astype(np.int32), face_map=np.arange(len(faces), dtype=np.int32),
                n_dyn_verts=n_dyn)


def input_digest(case):
    """sha256 over every input array of the case: the golden file records it, so a drift of the seeded
    generators (numpy / scipy version) fails loudly instead of comparing different scenes."""
    h = hashlib.sha256()
    sc = case["scene"]
    arrs = [sc.x, sc.v, sc.springs, sc.rest, sc.log_Y, sc.mass]
    if case["reset"] is not None:
        arrs += list(case["reset"])
    m = merged_mesh(case)
    if m is not None:
        arrs += [m["verts"], m["faces"], m["mesh_map"]]
    for f in case["frames"]:
        if f is not None:
            arrs += list(f)
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(repr(sorted(params_of(case).items())).encode())
    return h.hexdigest()


# ---------------------------------------------------------------------------- the reference's call sequence
def build_sim(case, cls, device="cpu", use_graph=True, **extra):
    """Construct `cls` -- the reference's SpringMassSystemWarp (under oracle/warp_exec.py) or this repo's drop-in
    of the same name -- on a case, with the constructor call of sim/physics/phystwin.py:336-357."""
    import torch
    from types import SimpleNamespace as NS
    sc, p = case["scene"], params_of(case)
    cfg = NS(dt=p["dt"], num_substeps=case["n_substeps"], init_spring_Y=3e4, dashpot_damping=p["dashpot_damping"],
             drag_damping=p["drag_damping"], collision_dist=p["collision_dist"], reverse_z=p["reverse_z"],
             spring_Y_min=p["spring_Y_min"], spring_Y_max=p["spring_Y_max"], self_collision=p["self_collision"],
             use_graph=use_graph, collide_elas=p["collide_elas"], collide_fric=p["collide_fric"],
             collide_eef_elas=p["collide_eef_elas"], collide_eef_fric=p["collide_eef_fric"],
             collide_self_elas=p["collide_self_elas"], collide_self_fric=p["collide_self_fric"],
             collision_requires_grad=False)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device=device)
    one = lambda k: torch.tensor([p[k]], dtype=torch.float32, device=device)
    kw = {}
    m = case["meshes"]
    if m is not None:
        mk = lambda vf: NS(vertices=vf[0], triangles=vf[1])
        kw = dict(dynamic_meshes=[mk(vf) for vf in m["dynamic"]], static_meshes=[mk(vf) for vf in m["static"]],
                  dynamic_points=t(np.concatenate([vf[0] for vf in m["dynamic"]], 0).astype(np.float32)),
                  use_pusher=case["use_pusher"])
    sim = cls(cfg, device, t(sc.x), t(sc.springs), t(sc.rest), t(sc.mass), sc.N, init_spring_Y=t(sc.log_Y),
              collide_elas=one("collide_elas"), collide_fric=one("collide_fric"),
              collide_eef_elas=one("collide_eef_elas"), collide_eef_fric=one("collide_eef_fric"),
              collide_self_elas=one("collide_self_elas"), collide_self_fric=one("collide_self_fric"),
              init_collision_mask=None, init_velocities=t(sc.v), **kw, **extra)
    if case["reset"] is not None:
        sim.set_init_state(t(case["reset"][0]), t(case["reset"][1]))
    return sim


def drive_frame(sim, case, k, wp, device="cpu", use_graph=True):
    """One frame as SpringMassDynamicsModule.step drives the simulator (phystwin.py:365-366, 455-460, 515-519);
    `wp` is the warp module in use (oracle.warp_exec for the reference, compat.warp for the drop-in)."""
    import torch
    if params_of(case)["self_collision"]:
        sim.update_collision_graph()
    tables = case["frames"][k]
    if tables is not None:
        sim.set_mesh_interactive(*[torch.tensor(np.ascontiguousarray(a), device=device) for a in tables])
    if use_graph:
        assert sim.graph is not None
        wp.capture_launch(sim.graph)
    else:
        sim.step()
