"""oracle/eef_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy; float32 where the reference computes in float32 torch tensors, float64 where it
computes in Python floats / numpy float64) of the per-frame end-effector step:
  SpringMassDynamicsModule.step   sim/physics/phystwin.py:362-510   (up to set_mesh_interactive)
  eef_pts_func                    sim/utils/robot/robot_pc_transformations.py:183-192 (scipy interp1d, linear)

Pinning.  `load_reference()` imports the reference's OWN sim/physics/phystwin.py from /root/reference and
`reference_step()` runs its unmodified `SpringMassDynamicsModule.step` on CPU against a fake simulator that
records what `set_mesh_interactive` receives.  Module-scope imports that cannot load here are stubbed: warp
(the simulator is faked), open3d, the sibling modules spring_mass_warp / kinematics_utils / robot_pc_sampler
(unused by `step`), and kornia.  kornia is a dependency the reference neither vendors nor pins
(pyproject.toml:26) and it is not installed in this image: its
`geometry.conversions.axis_angle_to_rotation_matrix` is restated below from its published algorithm and that
restatement is what the stub hands to the reference.  So: hysteresis, interpolation, per-substep tables and
velocities are pinned to the reference's code (and scipy's interp1d, which IS installed, is called as the
reference calls it); the axis-angle -> matrix conversion alone is PARITY UNPINNED.
The reference runs this on CUDA, where torch evaluates `tensor / python_scalar` as a multiplication by the
reciprocal; on CPU (how the goldens are made) it divides.  The difference is one float32 ulp of a
sub-millimetre quantity and is covered by the 1e-6 m tolerance.
tests/golden/eef_*.npz were generated that way by tests/golden/make_eef_golden.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_FILE = "/root/reference/sim/physics/phystwin.py"
f32 = np.float32


# --------------------------------------------------------------------------- kornia restated
def axis_angle_to_rotation_matrix(aa):
    """kornia.geometry.conversions.axis_angle_to_rotation_matrix, (N,3) -> (N,3,3), float32:
    theta2 = aa.aa; if theta2 > 1e-6: Rodrigues with w = aa / (sqrt(theta2) + 1e-6); else I + [aa]x."""
    aa = np.asarray(aa, f32).reshape(-1, 3)
    theta2 = (aa[:, 0] * aa[:, 0] + aa[:, 1] * aa[:, 1] + aa[:, 2] * aa[:, 2]).astype(f32)
    theta = np.sqrt(theta2).astype(f32)
    w = (aa / (theta + f32(1e-6))[:, None]).astype(f32)
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s = np.cos(theta).astype(f32), np.sin(theta).astype(f32)
    k = (f32(1.0) - c).astype(f32)
    normal = np.stack([c + wx * wx * k, wx * wy * k - wz * s, wy * s + wx * wz * k,
                       wz * s + wx * wy * k, c + wy * wy * k, -wx * s + wy * wz * k,
                       -wy * s + wx * wz * k, wx * s + wy * wz * k, c + wz * wz * k], -1).astype(f32)
    one = np.ones_like(theta2)
    rx, ry, rz = aa[:, 0], aa[:, 1], aa[:, 2]
    taylor = np.stack([one, -rz, ry, rz, one, -rx, -ry, rx, one], -1).astype(f32)
    mask = (theta2 > f32(1e-6))[:, None]
    return np.where(mask, normal, taylor).reshape(-1, 3, 3).astype(f32)


# --------------------------------------------------------------------------- scipy interp1d restated
def interp_table(table, x):
    """scipy.interpolate.interp1d(np.arange(n)/(n-1), table, axis=0)(x) for a scalar x -- linear,
    `_call_linear` of scipy >= 1.15 (the reference's pin, pyproject.toml:15; 1.18.1 installed here):
    searchsorted-left index clipped to [1, n-1];
    y = ((x - x_lo)/(x_hi - x_lo)) * y_hi + ((x_hi - x)/(x_hi - x_lo)) * y_lo in float64."""
    table = np.asarray(table)
    n = table.shape[0]
    grid = np.arange(n) / float(n - 1)
    idx = int(np.clip(np.searchsorted(grid, x), 1, n - 1))
    lo, hi = idx - 1, idx
    w_hi = (x - grid[lo]) / (grid[hi] - grid[lo])
    w_lo = (grid[hi] - x) / (grid[hi] - grid[lo])
    return w_hi * table[hi].astype(np.float64) + w_lo * table[lo].astype(np.float64)


def make_eef_pts_func(table):
    """The callable the reference builds (robot_pc_transformations.py:190): scipy's own interp1d."""
    import scipy.interpolate
    table = np.asarray(table, f32)
    return scipy.interpolate.interp1d(np.arange(table.shape[0]) / float(table.shape[0] - 1), table, axis=0)


# --------------------------------------------------------------------------- the restatement
def force_faces(mesh_map):
    """Rows of collision_forces the hysteresis reads (phystwin.py:384-391): the [18], [19], [1] entries of
    the faces with mesh_map == 0 (left finger) and == 1 (right finger)."""
    mesh_map = np.asarray(mesh_map)
    out = []
    for k in (0, 1):
        rows = np.nonzero(mesh_map == k)[0]
        out += [int(rows[18]), int(rows[19]), int(rows[1])]
    return out


def hysteresis(openness_cmd, current_openness, grasped, forces, faces, threshold):
    """phystwin.py:370-412 for one environment.  Returns (openness, openness_before, current_openness', grasped')."""
    openness = float(f32(openness_cmd))
    cur = openness if current_openness is None or np.isnan(current_openness) else float(current_openness)
    forces = np.asarray(forces, f32)
    filt = np.stack([forces[faces[0]] + forces[faces[1]] + forces[faces[2]],
                     forces[faces[3]] + forces[faces[4]] + forces[faces[5]]], 0)
    nrm = np.linalg.norm(filt, axis=1)
    before = cur
    if np.all(nrm < 100):
        grasped = False
    if openness < cur:
        if np.all(nrm > threshold):
            openness = cur
            grasped = True
        elif grasped:
            cur = max(openness, cur - 0.05)
            openness = cur
        else:
            cur = openness
    else:
        cur = openness
    return openness, before, cur, bool(grasped)


def eef_step(table, init_eef_xyz, eef_xyz, eef_vel, eef_rot, eef_rot_vel, openness_cmd, *, dt, n_substeps,
             current_openness=None, grasped=False, forces=None, faces=None, threshold=3e4, use_pusher=False):
    """One environment.  Returns dict(interp_pts (S,V,3), interp_center (S,3), dyn_vel (2|1,3), dyn_omega (1,3),
    current_openness, grasped)."""
    S = int(n_substeps)
    table = np.asarray(table, f32)
    xyz, vel = np.asarray(eef_xyz, f32).reshape(3), np.asarray(eef_vel, f32).reshape(3)
    rot, rvel = np.asarray(eef_rot, f32).reshape(3, 3), np.asarray(eef_rot_vel, f32).reshape(3)
    init = np.asarray(init_eef_xyz, f32).reshape(3)
    dts = (np.arange(1, S + 1, dtype=f32) * f32(dt)).astype(f32)
    xyz_next = (xyz[None] + (vel[None] * dts[:, None]).astype(f32)).astype(f32)                 # (S,3)
    D = axis_angle_to_rotation_matrix((rvel[None] * dts[:, None]).astype(f32))                  # (S,3,3)
    rot_next = np.einsum("ski,kj->sij", D, rot).astype(f32)                                     # D^T @ rot
    if use_pusher:
        cur, grasped = 1.0, grasped
        o_now = o_bef = 1.0
    else:
        if forces is None:
            forces, faces = np.zeros((1, 3), f32), [0] * 6
        o_now, o_bef, cur, grasped = hysteresis(openness_cmd, current_openness, grasped, forces, faces, threshold)
    pts = interp_table(table, float(np.clip(o_now, 0.0, 1.0))).astype(f32)
    bef = interp_table(table, float(np.clip(o_bef, 0.0, 1.0))).astype(f32)
    flip = np.array([1, -1, -1], f32)
    delta = ((pts - bef).astype(f32) * flip).astype(f32)
    rel = ((bef - init[None]).astype(f32) * flip).astype(f32)
    step = (delta / f32(dt * S)).astype(f32)
    rel_s = (rel[None] + (step[None] * dts[:, None, None]).astype(f32)).astype(f32)             # (S,V,3)
    interp = (xyz_next[:, None] + np.einsum("svj,sij->svi", rel_s, rot_next).astype(f32)).astype(f32)
    out = dict(interp_pts=interp, interp_center=xyz_next, current_openness=cur, grasped=grasped,
               dyn_omega=(-rvel * f32(0.5)).astype(f32)[None])
    half_v = (vel * f32(0.5)).astype(f32)
    if use_pusher:
        out["dyn_vel"] = half_v[None]
    else:
        cv = ((delta @ rot.T).astype(f32) / f32(2 * dt * S)).astype(f32)
        h = len(cv) // 2
        out["dyn_vel"] = (half_v[None] + np.stack([cv[:h].mean(0), cv[h:].mean(0)]).astype(f32)).astype(f32)
    return out


# --------------------------------------------------------------------------- the reference itself, on CPU
class _Arr:
    def __init__(self, a):
        self._a = np.asarray(a)

    def numpy(self):
        return self._a


class FakeSimulator:
    """What SpringMassDynamicsModule.step touches on `self.simulator` (phystwin.py:366, 383-386, 455-460, 519)."""

    def __init__(self, mesh_map, forces):
        self.mesh_map, self.collision_forces = _Arr(mesh_map), _Arr(forces)
        self.captured, self.graph = None, None

    def update_collision_graph(self):
        pass

    def set_mesh_interactive(self, pts, center, vel, omega):
        self.captured = [np.array(t.detach().cpu().numpy(), copy=True) for t in (pts, center, vel, omega)]

    def step(self):
        pass


def load_reference():
    """The reference's phystwin module (torch, CPU) with warp / kornia / open3d / sibling imports stubbed, or
    None when /root/reference is absent."""
    if not os.path.exists(REF_FILE):
        return None
    import torch

    def _aa2rm(aa):
        return torch.from_numpy(axis_angle_to_rotation_matrix(aa.detach().cpu().numpy())).to(aa.dtype)

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        return m

    kornia = pkg("kornia")
    kornia.geometry = pkg("kornia.geometry")
    kornia.geometry.conversions = types.ModuleType("kornia.geometry.conversions")
    kornia.geometry.conversions.axis_angle_to_rotation_matrix = _aa2rm
    stubs = {"kornia": kornia, "kornia.geometry": kornia.geometry,
             "kornia.geometry.conversions": kornia.geometry.conversions,
             "warp": types.ModuleType("warp"), "open3d": types.ModuleType("open3d"),
             "_r2s_ref_sim": pkg("_r2s_ref_sim"), "_r2s_ref_sim.physics": pkg("_r2s_ref_sim.physics"),
             "_r2s_ref_sim.utils": pkg("_r2s_ref_sim.utils"), "_r2s_ref_sim.utils.robot": pkg("_r2s_ref_sim.utils.robot"),
             "_r2s_ref_sim.physics.spring_mass_warp": types.ModuleType("_r2s_ref_sim.physics.spring_mass_warp"),
             "_r2s_ref_sim.utils.robot.kinematics_utils": types.ModuleType("_r2s_ref_sim.utils.robot.kinematics_utils"),
             "_r2s_ref_sim.utils.robot.robot_pc_sampler": types.ModuleType("_r2s_ref_sim.utils.robot.robot_pc_sampler")}
    stubs["_r2s_ref_sim.physics.spring_mass_warp"].SpringMassSystemWarp = object
    stubs["_r2s_ref_sim.utils.robot.kinematics_utils"].KinHelper = object
    stubs["_r2s_ref_sim.utils.robot.robot_pc_sampler"].RobotPcSampler = object
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_r2s_ref_sim.physics.phystwin", REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop("_r2s_ref_sim.physics.phystwin", None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class ReferenceModule:
    """A `self` for the reference's unbound SpringMassDynamicsModule.step: the attributes it reads and the two
    it keeps between frames (current_openness, grasped)."""

    def __init__(self, mod, *, dt, n_substeps, threshold, use_pusher, mesh_map, n_faces):
        self.mod = mod
        self.phystwin_cfg = types.SimpleNamespace(self_collision=False, num_substeps=int(n_substeps), dt=float(dt),
                                                  grasp_force_threshold=float(threshold), use_graph=False)
        self.use_pusher, self.device = bool(use_pusher), "cpu"
        self.current_openness, self.grasped = None, False
        self.simulator = FakeSimulator(mesh_map, np.zeros((n_faces, 3), f32))
        self.current_points = None

    def step(self, table_func, init_eef_xyz, eef_xyz, eef_vel, eef_rot, eef_rot_vel, openness_cmd, forces):
        import torch
        t = lambda a, shape: torch.tensor(np.asarray(a, f32).reshape(shape))
        self.simulator.collision_forces = _Arr(np.asarray(forces, f32))
        self.mod.SpringMassDynamicsModule.step(
            self, eef_xyz=t(eef_xyz, (1, 3)), eef_vel=t(eef_vel, (1, 3)), eef_rot=t(eef_rot, (1, 3, 3)),
            eef_rot_vel=t(eef_rot_vel, (1, 3)), gripper_openness=t(openness_cmd, (1, 1)), eef_pts_func=table_func,
            init_eef_xyz=t(init_eef_xyz, (3,)))
        pts, center, vel, omega = self.simulator.captured
        return dict(interp_pts=pts, interp_center=center, dyn_vel=vel, dyn_omega=omega,
                    current_openness=float(self.current_openness), grasped=bool(self.grasped))
