"""Host side of the per-frame end-effector step over include/r2s_eef.h.

`BatchedEefMotion` holds what SpringMassDynamicsModule keeps between frames (`current_openness`, `grasped`,
sim/physics/phystwin.py:358-359) for E environments on the device and, each frame, turns the end-effector
command into the per-substep collision-mesh tables -- the arithmetic of SpringMassDynamicsModule.step
(phystwin.py:362-510) up to `set_mesh_interactive`.  Bound to a BatchedSpringMass it writes the tables straight
into the physics handle's buffers (r2s_phys_motion_ptrs) and reads the finger forces of the previous frame
from the handle's `collision_forces`: no host round trip, no table upload.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def force_faces(mesh_map) -> list[int]:
    """Rows of collision_forces the grasp hysteresis sums (phystwin.py:384-391): entries [18], [19], [1] of
    the faces with mesh_map == 0 (left finger), then of mesh_map == 1 (right finger)."""
    mesh_map = np.asarray(mesh_map)
    out = []
    for k in (0, 1):
        rows = np.nonzero(mesh_map == k)[0]
        if len(rows) < 20:
            raise ValueError(f"finger {k} has {len(rows)} faces; the grasp test reads its faces 18, 19 and 1")
        out += [int(rows[18]), int(rows[19]), int(rows[1])]
    return out


class BatchedEefMotion:
    """E end effectors of one kind (a two-finger gripper, or a pusher).

    table:         (n_table, V, 3) float32 -- vertices of the dynamic collision mesh at opening k/(n_table-1),
                   the samples behind `eef_pts_func` (robot_pc_transformations.py:183-190); for a pusher the
                   rows are identical (:222-225).
    init_eef_xyz:  (3,) the end-effector position the table was sampled at.
    mesh_map:      (F,) of the merged collision mesh (gripper only: selects the force rows).
    phys:          optional BatchedSpringMass with the mesh set: outputs go into its motion tables in place and
                   the forces come from its collision_forces.
    """

    def __init__(self, E, table, init_eef_xyz, *, dt, n_substeps, grasp_force_threshold=3e4, use_pusher=False,
                 mesh_map=None, phys=None, device="cuda"):
        self.device = dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.R2SError("BatchedEefMotion needs a CUDA device: there is no CPU path")
        if dev.index is None:
            self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.E, self.S, self.dt = int(E), int(n_substeps), float(dt)
        self.use_pusher, self.threshold = bool(use_pusher), float(grasp_force_threshold)
        self.table = torch.as_tensor(np.asarray(table, np.float32)).to(dev).contiguous()
        self.n_table, self.V = int(self.table.shape[0]), int(self.table.shape[1])
        self.init_eef_xyz = torch.as_tensor(np.asarray(init_eef_xyz, np.float32).reshape(3)).to(dev)
        self.current_openness = torch.full((self.E,), float("nan"), dtype=torch.float64, device=dev)  # None
        self.grasped = torch.zeros((self.E,), dtype=torch.int32, device=dev)
        self.faces = [0] * 6 if use_pusher else force_faces(mesh_map)
        self.phys, self._forces, self.F = phys, None, 0
        if phys is not None:
            if phys.n_substeps != self.S:
                raise ValueError("the physics handle and the end-effector step disagree on n_substeps")
            self._bind()
        else:
            self.rows = 1 if use_pusher else 2
            self.interp_pts = torch.empty((self.E, self.S, self.V, 3), dtype=torch.float32, device=dev)
            self.interp_center = torch.empty((self.E, self.S, 3), dtype=torch.float32, device=dev)
            self.dyn_vel = torch.zeros((self.E, self.rows, 3), dtype=torch.float32, device=dev)
            self.dyn_omega = torch.empty((self.E, 1, 3), dtype=torch.float32, device=dev)
            self._motion = tuple(t.data_ptr() for t in (self.interp_pts, self.interp_center, self.dyn_vel, self.dyn_omega))

    def _bind(self):
        """(Re-)query the bound physics handle's motion tables and force array.  The handle re-allocates them when
        its mesh is replaced (r2s_phys_set_mesh) or when shared tables are installed (set_mesh_motion with
        [S,V,3] tables flips the per-env mode), so cached pointers would dangle: forward() asks again every frame
        (no re-allocation and no synchronisation happens while the mode stays per-env)."""
        phys = self.phys
        m = _lib.PhysMotion()
        _lib.check(self.lib.r2s_phys_motion_ptrs(phys.h, 1, C.byref(m)), "r2s_phys_motion_ptrs")
        if m.n_dyn_verts != self.V or m.n_env != self.E:
            raise ValueError(f"physics mesh has {m.n_dyn_verts} dynamic vertices x {m.n_env} envs, "
                             f"table has {self.V} x {self.E}")
        self._motion = (m.interp_pts, m.interp_center, m.dyn_vel, m.dyn_omega)
        self.rows = int(m.dyn_vel_rows)
        p = _lib.PhysPtrs()
        _lib.check(self.lib.r2s_phys_get_ptrs(phys.h, C.byref(p)), "r2s_phys_get_ptrs")
        if self._forces is None or self._forces.data_ptr() != p.collision_forces:
            phys._refresh_views()
            self._forces = phys.collision_forces
        self.F = int(self._forces.shape[1])

    def reset(self):
        """SpringMassDynamicsModule.__init__ state (phystwin.py:358-359)."""
        self.current_openness.fill_(float("nan"))
        self.grasped.zero_()

    def forward(self, eef_xyz, eef_vel, eef_rot, eef_rot_vel, gripper_openness=None, collision_forces=None):
        """eef_xyz/eef_vel/eef_rot_vel: [E,3]; eef_rot: [E,3,3]; gripper_openness: [E] (gripper only);
        collision_forces: [E,F,3] of the previous frame (default: the bound physics handle's, else zeros).
        Device float32 tensors.  Enqueues one kernel on the current stream."""
        if self.phys is not None:
            self._bind()
        f = collision_forces if collision_forces is not None else self._forces
        for name, t, n in (("eef_xyz", eef_xyz, 3), ("eef_vel", eef_vel, 3), ("eef_rot", eef_rot, 9),
                           ("eef_rot_vel", eef_rot_vel, 3), ("gripper_openness", gripper_openness, 1)):
            if t is None and name == "gripper_openness" and self.use_pusher:
                continue
            if t is None or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.E * n or \
                    t.device != self.device:
                raise ValueError(f"{name}: expected a contiguous float32 tensor of {self.E}x{n} on {self.device}")
        a = _lib.EefArgs()
        a.E, a.n_substeps, a.n_pts, a.n_table, a.use_pusher = self.E, self.S, self.V, self.n_table, int(self.use_pusher)
        a.F = int(f.shape[1]) if f is not None else 0
        for k in range(6):
            a.force_faces[k] = self.faces[k]
        a.dyn_vel_rows, a.grasp_force_threshold, a.dt = self.rows, self.threshold, self.dt
        a.table, a.init_eef_xyz = _ptr(self.table), _ptr(self.init_eef_xyz)
        a.eef_xyz, a.eef_vel, a.eef_rot, a.eef_rot_vel = _ptr(eef_xyz), _ptr(eef_vel), _ptr(eef_rot), _ptr(eef_rot_vel)
        a.openness_cmd, a.collision_forces = _ptr(gripper_openness), _ptr(f)
        a.current_openness, a.grasped = _ptr(self.current_openness), _ptr(self.grasped)
        a.interp_pts, a.interp_center, a.dyn_vel, a.dyn_omega = [C.c_void_p(p) for p in self._motion]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_eef_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "r2s_eef_forward")
