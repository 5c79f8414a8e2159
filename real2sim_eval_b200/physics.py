"""Host side of the PhysTwin spring-mass substep loop: the reference's Python
surface over the C ABI of include/r2s_phys.h.

`SpringMassSystemWarp` keeps the constructor, methods and attributes of
sim/physics/spring_mass_warp.py:477-995 that sim/physics/phystwin.py touches
(:336-357 ctor, :366 update_collision_graph, :383-386 mesh_map / collision_forces,
:455-460 set_mesh_interactive, :515-519 graph / step, :523-531 wp_state), with torch
tensors where Warp arrays were.  `BatchedSpringMass` is the B200-first entry:
E environments sharing one spring topology stepped by one persistent launch.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib

_NAN = float("nan")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _scalar(v, default=None) -> float:
    if v is None:
        return default
    if isinstance(v, torch.Tensor):
        return float(v.detach().reshape(-1)[0].item())
    return float(np.asarray(v).reshape(-1)[0])


class DeviceArray:
    """A torch tensor that also answers `.numpy()` like a Warp array does
    (phystwin.py:383,386 call `.numpy()` on device arrays)."""

    def __init__(self, t: torch.Tensor):
        self.t = t

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def torch(self):
        return self.t

    @property
    def shape(self):
        return tuple(self.t.shape)

    def __len__(self):
        return self.t.shape[0]


class _FromPtr:
    """__cuda_array_interface__ holder for zero-copy torch views of library buffers."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3)


def _view(ptr, shape, dtype, device):
    typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_FromPtr(ptr, shape, typestr), device=device)


class BatchedSpringMass:
    """E environments of one PhysTwin object (shared springs / stiffness / masses,
    per-environment state and, optionally, rest lengths and gripper motion)."""

    def __init__(self, E, springs, rest_lengths, *, num_particles, n_substeps, log_spring_Y=None, masses=None,
                 collision_mask=None, dt=5e-5, dashpot_damping=100.0, drag_damping=3.0, spring_Y_min=0.0,
                 spring_Y_max=1e5, collision_dist=0.005, self_collision=True, reverse_z=False,
                 collide_elas=0.5, collide_fric=0.3, collide_eef_elas=0.0, collide_eef_fric=1.0,
                 collide_self_elas=0.5, collide_self_fric=0.3, use_pusher=False, sign_mode=0, coll_row_cap=0,
                 threads=0, precise=False, mesh_accel=0, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.R2SError("BatchedSpringMass needs a CUDA device: there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        dev = self.device
        i32 = lambda t: None if t is None else torch.as_tensor(t, device=dev).to(torch.int32).contiguous()
        f32 = lambda t: None if t is None else torch.as_tensor(t, device=dev).to(torch.float32).contiguous()
        self.E, self.N, self.n_substeps = int(E), int(num_particles), int(n_substeps)
        springs = i32(springs).reshape(-1, 2)
        rest = f32(rest_lengths)
        self.S = springs.shape[0]
        rest_per_env = int(rest.dim() == 2)
        assert rest.shape[-1] == self.S and (not rest_per_env or rest.shape[0] == self.E)
        logY, masses, mask = f32(log_spring_Y), f32(masses), i32(collision_mask)
        d = _lib.PhysDesc()
        d.E, d.N, d.S, d.n_substeps = self.E, self.N, self.S, self.n_substeps
        d.self_collision, d.reverse_z, d.use_pusher = int(bool(self_collision)), int(bool(reverse_z)), int(bool(use_pusher))
        d.sign_mode, d.coll_row_cap, d.threads = int(sign_mode), int(coll_row_cap), int(threads)
        d.precise = int(bool(precise))
        d.mesh_accel = int(mesh_accel)
        d.dt, d.dashpot_damping, d.drag_damping = dt, dashpot_damping, drag_damping
        d.spring_Y_min, d.spring_Y_max, d.collision_dist = spring_Y_min, spring_Y_max, collision_dist
        d.collide_elas, d.collide_fric = collide_elas, collide_fric
        d.collide_eef_elas, d.collide_eef_fric = collide_eef_elas, collide_eef_fric
        d.collide_self_elas, d.collide_self_fric = collide_self_elas, collide_self_fric
        d.springs, d.rest_lengths, d.rest_per_env = _ptr(springs), _ptr(rest), rest_per_env
        d.log_spring_Y, d.masses, d.collision_mask = _ptr(logY), _ptr(masses), _ptr(mask)
        self.self_collision = bool(self_collision)
        self.use_pusher = bool(use_pusher)
        with torch.cuda.device(dev):
            torch.cuda.current_stream(dev).synchronize()
            self.h = self.lib.r2s_phys_create(C.byref(d))
        if not self.h:
            raise _lib.R2SError("r2s_phys_create failed: " + self.lib.r2s_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.F = 0
        self.n_dyn = 0
        self._refresh_views()

    def _refresh_views(self):
        p = _lib.PhysPtrs()
        _lib.check(self.lib.r2s_phys_get_ptrs(self.h, C.byref(p)), "r2s_phys_get_ptrs")
        dev = self.device
        self.x4 = _view(p.x4, (self.E, self.N, 4), torch.float32, dev)
        self.v4 = _view(p.v4, (self.E, self.N, 4), torch.float32, dev)
        self.status = _view(p.status, (self.E, 4), torch.int32, dev)
        self.coll_row_cap = p.coll_row_cap
        self.smem_state, self.smem_bytes = bool(p.smem_state), int(p.smem_bytes)
        self.collision_forces = _view(p.collision_forces, (self.E, p.F, 3), torch.float32, dev) if p.F else None
        self.mesh_map = _view(p.mesh_map, (p.F,), torch.int32, dev) if p.F else None
        if p.coll_num:
            self.coll_num = _view(p.coll_num, (self.E, self.N), torch.int32, dev)
            self.coll_idx = _view(p.coll_idx, (self.E, self.N, p.coll_row_cap), torch.int32, dev)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.r2s_phys_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- state
    @property
    def x(self) -> torch.Tensor:
        """(E,N,3) view of the float4 position storage (zero-copy, strided)."""
        return self.x4[..., :3]

    @property
    def v(self) -> torch.Tensor:
        return self.v4[..., :3]

    def set_state(self, x, v=None):
        dev = self.device
        x = torch.as_tensor(x, device=dev).to(torch.float32).contiguous()
        v = None if v is None else torch.as_tensor(v, device=dev).to(torch.float32).contiguous()
        broadcast = x.dim() == 2
        assert x.shape[-2:] == (self.N, 3) and (broadcast or x.shape[0] == self.E)
        stride = 0 if broadcast else self.N * 3
        with torch.cuda.device(dev):
            _lib.check(self.lib.r2s_phys_set_state(self.h, _ptr(x), _ptr(v), stride, _stream(dev)), "set_state")
        self._keep_state = (x, v)

    def get_state(self):
        dev = self.device
        x = torch.empty((self.E, self.N, 3), dtype=torch.float32, device=dev)
        v = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(self.lib.r2s_phys_get_state(self.h, _ptr(x), _ptr(v), _stream(dev)), "get_state")
        return x, v

    # ---- parameters
    def set_spring_Y(self, log_spring_Y):
        t = torch.as_tensor(log_spring_Y, device=self.device).to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_phys_set_spring_Y(self.h, _ptr(t), _stream(self.device)), "set_spring_Y")
        self._keep_Y = t

    def set_rest_lengths(self, rest):
        t = torch.as_tensor(rest, device=self.device).to(torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_phys_set_rest_lengths(self.h, _ptr(t), int(t.dim() == 2), _stream(self.device)),
                       "set_rest_lengths")
        self._keep_rest = t

    def set_collide(self, elas=None, fric=None, eef_elas=None, eef_fric=None, self_elas=None, self_fric=None):
        vals = [_scalar(v, _NAN) for v in (elas, fric, eef_elas, eef_fric, self_elas, self_fric)]
        _lib.check(self.lib.r2s_phys_set_collide(self.h, *vals), "set_collide")

    # ---- mesh
    def set_mesh(self, verts, faces, mesh_map, face_map, n_dyn_verts):
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        mesh_map = np.ascontiguousarray(mesh_map, dtype=np.int32)
        face_map = np.ascontiguousarray(face_map, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.r2s_phys_set_mesh(self.h, vp(verts), vp(faces), vp(mesh_map), vp(face_map), len(verts),
                                                  len(faces), int(n_dyn_verts)), "set_mesh")
        self.F, self.n_dyn = len(faces), int(n_dyn_verts)
        self._refresh_views()

    def set_mesh_motion(self, interp_pts, interp_center, dyn_vel, dyn_omega):
        dev = self.device
        f32 = lambda t: torch.as_tensor(t, device=dev).to(torch.float32).contiguous()
        interp_pts, interp_center, dyn_vel, dyn_omega = map(f32, (interp_pts, interp_center, dyn_vel, dyn_omega))
        per_env = int(interp_pts.dim() == 4)
        lead = (self.E,) if per_env else ()
        assert tuple(interp_pts.shape) == lead + (self.n_substeps, self.n_dyn, 3), interp_pts.shape
        interp_center = interp_center.reshape(lead + (self.n_substeps, 3))
        nv = 1 if self.use_pusher else 2
        dyn_vel = dyn_vel.reshape(lead + (nv, 3)).contiguous()
        dyn_omega = dyn_omega.reshape(lead + (1, 3)).contiguous()
        with torch.cuda.device(dev):
            _lib.check(self.lib.r2s_phys_set_mesh_motion(self.h, _ptr(interp_pts), _ptr(interp_center), _ptr(dyn_vel),
                                                         _ptr(dyn_omega), per_env, _stream(dev)), "set_mesh_motion")
        self._keep_motion = (interp_pts, interp_center, dyn_vel, dyn_omega)

    def motion_tables(self, per_env: bool = True):
        """Zero-copy views of the handle's own motion tables (r2s_phys_motion_ptrs): interp_pts
        [(E,) S, n_dyn, 3], interp_center [(E,) S, 3], dyn_vel [(E,) 2, 3], dyn_omega [(E,) 1, 3] -- what a
        device-side producer (r2s_eef_forward) fills in place of set_mesh_motion."""
        m = _lib.PhysMotion()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)   # switching per_env re-allocates
            _lib.check(self.lib.r2s_phys_motion_ptrs(self.h, int(per_env), C.byref(m)), "r2s_phys_motion_ptrs")
        lead = (m.n_env,) if per_env else ()
        dev = self.device
        return (_view(m.interp_pts, lead + (m.n_substeps, m.n_dyn_verts, 3), torch.float32, dev),
                _view(m.interp_center, lead + (m.n_substeps, 3), torch.float32, dev),
                _view(m.dyn_vel, lead + (m.dyn_vel_rows, 3), torch.float32, dev),
                _view(m.dyn_omega, lead + (1, 3), torch.float32, dev))

    # ---- collisions + stepping
    def create_resting_case(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_phys_create_resting_case(self.h, _stream(self.device)), "create_resting_case")

    def update_collision_graph(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_phys_update_collision_graph(self.h, _stream(self.device)),
                       "update_collision_graph")

    def step(self, n_substeps: int = 0):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_phys_step(self.h, int(n_substeps), _stream(self.device)), "r2s_phys_step")

    def algorithmic_bytes_per_env_substep(self) -> int:
        return int(self.lib.r2s_phys_algorithmic_bytes(self.h))


class _GraphToken:
    def __init__(self, owner):
        self.owner = owner


class _State:
    """wp_state stand-in: `.wp_x` / `.wp_v` are (N,3) torch views (PT:523-531)."""

    def __init__(self, sys: BatchedSpringMass):
        self._sys = sys

    @property
    def wp_x(self):
        return self._sys.x[0]

    @property
    def wp_v(self):
        return self._sys.v[0]


class SpringMassSystemWarp:
    """Drop-in for sim/physics/spring_mass_warp.py:477 (one environment).

    Constructor arguments, method names and the attributes read by
    sim/physics/phystwin.py are the reference's; `device` may be a torch device
    or the reference's Warp device string ('cuda:0')."""

    def __init__(self, phystwin_cfg, device, init_vertices, init_springs, init_rest_lengths, init_masses,
                 num_object_points, init_spring_Y=None, collide_elas=None, collide_fric=None, collide_eef_elas=None,
                 collide_eef_fric=None, collide_self_elas=None, collide_self_fric=None, init_collision_mask=None,
                 init_velocities=None, dynamic_meshes=None, static_meshes=None, dynamic_points=None,
                 use_pusher=False, sign_mode=0, precise=True):
        cfg = phystwin_cfg
        self.device = torch.device(str(device))
        self.dt, self.num_substeps = cfg.dt, int(cfg.num_substeps)
        self.dashpot_damping, self.drag_damping = cfg.dashpot_damping, cfg.drag_damping
        self.reverse_factor = 1.0 if not cfg.reverse_z else -1.0
        self.spring_Y_min, self.spring_Y_max = cfg.spring_Y_min, cfg.spring_Y_max
        self.self_collision, self.use_pusher = bool(cfg.self_collision), bool(use_pusher)
        self.collision_dist = cfg.collision_dist
        self.n_springs = init_springs.shape[0]
        self.num_object_points = num_object_points
        assert num_object_points == init_vertices.shape[0]
        if self.self_collision and init_collision_mask is not None:
            assert torch.unique(init_collision_mask).shape[0] > 1
            init_collision_mask = init_collision_mask[:num_object_points]
        g = lambda v, name: _scalar(v, float(getattr(cfg, name)))
        logY = init_spring_Y if init_spring_Y is not None else \
            torch.full((self.n_springs,), math.log(float(cfg.init_spring_Y)), dtype=torch.float32)
        self.sys = BatchedSpringMass(
            1, init_springs, init_rest_lengths, num_particles=num_object_points, n_substeps=self.num_substeps,
            log_spring_Y=logY, masses=init_masses[:num_object_points], collision_mask=init_collision_mask,
            dt=cfg.dt, dashpot_damping=cfg.dashpot_damping, drag_damping=cfg.drag_damping,
            spring_Y_min=cfg.spring_Y_min, spring_Y_max=cfg.spring_Y_max, collision_dist=cfg.collision_dist,
            self_collision=cfg.self_collision, reverse_z=cfg.reverse_z,
            collide_elas=g(collide_elas, "collide_elas"), collide_fric=g(collide_fric, "collide_fric"),
            collide_eef_elas=g(collide_eef_elas, "collide_eef_elas"),
            collide_eef_fric=g(collide_eef_fric, "collide_eef_fric"),
            collide_self_elas=g(collide_self_elas, "collide_self_elas"),
            collide_self_fric=g(collide_self_fric, "collide_self_fric"),
            use_pusher=use_pusher, sign_mode=sign_mode, precise=precise, coll_row_cap=500,   # SMW:544-549
            device=self.device)
        self.wp_state = _State(self.sys)
        v0 = None if init_velocities is None else init_velocities[:num_object_points]
        self.set_init_state(init_vertices, v0)

        # merged mesh, dynamic first (SMW:626-712)
        self.all_meshes_warp = None
        if static_meshes is not None or dynamic_meshes is not None:
            vertices, indices, mesh_map, face_map = [], [], [], []
            mesh_index = face_index = offset = 0
            for group, sign in ((dynamic_meshes or [], +1), (static_meshes or [], -1)):
                if sign < 0:
                    mesh_index = -1
                for mesh in group:
                    vertex = np.array(mesh.vertices, dtype=np.float32)
                    index = np.array(mesh.triangles, dtype=np.int32).reshape(-1, 3)
                    vertices.append(vertex)
                    indices.append(index + offset)
                    offset += vertex.shape[0]
                    mesh_map.append(np.full(len(index), mesh_index, dtype=np.int32))
                    face_map.append(np.arange(len(index), dtype=np.int32) + face_index)
                    mesh_index += sign
                    face_index += len(index)
            assert isinstance(dynamic_points, torch.Tensor)
            n_dyn_meshes = len(dynamic_meshes or [])
            self.num_eefs = n_dyn_meshes // 2 if not self.use_pusher else n_dyn_meshes
            assert self.num_eefs <= 1
            self.num_dynamic_points = len(dynamic_points)
            self._face_map = np.concatenate(face_map, 0)
            self.sys.set_mesh(np.concatenate(vertices, 0), np.concatenate(indices, 0), np.concatenate(mesh_map, 0),
                              self._face_map, self.num_dynamic_points)
            self.all_meshes_warp = self.sys  # truthiness only: "a mesh exists"
            self.num_dynamic_velocities = self.num_eefs * 2 if not self.use_pusher else self.num_eefs
        if self.self_collision:
            self.create_resting_case()
        # PT:515-517 replays `simulator.graph` through wp.capture_launch; the token points back here
        self.graph = _GraphToken(self) if getattr(cfg, "use_graph", True) else None

    # attributes phystwin.py reads through `.numpy()`
    @property
    def mesh_map(self):
        return DeviceArray(self.sys.mesh_map)

    @property
    def collision_forces(self):
        return DeviceArray(self.sys.collision_forces[0])

    @property
    def face_map(self):
        return DeviceArray(torch.as_tensor(self._face_map))

    @property
    def wp_collision_number(self):          # SMW:550-552
        return DeviceArray(self.sys.coll_num[0])

    @property
    def wp_collision_indices(self):         # SMW:544-549, (N, 500)
        return DeviceArray(self.sys.coll_idx[0])

    @property
    def collision_row_overflow(self) -> int:
        """Candidates dropped because a row was full (the reference would write out of bounds); blocking read."""
        return int(self.sys.status[0, 1])

    def create_resting_case(self):
        self.sys.create_resting_case()

    def set_init_state(self, x, v=None):
        x = torch.as_tensor(x).reshape(1, self.num_object_points, 3)
        v = None if v is None else torch.as_tensor(v).reshape(1, self.num_object_points, 3)
        self.sys.set_state(x, v)

    def set_mesh_interactive(self, interpolated_dynamic_points, interpolated_center, dynamic_velocity, dynamic_omega):
        self.sys.set_mesh_motion(interpolated_dynamic_points, interpolated_center.reshape(-1, 3), dynamic_velocity,
                                 dynamic_omega)

    def update_collision_graph(self):
        assert self.self_collision
        self.sys.update_collision_graph()

    def step(self):
        self.sys.step()

    def set_spring_Y(self, spring_Y):
        self.sys.set_spring_Y(spring_Y)

    def set_collide(self, collide_elas, collide_fric):
        self.sys.set_collide(elas=collide_elas, fric=collide_fric)

    def set_collide_eef(self, collide_eef_elas, collide_eef_fric):
        self.sys.set_collide(eef_elas=collide_eef_elas, eef_fric=collide_eef_fric)

    def set_collide_self(self, collide_self_elas, collide_self_fric):
        self.sys.set_collide(self_elas=collide_self_elas, self_fric=collide_self_fric)
