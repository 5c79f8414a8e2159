"""ctypes front-end of oracle/physics_ref.c (TEST INFRASTRUCTURE; pinned bit for bit to the reference's own
kernel source executed under oracle/warp_exec.py -- the three Warp built-ins stay restated, see the C file's header).  Mirrors the reference's
SpringMassSystemWarp surface (sim/physics/spring_mass_warp.py:477-995) closely
enough that tests read like calls into the reference."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib_path

_f = C.c_float
_i = C.c_int32
_pf = C.POINTER(C.c_float)
_pi = C.POINTER(C.c_int32)
_pu8 = C.POINTER(C.c_uint8)


class _Phys(C.Structure):
    _fields_ = [
        ("N", _i), ("S", _i), ("n_substeps", _i), ("self_collision", _i),
        ("dt", _f), ("dashpot_damping", _f), ("drag_damping", _f), ("reverse_factor", _f),
        ("spring_Y_min", _f), ("spring_Y_max", _f), ("collision_dist", _f),
        ("collide_elas", _f), ("collide_fric", _f), ("collide_eef_elas", _f), ("collide_eef_fric", _f),
        ("collide_self_elas", _f), ("collide_self_fric", _f),
        ("x", _pf), ("v", _pf), ("v_bc", _pf), ("v_bg", _pf), ("f", _pf),
        ("springs", _pi), ("rest", _pf), ("logY", _pf), ("mass", _pf), ("mask", _pi),
        ("coll_idx", _pi), ("coll_num", _pi), ("resting", _pu8),
        ("n_verts", _i), ("n_faces", _i), ("n_dyn_verts", _i), ("use_pusher", _i),
        ("mesh_pts", _pf), ("faces", _pi), ("mesh_map", _pi), ("face_map", _pi),
        ("collision_forces", _pf), ("interp_pts", _pf), ("interp_center", _pf),
        ("dyn_vel", _pf), ("dyn_omega", _pf), ("sign_mode", _i), ("pad_", _i),
    ]


class _Csr(C.Structure):
    _fields_ = [("row_ptr", _pi), ("nbr", _pi), ("sid", _pi)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(lib_path("libphysics_ref.so"))
        assert _lib.oracle_phys_struct_size() == C.sizeof(_Phys), "struct layout mismatch"
        _lib.oracle_phys_step.argtypes = [C.POINTER(_Phys), C.POINTER(_Csr)]
        _lib.oracle_phys_step_batch.argtypes = [C.POINTER(_Phys), C.c_int, C.POINTER(_Csr)]
        _lib.oracle_phys_update_collision_graph.argtypes = [C.POINTER(_Phys)]
        _lib.oracle_phys_create_resting_case.argtypes = [C.POINTER(_Phys)]
    return _lib


COLL_CAP = 500


def build_csr(n: int, springs: np.ndarray):
    """Directed adjacency in per-particle ascending spring-index order:
    row_ptr (N+1), nbr (2S), sid (2S).  This is the summation order of the CUDA
    spring-force gather."""
    springs = np.asarray(springs, dtype=np.int64)
    S = len(springs)
    owner = np.concatenate([springs[:, 0], springs[:, 1]])
    other = np.concatenate([springs[:, 1], springs[:, 0]])
    sid = np.concatenate([np.arange(S), np.arange(S)])
    order = np.lexsort((sid, owner))
    owner, other, sid = owner[order], other[order], sid[order]
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(row_ptr, owner + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    return row_ptr, other.astype(np.int32), sid.astype(np.int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else t()


class SpringMassOracle:
    """One environment.  Arrays are owned here as contiguous numpy buffers."""

    def __init__(self, x, v, springs, rest, log_Y, mass, *, n_substeps, dt=5e-5, dashpot_damping=100.0,
                 drag_damping=3.0, spring_Y_min=0.0, spring_Y_max=1e5, collision_dist=0.005,
                 self_collision=True, reverse_z=False, collide_elas=0.5, collide_fric=0.3,
                 collide_eef_elas=0.0, collide_eef_fric=1.0, collide_self_elas=0.5,
                 collide_self_fric=0.3, mask=None, mesh=None, use_pusher=False, sign_mode=0,
                 gather_order=True):
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).copy()
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32).copy()
        self.N, self.S = len(x), len(springs)
        self.x, self.v = f32(x), f32(v)
        self.v_bc, self.v_bg, self.f = np.zeros_like(self.x), np.zeros_like(self.x), np.zeros_like(self.x)
        self.springs, self.rest, self.log_Y, self.mass = i32(springs), f32(rest), f32(log_Y), f32(mass)
        self.mask = i32(mask if mask is not None else np.arange(self.N))
        self.n_substeps = int(n_substeps)
        self.self_collision = bool(self_collision)
        self.coll_idx = np.zeros((self.N, COLL_CAP), dtype=np.int32)
        self.coll_num = np.zeros(self.N, dtype=np.int32)
        self.resting = np.zeros((self.N, self.N), dtype=np.uint8) if self_collision else None
        s = self.s = _Phys()
        s.N, s.S, s.n_substeps, s.self_collision = self.N, self.S, self.n_substeps, int(self_collision)
        s.dt, s.dashpot_damping, s.drag_damping = dt, dashpot_damping, drag_damping
        s.reverse_factor = -1.0 if reverse_z else 1.0
        s.spring_Y_min, s.spring_Y_max, s.collision_dist = spring_Y_min, spring_Y_max, collision_dist
        s.collide_elas, s.collide_fric = collide_elas, collide_fric
        s.collide_eef_elas, s.collide_eef_fric = collide_eef_elas, collide_eef_fric
        s.collide_self_elas, s.collide_self_fric = collide_self_elas, collide_self_fric
        s.x, s.v, s.v_bc, s.v_bg, s.f = (_p(a, _pf) for a in (self.x, self.v, self.v_bc, self.v_bg, self.f))
        s.springs, s.rest, s.logY, s.mass, s.mask = (_p(self.springs, _pi), _p(self.rest, _pf),
                                                      _p(self.log_Y, _pf), _p(self.mass, _pf),
                                                      _p(self.mask, _pi))
        s.coll_idx, s.coll_num = _p(self.coll_idx, _pi), _p(self.coll_num, _pi)
        s.resting = _p(self.resting, _pu8)
        s.sign_mode, s.use_pusher = int(sign_mode), int(use_pusher)
        s.n_faces = 0
        if mesh is not None:
            self.set_mesh(**mesh)
        self.csr = None
        if gather_order:
            self._csr_arrays = build_csr(self.N, self.springs)
            self.csr = _Csr(*(_p(a, _pi) for a in self._csr_arrays))
        if self_collision:
            self.create_resting_case()

    # mesh = dict(verts (V,3), faces (F,3), mesh_map (F,), face_map (F,), n_dyn_verts)
    def set_mesh(self, verts, faces, mesh_map, face_map, n_dyn_verts):
        s = self.s
        self.mesh_pts = np.ascontiguousarray(verts, dtype=np.float32).copy()
        self.faces = np.ascontiguousarray(faces, dtype=np.int32).copy()
        self.mesh_map = np.ascontiguousarray(mesh_map, dtype=np.int32).copy()
        self.face_map = np.ascontiguousarray(face_map, dtype=np.int32).copy()
        self.collision_forces = np.zeros((len(faces), 3), dtype=np.float32)
        s.n_verts, s.n_faces, s.n_dyn_verts = len(verts), len(faces), int(n_dyn_verts)
        s.mesh_pts, s.faces = _p(self.mesh_pts, _pf), _p(self.faces, _pi)
        s.mesh_map, s.face_map = _p(self.mesh_map, _pi), _p(self.face_map, _pi)
        s.collision_forces = _p(self.collision_forces, _pf)
        # SMW:699-711 defaults: vertex table = rest pose repeated, zero velocities
        self.set_mesh_interactive(
            np.repeat(self.mesh_pts[None, :n_dyn_verts], self.n_substeps, 0),
            np.repeat(self.mesh_pts[:n_dyn_verts].mean(0)[None], self.n_substeps, 0),
            np.zeros((2, 3), np.float32), np.zeros((1, 3), np.float32))

    def set_mesh_interactive(self, interp_pts, interp_center, dyn_vel, dyn_omega):
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).copy()
        self.interp_pts, self.interp_center = f32(interp_pts), f32(interp_center)
        dv = f32(dyn_vel).reshape(-1, 3)
        if len(dv) < 2:
            dv = np.concatenate([dv, np.zeros((2 - len(dv), 3), np.float32)], 0)
        self.dyn_vel, self.dyn_omega = np.ascontiguousarray(dv), f32(dyn_omega)
        s = self.s
        s.interp_pts, s.interp_center = _p(self.interp_pts, _pf), _p(self.interp_center, _pf)
        s.dyn_vel, s.dyn_omega = _p(self.dyn_vel, _pf), _p(self.dyn_omega, _pf)

    def create_resting_case(self):
        lib().oracle_phys_create_resting_case(C.byref(self.s))

    def update_collision_graph(self):
        lib().oracle_phys_update_collision_graph(C.byref(self.s))

    def step(self):
        lib().oracle_phys_step(C.byref(self.s), C.byref(self.csr) if self.csr is not None else None)


def set_threads(n: int) -> int:
    """Use n OpenMP threads for step_batch (torchrun exports OMP_NUM_THREADS=1); returns the active maximum."""
    lib().oracle_set_threads(int(n))
    return int(lib().oracle_get_max_threads())


def step_batch(envs: list[SpringMassOracle]):
    """Step E environments with one OpenMP thread each (cpu baseline)."""
    arr = (_Phys * len(envs))(*[e.s for e in envs])
    csr = envs[0].csr
    lib().oracle_phys_step_batch(arr, len(envs), C.byref(csr) if csr is not None else None)
