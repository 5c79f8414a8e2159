"""Seeded synthetic workloads of the shapes BASELINE.json / SURVEY.md §8(d) name.

The reference ships no rope/sloth PhysTwin assets and no Gaussian scans (only
the T-block state in experiments/utils/T_final_state.pkl), so the bench and the
parity tests run on synthetic data of the same structure:

* spring graphs are built by the reference's own rule (sim/physics/phystwin.py:264-286:
  radius + max-neighbour KD search, first-come de-duplication, rest length > 1e-4,
  rest lengths recomputed in float32 from the posed cloud);
* the gripper is two closed 24-vertex / 44-triangle finger meshes (the size of the
  shipped left/right_finger_large_2.stl);
* cameras follow cfg/env/xarm_gripper.yaml:21-35 through the matrix conventions of
  sim/utils/gs/transform_utils.py:7-31.

Pure numpy/scipy host code; nothing here touches the GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")


# --------------------------------------------------------------------------
# spring graphs
# --------------------------------------------------------------------------
def build_springs(rest_pts: np.ndarray, posed_pts: np.ndarray, radius: float, max_nn: int):
    """Spring graph by the rule of sim/physics/phystwin.py:264-286.

    ``rest_pts`` is the un-posed cloud the KD tree is built on, ``posed_pts`` the
    aligned cloud rest lengths are measured on.  Returns (springs (S,2) int32,
    rest_lengths (S,) float32).
    """
    from scipy.spatial import cKDTree

    n = len(rest_pts)
    tree = cKDTree(rest_pts)
    dist, idx = tree.query(rest_pts, k=max_nn, distance_upper_bound=radius)
    posed64 = posed_pts.astype(np.float64)
    seen = set()
    springs = []
    for i in range(n):
        for j in idx[i][1:]:
            if j >= n:  # padding of the bounded query
                continue
            j = int(j)
            if (i, j) in seen or (j, i) in seen:
                continue
            if np.linalg.norm(posed64[i] - posed64[j]) > 1e-4:
                seen.add((i, j))
                springs.append((i, j))
    springs = np.asarray(springs, dtype=np.int32).reshape(-1, 2)
    p32 = posed_pts.astype(np.float32)
    d = p32[springs[:, 0]] - p32[springs[:, 1]]
    rest = np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)
    return springs, rest


def rest_lengths_f32(posed_pts: np.ndarray, springs: np.ndarray) -> np.ndarray:
    """float32 rest lengths of a posed cloud (phystwin.py:285-286)."""
    p32 = posed_pts.astype(np.float32)
    d = p32[springs[:, 0]] - p32[springs[:, 1]]
    return np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)


@dataclass
class Scene:
    """One PhysTwin-like object: particles + spring graph (+ defaults of cfg/physics/default.yaml)."""

    name: str
    x: np.ndarray          # (N,3) float32
    v: np.ndarray          # (N,3) float32
    springs: np.ndarray    # (S,2) int32
    rest: np.ndarray       # (S,) float32
    log_Y: np.ndarray      # (S,) float32  log stiffness (phystwin.py:344)
    mass: np.ndarray       # (N,) float32
    params: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.x.shape[0]

    @property
    def S(self):
        return self.springs.shape[0]


DEFAULT_PARAMS = dict(  # cfg/physics/default.yaml:7-38
    dt=5e-5, dashpot_damping=100.0, drag_damping=3.0, spring_Y_min=0.0, spring_Y_max=1e5,
    collision_dist=0.005, self_collision=True, reverse_z=False,
    collide_elas=0.5, collide_fric=0.3, collide_self_elas=0.5, collide_self_fric=0.3,
    collide_eef_elas=0.0, collide_eef_fric=1.0,
)


def _finish(name, pts, radius, max_nn, Y, seed, v_scale=0.0):
    pts = pts.astype(np.float32)
    springs, rest = build_springs(pts.astype(np.float64), pts.astype(np.float64), radius, max_nn)
    rng = np.random.default_rng(seed + 7919)
    v = (rng.uniform(-v_scale, v_scale, size=pts.shape)).astype(np.float32)
    return Scene(name, pts, v, springs, rest,
                 np.full(len(springs), np.log(np.float32(Y)), dtype=np.float32),
                 np.ones(len(pts), dtype=np.float32), dict(DEFAULT_PARAMS))


def make_rope(seed: int = 1234, n: int = 2048, v_scale: float = 0.0) -> Scene:
    """rope-synth: cylinder along x, length 1 m, radius 6 mm, axis at z = 12 mm."""
    rng = np.random.default_rng(seed)
    xs = rng.uniform(0.0, 1.0, n)
    r = 0.006 * np.sqrt(rng.uniform(0.0, 1.0, n))
    th = rng.uniform(0.0, 2 * np.pi, n)
    pts = np.stack([xs, r * np.cos(th), 0.012 + r * np.sin(th)], axis=1)
    return _finish("rope", pts, 0.02, 30, 3e4, seed, v_scale)


def make_sloth(seed: int = 1234, n: int = 3500, v_scale: float = 0.0) -> Scene:
    """sloth-synth: ellipsoid (0.10, 0.065, 0.125) m centred at z = 0.125."""
    rng = np.random.default_rng(seed)
    pts = np.empty((0, 3))
    while len(pts) < n:
        c = rng.uniform(-1.0, 1.0, (2 * n, 3))
        pts = np.concatenate([pts, c[(c * c).sum(1) <= 1.0]], axis=0)
    pts = pts[:n] * np.array([0.10, 0.065, 0.125]) + np.array([0.0, 0.0, 0.125])
    return _finish("sloth", pts, 0.02, 30, 3e4, seed, v_scale)


def load_tblock(v_scale: float = 0.0, seed: int = 1234) -> Scene:
    """The real T-block PhysTwin graph (N=2229, S=63100), from tests/golden/tblock.npz.

    The fixture is extracted from the reference's experiments/utils/T_final_state.pkl
    by tests/golden/make_tblock_fixture.py.
    """
    d = np.load(os.path.join(_GOLDEN, "tblock.npz"))
    x = d["x"].astype(np.float32)
    rng = np.random.default_rng(seed + 7919)
    v = rng.uniform(-v_scale, v_scale, size=x.shape).astype(np.float32)
    params = dict(DEFAULT_PARAMS)
    return Scene("tblock", x, v, d["springs"].astype(np.int32), d["rest"].astype(np.float32),
                 np.log(d["spring_Y"].astype(np.float32)), np.ones(len(x), dtype=np.float32), params)


def make_chain(n: int = 8, spacing: float = 0.01, z: float = 0.05, Y: float = 3e4) -> Scene:
    """Tiny 1-D chain for known-answer tests."""
    pts = np.stack([np.arange(n) * spacing, np.zeros(n), np.full(n, z)], axis=1).astype(np.float32)
    springs = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1).astype(np.int32)
    rest = rest_lengths_f32(pts, springs)
    return Scene("chain", pts, np.zeros_like(pts), springs, rest,
                 np.full(n - 1, np.log(np.float32(Y)), dtype=np.float32),
                 np.ones(n, dtype=np.float32), dict(DEFAULT_PARAMS))


def pose_transform(scene: Scene, seed: int):
    """Seeded rigid pose of an env's object: (R (3,3), centre (3,), shift (3,)); p -> (p - centre) R^T + centre + shift."""
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-0.3, 0.3)
    sh = rng.uniform(-0.02, 0.02, 2)
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return R, scene.x.astype(np.float64).mean(0), np.array([sh[0], sh[1], 0.0])


def pose_points(pts: np.ndarray, pose) -> np.ndarray:
    R, ctr, shift = pose
    return ((np.asarray(pts, np.float64) - ctr) @ R.T + ctr + shift).astype(np.float32)


def pose_scene(scene: Scene, seed: int) -> Scene:
    """Clone ``scene`` under a small seeded rigid pose (yaw + xy shift), as each
    parallel env randomises its object pose; rest lengths are recomputed in
    float32 from the posed cloud, so they differ per env (SURVEY §7 sizing note)."""
    R, ctr, shift = pose_transform(scene, seed)
    x = pose_points(scene.x, (R, ctr, shift))
    v = (scene.v.astype(np.float64) @ R.T).astype(np.float32)
    return Scene(scene.name, x, v, scene.springs, rest_lengths_f32(x, scene.springs), scene.log_Y,
                 scene.mass, dict(scene.params))


# --------------------------------------------------------------------------
# gripper meshes and per-substep motion tables
# --------------------------------------------------------------------------
def make_finger_mesh(length=0.045, half_w=0.011, half_t=0.004):
    """Closed 12-gon prism: 24 vertices / 44 triangles (outward-facing), the size of
    the shipped finger collision meshes.  Long axis = z, thin axis = y."""
    k = 12
    ang = 2 * np.pi * (np.arange(k) + 0.5) / k
    ring = np.stack([half_w * np.cos(ang), half_t * np.sin(ang)], axis=1)
    v = np.concatenate([np.concatenate([ring, np.zeros((k, 1))], 1),
                        np.concatenate([ring, np.full((k, 1), length)], 1)], 0)
    tris = []
    for i in range(k):
        j = (i + 1) % k
        tris += [(i, j, k + j), (i, k + j, k + i)]          # sides, outward
    for i in range(1, k - 1):
        tris.append((0, i + 1, i))                           # bottom cap, normal -z
        tris.append((k, k + i, k + i + 1))                   # top cap, normal +z
    return v.astype(np.float32), np.asarray(tris, dtype=np.int32)


def make_rod_mesh(radius=0.0375, length=0.205, n_circ=24, n_len=16):
    """Closed, outward-oriented triangulated cylinder along +z from z=0 (the shape of the shipped pusher,
    assets/robots/xarm/.../pusher_20cm.stl: a 75 mm x 205 mm rod; n_circ=112, n_len=112 gives ~25k triangles)."""
    ang = 2 * np.pi * np.arange(n_circ) / n_circ
    ring = np.stack([radius * np.cos(ang), radius * np.sin(ang)], 1)
    zs = np.linspace(0.0, length, n_len + 1)
    verts = [np.concatenate([ring, np.full((n_circ, 1), z)], 1) for z in zs]
    verts = np.concatenate(verts + [np.array([[0, 0, 0.0]]), np.array([[0, 0, length]])], 0)
    ib, it = (n_len + 1) * n_circ, (n_len + 1) * n_circ + 1
    tris = []
    for k in range(n_len):
        for i in range(n_circ):
            j = (i + 1) % n_circ
            a, b, c, d = k * n_circ + i, k * n_circ + j, (k + 1) * n_circ + j, (k + 1) * n_circ + i
            tris += [(a, b, c), (a, c, d)]
    for i in range(n_circ):
        j = (i + 1) % n_circ
        tris.append((ib, j, i))                                   # bottom cap, normal -z
        tris.append((it, n_len * n_circ + i, n_len * n_circ + j))  # top cap, normal +z
    return verts.astype(np.float32), np.asarray(tris, dtype=np.int32)


def make_pusher(center, tilt=0.0, **kw) -> "Gripper":
    """One rigid pusher rod hanging down with its tip at `center` (mesh_map 0 everywhere, use_pusher=True)."""
    v, f = make_rod_mesh(**kw)
    c, s = np.cos(tilt), np.sin(tilt)
    R = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    v = (v.astype(np.float64) @ R.T + np.asarray(center, np.float64)).astype(np.float32)
    return Gripper(v, f, np.zeros(len(f), np.int32), np.arange(len(f), dtype=np.int32))


def rigid_motion_tables(g: "Gripper", n_substeps, dt, vel, omega=(0.0, 0.0, 0.0)):
    """Per-substep tables of a RIGID tool (sim/physics/phystwin.py:462-510): rotation about the mesh centre with
    angular velocity `omega` plus translation `vel`; returns interp_pts, interp_center, dyn_vel (1,3), dyn_omega (1,3)."""
    vel, omega = np.asarray(vel, np.float64), np.asarray(omega, np.float64)
    ctr0 = g.verts.astype(np.float64).mean(0)
    pts, ctr = [], []
    for s in range(1, n_substeps + 1):
        t = s * dt
        th = np.linalg.norm(omega) * t
        if th > 0:
            k = omega / np.linalg.norm(omega)
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        else:
            R = np.eye(3)
        pts.append((g.verts.astype(np.float64) - ctr0) @ R.T + ctr0 + vel * t)
        ctr.append(ctr0 + vel * t)
    return (np.asarray(pts, np.float32), np.asarray(ctr, np.float32), (vel * 0.5)[None].astype(np.float32),
            (-omega * 0.5)[None].astype(np.float32))


@dataclass
class Gripper:
    verts: np.ndarray      # (48,3) float32 at rest pose (left finger first)
    faces: np.ndarray      # (88,3) int32 into verts
    mesh_map: np.ndarray   # (88,) 0 = left finger, 1 = right finger (SMW:643-645)
    face_map: np.ndarray   # (88,) arange (SMW:646-648)


def make_gripper(center, gap=0.03) -> Gripper:
    """Two fingers hanging down (tips at center z), separated by ``gap`` along y."""
    fv, ft = make_finger_mesh()
    c = np.asarray(center, dtype=np.float64)
    left = fv + np.array([c[0], c[1] - gap / 2, c[2]])
    right = fv + np.array([c[0], c[1] + gap / 2, c[2]])
    verts = np.concatenate([left, right], 0).astype(np.float32)
    faces = np.concatenate([ft, ft + len(fv)], 0).astype(np.int32)
    mesh_map = np.concatenate([np.zeros(len(ft)), np.ones(len(ft))]).astype(np.int32)
    return Gripper(verts, faces, mesh_map, np.arange(len(faces), dtype=np.int32))


EEF_ROT_DOWN = np.diag([1.0, -1.0, -1.0]).astype(np.float32)   # gripper frame = world with y, z flipped (phystwin.py:423-428)


def gripper_opening_table(center, gap_closed=0.008, gap_open=0.08, n=101) -> np.ndarray:
    """(n, 48, 3) float32: the finger vertices at opening k/(n-1), at the initial end-effector pose -- the
    samples `eef_pts_list` the reference builds by IK + FK at setup (robot_pc_transformations.py:183-189)
    and interpolates with scipy interp1d.  Opening 0 = closed (gap_closed), 1 = open (gap_open)."""
    return np.stack([make_gripper(center, gap_closed + (gap_open - gap_closed) * k / (n - 1)).verts
                     for k in range(n)]).astype(np.float32)


def gripper_motion(g: Gripper, n_substeps: int, dt: float, eef_vel, close_speed=0.0, omega=(0, 0, 0)):
    """Per-substep tables in the layout SpringMassSystemWarp.set_mesh_interactive takes
    (spring_mass_warp.py:769-804; produced by phystwin.py:374-460 in the reference):
    interp_pts (S,48,3), interp_center (S,3), dynamic_velocity (2,3), dynamic_omega (1,3).
    Linear translation at ``eef_vel`` plus symmetric closing along y."""
    eef_vel = np.asarray(eef_vel, dtype=np.float64)
    t = (np.arange(1, n_substeps + 1) * dt)[:, None, None]
    half = len(g.verts) // 2
    close = np.zeros((len(g.verts), 3))
    close[:half, 1] = +close_speed
    close[half:, 1] = -close_speed
    pts = g.verts[None].astype(np.float64) + (eef_vel[None, None] + close[None]) * t
    ctr0 = g.verts.astype(np.float64).mean(0)
    center = ctr0[None] + eef_vel[None] * t[:, 0]
    dyn_vel = np.stack([eef_vel * 0.5 + np.array([0, +close_speed / 2, 0]),
                        eef_vel * 0.5 + np.array([0, -close_speed / 2, 0])])
    dyn_omega = -np.asarray(omega, dtype=np.float64)[None] * 0.5
    return (pts.astype(np.float32), center.astype(np.float32), dyn_vel.astype(np.float32),
            dyn_omega.astype(np.float32))


# --------------------------------------------------------------------------
# Gaussians and cameras
# --------------------------------------------------------------------------
SIDE_CAM = dict(  # cfg/env/xarm_gripper.yaml:21-35
    w=848, h=480,
    intr=[427.2920227050781, 0., 429.9993591308594, 0., 426.7926940917969, 242.8115234375, 0., 0., 1.],
    c2w=[0.005258014128948334, 0.6125512321694572, -0.7904133989597472, 0.8830263898083726,
         0.9999860093046595, -0.0036779908994199082, 0.0038017861441641317, 0.05390846195611962,
         -0.000578344501100992, -0.7904223303719503, -0.6125620010799026, 0.3976033855145515,
         0.0, 0.0, 0.0, 1.0],
)
WRIST_LIKE_CAM = dict(  # second fixed view for the two-camera config: same intrinsics, top-down-ish
    w=848, h=480,
    intr=[433.2635498046875, 0., 425.69775390625, 0., 433.2635498046875, 244.70132446289062, 0., 0., 1.],
    c2w=[0.0, -1.0, 0.0, 0.45,
         -1.0, 0.0, 0.0, 0.0,
         0.0, 0.0, -1.0, 0.9,
         0.0, 0.0, 0.0, 1.0],
)


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    view: np.ndarray    # (16,) float32: w2c transposed, flattened (column-major w2c)
    proj: np.ndarray    # (16,) float32: (opengl_proj @ w2c) transposed, flattened
    campos: np.ndarray  # (3,) float32
    z_threshold: float = 0.05


def setup_camera(w, h, k, w2c, near=0.01, far=100.0, z_threshold=0.05) -> Camera:
    """numpy restatement of sim/utils/gs/transform_utils.py:7-31 (float32 throughout)."""
    k = np.asarray(k, dtype=np.float64).reshape(3, 3)
    fx, fy, cx, cy = k[0, 0], k[1, 1], k[0, 2], k[1, 2]
    w2c32 = np.asarray(w2c, dtype=np.float32).reshape(4, 4)
    campos = np.linalg.inv(w2c32)[:3, 3].astype(np.float32)
    view = w2c32.T.copy()
    opengl = np.array([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                       [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                       [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                       [0.0, 0.0, 1.0, 0.0]], dtype=np.float32)
    full = (view @ opengl.T).astype(np.float32)
    return Camera(int(w), int(h), float(w / (2 * fx)), float(h / (2 * fy)),
                  view.reshape(-1).copy(), full.reshape(-1).copy(), campos, z_threshold)


def make_camera(W: int, H: int, which: str = "side", jitter_seed: int | None = None) -> Camera:
    """Config camera with intrinsics scaled to (W, H); optional small seeded pose jitter."""
    cfg = SIDE_CAM if which == "side" else WRIST_LIKE_CAM
    k = np.asarray(cfg["intr"], dtype=np.float64).reshape(3, 3).copy()
    k[0] *= W / cfg["w"]
    k[1] *= H / cfg["h"]
    c2w = np.asarray(cfg["c2w"], dtype=np.float64).reshape(4, 4).copy()
    if jitter_seed is not None:
        rng = np.random.default_rng(jitter_seed)
        c2w[:3, 3] += rng.uniform(-0.02, 0.02, 3)
    return setup_camera(W, H, k, np.linalg.inv(c2w))


def spatial_order(pts: np.ndarray) -> np.ndarray:
    """Permutation that sorts points along a 30-bit Morton (Z-order) curve of their bounding box: consecutive points are
    spatial neighbours.  Used to lay the object Gaussians out so that the 32 Gaussians of a warp share most of their
    bones (the LBS blend then gathers a handful of distinct transforms per warp instead of 32 x 16)."""
    p = np.asarray(pts, np.float64)
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-12) * 1023.0).astype(np.int64), 0, 1023)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        return (v | (v << 2)) & 0x09249249

    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2), kind="stable")


@dataclass
class Gaussians:
    means3D: np.ndarray    # (P,3)
    scales: np.ndarray     # (P,3) (already exp-activated)
    rotations: np.ndarray  # (P,4) unit wxyz
    opacities: np.ndarray  # (P,1) (already sigmoid-activated)
    shs: np.ndarray        # (P,1,3) degree-0 coefficients (use_shs False, gs_renderer.py:945-947)
    n_object: int = 0
    bind_idx: np.ndarray | None = None   # (n_object, K) particle ids (object Gaussians)
    bind_w: np.ndarray | None = None     # (n_object, K) weights, rows sum to 1


def make_gaussians(seed: int, P: int = 200_000, particles: np.ndarray | None = None,
                   n_object: int | None = None, knn: int = 16) -> Gaussians:
    """Synthetic scan: n_object Gaussians bound to the particle cloud by
    K-NN inverse-distance weights + scene Gaussians in a 1.2 x 1.2 x 0.6 m table volume."""
    rng = np.random.default_rng(seed)
    if n_object is None:
        n_object = P // 10 if particles is not None else 0
    n_scene = P - n_object
    scene = rng.uniform([-0.1, -0.6, 0.0], [1.1, 0.6, 0.6], (n_scene, 3))
    bind_idx = bind_w = None
    if n_object:
        from scipy.spatial import cKDTree
        base = particles[rng.integers(0, len(particles), n_object)].astype(np.float64)
        obj = base + rng.normal(0.0, 0.002, (n_object, 3))
        k = min(knn, len(particles))
        dist, bind_idx = cKDTree(particles).query(obj, k=k)
        w = 1.0 / np.maximum(dist, 1e-6)
        bind_w = (w / w.sum(1, keepdims=True)).astype(np.float32)
        bind_idx = bind_idx.astype(np.int32)
        means = np.concatenate([obj, scene], 0)
    else:
        means = scene
    scales = np.exp(rng.normal(np.log(0.006), 0.5, (P, 3)))
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opa = 1.0 / (1.0 + np.exp(-rng.normal(1.5, 1.5, (P, 1))))
    rgb = rng.uniform(0.0, 1.0, (P, 1, 3))
    shs = (rgb - 0.5) / 0.28209479177387814
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return Gaussians(f32(means), f32(scales), f32(q), f32(opa), f32(shs), n_object, bind_idx, bind_w)


# ---------------------------------------------------------------------------- robot scan
XARM_LINK_IDS = (1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 14, 15, 16)   # robot_pc_transformations.py:35
XARM_N_LINKS = 18                                                      # :34 (0: world, 9: link_eef, 17: link_tcp)


def _rot_from_rotvec(r):
    r = np.asarray(r, np.float64)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _rigid(rotvec, t):
    m = np.eye(4)
    m[:3, :3] = _rot_from_rotvec(rotvec)
    m[:3, 3] = t
    return m


@dataclass
class RobotScan:
    """A synthetic stand-in for the scanned robot Gaussians and their link bookkeeping (the real scan and the
    URDF are not shipped): a serial chain of 15 moving links above the table."""
    link_names: list
    total_mask: np.ndarray      # (n,) float32 link index per Gaussian, as total_mask_path stores it (gs_renderer.py:505-506)
    link_id: np.ndarray         # (n,) int32 slot in XARM_LINK_IDS order, -1 if the link is not listed
    points: np.ndarray          # (n,3) float32 rest positions (at base_qpos)
    quats: np.ndarray           # (n,4) float32 w,x,y,z, un-normalised
    link_offset: np.ndarray     # (15,4,4) float64 tf_obj_to_link
    base_pose: np.ndarray       # (15,4,4) float64 FK pose at base_qpos


def make_robot_scan(n: int = 30_000, seed: int = 1234, origin=(-0.25, 0.0, 0.05), volume=None) -> RobotScan:
    """`volume=(lo, hi)`: rest positions uniform in that box instead of clustered around the links (keeps the
    bench's Gaussian density that of SURVEY.md §8d while the rows still move rigidly with their links)."""
    rng = np.random.default_rng(seed)
    names = ["world", "link_base"] + [f"link{i}" for i in range(1, 8)] + ["link_eef"] + \
            [f"finger{i}" for i in range(7)] + ["link_tcp"]
    L = len(XARM_LINK_IDS)
    base_pose = np.zeros((L, 4, 4))
    offs = np.zeros((L, 4, 4))
    cur = _rigid((0, 0, 0), origin)
    for s in range(L):
        cur = cur @ _rigid(rng.normal(0, 0.4, 3), rng.uniform(0.02, 0.09, 3) * (1 if s < 9 else 0.3))
        base_pose[s] = cur
        offs[s] = _rigid(rng.normal(0, 0.2, 3), rng.normal(0, 0.01, 3))
    # a few Gaussians belong to links that never move (world / link_eef / link_tcp): they keep their place
    mask = rng.choice(np.arange(XARM_N_LINKS), size=n, p=np.r_[0.03, np.full(8, 0.08), 0.01, np.full(7, 0.045), 0.005])
    slot_of = -np.ones(XARM_N_LINKS, np.int32)
    slot_of[list(XARM_LINK_IDS)] = np.arange(L, dtype=np.int32)
    link_id = slot_of[mask]
    local = rng.normal(0, 0.025, (n, 3))
    centre = np.where(link_id[:, None] >= 0, (base_pose @ offs)[np.maximum(link_id, 0), :3, 3], np.asarray(origin)[None])
    points = (centre + local).astype(np.float32)
    if volume is not None:
        lo, hi = np.asarray(volume[0], np.float64), np.asarray(volume[1], np.float64)
        points = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    quats = (rng.normal(0, 1, (n, 4)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
    return RobotScan(names, mask.astype(np.float32), link_id.astype(np.int32), points, quats, offs, base_pose)


def robot_link_poses(scan: RobotScan, seed: int, amount: float = 0.3) -> np.ndarray:
    """(15,4,4) float64 FK poses of one frame: every joint of the chain turned by a seeded angle, so the
    links move rigidly and coherently (what sapien's FK would return for some qpos)."""
    rng = np.random.default_rng(seed)
    L = len(scan.base_pose)
    out = np.zeros_like(scan.base_pose)
    prev_base, prev_new = np.eye(4), np.eye(4)
    for s in range(L):
        rel = np.linalg.inv(prev_base) @ scan.base_pose[s]           # joint frame at rest
        new = prev_new @ _rigid(rng.normal(0, amount, 3), rng.normal(0, 0.002, 3)) @ rel
        out[s] = new
        prev_base, prev_new = scan.base_pose[s], new
    return out
