"""oracle/links_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, float32) of the per-frame re-posing of the robot's Gaussians:
  transform_gs_xarm_gripper          sim/utils/robot/robot_pc_transformations.py:12-55  (mask gather / scatter)
  RobotPcSampler.transform_gs_torch  sim/utils/robot/robot_pc_sampler.py:119-162        (per-link transform)
  quat_mult_torch                    sim/utils/robot/robot_pc_sampler.py:17-24
  the final normalisation            sim/renderer/gs_renderer.py:905

Pinning.  `load_reference()` imports the reference's OWN robot_pc_sampler.py from /root/reference and runs its
`transform_gs_torch` / `quat_mult_torch` on CPU.  The module's heavy imports are stubbed: open3d, urdfpy and
sapien (asset loading and forward kinematics -- host-side, out of scope; the FK poses are INPUTS here, supplied
by a fake `robot_model`), and kornia.  kornia is a third-party dependency the reference neither vendors nor
pins (pyproject.toml:26 "kornia", no version) and it is not installed in this image: its
`geometry.conversions.rotation_matrix_to_quaternion` (w,x,y,z order, the one quat_mult_torch assumes) is
restated below from its published four-branch algorithm (`rotation_matrix_to_quaternion`), and that restatement
is what the stub hands to the reference.  So: the gather/transform/quaternion-product/scatter arithmetic is
pinned to the reference's code; the matrix->quaternion conversion alone is PARITY UNPINNED.
tests/golden/links_*.npz were generated that way by tests/golden/make_links_golden.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_FILE = "/root/reference/sim/utils/robot/robot_pc_sampler.py"
F32_TINY = np.float32(1.17549435e-38)


# --------------------------------------------------------------------------- kornia restated
def rotation_matrix_to_quaternion(R, eps=1e-8):
    """kornia.geometry.conversions.rotation_matrix_to_quaternion, (..., 3, 3) -> (..., 4) as (w, x, y, z).
    Branch on trace > 0, else on the largest diagonal entry; sq = 2 sqrt(1 + (+-)m00 (+-)m11 (+-)m22 + eps);
    the component matching the branch is sq / 4, the others (m_ij -+ m_ji) / clamp(sq, min=tiny)."""
    R = np.asarray(R, np.float32)
    m = R.reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[:, k] for k in range(9)]
    eps = np.float32(eps)
    one, two, quarter = np.float32(1), np.float32(2), np.float32(0.25)
    sdiv = lambda n, d: (n / np.maximum(d, F32_TINY)).astype(np.float32)
    with np.errstate(invalid="ignore"):
        trace = m00 + m11 + m22
        sq = np.sqrt(trace + one + eps) * two
        q_t = np.stack([quarter * sq, sdiv(m21 - m12, sq), sdiv(m02 - m20, sq), sdiv(m10 - m01, sq)], -1)
        sq = np.sqrt(one + m00 - m11 - m22 + eps) * two
        q_1 = np.stack([sdiv(m21 - m12, sq), quarter * sq, sdiv(m01 + m10, sq), sdiv(m02 + m20, sq)], -1)
        sq = np.sqrt(one + m11 - m00 - m22 + eps) * two
        q_2 = np.stack([sdiv(m02 - m20, sq), sdiv(m01 + m10, sq), quarter * sq, sdiv(m12 + m21, sq)], -1)
        sq = np.sqrt(one + m22 - m00 - m11 + eps) * two
        q_3 = np.stack([sdiv(m10 - m01, sq), sdiv(m02 + m20, sq), sdiv(m12 + m21, sq), quarter * sq], -1)
    w2 = np.where((m11 > m22)[:, None], q_2, q_3)
    w1 = np.where(((m00 > m11) & (m00 > m22))[:, None], q_1, w2)
    q = np.where((trace > 0)[:, None], q_t, w1)
    return q.reshape(R.shape[:-2] + (4,)).astype(np.float32)


# --------------------------------------------------------------------------- the restatement
def normalize(v, eps=1e-12):
    """torch.nn.functional.normalize(v, dim=-1): v / max(||v||, eps)."""
    v = np.asarray(v, np.float32)
    n = np.sqrt((v * v).sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (v / np.maximum(n, np.float32(eps))).astype(np.float32)


def quat_mult(q1, q2):
    """robot_pc_sampler.py:17-24, (w, x, y, z)."""
    w1, x1, y1, z1 = [q1[..., k] for k in range(4)]
    w2, x2, y2, z2 = [q2[..., k] for k in range(4)]
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1).astype(np.float32)


def link_matrices(link_pose, base_pose, link_offset):
    """mat_l = (pose_l @ offset_l) @ inverse(base_pose_l @ offset_l), float32 (robot_pc_sampler.py:138-150)."""
    pose = np.asarray(link_pose, np.float32).reshape(-1, 4, 4)
    base = np.asarray(base_pose, np.float32).reshape(-1, 4, 4)
    off = np.asarray(link_offset, np.float32).reshape(-1, 4, 4)
    rest_inv = np.linalg.inv((base @ off).astype(np.float32)).astype(np.float32)
    return ((pose @ off).astype(np.float32) @ rest_inv).astype(np.float32)


def transform_gs(points, quats, link_id, link_pose, base_pose, link_offset):
    """(points', quats') for one environment.  link_id[g] = slot of Gaussian g in the link tables, or -1 for a
    Gaussian no listed link owns (kept in place; its quaternion only normalised)."""
    points = np.asarray(points, np.float32)
    link_id = np.asarray(link_id).reshape(-1)
    mats = link_matrices(link_pose, base_pose, link_offset)
    ql = rotation_matrix_to_quaternion(mats[:, :3, :3])
    q = normalize(quats)                                   # robot_pc_transformations.py:29
    out_p, out_q = points.copy(), q.copy()
    mv = link_id >= 0
    M = mats[link_id[mv]]
    out_p[mv] = (np.einsum("gij,gj->gi", M[:, :3, :3], points[mv]) + M[:, :3, 3]).astype(np.float32)   # :151
    out_q[mv] = quat_mult(ql[link_id[mv]], q[mv])          # :154
    return out_p, normalize(out_q)                         # gs_renderer.py:905


# --------------------------------------------------------------------------- the reference itself, on CPU
class _FakePose:
    def __init__(self, m):
        self._m = m

    def to_transformation_matrix(self):
        return self._m


class _FakeRobotModel:
    """Stands in for sapien's pinocchio model: poses are looked up from tables the test supplies, keyed by
    which qpos was last passed to compute_forward_kinematics (robot_pc_sampler.py:131-136)."""

    def __init__(self, tables):
        self.tables, self.cur = tables, None

    def compute_forward_kinematics(self, qpos):
        self.cur = self.tables[float(np.asarray(qpos).reshape(-1)[0])]

    def get_link_pose(self, idx):
        return _FakePose(self.cur[idx])


class _FakeLink:
    def __init__(self, name):
        self.name = name


class _FakeRobot:
    def __init__(self, names):
        self._links = [_FakeLink(n) for n in names]

    def get_links(self):
        return self._links


def load_reference():
    """The reference's robot_pc_sampler module (torch, CPU) with its asset/FK imports stubbed, or None when
    /root/reference is absent."""
    if not os.path.exists(REF_FILE):
        return None
    import torch

    def _rm2q(rot):   # the kornia stub: this file's restatement (see the module docstring)
        return torch.from_numpy(rotation_matrix_to_quaternion(rot.detach().cpu().numpy())).to(rot.dtype)

    kornia = types.ModuleType("kornia")
    kornia.geometry = types.ModuleType("kornia.geometry")
    kornia.geometry.conversions = types.ModuleType("kornia.geometry.conversions")
    kornia.geometry.conversions.rotation_matrix_to_quaternion = _rm2q
    stubs = {"kornia": kornia, "kornia.geometry": kornia.geometry, "kornia.geometry.conversions": kornia.geometry.conversions,
             "open3d": types.ModuleType("open3d"), "urdfpy": types.ModuleType("urdfpy"), "sapien": types.ModuleType("sapien"),
             "sapien.core": types.ModuleType("sapien.core")}
    stubs["urdfpy"].URDF = object
    stubs["sapien"].core = stubs["sapien.core"]
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_ref_robot_pc_sampler", REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def reference_transform_gs(mod, points, quats, total_mask, link_id_list, link_names, link_pose, base_pose, link_offset):
    """Runs the reference's own code path for one environment: the mask gather / scatter of
    transform_gs_xarm_gripper (robot_pc_transformations.py:29, 44-52; restated here line for line because that
    file imports sapien at module scope and builds a URDF loader) around the reference's unmodified
    RobotPcSampler.transform_gs_torch, followed by the renderer's normalisation (gs_renderer.py:905)."""
    import torch
    n_links = len(link_names)
    sampler = mod.RobotPcSampler.__new__(mod.RobotPcSampler)
    sampler.sapien_robot = _FakeRobot(link_names)
    pose_t = {i: np.asarray(link_pose[s], np.float64) for s, i in enumerate(link_id_list)}
    base_t = {i: np.asarray(base_pose[s], np.float64) for s, i in enumerate(link_id_list)}
    sampler.robot_model = _FakeRobotModel({1.0: pose_t, 0.0: base_t})
    sampler.offsets = {link_names[i]: np.asarray(link_offset[s], np.float64) for s, i in enumerate(link_id_list)}
    scan_points = torch.tensor(np.asarray(points, np.float32))
    scan_quats = torch.nn.functional.normalize(torch.tensor(np.asarray(quats, np.float32)), dim=-1)
    total_mask = torch.tensor(np.asarray(total_mask, np.float32))
    links = sampler.sapien_robot.get_links()
    assert len(links) == n_links
    points_links = {links[i].name: scan_points[total_mask == i] for i in link_id_list}
    quats_links = {links[i].name: scan_quats[total_mask == i] for i in link_id_list}
    new_p, new_q = sampler.transform_gs_torch(points_links, quats_links, np.array([1.0]), base_qpos=np.array([0.0]))
    n = 0
    for i in link_id_list:
        k = len(points_links[links[i].name])
        scan_points[total_mask == i] = new_p[n:n + k]
        scan_quats[total_mask == i] = new_q[n:n + k]
        n += k
    quat = torch.nn.functional.normalize(scan_quats, dim=-1)
    return scan_points.numpy(), quat.numpy()
