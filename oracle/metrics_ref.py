"""oracle/metrics_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy) of the per-frame task-success tests and the episode rule:
  is_pusht_success   experiments/utils/calculate_success_T.py:17-29
  is_rope_success    experiments/utils/calculate_success_rope.py:38-75 (plane test), 77-134 (counts), 143-170
  is_sloth_success   experiments/utils/calculate_success_sloth.py:140-172
  episode rule       calculate_success_T.py:63-73, _rope.py:193-203, _sloth.py:194-204

Pinning.  `load_reference(task)` imports the reference's OWN scripts from /root/reference (pure numpy; the state
dictionaries they expect are rebuilt around plain arrays by `reference_frame_test`): the push-T and rope tests
are pinned to the reference's code that way (tests/golden/metrics_*.npz, made by
tests/golden/make_metrics_golden.py).  The sloth test calls Open3D (`get_minimal_oriented_bounding_box`,
`get_point_indices_within_bounding_box`); open3d is a dependency of the reference that is not installed in
this image, so the point-in-OBB test is restated from Open3D's published source
(OrientedBoundingBox::GetPointIndicesWithinBoundingBox: |(p - c) . axis_k| <= extent_k / 2) and is
PARITY UNPINNED; the box itself (a setup-time constant of the container mesh) is an input.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_DIR = "/root/reference/experiments/utils"
TASKS = {"pusht": 0, "rope": 1, "sloth": 2}
START_FRAME = {"pusht": 1700, "rope": 800, "sloth": 350}   # first counted pickle of an episode
NEED_FRAMES = 30
THRESHOLD = {"pusht": 0.002, "rope": 100, "sloth": 3050}


def rope_box():
    """The routing box of is_rope_success (calculate_success_rope.py:152-160) as (min_xyz, max_xyz)."""
    center = np.array([0.62, 0.05, 0.0])
    lo, hi = center.copy(), center.copy()
    lo[0] -= 0.035 / 2
    hi[0] += 0.035 / 2
    lo[1] -= 0.035 / 2
    hi[1] += 0.035 / 2
    lo[2] -= 0.0
    hi[2] += 0.03
    return lo, hi


def pusht_mse(x, target):
    """((x - x_target) ** 2).sum(1).mean() on float32 arrays (calculate_success_T.py:26)."""
    x, target = np.asarray(x, np.float32), np.asarray(target, np.float32)
    return ((x - target) ** 2).sum(1).mean()


def plane_hits(p0, p1, y_plane, x_min, x_max, z_min, z_max, eps=1e-12):
    """_segment_plane_intersections_xz (calculate_success_rope.py:38-75), float64."""
    y0, y1 = p0[:, 1], p1[:, 1]
    dy = y1 - y0
    parallel = np.abs(dy) <= eps
    t = np.zeros_like(dy)
    np.divide(y_plane - y0, dy, out=t, where=~parallel)
    on_seg = (~parallel) & (t >= -eps) & (t <= 1.0 + eps)
    xi = p0[:, 0] + t * (p1[:, 0] - p0[:, 0])
    zi = p0[:, 2] + t * (p1[:, 2] - p0[:, 2])
    rect = lambda x, z: (x >= x_min - eps) & (x <= x_max + eps) & (z >= z_min - eps) & (z <= z_max + eps)
    coplanar = parallel & (np.abs(y0 - y_plane) <= eps)
    return (on_seg & rect(xi, zi)) | (coplanar & (rect(p0[:, 0], p0[:, 2]) | rect(p1[:, 0], p1[:, 2])))


def rope_counts(x, springs, box=None):
    """(y_min_count, y_max_count) of count_xz_plane_intersections (calculate_success_rope.py:77-134)."""
    lo, hi = rope_box() if box is None else box
    V = np.asarray(x, dtype=float)
    E = np.asarray(springs, dtype=int)
    p0, p1 = V[E[:, 0]], V[E[:, 1]]
    a = plane_hits(p0, p1, lo[1], lo[0], hi[0], lo[2], hi[2])
    b = plane_hits(p0, p1, hi[1], lo[0], hi[0], lo[2], hi[2])
    return int(np.count_nonzero(a)), int(np.count_nonzero(b))


def obb_count(x, center, R, extent):
    """Open3D OrientedBoundingBox::GetPointIndicesWithinBoundingBox, float64 (restated; see module docstring)."""
    d = np.asarray(x, dtype=float) - np.asarray(center, dtype=float)[None]
    R, extent = np.asarray(R, dtype=float), np.asarray(extent, dtype=float)
    inside = np.ones(len(d), bool)
    for k in range(3):
        proj = d[:, 0] * R[0, k] + d[:, 1] * R[1, k] + d[:, 2] * R[2, k]
        inside &= np.abs(proj) <= extent[k] / 2
    return int(inside.sum())


def frame_test(task, x, *, target=None, springs=None, box=None, obb=None):
    """(passed, value0, value1) for one environment and frame."""
    if task == "pusht":
        mse = pusht_mse(x, target)
        return bool(mse < THRESHOLD[task]), float(mse), 0.0
    if task == "rope":
        c0, c1 = rope_counts(x, springs, box)
        return bool(c0 >= 100 and c1 >= 100), float(c0), float(c1)
    c = obb_count(x, *obb)
    return bool(c >= THRESHOLD[task]), float(c), 0.0


def episode_rule(passed_per_frame, start_frame, need=NEED_FRAMES):
    """hits counted from start_frame on; success once `need` frames passed."""
    hits, success, out = 0, False, []
    for f, p in enumerate(passed_per_frame):
        if f >= start_frame and p:
            hits += 1
            success = success or hits >= need
        out.append((hits, success))
    return out


# --------------------------------------------------------------------------- the reference itself
def load_reference(task):
    """The reference's calculate_success_<task>.py module, or None when /root/reference is absent.  The sloth
    script imports open3d at module scope (stubbed: only its numpy helpers are usable then)."""
    name = {"pusht": "calculate_success_T.py", "rope": "calculate_success_rope.py",
            "sloth": "calculate_success_sloth.py"}[task]
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        return None
    saved = sys.modules.get("open3d")
    if task == "sloth":
        sys.modules["open3d"] = types.ModuleType("open3d")
    try:
        spec = importlib.util.spec_from_file_location("_r2s_ref_" + name[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if task == "sloth":
            if saved is None:
                sys.modules.pop("open3d", None)
            else:
                sys.modules["open3d"] = saved
    return mod


def reference_frame_test(mod, task, x, *, target=None, springs=None):
    """Runs the reference's own is_pusht_success / is_rope_success on one state."""
    import torch
    state = {"renderer": {"x": torch.tensor(np.asarray(x, np.float32))}}
    if task == "pusht":
        init = {"physics": {"static_meshes": []}}
        return bool(mod.is_pusht_success(state, np.asarray(target, np.float32), init))
    init = {"physics": {"static_meshes": [{"vertices": np.zeros((1, 3)), "faces": np.zeros((0, 3), int)}],
                        "init_springs": torch.tensor(np.asarray(springs, np.int64))}}
    return bool(mod.is_rope_success(state, init))
