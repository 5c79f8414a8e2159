"""oracle/lbs_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, float32) of the linear-blend-skinning step that moves the object Gaussians
with the particles between physics and render: `interpolate_motions` of
sim/utils/gs/transform_utils.py:58-212 as sim/renderer/gs_renderer.py:732-749 calls it (quat=None,
precomputed relations / weights; only the transformed xyz is consumed).

Pinning: `load_reference()` imports the reference's OWN function from /root/reference (pure torch, runs on
CPU; `kornia` is only touched on the quat path and is stubbed).  tests/golden/lbs_*.npz were generated from
it by tests/golden/make_lbs_golden.py; tests/test_oracle_lbs.py checks this restatement against those vectors
and, when /root/reference is mounted, against the live function.

Reference behaviour restated, including its quirk: per bone F = sum_a (new_a - new_i)(old_a - old_i)^T over the
k_rel neighbours; R = U diag(1,1,+-1) V^T (Kabsch).  The reference writes `bone_transforms[:, :3, :3] = R` with R
computed for the rank >= 2 bones only, so if ANY bone has rank < 2 the assignment fails and its `except` branch
gives EVERY bone the identity rotation (transform_utils.py:159-167).  xyz' = sum_k w_k (R_b (xyz - bone_b) +
motion_b + bone_b).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_FILE = "/root/reference/sim/utils/gs/transform_utils.py"


def load_reference():
    """The reference's interpolate_motions (torch, CPU) or None when /root/reference is absent."""
    if not os.path.exists(REF_FILE):
        return None
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(here, "real2sim_eval_b200", "compat")   # `diff_gaussian_rasterization` import at file top
    if shim not in sys.path:
        sys.path.insert(0, shim)
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))  # only used when quat is not None
    spec = importlib.util.spec_from_file_location("_ref_transform_utils", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.interpolate_motions


def bone_rotations(bones, motions, relations):
    """(R (n,3,3) float32, all_rank_ok) -- transform_utils.py:73-167."""
    bones = np.asarray(bones, np.float32)
    motions = np.asarray(motions, np.float32)
    n = len(bones)
    adj = bones[relations] - bones[:, None]
    adj_new = (bones[relations] + motions[relations]) - (bones[:, None] + motions[:, None])
    F = (adj_new.transpose(0, 2, 1) @ adj).astype(np.float32)      # W is the identity
    rank = np.linalg.matrix_rank(F)
    if not (rank >= 2).all():
        return np.tile(np.eye(3, dtype=np.float32), (n, 1, 1)), False
    U, _, Vt = np.linalg.svd(F)
    S = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1))
    S[np.linalg.det(F) < 0, 2, 2] = -1
    R = U @ S @ Vt
    flip = np.abs(np.linalg.det(R) + 1) < 1e-3
    S[flip, 2, 2] *= -1
    R = U @ S @ Vt
    return R.astype(np.float32), True


def interpolate_motions(bones, motions, relations, xyz, weights, weights_indices):
    """Transformed xyz (P,3) float32 for quat=None -- transform_utils.py:183-189."""
    bones = np.asarray(bones, np.float32)
    motions = np.asarray(motions, np.float32)
    xyz = np.asarray(xyz, np.float32)
    weights = np.asarray(weights, np.float32)
    R, _ = bone_rotations(bones, motions, relations)
    b = bones[weights_indices]                                   # (P,k,3)
    t = np.einsum("pkij,pkj->pki", R[weights_indices], xyz[:, None] - b)
    t = t + motions[weights_indices] + b
    return (t * weights[:, :, None]).sum(axis=1).astype(np.float32)


def knn_relations(bones, k=8):
    """gs_renderer.py:195-200 (k_rel = 8): k nearest other bones."""
    from scipy.spatial import cKDTree
    _, idx = cKDTree(bones).query(bones, k=k + 1)
    return idx[:, 1:].astype(np.int64)


def knn_weights(bones, pts, k=16):
    """gs_renderer.py:202-211 (k_wgt = 16): inverse-distance weights over the k nearest bones."""
    from scipy.spatial import cKDTree
    dist, idx = cKDTree(bones).query(pts, k=k)
    w = 1.0 / (dist.astype(np.float32) + np.float32(1e-6))
    return (w / w.sum(-1, keepdims=True)).astype(np.float32), idx.astype(np.int64)
