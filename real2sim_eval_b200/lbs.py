"""Host side of the LBS step (object Gaussians follow the particles) over include/r2s_lbs.h.

`interpolate_motions` keeps the name and arguments of sim/utils/gs/transform_utils.py:58 for the way
sim/renderer/gs_renderer.py:740-749 calls it (precomputed relations and weights, quat=None);
`BatchedLBS` is the E-environment entry the batched env step uses.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def slot_layout(bone_positions, weights_indices, weights):
    """The layout hint of include/r2s_lbs.h from the bones' rest positions: `bone_slot` [N] = rank of each bone along
    a Morton curve, and per Gaussian the same (bone, weight) pairs with the bone replaced by its slot, ascending."""
    from .synth import spatial_order
    pos = np.asarray(bone_positions, np.float64).reshape(-1, 3)
    slot = np.empty(pos.shape[0], np.int64)
    slot[spatial_order(pos)] = np.arange(pos.shape[0])
    ws = slot[np.asarray(weights_indices, np.int64)]
    order = np.argsort(ws, axis=1, kind="stable")
    return (slot.astype(np.int32), np.take_along_axis(ws, order, 1).astype(np.int32),
            np.take_along_axis(np.asarray(weights, np.float32), order, 1))


class BatchedLBS:
    """Shared relations / weights (one PhysTwin), per-environment bones and Gaussians."""

    def __init__(self, E, N, P, n_obj, relations, weights, weights_indices, device="cuda", bone_positions=None):
        self.device = dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.R2SError("BatchedLBS needs a CUDA device: there is no CPU path")
        self.lib = _lib.load()
        self.E, self.N, self.P, self.n_obj = int(E), int(N), int(P), int(n_obj)
        as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))).to(device=dev, dtype=dt).contiguous()
        self.relations = as_t(relations, torch.int32)
        self.weights = as_t(weights, torch.float32)
        self.weights_indices = as_t(weights_indices, torch.int32)
        assert self.relations.shape[0] == self.N and self.weights.shape == self.weights_indices.shape
        assert self.weights.shape[0] == self.n_obj
        # layout hint (r2s_lbs.h): with the bones' rest positions known, their transforms are stored along a Morton
        # curve and every Gaussian lists its bones by ascending slot, so the Gaussians of a warp (themselves laid
        # out spatially by the caller) read neighbouring 48-byte rows of shared memory instead of 32 random ones
        self.bone_slot = self.weights_slots = self.weights_by_slot = None
        if bone_positions is not None and self.n_obj > 0:
            slot, ws, wbs = slot_layout(bone_positions, self.weights_indices.cpu().numpy(), self.weights.cpu().numpy())
            self.bone_slot, self.weights_slots = as_t(slot, torch.int32), as_t(ws, torch.int32)
            self.weights_by_slot = as_t(wbs, torch.float32)
        self.rot = torch.empty((self.E, self.N, 12), dtype=torch.float32, device=dev)
        self.rank_flags = torch.ones(self.E, dtype=torch.int32, device=dev)

    def forward(self, bones4, bones_new4, means3D):
        """bones4 / bones_new4: [E,N,4] float32 (particle state before / after the frame);
        means3D: [E,P,3] float32, rows < n_obj transformed in place."""
        assert bones4.is_contiguous() and bones_new4.is_contiguous() and means3D.is_contiguous()
        a = _lib.LbsArgs()
        a.E, a.N, a.P, a.n_obj = self.E, self.N, self.P, self.n_obj
        a.k_rel, a.k_wgt = self.relations.shape[1], self.weights.shape[1]
        a.relations, a.weights_indices, a.weights = _ptr(self.relations), _ptr(self.weights_indices), _ptr(self.weights)
        a.bones4, a.bones_new4, a.means3D = _ptr(bones4), _ptr(bones_new4), _ptr(means3D)
        a.rot_scratch, a.rank_flags = _ptr(self.rot), _ptr(self.rank_flags)
        a.bone_slot, a.weights_slots, a.weights_by_slot = _ptr(self.bone_slot), _ptr(self.weights_slots), _ptr(self.weights_by_slot)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_lbs_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "r2s_lbs_forward")
        return means3D


def interpolate_motions(bones, motions, relations, xyz, rot=None, quat=None, weights=None, weights_indices=None,
                        device="cuda", step="n/a"):
    """Drop-in for sim/utils/gs/transform_utils.py:58 on the path the renderer uses (quat=None): returns
    (xyz_transformed, None, weights)."""
    if quat is not None or rot is not None:
        raise NotImplementedError("only the xyz path (quat=None, rot=None) of interpolate_motions is provided; "
                                  "sim/renderer/gs_renderer.py:740-749 uses no other")
    dev = torch.device(device)
    bones = torch.as_tensor(bones, device=dev).to(torch.float32)
    motions = torch.as_tensor(motions, device=dev).to(torch.float32)
    xyz = torch.as_tensor(xyz, device=dev).to(torch.float32)
    if weights is None:  # transform_utils.py:170-179 (not on the per-frame path: the renderer precomputes them)
        dist = torch.norm(xyz[:, None] - bones, dim=-1)
        _, weights_indices = torch.topk(dist, 5, dim=-1, largest=False)
        dist = torch.norm(bones[weights_indices] - xyz[:, None], dim=-1)
        weights = 1 / (dist + 1e-6)
        weights = weights / weights.sum(dim=-1, keepdim=True)
    n, P = bones.shape[0], xyz.shape[0]
    pad = lambda t: torch.cat([t, torch.zeros_like(t[:, :1])], dim=1).contiguous()[None]
    lbs = BatchedLBS(1, n, P, P, relations, weights, weights_indices, device=dev)
    out = xyz.clone().contiguous()[None]
    lbs.forward(pad(bones), pad(bones + motions), out)
    return out[0], None, weights
