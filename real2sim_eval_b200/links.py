"""Host side of the robot-Gaussian re-posing step over include/r2s_links.h.

`BatchedLinkTransform` is the E-environment entry the batched env step uses: one robot scan (rest
positions / rotations / link slot per Gaussian, shared), one FK pose table per environment and frame.
`transform_gs` keeps the meaning of RobotPcSampler.transform_gs_torch + the mask scatter of
transform_gs_xarm_gripper (sim/utils/robot/robot_pc_sampler.py:119-162,
sim/utils/robot/robot_pc_transformations.py:12-55) for one environment, with the FK poses given by the
caller (forward kinematics itself stays on the host: sapien, out of scope).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

MAX_LINKS = 64  # R2S_LINKS_MAX


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def rest_inverse(base_pose, link_offset):
    """inverse(base_pose_l @ offset_l) per link in float32 (robot_pc_sampler.py:145-147), made once."""
    bp = torch.as_tensor(np.asarray(base_pose), dtype=torch.float32).reshape(-1, 4, 4)
    off = torch.as_tensor(np.asarray(link_offset), dtype=torch.float32).reshape(-1, 4, 4)
    return torch.linalg.inv(bp @ off).contiguous()


class BatchedLinkTransform:
    """Shared robot scan, per-environment link poses.  Rows [first, first + n_robot) of every
    environment's means3D / rotations are rewritten each frame from the rest scan."""

    def __init__(self, E, P, first, link_id, rest_means, rest_quats, link_offset, base_pose, device="cuda"):
        self.device = dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.R2SError("BatchedLinkTransform needs a CUDA device: there is no CPU path")
        self.lib = _lib.load()
        as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))).to(device=dev, dtype=dt).contiguous()
        self.link_id = as_t(link_id, torch.int32).reshape(-1)
        self.n_robot = int(self.link_id.numel())
        self.rest_means = as_t(rest_means, torch.float32).reshape(self.n_robot, 3)
        self.rest_quats = as_t(rest_quats, torch.float32).reshape(self.n_robot, 4)
        self.link_offset = as_t(link_offset, torch.float32).reshape(-1, 16)
        self.L = int(self.link_offset.shape[0])
        if not 0 < self.L <= MAX_LINKS:
            raise ValueError(f"1..{MAX_LINKS} links supported, got {self.L}")
        if self.n_robot and int(self.link_id.max()) >= self.L:
            raise ValueError("link_id refers to a slot outside the link table")
        self.rest_inv = rest_inverse(base_pose, self.link_offset.cpu()).to(dev).reshape(self.L, 16).contiguous()
        self.E, self.P, self.first = int(E), int(P), int(first)
        if self.first < 0 or self.first + self.n_robot > self.P:
            raise ValueError("robot rows do not fit in the P rows of an environment")
        self.scratch = torch.empty((self.E, self.L, 16), dtype=torch.float32, device=dev)

    def forward(self, link_pose, means3D, rotations):
        """link_pose: [E,L,4,4] float32 FK poses of this frame; means3D [E,P,3] / rotations [E,P,4]: rows
        first.. are overwritten with the re-posed scan (positions, normalised quaternions)."""
        assert link_pose.is_contiguous() and link_pose.numel() == self.E * self.L * 16 and link_pose.dtype == torch.float32
        assert means3D.is_contiguous() and rotations.is_contiguous()
        assert means3D.numel() == self.E * self.P * 3 and rotations.numel() == self.E * self.P * 4
        a = _lib.LinksArgs()
        a.E, a.L, a.P, a.first, a.n_robot = self.E, self.L, self.P, self.first, self.n_robot
        a.link_id, a.rest_means, a.rest_quats = _ptr(self.link_id), _ptr(self.rest_means), _ptr(self.rest_quats)
        a.link_pose, a.link_offset, a.rest_inv = _ptr(link_pose), _ptr(self.link_offset), _ptr(self.rest_inv)
        a.means3D, a.rotations, a.link_scratch = _ptr(means3D), _ptr(rotations), _ptr(self.scratch)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_links_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "r2s_links_forward")
        return means3D, rotations


def transform_gs(points, quats, total_mask, link_id_list, link_pose, base_pose, link_offset, device="cuda"):
    """One environment, reference argument meaning: `total_mask[g]` is the link index of Gaussian g
    (robot_pc_transformations.py:44-45), `link_id_list` the indices that move (:35), `link_pose` /
    `base_pose` / `link_offset` the 4x4 matrices of those links in list order.  Returns (points', quats')
    in the original row order, quaternions normalised as gs_renderer.py:905 leaves them."""
    dev = torch.device(device)
    total_mask = torch.as_tensor(np.asarray(total_mask)).to(torch.int64).reshape(-1)
    slot = torch.full((int(total_mask.max()) + 2,), -1, dtype=torch.int32)
    for s, i in enumerate(link_id_list):
        slot[int(i)] = s
    link_id = slot[total_mask.clamp(min=0)]
    link_id[total_mask < 0] = -1
    n = total_mask.numel()
    lt = BatchedLinkTransform(1, n, 0, link_id, points, quats, link_offset, base_pose, device=dev)
    pose = torch.as_tensor(np.asarray(link_pose), dtype=torch.float32).reshape(1, lt.L, 4, 4).to(dev).contiguous()
    out_p = torch.empty((1, n, 3), dtype=torch.float32, device=dev)
    out_q = torch.empty((1, n, 4), dtype=torch.float32, device=dev)
    lt.forward(pose, out_p, out_q)
    return out_p[0], out_q[0]
