"""Host side of the on-device success metrics over include/r2s_metrics.h.

`BatchedSuccess` keeps, for E environments of one task, what the reference's offline scripts derive from the
per-step pickle files (experiments/utils/calculate_success_{T,rope,sloth}.py): the per-frame test value, the
count of frames that passed since the task's start frame, and the episode's success flag -- plus an optional
ring buffer of the packed particle positions (the `x` of the pickled state, experiments/eval_policy.py:207-213).
One kernel per frame on the current stream; nothing is read back until `result()`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

TASKS = {"pusht": 0, "rope": 1, "sloth": 2}
START_FRAME = {"pusht": 1700, "rope": 800, "sloth": 350}   # calculate_success_T.py:66, _rope.py:196, _sloth.py:197
THRESHOLD = {"pusht": 0.002, "rope": 100.0, "sloth": 3050.0}
NEED_FRAMES = 30


def rope_box():
    """The routing box of is_rope_success (calculate_success_rope.py:152-160): (min_xyz, max_xyz), float64."""
    c = np.array([0.62, 0.05, 0.0])
    lo, hi = c.copy(), c.copy()
    lo[0] -= 0.035 / 2
    hi[0] += 0.035 / 2
    lo[1] -= 0.035 / 2
    hi[1] += 0.035 / 2
    lo[2] -= 0.0
    hi[2] += 0.03
    return lo, hi


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedSuccess:
    def __init__(self, task, E, N, *, target=None, springs=None, box=None, obb=None, start_frame=None,
                 need_frames=NEED_FRAMES, threshold=None, shift=None, ring_slots=0, device="cuda"):
        """task: 'pusht' (target: (N,3) positions), 'rope' (springs: (S,2); box: (min_xyz, max_xyz), default the
        reference's), 'sloth' (obb: (center (3,), R (3,3), extent (3,)) of the container, already scaled by 1.05).
        shift: (3,) added to the positions first (-global_translation).  ring_slots: frames of packed positions
        kept on the device (0 = none)."""
        if task not in TASKS:
            raise ValueError(f"unknown task {task!r}")
        self.device = dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.R2SError("BatchedSuccess needs a CUDA device: there is no CPU path")
        self.lib = _lib.load()
        self.task, self.E, self.N = task, int(E), int(N)
        self.start_frame = START_FRAME[task] if start_frame is None else int(start_frame)
        self.need_frames, self.threshold = int(need_frames), float(THRESHOLD[task] if threshold is None else threshold)
        self.box = np.zeros(15, np.float64)
        self.target = self.springs = None
        self.S = 0
        if task == "pusht":
            t = np.asarray(target, np.float32).reshape(-1, 3)
            if len(t) != self.N:
                raise ValueError("target must hold one position per particle")    # calculate_success_T.py:25
            self.target = torch.as_tensor(t).to(dev).contiguous()
        elif task == "rope":
            s = np.asarray(springs, np.int64).reshape(-1, 2)
            if s.min() < 0 or s.max() >= self.N:
                raise ValueError("springs contain out-of-range vertex indices.")  # calculate_success_rope.py:109
            self.springs, self.S = torch.as_tensor(s.astype(np.int32)).to(dev).contiguous(), len(s)
            lo, hi = rope_box() if box is None else box
            if np.any(np.asarray(lo) > np.asarray(hi)):
                raise ValueError("bbox min must be <= max component-wise.")       # calculate_success_rope.py:35
            self.box[:3], self.box[3:6] = lo, hi
        else:
            c, R, ext = obb
            self.box[:3], self.box[3:12], self.box[12:15] = c, np.asarray(R, np.float64).reshape(9), ext
        self.shift = None if shift is None else torch.as_tensor(np.asarray(shift, np.float32).reshape(3)).to(dev)
        self.value = torch.zeros((self.E, 2), dtype=torch.float32, device=dev)
        self.passed = torch.zeros((self.E,), dtype=torch.int32, device=dev)
        self.hits = torch.zeros((self.E,), dtype=torch.int32, device=dev)
        self.success = torch.zeros((self.E,), dtype=torch.int32, device=dev)
        self.ring_slots = int(ring_slots)
        self.ring = torch.zeros((self.ring_slots, self.E, self.N, 3), dtype=torch.float32, device=dev) if ring_slots else None
        self.frame = 0

    def reset(self):
        self.hits.zero_(); self.success.zero_(); self.passed.zero_()
        self.frame = 0

    def update(self, x4, frame=None):
        """x4: [E,N,4] float32 device tensor (the physics handle's positions).  `frame` defaults to an internal
        counter (the pickle file number of the reference's episode)."""
        if x4.dtype != torch.float32 or not x4.is_contiguous() or x4.numel() != self.E * self.N * 4:
            raise ValueError(f"x4: expected a contiguous float32 [{self.E},{self.N},4] tensor")
        f = self.frame if frame is None else int(frame)
        a = _lib.SuccessArgs()
        a.E, a.N, a.S, a.task, a.frame = self.E, self.N, self.S, TASKS[self.task], f
        a.start_frame, a.need_frames, a.ring_slots = self.start_frame, self.need_frames, self.ring_slots
        a.x4, a.shift, a.target, a.springs = _ptr(x4), _ptr(self.shift), _ptr(self.target), _ptr(self.springs)
        for k in range(15):
            a.box[k] = float(self.box[k])
        a.threshold = self.threshold
        a.value, a.passed, a.hits, a.success, a.ring = (_ptr(self.value), _ptr(self.passed), _ptr(self.hits),
                                                        _ptr(self.success), _ptr(self.ring))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.r2s_success_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "r2s_success_forward")
        self.frame = f + 1

    def result(self):
        """(success (E,) bool, hits (E,) int) on the host -- the one synchronising read."""
        return self.success.bool().cpu().numpy(), self.hits.cpu().numpy()
