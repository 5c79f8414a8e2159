"""GPU parity of r2s_success_forward (through the C ABI) against the golden vectors made from the reference's
calculate_success_T.py / calculate_success_rope.py and against oracle/metrics_ref.py.  Counts and flags are
exact; the push-T mse within 1e-6 relative (summation order of the mean)."""
import os

import numpy as np
import pytest

from real2sim_eval_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def x4_of(x):
    import torch
    x = np.asarray(x, np.float32)
    out = torch.zeros(x.shape[:-1] + (4,), device="cuda")
    out[..., :3] = torch.tensor(x, device="cuda")
    out[..., 3] = 123.0          # the pad lane must be ignored
    return out.contiguous()


def test_pusht_matches_reference_golden():
    from real2sim_eval_b200.metrics import BatchedSuccess
    d = np.load(os.path.join(G, "metrics_pusht.npz"))
    K, N = d["x"].shape[:2]
    m = BatchedSuccess("pusht", K, N, target=d["target"], start_frame=0)
    m.update(x4_of(d["x"]))
    assert np.array_equal(m.passed.cpu().numpy().astype(bool), d["ref_pass"])
    mse = m.value[:, 0].cpu().numpy()
    assert np.all(np.abs(mse - d["ref_mse"]) <= 1e-6 * d["ref_mse"] + 1e-12)


def test_rope_counts_match_reference_golden_exactly():
    from real2sim_eval_b200.metrics import BatchedSuccess
    d = np.load(os.path.join(G, "metrics_rope.npz"))
    K, N = d["x"].shape[:2]
    m = BatchedSuccess("rope", K, N, springs=d["springs"], start_frame=0)
    m.update(x4_of(d["x"]))
    assert np.array_equal(m.value.cpu().numpy().astype(np.int32), d["ref_counts"])
    assert np.array_equal(m.passed.cpu().numpy().astype(bool), d["ref_pass"])


def test_sloth_obb_shift_ring_and_episode_rule_match_the_oracle():
    import torch
    from oracle import metrics_ref
    from real2sim_eval_b200.metrics import BatchedSuccess
    sloth = synth.make_sloth()
    E, N = 4, sloth.N
    rng = np.random.default_rng(2)
    R = synth._rot_from_rotvec([0.1, -0.2, 0.7])
    obb = (np.array([0.02, -0.01, 0.12]), R, np.array([0.2, 0.13, 0.27]) * 1.05)
    shift = np.float32([0.01, -0.02, 0.0])
    thr = 0.6 * N
    m = BatchedSuccess("sloth", E, N, obb=obb, shift=shift, start_frame=3, need_frames=4, threshold=thr, ring_slots=5)
    passed_hist = [[] for _ in range(E)]
    frames = 12
    states = []
    for f in range(frames):
        x = np.stack([sloth.x + rng.normal(0, 0.01, 3) + (0.08 if (f + e) % 5 == 0 else 0.0) for e in range(E)]).astype(np.float32)
        states.append(x)
        m.update(x4_of(x))
        for e in range(E):
            world = x[e] + shift
            ok, c, _ = metrics_ref.frame_test("sloth", world, obb=obb)
            ok = c >= thr
            passed_hist[e].append(ok)
            assert int(m.value[e, 0]) == int(c) and bool(m.passed[e]) == ok, (f, e)
    success, hits = m.result()
    for e in range(E):
        h, s = metrics_ref.episode_rule(passed_hist[e], 3, 4)[-1]
        assert (int(hits[e]), bool(success[e])) == (h, s)
    assert success.any()
    # ring: slot f % 5 holds frame f's world-frame positions, packed
    for f in range(frames - 5, frames):
        assert np.array_equal(m.ring[f % 5].cpu().numpy(), states[f] + shift)


def test_argument_checks():
    from real2sim_eval_b200 import _lib
    from real2sim_eval_b200.metrics import BatchedSuccess
    with pytest.raises(ValueError, match="out-of-range"):
        BatchedSuccess("rope", 1, 4, springs=[[0, 4]])
    with pytest.raises(ValueError, match="one position per particle"):
        BatchedSuccess("pusht", 1, 4, target=np.zeros((3, 3)))
    with pytest.raises(_lib.R2SError, match="no CPU path"):
        BatchedSuccess("pusht", 1, 4, target=np.zeros((4, 3)), device="cpu")
    import torch
    m = BatchedSuccess("pusht", 2, 4, target=np.zeros((4, 3)))
    with pytest.raises(ValueError):
        m.update(torch.zeros((2, 4, 3), device="cuda"))
