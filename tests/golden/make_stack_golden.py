"""Generates tests/golden/stack_{gripper,pusher}.npz from the COMPLETE reference physics stack on the CPU:
/root/reference/sim/physics/phystwin.py (SpringMassDynamicsModule.__init__ / .step, unmodified) on top of
/root/reference/sim/physics/spring_mass_warp.py (unmodified, executed under oracle/warp_exec.py), driven over three
closed-loop frames by tests/ref_harness.py (which says what is faked and why: kornia's axis-angle conversion,
Open3D's KD-tree, the sapien-backed robot sampler, synthetic checkpoint files).  Stored: the particle positions the
reference returns after every frame, its `current_openness` / `grasped`, and digests of its spring graph and rest lengths; the inputs are rebuilt
by the tests from ref_harness.stack_scenario (same seeded builders).
    python tests/golden/make_stack_golden.py          # needs /root/reference, ~3 min"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)
import ref_harness      # noqa: E402


def main():
    root = ref_harness.reference_root()
    assert root, "needs the reference tree"
    pt = ref_harness.load_phystwin(root, "interp")
    for use_pusher in (False, True):
        with tempfile.TemporaryDirectory() as tmp:
            mod, errs, o = ref_harness.drive_and_compare(pt, "cpu", use_pusher, tmp, frames=3)
        name = "pusher" if use_pusher else "gripper"
        np.savez_compressed(os.path.join(HERE, f"stack_{name}.npz"), x=np.stack(mod._r2s_x_frames),
                            current_openness=np.float64(mod.current_openness), grasped=np.bool_(mod.grasped),
                            springs_sha=np.array(hashlib.sha256(np.ascontiguousarray(mod.init_springs.numpy()).tobytes()).hexdigest()),
                            rest_sha=np.array(hashlib.sha256(np.ascontiguousarray(mod.init_rest_lengths.numpy()).tobytes()).hexdigest()),
                            oracle_err=np.array([e.max() for e in errs]))
        print(name, [float(e.max()) for e in errs], mod.current_openness, mod.grasped)


if __name__ == "__main__":
    main()
