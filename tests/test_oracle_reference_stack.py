"""CPU: the COMPLETE reference physics stack -- sim/physics/phystwin.py (SpringMassDynamicsModule.__init__ / .step)
on top of sim/physics/spring_mass_warp.py, both unmodified, the latter executed under oracle/warp_exec.py -- against
the composition of this repository's oracles (oracle/eef_ref.py + oracle/physics_ref.c) over three closed-loop frames
(grasp hysteresis fed by the previous frame's finger forces).  Needs the reference tree; skipped on the GPU box.
The same harness (tests/ref_harness.py) with backend "cuda" is tests/test_gpu_dropin.py."""
import pytest

import ref_harness

ROOT = ref_harness.reference_root()
pytestmark = pytest.mark.skipif(ROOT is None, reason="reference tree not available")


@pytest.mark.parametrize("use_pusher", [False, True])
def test_full_reference_stack_matches_the_oracle_composition(tmp_path, use_pusher):
    pt = ref_harness.load_phystwin(ROOT, "interp")
    mod, errs, o = ref_harness.drive_and_compare(pt, "cpu", use_pusher, tmp_path, frames=2)
    assert type(mod.simulator).__module__ == "_ref_spring_mass_warp"
    for f, dx in enumerate(errs):
        assert dx.max() <= 1e-6, f"frame {f}: |dx|max = {dx.max()}"
