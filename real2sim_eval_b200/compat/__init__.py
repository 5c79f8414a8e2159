"""Import shims that let the UNMODIFIED reference tree (sim/, experiments/) run on
this library: put this directory first on sys.path and

    from diff_gaussian_rasterization import GaussianRasterizer, GaussianRasterizationSettings
    import warp as wp        # only the five calls sim/physics/phystwin.py makes

resolve to the B200 implementations; sim/physics/spring_mass_warp.py is replaced
by real2sim_eval_b200.physics.SpringMassSystemWarp."""
import os
import sys


def install(reference_root: str | None = None) -> None:
    """Prepend the shim directory to sys.path and, if `reference_root` holds the
    reference checkout, alias sim.physics.spring_mass_warp to our class."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import types
    from .. import physics
    mod = types.ModuleType("sim.physics.spring_mass_warp")
    mod.SpringMassSystemWarp = physics.SpringMassSystemWarp
    sys.modules["sim.physics.spring_mass_warp"] = mod
    if reference_root and reference_root not in sys.path:
        sys.path.append(reference_root)
