"""real2sim_eval_b200 -- B200-native (sm_100a) inner loop for kywind/real2sim-eval:
the PhysTwin spring-mass substep loop and the Gaussian-splat forward rasterizer,
hand-written CUDA behind a C ABI (include/r2s_*.h), with host classes that keep the
reference's Python call signatures.  See DESIGN.md / INTEGRATION.md.

Importing the package does not touch the GPU; the CUDA library is loaded on first
use and there is no CPU fallback.
"""
from . import synth  # noqa: F401  (pure numpy)

__all__ = ["synth", "load_library", "GaussianRasterizer", "GaussianRasterizationSettings", "BatchedRasterizer",
           "SpringMassSystemWarp", "BatchedSpringMass"]


def load_library():
    from . import _lib
    return _lib.load()


def __getattr__(name):  # lazy: torch is only imported when a device class is requested
    if name in ("GaussianRasterizer", "GaussianRasterizationSettings", "BatchedRasterizer", "rasterize_gaussians"):
        from . import rasterizer
        return getattr(rasterizer, name)
    if name in ("SpringMassSystemWarp", "BatchedSpringMass"):
        from . import physics
        return getattr(physics, name)
    raise AttributeError(name)
