"""Drop-in module name for the reference's rasterizer package
(third-party/diff-gaussian-rasterization-w-depth/diff_gaussian_rasterization/__init__.py)."""
from real2sim_eval_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)
