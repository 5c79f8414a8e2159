"""ctypes front-end of oracle/_ref/libref_raster.so: the reference's own CUDA
rasterizer (compiled unmodified by oracle/Makefile), driven with torch buffers.
TEST INFRASTRUCTURE (also the reference-rasterizer timing baseline in bench.py)."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_raster.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        _lib.ref_raster_forward.argtypes = [i, i, i, vp, i, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, i, f,
                                            vp, vp, vp]
        _lib.ref_raster_forward.restype = i
        if hasattr(_lib, "ref_raster_lists"):
            _lib.ref_raster_lists.argtypes = [i, i, i, vp, vp]
            _lib.ref_raster_lists.restype = i
    return _lib


def forward_torch(t, view, proj, campos, bg, W, H, tanfovx, tanfovy, sh_degree, z_threshold, out_color, out_depth,
                  radii, scale_modifier=1.0):
    """All tensors CUDA float32 contiguous; returns num_rendered (host int; the reference syncs)."""
    p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
    P = t["means3D"].shape[0]
    shs = t.get("shs")
    M = 0 if shs is None else shs.shape[1]
    n = lib().ref_raster_forward(P, int(sh_degree), M, p(bg), W, H, p(t["means3D"]), p(shs), p(t.get("colors_precomp")),
                                 p(t["opacities"]), p(t.get("scales")), scale_modifier, p(t.get("rotations")),
                                 p(t.get("cov3D_precomp")), p(view), p(proj), p(campos), tanfovx, tanfovy, 0,
                                 z_threshold, p(out_color), p(out_depth), p(radii))
    if n < 0:
        raise RuntimeError(f"reference rasterizer CUDA error {-n}")
    return n


def forward(g, cam, sh_degree=0, bg=(0.0, 0.0, 0.0)):
    import torch
    t = {k: torch.tensor(np.ascontiguousarray(v, dtype=np.float32)).cuda() for k, v in g.items()}
    P = t["means3D"].shape[0]
    color = torch.zeros((3, cam.H, cam.W), device="cuda")
    depth = torch.zeros((1, cam.H, cam.W), device="cuda")
    radii = torch.zeros(P, dtype=torch.int32, device="cuda")
    tt = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    torch.cuda.synchronize()
    n = forward_torch(t, tt(cam.view), tt(cam.proj), tt(cam.campos), tt(bg), cam.W, cam.H, cam.tanfovx, cam.tanfovy,
                      sh_degree, cam.z_threshold, color, depth, radii)
    torch.cuda.synchronize()
    return color.cpu().numpy(), radii.cpu().numpy(), depth.cpu().numpy(), n


def lists(num_rendered, W, H):
    """The reference's own sorted point_list [num_rendered] and tile ranges [tiles, 2] of the last forward()
    (read out of its binning / image buffers with its own fromChunk layout)."""
    import torch
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    pl = torch.zeros(max(num_rendered, 1), dtype=torch.int32, device="cuda")
    rg = torch.zeros((tiles, 2), dtype=torch.int32, device="cuda")
    rc = lib().ref_raster_lists(int(num_rendered), W, H, C.c_void_p(pl.data_ptr()), C.c_void_p(rg.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"ref_raster_lists failed: {rc}")
    torch.cuda.synchronize()
    return pl[:num_rendered].cpu().numpy().view(np.uint32), rg.cpu().numpy().view(np.uint32)
