"""ctypes front-end of oracle/raster_ref.c (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib_path

_f = C.c_float
_i = C.c_int32
_pf = C.POINTER(C.c_float)
_pi = C.POINTER(C.c_int32)
_pu = C.POINTER(C.c_uint32)
_pu64 = C.POINTER(C.c_uint64)


class _Raster(C.Structure):
    _fields_ = [
        ("P", _i), ("D", _i), ("M", _i), ("W", _i), ("H", _i), ("prefiltered", _i),
        ("scale_modifier", _f), ("tanfovx", _f), ("tanfovy", _f), ("z_threshold", _f),
        ("means3D", _pf), ("scales", _pf), ("rotations", _pf), ("opacities", _pf), ("shs", _pf),
        ("colors_precomp", _pf), ("cov3D_precomp", _pf), ("viewmatrix", _pf), ("projmatrix", _pf),
        ("campos", _pf), ("bg", _pf),
        ("out_color", _pf), ("out_depth", _pf), ("radii", _pi),
        ("depths", _pf), ("means2D", _pf), ("cov3D", _pf), ("conic_opacity", _pf), ("rgb", _pf),
        ("tiles_touched", _pu), ("final_T", _pf), ("n_contrib", _pu), ("ranges", _pu),
        ("point_list", _pu), ("point_keys", _pu64), ("point_list_cap", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(lib_path("libraster_ref.so"))
        assert _lib.oracle_raster_struct_size() == C.sizeof(_Raster), "struct layout mismatch"
        _lib.oracle_raster_forward.argtypes = [C.POINTER(_Raster)]
        _lib.oracle_raster_forward.restype = C.c_int64
        _lib.oracle_get_higher_msb.argtypes = [C.c_uint32]
        _lib.oracle_get_higher_msb.restype = C.c_uint32
    return _lib


def _c(a, dt=np.float32):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else t()


def get_higher_msb(n: int) -> int:
    return int(lib().oracle_get_higher_msb(n))


def mark_visible(means3D, viewmatrix, projmatrix):
    m, v, p = _c(means3D), _c(viewmatrix).reshape(-1), _c(projmatrix).reshape(-1)
    out = np.zeros(len(m), dtype=np.uint8)
    lib().oracle_mark_visible(C.c_int(len(m)), _p(m, _pf), _p(v, _pf), _p(p, _pf),
                              out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.astype(bool)


def rasterize(means3D, opacities, *, viewmatrix, projmatrix, campos, bg, W, H, tanfovx, tanfovy,
              shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
              sh_degree=0, scale_modifier=1.0, z_threshold=0.05, prefiltered=False, aux=False,
              list_cap=None):
    """Same contract as the reference's GaussianRasterizer.forward
    (diff_gaussian_rasterization/__init__.py:165-198): returns color (3,H,W), radii (P,),
    depth (1,H,W); with aux=True also a dict of intermediates."""
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    a = _Raster()
    keep = []

    def own(x, dt=np.float32):
        x = _c(x, dt)
        keep.append(x)
        return x

    means3D = own(means3D).reshape(-1, 3)
    P = len(means3D)
    a.P, a.W, a.H, a.D = P, int(W), int(H), int(sh_degree)
    a.M = 0 if shs is None else int(np.asarray(shs).reshape(P, -1, 3).shape[1]) if P else 0
    a.prefiltered = int(prefiltered)
    a.scale_modifier, a.tanfovx, a.tanfovy, a.z_threshold = scale_modifier, tanfovx, tanfovy, z_threshold
    a.means3D = _p(means3D, _pf)
    a.scales, a.rotations = _p(own(scales), _pf), _p(own(rotations), _pf)
    a.opacities, a.shs = _p(own(opacities), _pf), _p(own(shs), _pf)
    a.colors_precomp, a.cov3D_precomp = _p(own(colors_precomp), _pf), _p(own(cov3D_precomp), _pf)
    a.viewmatrix, a.projmatrix = _p(own(viewmatrix).reshape(-1), _pf), _p(own(projmatrix).reshape(-1), _pf)
    a.campos, a.bg = _p(own(campos), _pf), _p(own(bg), _pf)
    color = np.zeros((3, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    radii = np.zeros(P, np.int32)
    a.out_color, a.out_depth, a.radii = _p(color, _pf), _p(depth, _pf), _p(radii, _pi)
    out = {}
    if aux:
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        cap = int(list_cap if list_cap is not None else max(1, 64 * P))
        out = dict(depths=np.zeros(P, np.float32), means2D=np.zeros((P, 2), np.float32),
                   cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
                   rgb=np.zeros((P, 3), np.float32), tiles_touched=np.zeros(P, np.uint32),
                   final_T=np.zeros((H, W), np.float32), n_contrib=np.zeros((H, W), np.uint32),
                   ranges=np.zeros((tiles, 2), np.uint32), point_list=np.zeros(cap, np.uint32),
                   point_keys=np.zeros(cap, np.uint64))
        a.depths, a.means2D, a.cov3D = _p(out["depths"], _pf), _p(out["means2D"], _pf), _p(out["cov3D"], _pf)
        a.conic_opacity, a.rgb = _p(out["conic_opacity"], _pf), _p(out["rgb"], _pf)
        a.tiles_touched, a.final_T = _p(out["tiles_touched"], _pu), _p(out["final_T"], _pf)
        a.n_contrib, a.ranges = _p(out["n_contrib"], _pu), _p(out["ranges"], _pu)
        a.point_list, a.point_keys, a.point_list_cap = _p(out["point_list"], _pu), _p(out["point_keys"], _pu64), cap
    R = lib().oracle_raster_forward(C.byref(a))
    if aux:
        if R < 0:
            raise RuntimeError("point_list capacity too small")
        out["num_rendered"] = int(R)
        out["point_list"] = out["point_list"][:R]
        out["point_keys"] = out["point_keys"][:R]
        return color, radii, depth, out
    return color, radii, depth
