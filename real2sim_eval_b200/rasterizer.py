"""Host side of the Gaussian-splat forward rasterizer: the reference's Python API
over the C ABI of include/r2s_raster.h.

`GaussianRasterizationSettings` / `GaussianRasterizer` / `rasterize_gaussians` keep
the names, argument meaning and error behaviour of
third-party/diff-gaussian-rasterization-w-depth/diff_gaussian_rasterization/__init__.py
(:135-147, :149-198, :17-38) so sim/renderer/gs_renderer.py:941-1041 and
sim/utils/gs/transform_utils.py:17-30 run unchanged (see compat/).  Forward only:
every reference call site is under torch.no_grad (gs_renderer.py:923,952,1018).

`BatchedRasterizer` is the B200-first entry: B independent (scene, camera) views
per enqueue, caller-visible workspace, no host synchronisation.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    z_threshold: float


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    return t.contiguous()


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class BatchedRasterizer:
    """B views per call over one caller-owned workspace (grows on demand)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.R2SError("BatchedRasterizer needs a CUDA device: there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.ws: Optional[torch.Tensor] = None
        self.max_instances = 0
        self.shape = None
        self.layout = None
        # sticky device counter: +1 for every forward whose instance lists overflowed the workspace (that frame
        # holds background only); read with overflows() when convenient -- no per-frame synchronisation
        self.overflow_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._composited = None     # event after the last compositing kernel launched on a separate stream

    # ---- workspace
    def reserve(self, B, P, W, H, max_instances):
        key = (B, P, W, H, int(max_instances))
        if self.shape == key and self.ws is not None:
            return
        L = _lib.RasterLayout()
        _lib.check(self.lib.r2s_raster_workspace_layout(B, P, W, H, int(max_instances), C.byref(L)), "workspace_layout")
        if self.ws is None or self.ws.numel() < L.total:
            self.ws = None
            self.ws = torch.empty(int(L.total), dtype=torch.uint8, device=self.device)
        self.layout, self.shape, self.max_instances = L, key, int(max_instances)

    def forward(self, means3D, opacities, *, viewmatrix, projmatrix, campos, bg, W, H, tanfovx, tanfovy,
                shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, sh_degree=0,
                scale_modifier=1.0, z_threshold=0.05, prefiltered=False, views_per_scene=1,
                max_instances=None, out_color=None, out_depth=None, radii=None, want_radii=True, out_rgb8=None,
                fast=False, composite_stream=None):
        """tanfovx / tanfovy: floats shared by every view, or per-view sequences / tensors of length B (the
        reference builds one settings tuple per camera, transform_utils.py:17-30).  fast=True selects the
        ex2.approx compositing variant (within the 1e-4 relative contract, not bit-identical)."""
        dev = self.device
        means3D = _f32c(means3D, dev)
        viewmatrix = _f32c(viewmatrix, dev).reshape(-1, 16)
        projmatrix = _f32c(projmatrix, dev).reshape(-1, 16)
        B = viewmatrix.shape[0]
        campos = _f32c(campos, dev).reshape(B, 3)
        bg = _f32c(bg, dev).reshape(3)
        n_scenes = B // views_per_scene
        P = 0 if means3D is None else means3D.reshape(n_scenes, -1, 3).shape[1]
        shs, colors_precomp = _f32c(shs, dev), _f32c(colors_precomp, dev)
        scales, rotations, cov3D_precomp = _f32c(scales, dev), _f32c(rotations, dev), _f32c(cov3D_precomp, dev)
        opacities = _f32c(opacities, dev)
        if P > 0:
            if (shs is None) == (colors_precomp is None):
                raise Exception("Please provide excatly one of either SHs or precomputed colors!")
            if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                    ((scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        M = 0 if shs is None else shs.reshape(n_scenes, P, -1, 3).shape[2]
        if max_instances is None:
            max_instances = max(self.max_instances, 8 * B * max(P, 1) + 65536)
        self.reserve(B, P, W, H, max_instances)
        if out_color is None:
            out_color = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        if out_depth is None:
            out_depth = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        if radii is None and want_radii:
            radii = torch.empty((B, P), dtype=torch.int32, device=dev)
        a = _lib.RasterArgs()
        a.B, a.views_per_scene, a.P, a.D, a.M, a.W, a.H = B, views_per_scene, P, int(sh_degree), M, W, H
        a.prefiltered = int(bool(prefiltered))
        tanfov_views = None
        if not isinstance(tanfovx, (int, float)) or not isinstance(tanfovy, (int, float)):
            tx = torch.as_tensor(tanfovx, dtype=torch.float32, device=dev).reshape(-1).expand(B) if not isinstance(tanfovx, (int, float)) \
                else torch.full((B,), float(tanfovx), dtype=torch.float32, device=dev)
            ty = torch.as_tensor(tanfovy, dtype=torch.float32, device=dev).reshape(-1).expand(B) if not isinstance(tanfovy, (int, float)) \
                else torch.full((B,), float(tanfovy), dtype=torch.float32, device=dev)
            tanfov_views = torch.stack([tx, ty], 1).contiguous()
            tanfovx, tanfovy = 1.0, 1.0          # unused when the per-view table is given
        a.scale_modifier, a.tanfovx, a.tanfovy, a.z_threshold = scale_modifier, float(tanfovx), float(tanfovy), z_threshold
        a.tanfov_views, a.overflow_count = _ptr(tanfov_views), _ptr(self.overflow_count)
        a.composite_mode = 1 if fast else 0
        cur = torch.cuda.current_stream(dev)
        if composite_stream is not None:
            # the compositing kernel of the previous forward on this workspace may still be reading the lists and
            # records this forward rewrites: order this forward's binning after it
            if self._composited is not None:
                cur.wait_event(self._composited)
            a.composite_stream = C.c_void_p(composite_stream.cuda_stream)
        a.means3D, a.scales, a.rotations, a.opacities = _ptr(means3D), _ptr(scales), _ptr(rotations), _ptr(opacities)
        a.shs, a.colors_precomp, a.cov3D_precomp = _ptr(shs), _ptr(colors_precomp), _ptr(cov3D_precomp)
        a.viewmatrix, a.projmatrix, a.campos, a.bg = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), _ptr(bg)
        a.out_color, a.out_depth, a.radii = _ptr(out_color), _ptr(out_depth), _ptr(radii)
        if out_rgb8 is not None:   # optional [B,H,W,3] uint8 image (the reference's host-side format)
            if out_rgb8.dtype != torch.uint8 or not out_rgb8.is_contiguous() or out_rgb8.numel() != B * H * W * 3:
                raise ValueError("out_rgb8 must be a contiguous uint8 tensor of shape [B,H,W,3]")
            a.out_rgb8 = _ptr(out_rgb8)
        a.workspace, a.workspace_bytes, a.max_instances = _ptr(self.ws), self.ws.numel(), self.max_instances
        with torch.cuda.device(dev):
            _lib.check(self.lib.r2s_raster_forward(C.byref(a), _stream_ptr(dev)), "r2s_raster_forward")
        if composite_stream is not None:
            self._composited = torch.cuda.Event()
            self._composited.record(composite_stream)
        self._keep = (means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp, viewmatrix,
                      projmatrix, campos, bg, tanfov_views)
        return out_color, radii, out_depth

    def status(self):
        """(num_rendered, overflow) of the last forward; synchronises the stream."""
        n, o = C.c_int64(0), C.c_int32(0)
        _lib.check(self.lib.r2s_raster_status(_ptr(self.ws), _stream_ptr(self.device), C.byref(n), C.byref(o)),
                   "r2s_raster_status")
        return int(n.value), bool(o.value)

    def overflows(self, reset: bool = False) -> int:
        """Number of forwards since the last reset whose instance lists overflowed max_instances (those frames
        hold background only); synchronises the stream."""
        n = int(self.overflow_count.item())
        if reset:
            self.overflow_count.zero_()
        return n

    def intermediates(self):
        """Typed torch views of the workspace arrays (parity tests, accounting)."""
        B, P, W, H, cap = self.shape
        L, ws = self.layout, self.ws
        T = L.super_x * L.super_y   # lists are kept per super-tile (4x4 tiles)

        def view(off, nbytes, dtype):
            return ws[off:off + nbytes].view(dtype)

        return dict(
            depths=view(L.depths, 4 * B * P, torch.float32).view(B, P),
            radii=view(L.radii, 4 * B * P, torch.int32).view(B, P),
            tiles_touched=view(L.tiles_touched, 4 * B * P, torch.int32).view(B, P),
            rec_a=view(L.rec_a, 32 * B * P, torch.float32).view(B, P, 8)[..., :4],   # 32-byte records {rec_a, rec_b}
            rec_b=view(L.rec_a, 32 * B * P, torch.float32).view(B, P, 8)[..., 4:],
            rec_c=view(L.rec_c, 4 * B * P, torch.float32).view(B, P),
            rects=view(L.rects, 4 * B * P, torch.int32).view(B, P),
            super_count=view(L.tile_count, 4 * B * T, torch.int32).view(B, T),
            super_offset=view(L.tile_offset, 4 * (B * T + 1), torch.int32),
            keys=view(L.keys, 8 * cap, torch.int64),
            sorted_rect=view(L.sorted_rect, 4 * cap, torch.int32),   # filled only when id_shift == 0
            # low word of a key: id << id_shift | rectangle local to the super-tile (x0 | y0 << 3 | x1 << 6 | y1 << 9)
            # when P <= 2^20 (id_shift = 12); else the id alone, with the rectangles in sorted_rect
            id_shift=12 if P <= (1 << 20) else 0,
            tiles=(L.tiles_x, L.tiles_y),
            supers=(L.super_x, L.super_y),
        )


_default_rasterizers: dict = {}


def _default(device) -> BatchedRasterizer:
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.R2SError("GaussianRasterizer inputs must be CUDA tensors: there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _default_rasterizers:
        _default_rasterizers[idx] = BatchedRasterizer(torch.device("cuda", idx))
    return _default_rasterizers[idx]


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    """diff_gaussian_rasterization.rasterize_gaussians (__init__.py:17-38), forward only.
    Returns (color (3,H,W), radii (P,) int32, depth (1,H,W))."""
    rs = raster_settings
    r = _default(means3D.device)
    empty = lambda t: None if t is None or t.numel() == 0 else t
    kw = dict(viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix, campos=rs.campos, bg=rs.bg, W=int(rs.image_width),
              H=int(rs.image_height), tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), shs=empty(sh),
              colors_precomp=empty(colors_precomp), scales=empty(scales), rotations=empty(rotations),
              cov3D_precomp=empty(cov3Ds_precomp), sh_degree=int(rs.sh_degree),
              scale_modifier=float(rs.scale_modifier), z_threshold=float(rs.z_threshold),
              prefiltered=bool(rs.prefiltered))
    with torch.no_grad():
        color, radii, depth = r.forward(means3D, opacities, **kw)
        total, overflow = r.status()  # one sync, as the reference has (rasterizer_impl.cu:283-284)
        if overflow:
            color, radii, depth = r.forward(means3D, opacities, max_instances=int(total * 1.25) + 4096, **kw)
    return color[0], radii[0], depth[0]


class GaussianRasterizer(torch.nn.Module):
    """Drop-in for diff_gaussian_rasterization.GaussianRasterizer (__init__.py:149-198)."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions, positions.device)
            if pos.device.type != "cuda":
                raise _lib.R2SError("markVisible needs CUDA tensors: there is no CPU path")
            out = torch.empty(pos.shape[0], dtype=torch.bool, device=pos.device)
            view = _f32c(rs.viewmatrix, pos.device)
            proj = _f32c(rs.projmatrix, pos.device)
            with torch.cuda.device(pos.device):
                _lib.check(_lib.load().r2s_mark_visible(pos.shape[0], _ptr(pos), _ptr(view), _ptr(proj), _ptr(out),
                                                        _stream_ptr(pos.device)), "r2s_mark_visible")
        return out

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs)
