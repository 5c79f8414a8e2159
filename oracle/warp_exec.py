"""oracle/warp_exec.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A small interpreter of the part of the `warp` (warp-lang 1.7.0, pyproject.toml:19) surface that
/root/reference/sim/physics/spring_mass_warp.py uses, so that the reference's OWN, UNMODIFIED kernel
source can be executed in this container (warp-lang is not installable offline): `load_reference()`
installs this module as `warp` in sys.modules and imports the reference file from where it lies.

How the kernels run.  The bodies of the `@wp.kernel` / `@wp.func` functions are plain Python arithmetic on
`wp.*` calls, so they are executed as Python, one thread index at a time (`wp.launch` loops `tid` over `dim`
in ascending order), on float32 numpy scalars and 3-vectors: every operation rounds to float32 exactly where
the generated CUDA does (NEP-50 scalar rules keep Python literals "weak"), without FMA contraction.  What
differs from a real Warp run and cannot matter beyond float rounding: the order of the float atomics of
`eval_springs` (ascending spring index here, unordered on the GPU), `expf`/`sqrtf` being numpy's, and
FMA contraction (nvrtc contracts, this does not).  `wp.ScopedCapture` records launches / memsets instead of
executing them (a CUDA stream capture does not execute either) and `wp.capture_launch` replays the list, so
the reference's `use_graph=True` constructor + `capture_launch` call sequence (phystwin.py:515-517) is
reproduced, including the substep index being frozen into each node.

What is NOT the reference's code: three Warp built-ins whose arithmetic lives inside warp-lang's native
library, restated here from their published source (warp/native/hashgrid.h, mesh.h, bvh.h of Warp 1.x):
  * wp.HashGrid.build / wp.hash_grid_point_id / wp.hash_grid_query      -> class HashGrid below
  * wp.Mesh / refit / wp.mesh_query_point_sign_winding_number            -> mesh_query_point_sign_winding_number
  * wp.mesh_eval_position                                                -> mesh_eval_position
Those three stay PARITY UNPINNED (candidate iteration order, closest-face ties in BVH order, the
accuracy=3.0 far-field winding approximation / the ray-parity fallback of a mesh built without
support_winding_number); everything else the goldens made with this module contain is the reference's own
arithmetic in the reference's own statement order.
"""
from __future__ import annotations

import ast
import builtins as _bi
import importlib.util
import inspect
import sys
import textwrap
import types

import numpy as np

F32 = np.float32
REF_SMW = "/root/reference/sim/physics/spring_mass_warp.py"

# ----------------------------------------------------------------------------- dtypes


class _VecType:
    """wp.vec3 / wp.vec2i: a dtype tag that is also the constructor of a value."""

    def __init__(self, name, n, scalar):
        self.name, self.n, self.scalar = name, n, scalar

    def __call__(self, *a):
        if len(a) == 0:
            return np.zeros(self.n, self.scalar)
        if len(a) == 1 and np.ndim(a[0]) == 0:
            return np.full(self.n, a[0], self.scalar)
        if len(a) == 1:
            return np.asarray(a[0], self.scalar).reshape(self.n).copy()
        assert len(a) == self.n
        return np.array(a, self.scalar)

    def __repr__(self):
        return f"wp.{self.name}"


vec3 = _VecType("vec3", 3, np.float32)
vec2i = _VecType("vec2i", 2, np.int32)
float32 = np.float32
int32 = np.int32
uint64 = np.uint64


class _BoolType:
    def __call__(self, v=False):
        return _bi.bool(v)


bool = _BoolType()          # noqa: A001  (wp.bool)
_py_bool = _bi.bool


def _np_scalar(dtype):
    if dtype is None or dtype is float or dtype is float32:
        return np.float32
    if dtype is int or dtype is int32:
        return np.int32
    if dtype is bool or dtype is _py_bool:
        return np.bool_
    if isinstance(dtype, _VecType):
        return dtype.scalar
    raise TypeError(f"unsupported dtype {dtype!r}")


# ----------------------------------------------------------------------------- arrays
_capture = None            # list of thunks while a ScopedCapture is open


def _do(thunk):
    if _capture is not None:
        _capture.append(thunk)
    else:
        thunk()


class array:
    """wp.array: `array(dtype=...)` (annotation), or `array(data, dtype=..., device=...)` (value).
    `data` holds the numpy buffer, with a trailing axis of length n for vector dtypes."""

    def __init__(self, data=None, dtype=None, shape=None, device=None, requires_grad=False, ndim=None, **_):
        self.dtype = dtype
        self.requires_grad = requires_grad
        self.device = device
        self.vec = dtype.n if isinstance(dtype, _VecType) else 0
        self.ndim_annot = ndim
        self.data = None
        if data is not None:
            a = np.asarray(data)
            a = np.ascontiguousarray(a, dtype=_np_scalar(dtype))
            if self.vec and (a.ndim == 0 or a.shape[-1] != self.vec):
                a = a.reshape(-1, self.vec)
            self.data = a
        elif shape is not None:
            shape = (shape,) if np.ndim(shape) == 0 else tuple(shape)
            self.data = np.zeros(shape + ((self.vec,) if self.vec else ()), _np_scalar(dtype))

    # -- the attributes / methods the reference touches
    @property
    def shape(self):
        return self.data.shape[:-1] if self.vec else self.data.shape

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]

    def numpy(self):
        return self.data

    def zero_(self):
        _do(lambda: self.data.fill(0))

    def __getitem__(self, idx):
        r = self.data[idx]
        # one element of a vector array: value semantics (a register copy in the generated code)
        if isinstance(r, np.ndarray) and self.vec and r.ndim == 1:
            return r.copy()
        return r

    def __setitem__(self, idx, val):
        self.data[idx] = val


def array2d(dtype=None, **kw):
    return array(dtype=dtype, ndim=2, **kw)


def zeros(shape, dtype=float32, device=None, requires_grad=False, **_):
    return array(dtype=dtype, shape=shape, device=device, requires_grad=requires_grad)


def zeros_like(a, requires_grad=False, **_):
    return array(dtype=a.dtype, shape=a.shape, device=a.device, requires_grad=requires_grad)


def from_torch(t, dtype=None, requires_grad=False, **_):
    """Shares memory with the (CPU, contiguous) tensor, as wp.from_torch does on the device."""
    a = t.detach().numpy()
    if dtype is None:
        dtype = {np.dtype(np.float32): float32, np.dtype(np.int32): int32}[a.dtype]
    want = _np_scalar(dtype)
    if a.dtype != want or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=want)
    out = array(dtype=dtype, requires_grad=requires_grad)
    out.data = a
    return out


def to_torch(a, requires_grad=None):
    import torch
    return torch.from_numpy(a.data)


def _as_array(value, annot):
    """Kernel argument packing: wp arrays pass through; torch tensors / numpy arrays are viewed with the
    annotated dtype (Warp accepts __cuda_array_interface__ objects for array parameters), a missing outer axis
    of an array2d parameter is a unit axis (copy_2dvec3 on a (S,3) centre table, SMW:785-790)."""
    if isinstance(value, array):
        return value
    if hasattr(value, "detach"):
        value = value.detach().numpy()
    out = array(dtype=annot.dtype)
    a = np.ascontiguousarray(np.asarray(value), dtype=_np_scalar(annot.dtype))
    if out.vec and a.shape[-1] != out.vec:
        a = a.reshape(-1, out.vec)
    want_nd = (annot.ndim_annot or 1) + (1 if out.vec else 0)
    while a.ndim < want_nd:
        a = np.expand_dims(a, a.ndim - (1 if out.vec else 0))
    out.data = a
    return out


# ----------------------------------------------------------------------------- kernels
_tid = None


class Kernel:
    def __init__(self, fn):
        self.fn = fn
        self.__name__ = fn.__name__
        self.sig = list(inspect.signature(fn).parameters.values())
        # how many indices `wp.tid()` yields in this kernel (i, j = wp.tid())
        self.tid_arity = 1
        tree = ast.parse(textwrap.dedent(inspect.getsource(fn)))
        for node in ast.walk(tree):
            if (isinstance(node, ast.Assign) and isinstance(node.value, ast.Call)
                    and isinstance(node.value.func, ast.Attribute) and node.value.func.attr == "tid"
                    and isinstance(node.targets[0], ast.Tuple)):
                self.tid_arity = len(node.targets[0].elts)

    def pack(self, args):
        assert len(args) == len(self.sig), f"{self.__name__}: {len(args)} args for {len(self.sig)} params"
        out = []
        for p, v in zip(self.sig, args):
            an = p.annotation
            if isinstance(an, array):
                out.append(_as_array(v, an))
            elif an is float or an is float32:
                out.append(F32(v))
            elif an is int or an is int32:
                out.append(int(v))
            elif an is _py_bool or an is bool:
                out.append(_py_bool(v))
            else:                      # wp.uint64 handles (grid / mesh ids)
                out.append(v)
        return out


def kernel(fn=None, **_):
    if fn is not None:
        return Kernel(fn)
    return lambda f: Kernel(f)


def func(fn):
    return fn


def tid():
    return _tid


def launch(kernel=None, dim=None, inputs=(), outputs=(), device=None, **_):
    k = kernel
    args = k.pack(list(inputs) + list(outputs))
    dims = (int(dim),) if np.ndim(dim) == 0 else tuple(int(d) for d in dim)
    dims = dims + (1,) * (k.tid_arity - len(dims))

    def run():
        global _tid
        if k.tid_arity == 1:
            for t in range(int(np.prod(dims))):
                _tid = t
                k.fn(*args)
        else:
            for t in np.ndindex(*dims):
                _tid = t[:k.tid_arity]
                k.fn(*args)
        _tid = None

    _do(run)


class ScopedTimer:
    enabled = False

    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Graph:
    def __init__(self, nodes):
        self.nodes = nodes


class ScopedCapture:
    def __init__(self, *a, **k):
        self.graph = None

    def __enter__(self):
        global _capture
        assert _capture is None
        _capture = []
        return self

    def __exit__(self, *a):
        global _capture
        self.graph = _Graph(_capture)
        _capture = None
        return False


def capture_launch(graph):
    assert _capture is None
    for node in graph.nodes:
        node()


def init():
    pass


def set_module_options(*a, **k):
    pass


config = types.SimpleNamespace(mode="release", verify_cuda=False, quiet=True)

# ----------------------------------------------------------------------------- scalar / vector built-ins


def exp(x):
    return np.exp(F32(x))


def _dot3(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def dot(a, b):
    return F32(_dot3(a, b))


def length(a):
    return np.sqrt(F32(_dot3(a, a)))


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F32)


def normalize(a):
    l = length(a)                       # warp/native/vec.h normalize: a / l if l > 0 else 0
    if l > F32(0.0):
        return a / l
    return np.zeros(3, F32)


def max(a, b):                          # noqa: A001
    return F32(a) if a > b else F32(b)


def min(a, b):                          # noqa: A001
    return F32(a) if a < b else F32(b)


def clamp(x, low, high):
    return min(max(x, low), high)


def atomic_add(arr, idx, val):
    arr.data[idx] += val


def atomic_sub(arr, idx, val):
    arr.data[idx] -= val


# ----------------------------------------------------------------------------- HashGrid (warp/native/hashgrid.h)
class HashGrid:
    """Restated: cell = int(p * cell_width_inv) per axis (C truncation), + 2^20, clamp >= 0, mod dim;
    id = cz*dx*dy + cy*dx + cx; points radix-sorted (stable) by cell id; a query walks the cell box
    [int((p-r)*inv), min(int((p+r)*inv), start+dim-1)] x fastest, then y, then z, and yields every point of
    every visited cell."""

    def __init__(self, dim_x, dim_y, dim_z, device=None):
        self.dim = (int(dim_x), int(dim_y), int(dim_z))
        self.id = self
        self.point_ids = None

    def _index(self, x, y, z):
        o = 1 << 20
        x, y, z = _bi.max(0, x + o), _bi.max(0, y + o), _bi.max(0, z + o)
        dx, dy, dz = self.dim
        return (z % dz) * (dx * dy) + (y % dy) * dx + (x % dx)

    def build(self, points, radius):
        def run():
            pts = points.data if isinstance(points, array) else np.asarray(points, F32)
            self.cell_width = F32(radius)
            self.inv = F32(1.0) / self.cell_width
            n = len(pts)
            cells = np.empty(n, np.int64)
            for i in range(n):
                p = pts[i]
                cells[i] = self._index(int(p[0] * self.inv), int(p[1] * self.inv), int(p[2] * self.inv))
            order = np.argsort(cells, kind="stable")
            self.point_ids = order.astype(np.int32)
            sc = cells[order]
            self.cell_lists = {}
            start = 0
            for k in range(1, n + 1):
                if k == n or sc[k] != sc[start]:
                    self.cell_lists[int(sc[start])] = self.point_ids[start:k]
                    start = k
        _do(run)


def hash_grid_point_id(grid, index):
    return int(grid.point_ids[index])


def hash_grid_query(grid, pos, radius):
    radius = F32(radius)
    inv = grid.inv
    s = [int((pos[c] - radius) * inv) for c in range(3)]
    e = [_bi.min(int((pos[c] + radius) * inv), s[c] + grid.dim[c] - 1) for c in range(3)]
    for z in range(s[2], e[2] + 1):
        for y in range(s[1], e[1] + 1):
            for x in range(s[0], e[0] + 1):
                lst = grid.cell_lists.get(grid._index(x, y, z))
                if lst is not None:
                    for j in lst:
                        yield int(j)


# ----------------------------------------------------------------------------- Mesh (warp/native/mesh.h)
class Mesh:
    def __init__(self, points, indices, velocities=None, support_winding_number=False, **_):
        self.points = points
        self.indices = indices
        self.faces = indices.data.reshape(-1, 3)
        self.id = self

    def refit(self):
        pass                            # brute-force queries read the current points


class _Query:
    __slots__ = ("result", "face", "u", "v", "sign")


def _closest_bary_all(a, b, c, p):
    """closest_point_to_triangle (Ericson 5.1.5 as in warp/native/mesh.h) for all faces at once, float32,
    first matching region wins; returns barycentrics (u, v) with point = u*a + v*b + (1-u-v)*c."""
    ab, ac, ap = b - a, c - a, p - a
    d1 = ab[:, 0] * ap[:, 0] + ab[:, 1] * ap[:, 1] + ab[:, 2] * ap[:, 2]
    d2 = ac[:, 0] * ap[:, 0] + ac[:, 1] * ap[:, 1] + ac[:, 2] * ap[:, 2]
    bp = p - b
    d3 = ab[:, 0] * bp[:, 0] + ab[:, 1] * bp[:, 1] + ab[:, 2] * bp[:, 2]
    d4 = ac[:, 0] * bp[:, 0] + ac[:, 1] * bp[:, 1] + ac[:, 2] * bp[:, 2]
    cp = p - c
    d5 = ab[:, 0] * cp[:, 0] + ab[:, 1] * cp[:, 1] + ab[:, 2] * cp[:, 2]
    d6 = ac[:, 0] * cp[:, 0] + ac[:, 1] * cp[:, 1] + ac[:, 2] * cp[:, 2]
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    one, zero = np.ones_like(d1), np.zeros_like(d1)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ab = d1 / (d1 - d3)
        t_ac = d2 / (d2 - d6)
        t_bc = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        denom = F32(1.0) / (va + vb + vc)
    conds = [
        (d1 <= 0) & (d2 <= 0),
        (d3 >= 0) & (d4 <= d3),
        (vc <= 0) & (d1 >= 0) & (d3 <= 0),
        (d6 >= 0) & (d5 <= d6),
        (vb <= 0) & (d2 >= 0) & (d6 <= 0),
        (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0),
    ]
    # (u, v) per region, third weight = 1 - u - v as mesh_eval_position forms it
    us = [one, zero, one - t_ab, zero, one - t_ac, zero]
    vs = [zero, one, t_ab, zero, zero, one - t_bc]
    vv, ww = vb * denom, vc * denom
    u = np.select(conds, us, default=one - vv - ww).astype(F32)
    v = np.select(conds, vs, default=vv).astype(F32)
    return u, v


def _solid_angles(a, b, c, p):
    """Van Oosterom & Strackee, float32, per face."""
    a, b, c = a - p, b - p, c - p
    la = np.sqrt(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2])
    lb = np.sqrt(b[:, 0] * b[:, 0] + b[:, 1] * b[:, 1] + b[:, 2] * b[:, 2])
    lc = np.sqrt(c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2])
    bc = np.stack([b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1], b[:, 2] * c[:, 0] - b[:, 0] * c[:, 2],
                   b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]], 1)
    det = a[:, 0] * bc[:, 0] + a[:, 1] * bc[:, 1] + a[:, 2] * bc[:, 2]
    ab = a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]
    bcd = b[:, 0] * c[:, 0] + b[:, 1] * c[:, 1] + b[:, 2] * c[:, 2]
    ca = c[:, 0] * a[:, 0] + c[:, 1] * a[:, 1] + c[:, 2] * a[:, 2]
    den = la * lb * lc + ab * lc + bcd * la + ca * lb
    return (F32(2.0) * np.arctan2(det, den)).astype(F32)


def mesh_query_point_sign_winding_number(mesh, point, max_dist=0.0, accuracy=2.0, threshold=0.5):
    """Restated brute force over all faces: strictly-smaller squared distance starting from max_dist^2, lowest
    face index on exact ties (Warp: BVH traversal order); sign = -1 if the exact winding number (sum of solid
    angles / 4 pi; Warp: far-field approximation controlled by `accuracy`) exceeds `threshold`, else +1."""
    q = _Query()
    pts = mesh.points.data
    p = np.asarray(point, F32)
    md = F32(max_dist)
    lo, hi = pts.min(0), pts.max(0)
    q.result, q.face, q.u, q.v, q.sign = False, 0, F32(0), F32(0), F32(1)
    if np.any(p < lo - md * F32(1.001)) or np.any(p > hi + md * F32(1.001)):
        return q                        # what the BVH prunes: every triangle is farther than max_dist
    f = mesh.faces
    a, b, c = pts[f[:, 0]], pts[f[:, 1]], pts[f[:, 2]]
    u, v = _closest_bary_all(a, b, c, p[None])
    w = (F32(1.0) - u - v).astype(F32)
    cpt = (u[:, None] * a + v[:, None] * b + w[:, None] * c).astype(F32)
    d = cpt - p[None]
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    k = int(np.argmin(d2))              # first minimum = lowest face index among exact ties
    if not d2[k] < md * md:
        return q
    q.result, q.face, q.u, q.v = True, k, F32(u[k]), F32(v[k])
    wn = F32(0.0)
    for s in _solid_angles(a, b, c, p[None]):
        wn = F32(wn + s)
    wn = F32(wn * F32(0.25) * F32(0.31830988618379067))
    q.sign = F32(-1.0) if wn > F32(threshold) else F32(1.0)
    return q


def mesh_eval_position(mesh, face, u, v):
    """warp/native/mesh.h mesh_eval_position: u*p + v*q + (1-u-v)*r."""
    pts = mesh.points.data
    i, j, k = mesh.faces[face]
    return (pts[i] * u + pts[j] * v + pts[k] * (F32(1.0) - u - v)).astype(F32)


# ----------------------------------------------------------------------------- loading the reference
def install():
    """Make `import warp` resolve to this module."""
    sys.modules["warp"] = sys.modules[__name__]
    return sys.modules[__name__]


_smw = None


def load_reference(path: str = REF_SMW):
    """Import the reference's sim/physics/spring_mass_warp.py, unmodified, under this interpreter."""
    global _smw
    if _smw is None:
        prev = sys.modules.get("warp")
        install()
        try:
            spec = importlib.util.spec_from_file_location("_ref_spring_mass_warp", path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod          # inspect.getsource needs the module registered
            spec.loader.exec_module(mod)
        finally:
            if prev is not None:
                sys.modules["warp"] = prev
            else:
                sys.modules.pop("warp", None)
        _smw = mod
    return _smw
