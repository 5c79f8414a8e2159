"""ctypes binding of libr2s.so (the C ABI declared in include/r2s_*.h).

There is no fallback: if the CUDA library cannot be loaded every product entry
point raises.  The library is built in-tree by real2sim_eval_b200/build.py.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("R2S_LIB", os.path.join(_HERE, "libr2s.so"))

c_i32, c_i64, c_f, c_vp, c_sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class R2SError(RuntimeError):
    pass


class PhysDesc(C.Structure):  # include/r2s_phys.h: r2s_phys_desc
    _fields_ = [
        ("E", c_i32), ("N", c_i32), ("S", c_i32), ("n_substeps", c_i32), ("self_collision", c_i32),
        ("reverse_z", c_i32), ("use_pusher", c_i32), ("sign_mode", c_i32), ("coll_row_cap", c_i32),
        ("threads", c_i32), ("precise", c_i32), ("mesh_accel", c_i32), ("pad0_", c_i32),
        ("dt", c_f), ("dashpot_damping", c_f), ("drag_damping", c_f),
        ("spring_Y_min", c_f), ("spring_Y_max", c_f), ("collision_dist", c_f),
        ("collide_elas", c_f), ("collide_fric", c_f), ("collide_eef_elas", c_f), ("collide_eef_fric", c_f),
        ("collide_self_elas", c_f), ("collide_self_fric", c_f),
        ("springs", c_vp), ("rest_lengths", c_vp), ("rest_per_env", c_i32),
        ("log_spring_Y", c_vp), ("masses", c_vp), ("collision_mask", c_vp),
    ]


class PhysPtrs(C.Structure):  # r2s_phys_ptrs
    _fields_ = [
        ("x4", c_vp), ("v4", c_vp), ("collision_forces", c_vp), ("mesh_map", c_vp), ("coll_num", c_vp),
        ("coll_idx", c_vp), ("status", c_vp),
        ("F", c_i32), ("coll_row_cap", c_i32), ("smem_state", c_i32), ("smem_bytes", c_i32),
    ]


class RasterArgs(C.Structure):  # include/r2s_raster.h: r2s_raster_args
    _fields_ = [
        ("B", c_i32), ("views_per_scene", c_i32), ("P", c_i32), ("D", c_i32), ("M", c_i32), ("W", c_i32),
        ("H", c_i32), ("prefiltered", c_i32),
        ("scale_modifier", c_f), ("tanfovx", c_f), ("tanfovy", c_f), ("z_threshold", c_f),
        ("means3D", c_vp), ("scales", c_vp), ("rotations", c_vp), ("opacities", c_vp), ("shs", c_vp),
        ("colors_precomp", c_vp), ("cov3D_precomp", c_vp),
        ("viewmatrix", c_vp), ("projmatrix", c_vp), ("campos", c_vp), ("bg", c_vp),
        ("out_color", c_vp), ("out_depth", c_vp), ("radii", c_vp), ("out_rgb8", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_sz), ("max_instances", c_i64),
        ("tanfov_views", c_vp), ("overflow_count", c_vp), ("composite_mode", c_i32), ("pad0_", c_i32),
        ("composite_stream", c_vp),
    ]


class LinksArgs(C.Structure):  # include/r2s_links.h: r2s_links_args
    _fields_ = [
        ("E", c_i32), ("L", c_i32), ("P", c_i32), ("first", c_i32), ("n_robot", c_i32), ("pad0_", c_i32),
        ("link_id", c_vp), ("rest_means", c_vp), ("rest_quats", c_vp), ("link_pose", c_vp), ("link_offset", c_vp),
        ("rest_inv", c_vp), ("means3D", c_vp), ("rotations", c_vp), ("link_scratch", c_vp),
    ]


class EefArgs(C.Structure):  # include/r2s_eef.h: r2s_eef_args
    _fields_ = [
        ("E", c_i32), ("n_substeps", c_i32), ("n_pts", c_i32), ("n_table", c_i32), ("use_pusher", c_i32), ("F", c_i32),
        ("force_faces", c_i32 * 6), ("dyn_vel_rows", c_i32), ("pad0_", c_i32),
        ("grasp_force_threshold", c_f), ("pad1_", c_f), ("dt", C.c_double),
        ("table", c_vp), ("init_eef_xyz", c_vp), ("eef_xyz", c_vp), ("eef_vel", c_vp), ("eef_rot", c_vp),
        ("eef_rot_vel", c_vp), ("openness_cmd", c_vp), ("collision_forces", c_vp), ("current_openness", c_vp),
        ("grasped", c_vp), ("interp_pts", c_vp), ("interp_center", c_vp), ("dyn_vel", c_vp), ("dyn_omega", c_vp),
    ]


class SuccessArgs(C.Structure):  # include/r2s_metrics.h: r2s_success_args
    _fields_ = [
        ("E", c_i32), ("N", c_i32), ("S", c_i32), ("task", c_i32), ("frame", c_i32), ("start_frame", c_i32),
        ("need_frames", c_i32), ("ring_slots", c_i32),
        ("x4", c_vp), ("shift", c_vp), ("target", c_vp), ("springs", c_vp),
        ("box", C.c_double * 15), ("threshold", C.c_double),
        ("value", c_vp), ("passed", c_vp), ("hits", c_vp), ("success", c_vp), ("ring", c_vp),
    ]


class PhysMotion(C.Structure):  # include/r2s_phys.h: r2s_phys_motion
    _fields_ = [
        ("interp_pts", c_vp), ("interp_center", c_vp), ("dyn_vel", c_vp), ("dyn_omega", c_vp),
        ("n_env", c_i32), ("n_substeps", c_i32), ("n_dyn_verts", c_i32), ("dyn_vel_rows", c_i32),
    ]


class LbsArgs(C.Structure):  # include/r2s_lbs.h: r2s_lbs_args
    _fields_ = [
        ("E", c_i32), ("N", c_i32), ("P", c_i32), ("n_obj", c_i32), ("k_rel", c_i32), ("k_wgt", c_i32),
        ("relations", c_vp), ("weights_indices", c_vp), ("weights", c_vp), ("bones4", c_vp), ("bones_new4", c_vp),
        ("means3D", c_vp), ("rot_scratch", c_vp), ("rank_flags", c_vp),
        ("bone_slot", c_vp), ("weights_slots", c_vp), ("weights_by_slot", c_vp),
    ]


class RasterLayout(C.Structure):  # r2s_raster_layout
    _fields_ = [(n, c_sz) for n in ("status", "depths", "radii", "tiles_touched", "rec_a", "rec_b", "rec_c", "rects",
                                    "tile_count", "tile_offset", "tile_fill", "keys", "keys_alt", "sorted_rect",
                                    "total")] + \
               [("tiles_x", c_i32), ("tiles_y", c_i32), ("super_x", c_i32), ("super_y", c_i32)]


# every symbol include/*.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("r2s_last_error", C.c_char_p, []),
    ("r2s_version", C.c_int, []),
    ("r2s_launch_count", c_i64, []),
    ("r2s_phys_create", c_vp, [C.POINTER(PhysDesc)]),
    ("r2s_phys_destroy", C.c_int, [c_vp]),
    ("r2s_phys_set_state", C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    ("r2s_phys_get_state", C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    ("r2s_phys_set_spring_Y", C.c_int, [c_vp, c_vp, c_vp]),
    ("r2s_phys_set_rest_lengths", C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    ("r2s_phys_set_collide", C.c_int, [c_vp, c_f, c_f, c_f, c_f, c_f, c_f]),
    ("r2s_phys_set_mesh", C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32]),
    ("r2s_phys_set_mesh_motion", C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp]),
    ("r2s_phys_motion_ptrs", C.c_int, [c_vp, C.c_int, C.POINTER(PhysMotion)]),
    ("r2s_phys_create_resting_case", C.c_int, [c_vp, c_vp]),
    ("r2s_phys_update_collision_graph", C.c_int, [c_vp, c_vp]),
    ("r2s_phys_step", C.c_int, [c_vp, c_i32, c_vp]),
    ("r2s_phys_get_ptrs", C.c_int, [c_vp, C.POINTER(PhysPtrs)]),
    ("r2s_phys_algorithmic_bytes", c_i64, [c_vp]),
    ("r2s_raster_workspace_bytes", c_sz, [c_i32, c_i32, c_i32, c_i32, c_i64]),
    ("r2s_raster_workspace_layout", C.c_int, [c_i32, c_i32, c_i32, c_i32, c_i64, C.POINTER(RasterLayout)]),
    ("r2s_raster_forward", C.c_int, [C.POINTER(RasterArgs), c_vp]),
    ("r2s_raster_status", C.c_int, [c_vp, c_vp, C.POINTER(c_i64), C.POINTER(c_i32)]),
    ("r2s_mark_visible", C.c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    ("r2s_raster_set_profile", C.c_int, [c_i32]),
    ("r2s_raster_get_profile", C.c_int, [C.POINTER(c_f * 5)]),
    ("r2s_lbs_forward", C.c_int, [C.POINTER(LbsArgs), c_vp]),
    ("r2s_links_forward", C.c_int, [C.POINTER(LinksArgs), c_vp]),
    ("r2s_eef_forward", C.c_int, [C.POINTER(EefArgs), c_vp]),
    ("r2s_success_forward", C.c_int, [C.POINTER(SuccessArgs), c_vp]),
]

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libr2s.so, binding every declared symbol.  Raises R2SError if it is
    missing and cannot be built -- there is no CPU or library fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise R2SError(f"{LIB_PATH} is missing (run python -m real2sim_eval_b200.build)")
        from . import build as _build
        _build.build()
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise R2SError(f"cannot load {LIB_PATH}: {e}") from e
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().r2s_last_error().decode(errors="replace")
        raise R2SError(f"{what or 'libr2s'} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(load().r2s_launch_count())
