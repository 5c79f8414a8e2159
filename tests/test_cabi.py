"""The C-ABI shared library: loads without a GPU, exports every symbol include/*.h declares, the
ctypes mirrors have the C structs' sizes, and argument validation reports through the error API
(no compute is launched here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared():
    names = set()
    for f in os.listdir(INC):
        src = open(os.path.join(INC, f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(r2s_[A-Za-z_0-9]+)\s*\(", src))
    return names


def test_library_loads_and_exports_every_declared_symbol():
    from real2sim_eval_b200 import _lib
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 25
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert declared == bound, f"header/binding mismatch: {declared ^ bound}"
    for n in declared:
        assert hasattr(lib, n), n
    assert lib.r2s_version() == 100 and lib.r2s_last_error() is not None


def test_ctypes_struct_mirrors_match_the_c_headers(tmp_path):
    from real2sim_eval_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "r2s_phys.h"\n#include "r2s_raster.h"\n#include "r2s_lbs.h"\n'
                   '#include "r2s_links.h"\n#include "r2s_eef.h"\n#include "r2s_metrics.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(r2s_phys_desc), sizeof(r2s_phys_ptrs),'
                   ' sizeof(r2s_raster_args), sizeof(r2s_raster_layout), sizeof(r2s_lbs_args), sizeof(r2s_links_args),'
                   ' sizeof(r2s_eef_args), sizeof(r2s_phys_motion), sizeof(r2s_success_args)); return 0;}\n')
    exe = tmp_path / "sz"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-I", INC, str(src), "-o", str(exe)], check=True)   # the headers are plain C
    got = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    want = [C.sizeof(_lib.PhysDesc), C.sizeof(_lib.PhysPtrs), C.sizeof(_lib.RasterArgs), C.sizeof(_lib.RasterLayout),
            C.sizeof(_lib.LbsArgs), C.sizeof(_lib.LinksArgs), C.sizeof(_lib.EefArgs), C.sizeof(_lib.PhysMotion),
            C.sizeof(_lib.SuccessArgs)]
    assert got == want


def test_argument_validation_reports_through_last_error():
    from real2sim_eval_b200 import _lib
    lib = _lib.load()
    assert lib.r2s_raster_forward(None, None) == -1
    assert b"null args" in lib.r2s_last_error()
    a = _lib.RasterArgs()
    a.B, a.views_per_scene, a.P, a.W, a.H = 3, 2, 10, 64, 64
    assert lib.r2s_raster_forward(C.byref(a), None) == -1
    assert b"views_per_scene" in lib.r2s_last_error()
    assert lib.r2s_phys_step(None, 0, None) == -1 and b"null handle" in lib.r2s_last_error()
    d = _lib.PhysDesc()
    assert not lib.r2s_phys_create(C.byref(d)) and b"bad descriptor" in lib.r2s_last_error()
    assert lib.r2s_raster_get_profile(None) == -1
    assert lib.r2s_lbs_forward(None, None) == -1 and b"null args" in lib.r2s_last_error()
    assert lib.r2s_eef_forward(None, None) == -1 and b"null args" in lib.r2s_last_error()
    assert lib.r2s_success_forward(None, None) == -1 and b"null args" in lib.r2s_last_error()
    e = _lib.EefArgs()
    e.E, e.n_substeps, e.n_pts, e.n_table = 1, 4096, 48, 101
    assert lib.r2s_eef_forward(C.byref(e), None) == -1 and b"bad sizes" in lib.r2s_last_error()


def test_workspace_layout_arithmetic():
    from real2sim_eval_b200 import _lib
    lib = _lib.load()
    L = _lib.RasterLayout()
    assert lib.r2s_raster_workspace_layout(4, 1000, 512, 480, 50000, C.byref(L)) == 0
    assert (L.tiles_x, L.tiles_y, L.super_x, L.super_y) == (32, 30, 8, 8)
    offs = [L.status, L.depths, L.radii, L.tiles_touched, L.rec_a, L.rec_c, L.rects, L.tile_count,
            L.tile_offset, L.tile_fill, L.keys, L.keys_alt, L.sorted_rect, L.total]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs[:-1])
    assert L.rec_b == L.rec_a + 16 and L.rec_c - L.rec_a >= 32 * 4 * 1000, "32-byte records {rec_a, rec_b}"
    assert L.keys_alt - L.keys >= 8 * 50000
    assert lib.r2s_raster_workspace_bytes(4, 1000, 512, 480, 50000) == L.total
    assert lib.r2s_raster_workspace_layout(1, 10, 16 * 300, 64, 10, C.byref(L)) != 0, "tile rectangles are 8-bit"
    assert lib.r2s_raster_workspace_bytes(0, 10, 64, 64, 10) == 0


def test_product_path_has_no_cpu_fallback_and_never_imports_the_oracle():
    import torch
    from real2sim_eval_b200 import _lib
    from real2sim_eval_b200.physics import BatchedSpringMass
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    with pytest.raises(_lib.R2SError, match="no CPU path"):
        BatchedRasterizer("cpu")
    with pytest.raises(_lib.R2SError, match="no CPU path"):
        BatchedSpringMass(1, [[0, 1]], [0.1], num_particles=2, n_substeps=1, device="cpu")
    pkg = os.path.join(ROOT, "real2sim_eval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
    code = "import sys, real2sim_eval_b200, real2sim_eval_b200.physics, real2sim_eval_b200.rasterizer, " \
           "real2sim_eval_b200.envs; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_preprocess_fp_pairing_signature_is_the_one_verified_on_the_gpu():
    """preprocess_kernel is bit-identical to the reference build only while ptxas fuses the same multiplies with the
    same adds (moving loads or changing the register cap has changed conics / SH colours by an ulp
    before).  tests/golden/sass_signature_preprocess.json holds the floating-point instruction signature of the build
    whose images were `array_equal` to the live reference on a B200; a different signature from the same toolchain
    means: re-run `pytest -m gpu tests/test_gpu_raster.py`, then refresh the file with tools/sass_signature.py --write."""
    import json
    import shutil
    from real2sim_eval_b200 import _lib
    root = ROOT
    sys.path.insert(0, os.path.join(root, "tools"))
    import sass_signature
    _lib.load()   # builds the library if needed
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    want = json.load(open(os.path.join(root, "tests", "golden", "sass_signature_preprocess.json")))
    if sass_signature.toolchain() != want["nvcc"]:
        pytest.skip(f"signature recorded with nvcc {want['nvcc']}")
    got = sass_signature.signature(_lib.LIB_PATH, "preprocess_kernel")
    assert got["by_opcode"] == want["by_opcode"] and got["digest"] == want["digest"], \
        "FP instruction pairing of preprocess_kernel changed: re-verify bit-identity on a GPU, then refresh the golden"
