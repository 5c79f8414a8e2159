"""Host-side logic that needs no GPU: synthetic workload generators, camera conventions, the
reference-compatible import shims, env sharding, and the N>1 metrics all-gather over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from real2sim_eval_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rope_graph_follows_the_reference_rule():
    """phystwin.py:264-286 with radius 0.02 / 30 neighbours on the seeded rope: the sizes SURVEY §8d records."""
    r = synth.make_rope()
    assert (r.N, r.S) == (2048, 32889)
    deg = np.bincount(r.springs.reshape(-1), minlength=r.N)
    assert deg.min() == 29 and deg.max() == 49
    assert len({tuple(sorted(s)) for s in r.springs.tolist()}) == r.S, "first-come de-duplication"
    assert (r.rest > 1e-4).all()
    posed = synth.pose_scene(r, 7)
    assert np.array_equal(posed.springs, r.springs) and not np.array_equal(posed.rest, r.rest)
    assert np.abs(posed.rest - r.rest).max() < 1e-6, "rest lengths differ per env only by float32 rounding"


def test_finger_mesh_is_closed_and_outward():
    v, f = synth.make_finger_mesh()
    assert v.shape == (24, 3) and f.shape == (44, 3)
    edges = {}
    for a, b, c in f.tolist():
        for e in ((a, b), (b, c), (c, a)):
            edges[e] = edges.get(e, 0) + 1
    assert all(edges.get((b, a), 0) == 1 and n == 1 for (a, b), n in edges.items()), "2-manifold, consistent winding"
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol = np.einsum("ij,ij->i", p0, np.cross(p1, p2)).sum() / 6.0
    assert vol > 0, "outward orientation (positive signed volume)"
    g = synth.make_gripper((0.5, 0.0, 0.01))
    assert g.verts.shape == (48, 3) and g.faces.shape == (88, 3) and g.mesh_map.tolist() == [0] * 44 + [1] * 44


def test_setup_camera_matches_the_reference_formulas():
    """sim/utils/gs/transform_utils.py:7-31 restated with torch, as the reference computes it."""
    import torch
    w, h = 848, 480
    k = np.asarray(synth.SIDE_CAM["intr"]).reshape(3, 3)
    w2c = np.linalg.inv(np.asarray(synth.SIDE_CAM["c2w"]).reshape(4, 4))
    cam = synth.setup_camera(w, h, k, w2c, near=0.01, far=100.0)
    fx, fy, cx, cy = k[0][0], k[1][1], k[0][2], k[1][2]
    t_w2c = torch.tensor(w2c).float()
    center = torch.inverse(t_w2c)[:3, 3]
    t_view = t_w2c.unsqueeze(0).transpose(1, 2)
    proj = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0], [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                         [0.0, 0.0, 100.0 / (100.0 - 0.01), -(100.0 * 0.01) / (100.0 - 0.01)],
                         [0.0, 0.0, 1.0, 0.0]]).float().unsqueeze(0).transpose(1, 2)
    full = t_view.bmm(proj)
    assert np.allclose(cam.view, t_view.reshape(-1).numpy(), atol=1e-7)
    assert np.allclose(cam.proj, full.reshape(-1).numpy(), rtol=1e-6, atol=1e-7)
    assert np.allclose(cam.campos, center.numpy(), atol=1e-6)
    assert cam.tanfovx == pytest.approx(w / (2 * fx)) and cam.tanfovy == pytest.approx(h / (2 * fy))


def test_compat_shims_expose_the_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "real2sim_eval_b200", "compat"))
    try:
        import importlib
        dgr = importlib.import_module("diff_gaussian_rasterization")
        assert dgr.GaussianRasterizationSettings._fields == (
            "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
            "sh_degree", "campos", "prefiltered", "z_threshold")          # __init__.py:135-147
        import inspect
        sig = inspect.signature(dgr.GaussianRasterizer.forward)
        assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                            "rotations", "cov3D_precomp"]  # __init__.py:165
        wp = importlib.import_module("warp")
        for name in ("init", "set_module_options", "ScopedTimer", "to_torch", "capture_launch"):
            assert hasattr(wp, name)
    finally:
        sys.path.pop(0)
        for m in ("diff_gaussian_rasterization", "warp"):
            sys.modules.pop(m, None)
    from real2sim_eval_b200.physics import SpringMassSystemWarp
    import inspect
    params = list(inspect.signature(SpringMassSystemWarp.__init__).parameters)[1:21]
    assert params == ["phystwin_cfg", "device", "init_vertices", "init_springs", "init_rest_lengths", "init_masses",
                      "num_object_points", "init_spring_Y", "collide_elas", "collide_fric", "collide_eef_elas",
                      "collide_eef_fric", "collide_self_elas", "collide_self_fric", "init_collision_mask",
                      "init_velocities", "dynamic_meshes", "static_meshes", "dynamic_points", "use_pusher"]  # SMW:478-500


def test_env_sharding_partitions_exactly():
    for total, world in ((256, 1), (256, 8), (2048, 8), (10, 4), (3, 8)):
        parts = [shard.shard_envs(total, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(total))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    mix = shard.interleave_scene_types({"rope": 1024, "sloth": 512, "tblock": 512}, 8)
    assert all(len(m) == 256 and m.count("rope") == 128 and m.count("sloth") == 64 for m in mix)
    with pytest.raises(ValueError):
        shard.shard_envs(8, 2, 2)
    assert shard.gather_metrics([1.0, 2.0]) == [[1.0, 2.0]] and shard.max_over_ranks(3.5) == 3.5


_WORKER = r"""
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from real2sim_eval_b200 import shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
mine = shard.shard_envs(10, world, rank)
got = shard.gather_metrics([float(rank), float(len(mine)), float(sum(mine))])
mx = shard.max_over_ranks(10.0 + rank)
dist.barrier()
if rank == 0:
    print(json.dumps({"got": got, "max": mx}))
dist.destroy_process_group()
"""


def test_metrics_allgather_world_size_2_gloo(tmp_path):
    """The N>1 path of bench.py (shard -> run -> all-gather -> max over ranks) on CPU with gloo."""
    import json
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["got"] == [[0.0, 5.0, 10.0], [1.0, 5.0, 35.0]] and res["max"] == 11.0


def test_eef_and_metrics_host_helpers_need_no_gpu():
    """Host-side pieces of the N3 / N4 rows: the force rows the grasp test reads, the routing box, argument
    checks that fire before any device work."""
    import numpy as np
    import pytest
    from oracle import eef_ref, metrics_ref
    from real2sim_eval_b200 import _lib, eef, metrics, synth
    g = synth.make_gripper((0.5, 0.0, 0.03))
    assert eef.force_faces(g.mesh_map) == eef_ref.force_faces(g.mesh_map) == [18, 19, 1, 44 + 18, 44 + 19, 44 + 1]
    with pytest.raises(ValueError, match="faces 18, 19 and 1"):
        eef.force_faces(np.zeros(10, np.int32))
    lo, hi = metrics.rope_box()
    rlo, rhi = metrics_ref.rope_box()
    assert np.array_equal(lo, rlo) and np.array_equal(hi, rhi)
    assert metrics.START_FRAME == metrics_ref.START_FRAME and metrics.NEED_FRAMES == metrics_ref.NEED_FRAMES == 30
    with pytest.raises(_lib.R2SError, match="no CPU path"):
        eef.BatchedEefMotion(1, synth.gripper_opening_table((0, 0, 0)), (0, 0, 0), dt=5e-5, n_substeps=4,
                             mesh_map=g.mesh_map, device="cpu")
    with pytest.raises(ValueError, match="unknown task"):
        metrics.BatchedSuccess("fold", 1, 4)
    t = synth.gripper_opening_table((0.5, 0.0, 0.03))
    assert t.shape == (101, 48, 3) and t.dtype == np.float32
    gap = lambda k: t[k, 24:, 1].mean() - t[k, :24, 1].mean()
    assert abs(gap(0) - 0.008) < 1e-6 and abs(gap(100) - 0.08) < 1e-6 and gap(50) > gap(49)


def test_launch_summary_tool_reads_the_committed_launch_list():
    """tools/launch_summary.py on profiles/r01b_launches_256envs.csv: the compositing kernel dominates a step."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r01b_launches_256envs.csv")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "launch_summary.py"), csv_path, "--steps", "3"],
                         capture_output=True, text=True, check=True).stdout
    rows = [l for l in out.splitlines() if "share of one step" in l and "%" in l]
    assert rows[0].startswith("composite_kernel") and len(rows) >= 10
    shares = [float(l.split("share of one step")[1].replace("%", "")) for l in rows]
    assert abs(sum(shares) - 100.0) < 0.5 and shares[0] > 50.0


def test_super_tile_local_rectangle_is_equivalent_to_the_global_one():
    """The invariant behind the packed sort keys (raster.cu emit_kernel / composite_kernel): for every tile
    rectangle [minx, maxx) x [miny, maxy) and every 4x4 super-tile it touches, a tile of that super-tile lies in
    the rectangle iff its local coordinates lie in the rectangle clipped to the super-tile and made local --
    computed as the kernel does (only the first / last super-tile of a row or column is cut)."""
    K = 4
    n = 0
    for minx in range(0, 13):
        for maxx in range(minx + 1, 14):
            sx0, sx1 = minx // K, (maxx + K - 1) // K
            for sx in range(sx0, sx1):
                x0 = minx - K * sx0 if sx == sx0 else 0
                x1 = maxx - K * sx if sx == sx1 - 1 else K
                assert 0 <= x0 <= 3 and 1 <= x1 <= 4, (minx, maxx, sx)
                # the same values as the straightforward clip
                assert x0 == min(max(minx - K * sx, 0), K) and x1 == min(max(maxx - K * sx, 0), K)
                for lx in range(K):
                    tx = K * sx + lx
                    assert (minx <= tx < maxx) == (x0 <= lx < x1)
                    n += 1
    assert n > 500
    # packing: four 3-bit fields under the id
    low = (123456 << 12) | (1 | (2 << 3) | (4 << 6) | (3 << 9))
    assert (low & 7, (low >> 3) & 7, (low >> 6) & 7, (low >> 9) & 7, low >> 12) == (1, 2, 4, 3, 123456)
    assert ((1 << 20) - 1) << 12 < (1 << 32), "ids below 2^20 fit above the 12 rectangle bits"


def test_bench_config_label_and_numa_binding_are_safe_without_a_gpu():
    """bench.py helpers: the BASELINE config a command line describes, and the NUMA binding, which must degrade to a
    report (never raise) on a box without GPUs or sysfs entries."""
    import types
    import bench
    mk = lambda **kw: types.SimpleNamespace(**{**dict(scene="rope", envs=256, res=[512, 512], cameras=1, substeps=10), **kw})
    assert bench.config_label(mk()) == "BASELINE configs[1]"
    assert bench.config_label(mk(scene="sloth", envs=64, res=[640, 480], cameras=2)) == "BASELINE configs[2]"
    assert bench.config_label(mk(envs=128)) == "custom shape"
    info = bench.bind_to_gpu_numa_node(0)
    assert info["bound"] is False and ("error" in info or "numa_node" in info)


def test_spatial_order_is_a_permutation_that_keeps_neighbours_together():
    from real2sim_eval_b200 import synth
    p = np.random.default_rng(0).uniform(0, 1, (2000, 3))
    o = synth.spatial_order(p)
    assert sorted(o.tolist()) == list(range(2000))
    step_sorted = np.linalg.norm(np.diff(p[o], axis=0), axis=1).mean()
    step_random = np.linalg.norm(np.diff(p, axis=0), axis=1).mean()
    assert step_sorted < 0.3 * step_random


def test_lbs_slot_layout_keeps_every_bone_weight_pair():
    """The LBS layout hint (include/r2s_lbs.h: bone_slot / weights_slots / weights_by_slot) only moves data: the slots
    are a permutation, every Gaussian keeps its (bone, weight) pairs, rows ascend, and neighbouring bones get
    neighbouring slots; the oracle's blend over the re-ordered pairs is the same sum to rounding."""
    from oracle import lbs_ref
    from real2sim_eval_b200.lbs import slot_layout
    rng = np.random.default_rng(3)
    n, n_obj, k = 500, 300, 16
    base = rng.uniform(0, 0.2, (n, 3)).astype(np.float32)
    pts = (base[rng.integers(0, n, n_obj)] + rng.normal(0, 0.003, (n_obj, 3))).astype(np.float32)
    w, wi = lbs_ref.knn_weights(base, pts, k)
    slot, ws, wbs = slot_layout(base, wi, w)
    assert sorted(slot.tolist()) == list(range(n))
    assert (np.diff(ws, axis=1) >= 0).all()
    inv = np.empty(n, np.int64); inv[slot] = np.arange(n)
    for g in range(n_obj):
        assert sorted(zip(inv[ws[g]].tolist(), wbs[g].tolist())) == sorted(zip(wi[g].tolist(), w[g].tolist()))
    spread = lambda rows: np.mean(rows.max(1) - rows.min(1))
    assert spread(ws) < 0.5 * spread(np.asarray(wi)), "a Gaussian's bones sit closer together in slot order"
    rel = lbs_ref.knn_relations(base, 8)
    motions = (0.02 * np.sin(30 * base[:, [1, 2, 0]])).astype(np.float32)
    a = lbs_ref.interpolate_motions(base, motions, rel, pts, w, wi)
    b = lbs_ref.interpolate_motions(base, motions, rel, pts, wbs, inv[ws])
    assert np.abs(a - b).max() < 1e-6


def test_ncu_sass_tool_summarises_a_source_page(tmp_path):
    """tools/ncu_sass.py on a hand-made `ncu --page source --csv --print-source sass` excerpt: opcode mix weighted by
    executions (predicated instructions counted under their opcode), the bulk-copy / mbarrier list, the hot-loop window."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = '"Address","Source","Warp Stall Sampling (All Samples)","Warp Stall Sampling (Not-issued Samples)","# Samples","Instructions Executed"'
    rows = [("LDC R1, c[0x0][0x37c]", 10, 1), ("UBLKCP.S.G [UR8], [UR4], UR6", 4, 2), ("FFMA R1, R2, R3, R4", 600, 5),
            ("@!P0 BRA 0x10", 100, 7), ("MUFU.EX2 R5, R5", 200, 9), ("FMUL R5, R5, R6", 90, 3)]
    src = tmp_path / "sass.csv"
    src.write_text('"Kernel Name","k",\n' + hdr + "\n" +
                   "".join(f'"0x{16 * i:x}","      {s}","{n}","0","{n}","{ex}"\n' for i, (s, ex, n) in enumerate(rows)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_sass.py"), str(src), "--nth", "1", "--before", "1",
                          "--window", "2"], capture_output=True, text=True, check=True).stdout
    assert "total warp instructions 1,004, stall samples 27" in out
    mix = [l for l in out.splitlines() if l.startswith("#   ")]
    assert mix[0].split()[1] == "FFMA" and any(l.split()[1] == "BRA" and "10.0 %" in l for l in mix)
    assert sum("UBLKCP.S.G" in l for l in out.splitlines() if l.startswith("  [")) == 1
    win = out.split("(3)")[1].splitlines()[1:]
    assert len(win) == 2 and "BRA" in win[0] and "MUFU.EX2" in win[1]


def test_exponent_field_arithmetic_of_the_compositing_kernel():
    """The IEEE facts raster.cu's opacity_expf_seq rests on, checked in numpy float32 (no GPU):
    (1) expf's range reduction t = floor(252 * s + magic) keeps the biased exponent j + 127 in its low bits; with the
        magic lowered by 127 the low 9 bits of t's bit pattern are j itself in two's complement, and t - 12582912 == j;
    (2) fl(op * (2^j * r)) == fl(op * r) scaled by 2^j (an integer add of j << 23 on the bit pattern) whenever the
        product stays normal -- which holds wherever an entry can be blended: op * e^x >= 0.998/255;
    (3) below power_min = -log(255 * op) - 2e-3 the alpha of forward.cu:350 is below 1/255, so the warp vote may skip."""
    rng = np.random.default_rng(8)
    f32 = np.float32
    # (1): fma.rm to a float32 in [2^23, 2^24) is floor() of the exact value; 252 * s + magic is exact in float64
    s = rng.uniform(0, 1, 20000).astype(f32)
    s[:2] = [0.0, 1.0]
    t_lib = np.floor(252.0 * s.astype(np.float64) + 12582913.0).astype(f32)        # libdevice's constant
    t_new = np.floor(252.0 * s.astype(np.float64) + 12582786.0).astype(f32)        # lowered by 127
    j = (t_lib - f32(12583039.0)).astype(np.int64)
    assert np.array_equal((t_new - f32(12582912.0)).astype(np.int64), j) and j.min() == -126 and j.max() == 126
    assert np.array_equal(t_lib.view(np.uint32) & 0xFF, (j + 127).astype(np.uint32)), "libdevice: biased exponent in the low byte"
    assert np.array_equal(t_new.view(np.uint32) & 0x1FF, (j % 512).astype(np.uint32)), "lowered magic: j in two's complement"
    assert np.array_equal(((t_new.view(np.uint32).astype(np.uint64) << 23) & 0xFFFFFFFF).astype(np.uint32),
                          ((j.astype(np.int64) << 23) & 0xFFFFFFFF).astype(np.uint32))
    # (2): scaling by 2^j commutes with the rounding of op * r while nothing underflows
    n = 200000
    op = np.concatenate([rng.uniform(1e-3, 1.0, n // 2), 10 ** rng.uniform(-2.4, 0.6, n // 2)]).astype(f32)
    r = rng.uniform(0.7, 1.4143, n).astype(f32)                                    # ex2 of the reduced argument
    jj = rng.integers(-9, 1, n)
    two_j = np.ldexp(f32(1.0), jj).astype(f32)
    ref = op * (two_j * r)                                                         # fmul(opacity, expf(x))
    live = ref >= f32(0.998 / 255.0)
    bits = (op * r).view(np.uint32).astype(np.int64) + (jj.astype(np.int64) << 23)
    got = (bits & 0xFFFFFFFF).astype(np.uint32).view(f32)
    assert live.sum() > n // 3 and np.array_equal(got[live], ref[live])
    # (3): the lower bound of the warp vote
    opv = 10 ** rng.uniform(-2.3, 0.5, 50000)
    pmin = (-np.log(255.0 * opv) - 2e-3).astype(f32)
    x = (pmin - np.abs(rng.normal(0, 0.5, opv.size)).astype(f32) - f32(1e-6)).astype(f32)     # just below the bound and further down
    alpha = np.minimum(f32(0.99), opv.astype(f32) * np.exp(x.astype(f32)))
    assert (alpha < f32(1.0 / 255.0)).all()
