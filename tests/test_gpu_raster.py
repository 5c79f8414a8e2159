"""GPU parity: CUDA rasterizer (through the C ABI) vs the CPU oracle
(oracle/raster_ref.c), vs the UNMODIFIED reference CUDA rasterizer when
oracle/_ref/libref_raster.so travelled with the snapshot, and vs the committed
golden fixtures made from that reference (tests/golden/raster_*.npz).

Contract (BASELINE.json north_star): RGB / depth within 1e-4 relative per pixel.
Integer stages (radii, tile ranges, sorted instance lists) are compared bit-exactly
whenever the float stage feeding them agrees bit-exactly, which the tests check
first.  Per-pixel thresholds (alpha < 1/255, T < 1e-4, the median-depth crossing)
turn 1-ulp differences in exp() into isolated jumps, so image comparisons allow a
small counted budget of outlier pixels and say so."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import r2s_testutil as _util
from real2sim_eval_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4
ATOL = 1e-5
OUTLIER_FRAC = 2e-3   # pixels allowed to differ by a threshold flip


def _torchify(d, dev="cuda"):
    import torch
    return {k: torch.tensor(v, device=dev) for k, v in d.items()}


def _run_cuda(g, cam, sh_degree=0, bg=(0.0, 0.0, 0.0), **kw):
    import torch
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    r = BatchedRasterizer("cuda")
    t = _torchify(g)
    kw.setdefault("max_instances", 64 * len(g["means3D"]) + 4096)
    color, radii, depth = r.forward(
        t["means3D"], t["opacities"], viewmatrix=torch.tensor(cam.view).cuda(), projmatrix=torch.tensor(cam.proj).cuda(),
        campos=torch.tensor(cam.campos).cuda(), bg=torch.tensor(bg, dtype=torch.float32).cuda(), W=cam.W, H=cam.H,
        tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, shs=t.get("shs"), colors_precomp=t.get("colors_precomp"),
        scales=t.get("scales"), rotations=t.get("rotations"), cov3D_precomp=t.get("cov3D_precomp"),
        sh_degree=sh_degree, z_threshold=cam.z_threshold, **kw)
    total, overflow = r.status()
    return r, color[0].cpu().numpy(), radii[0].cpu().numpy(), depth[0].cpu().numpy(), total, overflow


def _run_oracle(g, cam, sh_degree=0, bg=(0.0, 0.0, 0.0), aux=True):
    from oracle import raster_ref
    return raster_ref.rasterize(
        g["means3D"], g["opacities"], viewmatrix=cam.view, projmatrix=cam.proj, campos=cam.campos,
        bg=np.asarray(bg, np.float32), W=cam.W, H=cam.H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, shs=g.get("shs"),
        colors_precomp=g.get("colors_precomp"), scales=g.get("scales"), rotations=g.get("rotations"),
        cov3D_precomp=g.get("cov3D_precomp"), sh_degree=sh_degree, z_threshold=cam.z_threshold, aux=aux)


def _close_images(a, b, what):
    bad = np.abs(a - b) > (ATOL + RTOL * np.abs(b))
    frac = bad.mean()
    assert frac <= OUTLIER_FRAC, f"{what}: {bad.sum()} of {bad.size} pixels beyond rtol={RTOL} (max |d|={np.abs(a - b).max()})"
    return frac


def _super_lists(r, view=0):
    """(offsets [ST+1], keys uint64, rects uint32) of the sorted super-tile lists of one view."""
    it = r.intermediates()
    ST = it["supers"][0] * it["supers"][1]
    off = it["super_offset"].cpu().numpy().astype(np.int64)[view * ST: (view + 1) * ST + 1]
    keys = it["keys"].cpu().numpy().view(np.uint64)
    rects = it["sorted_rect"].cpu().numpy().view(np.uint32)
    return off, keys, rects, it


def _tile_lists(r, view=0):
    """Per-tile Gaussian id lists, rebuilt on the host the way composite_kernel walks them: the
    order-preserving subsequence of the tile's super-tile list whose rectangle contains the tile."""
    off, keys, rects, it = _super_lists(r, view)
    gx, gy = it["tiles"]
    sgx = it["supers"][0]
    sh = it["id_shift"]
    out = []
    for ty in range(gy):
        for tx in range(gx):
            s = (ty // 4) * sgx + tx // 4
            k = keys[off[s]:off[s + 1]]
            low = (k & np.uint64(0xffffffff)).astype(np.uint32)
            if sh:   # rectangle local to the super-tile, packed under the id
                lx, ly = tx % 4, ty % 4
                keep = (lx >= (low & 7)) & (lx < ((low >> 6) & 7)) & (ly >= ((low >> 3) & 7)) & (ly < ((low >> 9) & 7))
            else:
                rc = rects[off[s]:off[s + 1]]
                keep = (tx >= (rc & 255)) & (tx < ((rc >> 16) & 255)) & (ty >= ((rc >> 8) & 255)) & (ty < (rc >> 24))
            out.append(low[keep] >> np.uint32(sh))
    return out


REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libref_raster.so")
LIST_CHECKS = {"tiles": 0, "calls": 0}   # how many per-tile list comparisons against the live reference really ran


def test_reference_library_travelled():
    """The live comparisons below need the unmodified reference rasterizer built by oracle/Makefile in the build
    container (oracle/_ref/, git-ignored but shipped by gpurun): its absence on a GPU box must be loud."""
    assert os.path.exists(REF_SO), "oracle/_ref/libref_raster.so missing: run `make -C oracle` where /root/reference is mounted"
    import ref_raster
    assert hasattr(ref_raster.lib(), "ref_raster_lists")


def _assert_lists_equal_reference(r, g, cam, sh_degree=0, bg=(0.0, 0.0, 0.0), view=0, total=None):
    """UNCONDITIONAL list equality: the sequence composite_kernel walks for every tile (order-preserving filter of
    its super-tile's sorted list) equals the reference's own sorted point_list between its own tile ranges, read
    out of the live reference's binning / image buffers (identifyTileRanges + SortPairs, rasterizer_impl.cu:303-321)."""
    import ref_raster
    rc, rr, rd, n = ref_raster.forward(g, cam, sh_degree=sh_degree, bg=bg)
    pl, rg = ref_raster.lists(n, cam.W, cam.H)
    if total is not None:
        assert total == n, "instance count equals the reference's num_rendered"
    lists = _tile_lists(r, view)
    assert len(lists) == len(rg)
    for t, ids in enumerate(lists):
        assert np.array_equal(ids, pl[rg[t, 0]:rg[t, 1]]), f"tile {t}: list differs from the reference's"
    LIST_CHECKS["tiles"] += len(lists)
    LIST_CHECKS["calls"] += 1
    return rc, rr, rd, n


@pytest.mark.parametrize("W,H,P,seed", [(64, 64, 1500, 1), (128, 96, 4000, 2), (200, 120, 2500, 3)])
def test_matches_oracle_stagewise(W, H, P, seed):
    g = _util.small_gaussians(seed, P)
    cam = _util.make_test_camera(W, H)
    r, color, radii, depth, total, overflow = _run_cuda(g, cam, bg=(0.1, 0.2, 0.3))
    oc, orad, od, aux = _run_oracle(g, cam, bg=(0.1, 0.2, 0.3))
    assert not overflow
    it = r.intermediates()
    # --- preprocess: float stage within a few ulp, integer stage exact wherever floats agree
    depths = it["depths"][0].cpu().numpy()
    assert np.allclose(depths, aux["depths"], rtol=2e-6, atol=1e-7)
    vis = orad > 0
    same_vis = (radii > 0) == vis
    assert same_vis.mean() > 0.999
    ra = it["rec_a"][0].cpu().numpy()
    assert np.allclose(ra[vis & same_vis, :2], aux["means2D"][vis & same_vis], rtol=1e-5, atol=2e-4)
    agree = same_vis & (radii == orad)
    assert agree.mean() > 0.995, "radii (ceil of 3 sigma) may flip only on exact ties"
    tt = it["tiles_touched"][0].cpu().numpy().astype(np.int64)
    assert (tt[agree] == aux["tiles_touched"][agree].astype(np.int64)).mean() > 0.999
    # --- binning + sort: exact when the float stage agreed exactly
    float_exact = np.array_equal(depths, aux["depths"]) and np.array_equal(radii, orad) and \
        np.array_equal(tt, aux["tiles_touched"].astype(np.int64)) and \
        np.array_equal(ra[vis, :2], aux["means2D"][vis])
    off, keys, rects, _ = _super_lists(r)
    assert total == int(tt.sum()), "status reports the reference's num_rendered (sum of tiles_touched)"
    for t in range(len(off) - 1):      # every super-tile list is strictly ascending in (depth bits, id)
        seg = keys[off[t]:off[t + 1]]
        assert (np.diff(seg.view(np.int64)) > 0).all()
    if float_exact:                     # vs the CPU oracle only when its (uncontracted) float stage agrees bitwise
        assert total == aux["num_rendered"]
        lists = _tile_lists(r)          # what the composite kernel walks == the reference's sorted tile lists
        for t, ids in enumerate(lists):
            r0, r1 = aux["ranges"][t]
            assert np.array_equal(ids, aux["point_list"][r0:r1]), f"tile {t}"
    _assert_lists_equal_reference(r, g, cam, bg=(0.1, 0.2, 0.3), total=total)   # vs the live reference: always
    # --- images
    _close_images(color, oc, "color vs oracle")
    _close_images(depth, od, "depth vs oracle")


@pytest.mark.parametrize("deg", [1, 2, 3])
def test_sh_degrees_and_precomputed_inputs(deg):
    P, M = 800, (deg + 1) ** 2
    g = _util.small_gaussians(10 + deg, P, sh_coeffs=M)
    cam = _util.make_test_camera(96, 64)
    _, color, radii, depth, _, _ = _run_cuda(g, cam, sh_degree=deg)
    oc, orad, od, aux = _run_oracle(g, cam, sh_degree=deg)
    _close_images(color, oc, f"SH degree {deg}")
    # precomputed colour + covariance path (colors_precomp / cov3D_precomp)
    g2 = dict(means3D=g["means3D"], opacities=g["opacities"], colors_precomp=aux["rgb"], cov3D_precomp=aux["cov3D"])
    _, color2, radii2, depth2, _, _ = _run_cuda(g2, cam)
    oc2, orad2, od2, _ = _run_oracle(g2, cam)
    _close_images(color2, oc2, "precomputed inputs")
    assert (radii2 == orad2).mean() > 0.995


def test_argument_errors_match_reference_messages():
    import torch
    from real2sim_eval_b200.rasterizer import GaussianRasterizer, GaussianRasterizationSettings
    cam = _util.make_test_camera(32, 32)
    t = lambda a: torch.tensor(a).cuda()
    rs = GaussianRasterizationSettings(32, 32, cam.tanfovx, cam.tanfovy, t(np.zeros(3, np.float32)), 1.0,
                                       t(cam.view).reshape(1, 4, 4), t(cam.proj).reshape(1, 4, 4), 0, t(cam.campos),
                                       False, 0.05)
    g = _torchify(_util.small_gaussians(5, 50))
    rast = GaussianRasterizer(rs)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(means3D=g["means3D"], means2D=None, opacities=g["opacities"], scales=g["scales"], rotations=g["rotations"])
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        rast(means3D=g["means3D"], means2D=None, opacities=g["opacities"], shs=g["shs"])
    color, radii, depth = rast(means3D=g["means3D"], means2D=torch.zeros_like(g["means3D"]), opacities=g["opacities"],
                               shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
    assert tuple(color.shape) == (3, 32, 32) and tuple(depth.shape) == (1, 32, 32) and radii.dtype == torch.int32
    vis = rast.markVisible(g["means3D"])
    from oracle import raster_ref
    assert np.array_equal(vis.cpu().numpy(), raster_ref.mark_visible(g["means3D"].cpu().numpy(), cam.view, cam.proj))


def test_empty_and_fully_culled_inputs():
    cam = _util.make_test_camera(48, 40)
    g = _util.small_gaussians(3, 64)
    g["means3D"][:] += np.array([5.0, 0.0, 0.0], np.float32)  # behind the camera
    _, color, radii, depth, total, _ = _run_cuda(g, cam, bg=(0.25, 0.5, 0.75))
    assert total == 0 and (radii == 0).all()
    assert np.allclose(color, np.array([0.25, 0.5, 0.75], np.float32)[:, None, None]) and (depth == 15.0).all()
    import torch
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    r = BatchedRasterizer("cuda")
    z = lambda *s: torch.zeros(*s, device="cuda")
    color, radii, depth = r.forward(z(0, 3), z(0, 1), viewmatrix=torch.tensor(cam.view).cuda(),
                                    projmatrix=torch.tensor(cam.proj).cuda(), campos=torch.tensor(cam.campos).cuda(),
                                    bg=z(3), W=48, H=40, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, shs=z(0, 1, 3),
                                    scales=z(0, 3), rotations=z(0, 4))
    assert float(color.abs().max()) == 0.0 and float(depth.min()) == 15.0


def test_overflow_is_reported_and_dropin_regrows():
    import torch
    from real2sim_eval_b200.rasterizer import BatchedRasterizer, GaussianRasterizer, GaussianRasterizationSettings
    g = _util.small_gaussians(7, 3000, scale=0.05)
    cam = _util.make_test_camera(96, 96)
    r, color, _, _, total, overflow = _run_cuda(g, cam, bg=(1.0, 0.0, 0.0), max_instances=100)
    assert overflow and total > 100
    assert np.allclose(color[0], 1.0) and np.allclose(color[1:], 0.0), "overflow renders background only"
    # the drop-in path notices the overflow of its default workspace and re-runs with room
    t = lambda a: torch.tensor(a).cuda()
    rs = GaussianRasterizationSettings(96, 96, cam.tanfovx, cam.tanfovy, t(np.zeros(3, np.float32)), 1.0, t(cam.view),
                                       t(cam.proj), 0, t(cam.campos), False, cam.z_threshold)
    tg = _torchify(g)
    color, radii, depth = GaussianRasterizer(rs)(means3D=tg["means3D"], means2D=None, opacities=tg["opacities"],
                                                 shs=tg["shs"], scales=tg["scales"], rotations=tg["rotations"])
    oc, _, od, _ = _run_oracle(g, cam)
    _close_images(color.cpu().numpy(), oc, "drop-in after regrow")


def test_long_tile_lists_take_the_merge_path():
    """> 4096 instances in one tile: chunked shared-memory sort + merge-path passes."""
    g = _util.small_gaussians(11, 12000, box=((-0.05, -0.05, 0.1), (0.05, 0.05, 0.2)), scale=0.01)
    cam = _util.make_test_camera(32, 32)
    r, color, radii, depth, total, overflow = _run_cuda(g, cam, max_instances=200000)
    oc, orad, od, aux = _run_oracle(g, cam)
    off, keys, rects, _ = _super_lists(r)
    assert np.diff(off).max() > 4096, "scenario must exceed one shared-memory chunk"
    for t in range(len(off) - 1):
        seg = keys[off[t]:off[t + 1]]
        assert (np.diff(seg.view(np.int64)) > 0).all()
    if np.array_equal(radii, orad) and np.array_equal(r.intermediates()["depths"][0].cpu().numpy(), aux["depths"]):
        for t, ids in enumerate(_tile_lists(r)):
            r0, r1 = aux["ranges"][t]
            assert np.array_equal(ids, aux["point_list"][r0:r1])
    rc, rr, rd, _ = _assert_lists_equal_reference(r, g, cam, total=total)
    assert np.array_equal(color, rc) and np.array_equal(depth, rd) and np.array_equal(radii, rr)
    _close_images(color, oc, "long lists")
    _close_images(depth, od, "long lists depth")


def test_equal_depths_resolve_by_gaussian_id():
    """Many Gaussians with the SAME depth bits: the stable radix sort of the reference keeps emission
    (= id) order among them; here the bucket sort bails out (crowded bucket) and the radix path + id
    tie-break must give the same lists."""
    g = _util.small_gaussians(21, 900, scale=0.03)
    g["means3D"][100:500] = g["means3D"][100]            # 400 coincident centres -> identical depth
    g["opacities"][100:500] = 0.05
    cam = _util.make_test_camera(64, 48)
    r, color, radii, depth, total, _ = _run_cuda(g, cam)
    oc, orad, od, aux = _run_oracle(g, cam)
    dcu = r.intermediates()["depths"][0].cpu().numpy()
    assert len(np.unique(dcu[100:500])) == 1, "the coincident Gaussians share one depth bit pattern"
    if np.array_equal(radii, orad) and np.array_equal(dcu, aux["depths"]):
        for t, ids in enumerate(_tile_lists(r)):
            r0, r1 = aux["ranges"][t]
            assert np.array_equal(ids, aux["point_list"][r0:r1]), f"tile {t}"
    off, keys, rects, _ = _super_lists(r)
    for t in range(len(off) - 1):
        assert (np.diff(keys[off[t]:off[t + 1]].view(np.int64)) > 0).all(), "ties ordered by ascending id"
    _close_images(color, oc, "equal depths")
    rc, rr, rd, n = _assert_lists_equal_reference(r, g, cam, total=total)   # ties in the reference's stable-sort order
    assert np.array_equal(color, rc) and np.array_equal(depth, rd), "bit-identical to the reference CUDA rasterizer"


def test_batch_equals_single_views_and_shared_scene():
    """B views in one enqueue == B separate calls (bitwise); views_per_scene shares Gaussians."""
    import torch
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    W, H, P = 80, 48, 1200
    cams = [_util.make_test_camera(W, H, eye=(0.9, 0.05 + 0.1 * i, 0.5)) for i in range(4)]
    gs = [_util.small_gaussians(20 + s, P) for s in range(2)]           # 2 scenes x 2 cameras
    singles = []
    for b, cam in enumerate(cams):
        _, color, radii, depth, _, _ = _run_cuda(gs[b // 2], cam)
        singles.append((color, radii, depth))
    r = BatchedRasterizer("cuda")
    st = lambda key: torch.tensor(np.stack([g[key] for g in gs])).cuda()
    color, radii, depth = r.forward(
        st("means3D"), st("opacities"), viewmatrix=torch.tensor(np.stack([c.view for c in cams])).cuda(),
        projmatrix=torch.tensor(np.stack([c.proj for c in cams])).cuda(),
        campos=torch.tensor(np.stack([c.campos for c in cams])).cuda(), bg=torch.zeros(3).cuda(), W=W, H=H,
        tanfovx=cams[0].tanfovx, tanfovy=cams[0].tanfovy, shs=st("shs"), scales=st("scales"), rotations=st("rotations"),
        views_per_scene=2)
    for b in range(4):
        assert np.array_equal(color[b].cpu().numpy(), singles[b][0])
        assert np.array_equal(radii[b].cpu().numpy(), singles[b][1])
        assert np.array_equal(depth[b].cpu().numpy(), singles[b][2])


def test_known_answers_single_gaussian():
    """Centre pixel alpha = min(0.99, opacity) (forward.cu:350) and the median depth rule."""
    import torch
    cam = _util.make_test_camera(33, 33, eye=(1.0, 0.0, 0.0), target=(0.0, 0.0, 0.0))
    g = dict(means3D=np.zeros((1, 3), np.float32), scales=np.full((1, 3), 0.05, np.float32),
             rotations=np.array([[1, 0, 0, 0]], np.float32), opacities=np.array([[0.6]], np.float32),
             colors_precomp=np.array([[1.0, 0.5, 0.25]], np.float32))
    _, color, radii, depth, total, _ = _run_cuda(g, cam)
    oc, _, od, aux = _run_oracle(g, cam)
    cx, cy = np.round(aux["means2D"][0]).astype(int)
    # the projected centre is within half a pixel of (cx, cy): alpha there is opacity * exp(-small)
    assert 0.55 < color[0, cy, cx] <= 0.6 and abs(color[1, cy, cx] - 0.5 * color[0, cy, cx]) < 1e-6
    assert color.max() <= 0.6 + 1e-6
    assert depth[0, cy, cx] == np.float32(aux["depths"][0]), "T crosses 0.5 at this Gaussian -> its depth"
    assert depth[0, 0, 0] == 15.0
    _close_images(color, oc, "single gaussian")


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libref_raster.so")),
                    reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("W,H,P,seed,deg", [(64, 64, 2000, 1, 0), (160, 96, 6000, 4, 0), (96, 96, 1500, 5, 2)])
def test_matches_unmodified_reference_cuda(W, H, P, seed, deg):
    """Live comparison with the reference's own CUDA rasterizer on the same device."""
    import ref_raster
    g = _util.small_gaussians(seed, P, sh_coeffs=(deg + 1) ** 2)
    cam = _util.make_test_camera(W, H)
    r, color, radii, depth, total, _ = _run_cuda(g, cam, sh_degree=deg, bg=(0.1, 0.2, 0.3))
    rc, rr, rd, n = _assert_lists_equal_reference(r, g, cam, sh_degree=deg, bg=(0.1, 0.2, 0.3), total=total)
    assert total == n, "instance count equals the reference's num_rendered"
    assert np.array_equal(radii, rr)
    assert np.array_equal(color, rc), f"colour not bit-identical: {(color != rc).sum()} values, max |d|={np.abs(color - rc).max()}"
    assert np.array_equal(depth, rd), f"depth not bit-identical: {(depth != rd).sum()} values"
    print(f"[{W}x{H} P={P} deg={deg}] vs reference CUDA: color bit-identical on {(color == rc).mean() * 100:.4f}% of values, "
          f"max |d|={np.abs(color - rc).max():.3e}; depth identical on {(depth == rd).mean() * 100:.4f}%")
    # pin the CPU oracle against the real reference.  The oracle evaluates without FMA contraction,
    # the reference with it: a Gaussian whose 3-sigma radius sits on an integer can round the other
    # way (ceil), which changes its tile rectangle and touches a few hundred pixels by < 1e-2.
    oc, orad, od, _ = _run_oracle(g, cam, sh_degree=deg, bg=(0.1, 0.2, 0.3))
    assert (orad == rr).mean() >= 0.995
    bad = np.abs(oc - rc) > (ATOL + RTOL * np.abs(rc))
    assert bad.mean() <= 0.05 and np.abs(oc - rc).max() <= 2e-2, (bad.mean(), np.abs(oc - rc).max())


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libref_raster.so")),
                    reason="oracle/_ref not built (needs /root/reference at build time)")
def test_non_positive_tiny_and_large_opacities_match_reference():
    """The compositing kernel applies expf's power-of-two factor on the exponent field of opacity * r, which is exact
    only for a positive product that does not underflow.  Opacities outside (0, 1] -- negative, zero, denormal-small,
    above one -- must still give the reference's image: the reference skips alpha < 1/255 and clamps at 0.99."""
    import ref_raster
    W, H, P = 96, 96, 3000
    g = _util.small_gaussians(11, P)
    op = g["opacities"].copy()
    rng = np.random.default_rng(5)
    kind = rng.integers(0, 8, P)
    op[kind == 0] = -0.7
    op[kind == 1] = 0.0
    op[kind == 2] = 1e-30
    op[kind == 3] = 3.5
    op[kind == 4] = 1.0 / 255.0
    g = dict(g, opacities=op.astype(np.float32))
    cam = _util.make_test_camera(W, H)
    r, color, radii, depth, total, _ = _run_cuda(g, cam, bg=(0.3, 0.1, 0.2))
    rc, rr, rd, n = _assert_lists_equal_reference(r, g, cam, bg=(0.3, 0.1, 0.2), total=total)
    assert total == n and np.array_equal(radii, rr)
    assert np.array_equal(color, rc), f"colour not bit-identical: {(color != rc).sum()} values"
    assert np.array_equal(depth, rd), f"depth not bit-identical: {(depth != rd).sum()} values"


def test_golden_fixtures_from_reference():
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "raster_*.npz")))
    if not files:
        pytest.skip("golden fixtures not generated yet (tests/golden/make_raster_golden.py on a GPU box)")
    for f in files:
        d = np.load(f)
        g = {k: d[k] for k in ("means3D", "scales", "rotations", "opacities", "shs")}
        cam = synth.Camera(int(d["W"]), int(d["H"]), float(d["tanfovx"]), float(d["tanfovy"]), d["view"], d["proj"],
                           d["campos"], float(d["z_threshold"]))
        _, color, radii, depth, total, _ = _run_cuda(g, cam, sh_degree=int(d["sh_degree"]), bg=tuple(d["bg"]))
        assert total == int(d["num_rendered"])
        assert np.array_equal(radii, d["radii"])
        _close_images(color, d["color"], f"{os.path.basename(f)} color")
        _close_images(depth, d["depth"], f"{os.path.basename(f)} depth")


def test_full_size_properties_512():
    """BASELINE config 2 render size (512x512, 200k Gaussians/scene), 4 views: properties that need no
    oracle at this size -- per-tile lists sorted, instance totals consistent, colours in range,
    a permutation of the Gaussian order leaves the image unchanged up to depth ties."""
    import torch
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    P, W, H, B = 200_000, 512, 512, 4
    rope = synth.make_rope()
    gs = [synth.make_gaussians(1234 + e, P, rope.x) for e in range(B)]
    cams = [synth.make_camera(W, H, "side", jitter_seed=e) for e in range(B)]
    st = lambda key: torch.tensor(np.stack([getattr(g, key) for g in gs])).cuda()
    r = BatchedRasterizer("cuda")
    color, radii, depth = r.forward(
        st("means3D"), st("opacities"), viewmatrix=torch.tensor(np.stack([c.view for c in cams])).cuda(),
        projmatrix=torch.tensor(np.stack([c.proj for c in cams])).cuda(),
        campos=torch.tensor(np.stack([c.campos for c in cams])).cuda(), bg=torch.zeros(3).cuda(), W=W, H=H,
        tanfovx=cams[0].tanfovx, tanfovy=cams[0].tanfovy, shs=st("shs"), scales=st("scales"), rotations=st("rotations"),
        max_instances=8 * B * P)
    total, overflow = r.status()
    assert not overflow and total > 0
    it = r.intermediates()
    off = it["super_offset"].cpu().numpy().astype(np.int64)
    coarse = int(off[-1])
    assert total == int(it["tiles_touched"].sum()) and 0 < coarse <= total
    keys = it["keys"][:coarse].cpu().numpy()
    brk = np.zeros(coarse, bool)
    brk[off[1:-1][off[1:-1] < coarse]] = True
    assert (np.diff(keys)[~brk[1:]] > 0).all(), "every super-tile list strictly ascending in (depth, id)"
    assert torch.isfinite(color).all() and float(color.min()) >= 0.0
    assert float(depth.min()) > 0.0 and float(depth.max()) <= 15.0
    print(f"R/P = {total / (B * P):.3f}, mean tile list = {total / (B * (W // 16) * (H // 16)):.1f}, "
          f"(Gaussian, super-tile) instances / (Gaussian, tile) instances = {coarse / total:.3f}")


def test_scale_modifier_and_prefiltered_flag():
    g = _util.small_gaussians(31, 900)
    cam = _util.make_test_camera(72, 56)
    from oracle import raster_ref
    _, color, radii, depth, _, _ = _run_cuda(g, cam, scale_modifier=1.7, prefiltered=False)
    oc, orad, od = raster_ref.rasterize(g["means3D"], g["opacities"], viewmatrix=cam.view, projmatrix=cam.proj,
                                        campos=cam.campos, bg=np.zeros(3, np.float32), W=cam.W, H=cam.H,
                                        tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, shs=g["shs"], scales=g["scales"],
                                        rotations=g["rotations"], scale_modifier=1.7, z_threshold=cam.z_threshold)
    assert (radii == orad).mean() > 0.995 and radii.max() > 0
    _close_images(color, oc, "scale_modifier 1.7")
    _close_images(depth, od, "scale_modifier 1.7 depth")


def test_rgb8_output_is_the_reference_host_conversion():
    """out_rgb8 == (clamp(color, 0, 1).permute(1, 2, 0) * 255).astype(uint8), byte for byte
    (sim/renderer/gs_renderer.py:949 + experiments/eval_policy.py:248), including colours above 1
    (bright SH terms) and a non-multiple-of-16 image size."""
    import torch
    g = _util.small_gaussians(37, 3000)
    g["shs"] = (g["shs"] * 3.0).astype(np.float32)      # push some pixels outside [0, 1]
    cam = _util.make_test_camera(136, 88)
    B = 2
    rgb8 = torch.full((B, cam.H, cam.W, 3), 7, dtype=torch.uint8, device="cuda")
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    r = BatchedRasterizer("cuda")
    t = _torchify(g)
    view = torch.tensor(np.stack([cam.view] * B)).cuda()
    proj = torch.tensor(np.stack([cam.proj] * B)).cuda()
    campos = torch.tensor(np.stack([cam.campos] * B)).cuda()
    color, _, _ = r.forward(t["means3D"], t["opacities"], viewmatrix=view, projmatrix=proj, campos=campos,
                            bg=torch.tensor([0.2, 0.4, 1.5]).cuda(), W=cam.W, H=cam.H, tanfovx=cam.tanfovx,
                            tanfovy=cam.tanfovy, shs=t["shs"], scales=t["scales"], rotations=t["rotations"],
                            z_threshold=cam.z_threshold, views_per_scene=B,
                            max_instances=64 * B * len(g["means3D"]) + 4096, out_rgb8=rgb8)
    assert r.status()[1] == 0
    c = color.cpu().numpy()
    assert (c > 1.0).any() and (c < 0.0).any() or (c > 1.0).any()
    want = (np.clip(c, 0.0, 1.0).transpose(0, 2, 3, 1) * 255).astype(np.uint8)
    got = rgb8.cpu().numpy()
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        r.forward(t["means3D"], t["opacities"], viewmatrix=view, projmatrix=proj, campos=campos,
                  bg=torch.zeros(3).cuda(), W=cam.W, H=cam.H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                  shs=t["shs"], scales=t["scales"], rotations=t["rotations"], views_per_scene=B,
                  out_rgb8=torch.empty((B, cam.H, cam.W, 3), dtype=torch.float32, device="cuda"))


def test_more_than_2_pow_20_gaussians_use_the_gathered_rectangle_path():
    """Keys carry id << 12 | local rectangle only while ids fit 20 bits; beyond that the id fills the low word
    and the rectangles are gathered after the sort.  1,053,576 Gaussians (most behind the camera): the lists
    and images must still be the oracle's."""
    P_vis, P_far = 5000, (1 << 20)
    g = _util.small_gaussians(31, P_vis, scale=0.02)
    far = _util.small_gaussians(32, P_far, box=((5.0, -0.5, 0.0), (8.0, 0.5, 0.5)), scale=0.01)   # behind the eye
    order = np.random.default_rng(5).permutation(P_vis + P_far)
    g = {k: np.concatenate([g[k], far[k]])[order] for k in g}
    cam = _util.make_test_camera(64, 64)
    r, color, radii, depth, total, overflow = _run_cuda(g, cam, max_instances=400000)
    assert not overflow and r.intermediates()["id_shift"] == 0
    oc, orad, od, aux = _run_oracle(g, cam, aux=True)
    assert (orad > 0).sum() > 1000 and (orad == 0).sum() > P_far // 2
    assert np.array_equal(radii, orad)
    if np.array_equal(r.intermediates()["depths"][0].cpu().numpy(), aux["depths"]):
        for t, ids in enumerate(_tile_lists(r)):
            r0, r1 = aux["ranges"][t]
            assert np.array_equal(ids, aux["point_list"][r0:r1]), f"tile {t}"
    rc, rr, rd, _ = _assert_lists_equal_reference(r, g, cam, total=total)
    assert np.array_equal(color, rc) and np.array_equal(depth, rd) and np.array_equal(radii, rr)
    _close_images(color, oc, "P > 2^20")
    _close_images(depth, od, "P > 2^20 depth")


def _bench_scene(e, P=200_000):
    rope = synth.make_rope()
    g = synth.make_gaussians(1234 + e, P, rope.x)
    return dict(means3D=g.means3D, scales=g.scales, rotations=g.rotations, opacities=g.opacities, shs=g.shs)


@pytest.mark.parametrize("W,H,cams", [(512, 512, ("side",)), (640, 480, ("side", "wrist"))])
def test_bench_configurations_bit_identical_to_reference(W, H, cams):
    """BASELINE configs[1] (512x512) and configs[2] (640x480, two cameras) at the bench's 200 k Gaussians per scene,
    two scenes each, batched in ONE enqueue (views_per_scene = cameras) against the live reference run view by
    view: num_rendered, radii, every per-tile list, colour, depth and the uint8 HWC image must be IDENTICAL."""
    import torch
    import ref_raster
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    S = 2
    scenes = [_bench_scene(e) for e in range(S)]
    P = len(scenes[0]["means3D"])
    camlist = [synth.make_camera(W, H, c, jitter_seed=e) for e in range(S) for c in cams]
    B = len(camlist)
    st = lambda key: torch.tensor(np.stack([g[key] for g in scenes])).cuda()
    rgb8 = torch.zeros((B, H, W, 3), dtype=torch.uint8, device="cuda")
    r = BatchedRasterizer("cuda")
    color, radii, depth = r.forward(
        st("means3D"), st("opacities"), viewmatrix=torch.tensor(np.stack([c.view for c in camlist])).cuda(),
        projmatrix=torch.tensor(np.stack([c.proj for c in camlist])).cuda(),
        campos=torch.tensor(np.stack([c.campos for c in camlist])).cuda(), bg=torch.zeros(3).cuda(), W=W, H=H,
        tanfovx=[c.tanfovx for c in camlist], tanfovy=[c.tanfovy for c in camlist],   # the cameras differ in intrinsics
        shs=st("shs"), scales=st("scales"), rotations=st("rotations"), views_per_scene=len(cams),
        z_threshold=camlist[0].z_threshold, max_instances=8 * B * P, out_rgb8=rgb8)
    if len(cams) > 1:
        assert camlist[0].tanfovx != camlist[1].tanfovx
    total, overflow = r.status()
    assert not overflow
    n_sum = 0
    for b, cam in enumerate(camlist):
        rc, rr, rd, n = _assert_lists_equal_reference(r, scenes[b // len(cams)], cam, view=b)
        n_sum += n
        assert np.array_equal(radii[b].cpu().numpy(), rr), f"view {b}: radii"
        cb, db = color[b].cpu().numpy(), depth[b].cpu().numpy()
        assert np.array_equal(cb, rc), f"view {b}: {(cb != rc).sum()} colour values differ, max |d| = {np.abs(cb - rc).max()}"
        assert np.array_equal(db, rd), f"view {b}: {(db != rd).sum()} depth values differ"
        want8 = (np.clip(rc, 0.0, 1.0).transpose(1, 2, 0) * 255).astype(np.uint8)   # gs_renderer.py:949, eval_policy.py:248
        assert np.array_equal(rgb8[b].cpu().numpy(), want8), f"view {b}: uint8 image"
    assert total == n_sum, "status total = sum of the reference's num_rendered over the batch"


@pytest.mark.parametrize("W,H,cam", [(512, 512, "side"), (640, 480, "wrist")])
def test_fast_composite_variant_within_contract(W, H, cam):
    """composite_mode = FAST (log2(e) folded into the staged conic, ex2.approx instead of expf) at the bench
    configurations against the live reference: RGB and depth within 1e-4 relative per pixel (north_star's contract)
    apart from the counted threshold-flip budget; radii / lists are untouched by the mode."""
    import torch
    import ref_raster
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    g = _bench_scene(3)
    c = synth.make_camera(W, H, cam, jitter_seed=3)
    rc, rr, rd, n = ref_raster.forward(g, c)
    t = _torchify(g)
    out = {}
    for fast in (False, True):
        r = BatchedRasterizer("cuda")
        color, radii, depth = r.forward(
            t["means3D"], t["opacities"], viewmatrix=torch.tensor(c.view).cuda(), projmatrix=torch.tensor(c.proj).cuda(),
            campos=torch.tensor(c.campos).cuda(), bg=torch.zeros(3).cuda(), W=W, H=H, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
            shs=t["shs"], scales=t["scales"], rotations=t["rotations"], z_threshold=c.z_threshold,
            max_instances=8 * len(g["means3D"]), fast=fast)
        assert r.status() == (n, False) and np.array_equal(radii[0].cpu().numpy(), rr)
        out[fast] = (color[0].cpu().numpy(), depth[0].cpu().numpy())
    assert np.array_equal(out[False][0], rc) and np.array_equal(out[False][1], rd), "precise stays bit-identical"
    fc = _close_images(out[True][0], rc, "fast colour vs reference CUDA")
    fd = _close_images(out[True][1], rd, "fast depth vs reference CUDA")
    rel = np.abs(out[True][0] - rc) / (np.abs(rc) + 1e-3)
    print(f"[fast {W}x{H}] colour: max |d| = {np.abs(out[True][0] - rc).max():.2e}, 99.9th pct rel = "
          f"{np.quantile(rel, 0.999):.2e}, outlier pixels {fc:.2e} / depth {fd:.2e}")


def test_overflow_counter_is_sticky():
    """args.overflow_count: +1 per overflowing forward, surviving later good frames, read without a per-frame sync."""
    g = _util.small_gaussians(7, 3000, scale=0.05)
    cam = _util.make_test_camera(96, 96)
    r, *_ = _run_cuda(g, cam, max_instances=100)
    assert r.overflows() == 1
    import torch
    t = _torchify(g)
    kw = dict(viewmatrix=torch.tensor(cam.view).cuda(), projmatrix=torch.tensor(cam.proj).cuda(),
              campos=torch.tensor(cam.campos).cuda(), bg=torch.zeros(3).cuda(), W=cam.W, H=cam.H, tanfovx=cam.tanfovx,
              tanfovy=cam.tanfovy, shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    r.forward(t["means3D"], t["opacities"], max_instances=400000, **kw)
    assert r.status()[1] is False and r.overflows() == 1, "a good frame does not clear the sticky count"
    r.forward(t["means3D"], t["opacities"], max_instances=50, **kw)
    assert r.overflows(reset=True) == 2 and r.overflows() == 0


def test_list_equality_checks_really_ran():
    """list comparisons must not be silently skipped.  Collected last in this file: by now the live
    reference's lists have been compared tile by tile in every test that claims it."""
    assert LIST_CHECKS["calls"] >= 12 and LIST_CHECKS["tiles"] >= 6000, LIST_CHECKS
