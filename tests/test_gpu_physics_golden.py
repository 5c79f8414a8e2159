"""GPU parity against the REFERENCE's own physics source: the CUDA frame / grid kernels (through the C ABI) vs
tests/golden/phys_*.npz, the outputs of the unmodified sim/physics/spring_mass_warp.py executed under
oracle/warp_exec.py (tests/golden/make_physics_golden.py).  Contract (BASELINE.json north_star): particle
positions within 1e-5 m absolute after N substeps.

The kernel evaluates the same expressions with nvcc's FMA contraction and a per-particle (adjacency-order)
summation instead of the reference's unordered float atomics, so agreement is to rounding, not bitwise;
thresholds in the contact code (approach speed, closest-face ties) turn rounding into a different branch for a
few particles, which is bounded by a counted budget as in test_gpu_physics.py."""
import numpy as np
import pytest

import phys_cases
import r2s_testutil as util

pytestmark = pytest.mark.gpu

TOL_X = 1e-5          # the contract
# (tol_x, tol_v, outlier_frac, hard_x) per case: what is actually held
HOLD = dict(
    rope_s10=(2e-6, 5e-3, 0.0, None), tblock_s100=(5e-6, 1e-3, 0.0, None),
    two_ropes_collide=(5e-6, 2e-2, 0.01, 1e-3), chain_ground=(2e-6, 5e-3, 0.0, None),
    chain_reverse_z=(2e-6, 5e-3, 0.0, None), gripper_graze=(5e-6, 5e-3, 0.0, None),
    gripper_inside=(5e-6, 2e-2, 0.01, 2e-2), static_and_gripper=(5e-6, 2e-2, 0.01, 2e-3),
    pusher_tblock=(5e-6, 2e-2, 0.01, 2e-3), pusher_static_tblock=(5e-6, 2e-2, 0.01, 2e-3))


def _check_state(x, v, g, k, name):
    tol_x, tol_v, frac, hard = HOLD[name]
    assert tol_x <= TOL_X
    dx = np.abs(x - g[f"f{k}_x"]).max(axis=1)
    dv = np.abs(v - g[f"f{k}_v"]).max(axis=1)
    bad = (dx > tol_x) | (dv > tol_v)
    assert bad.mean() <= frac, f"{name} frame {k}: {bad.sum()} of {bad.size} particles off (|dx|max={dx.max()}, |dv|max={dv.max()})"
    if hard is not None:
        assert dx.max() <= hard
    return float(dx.max())


def _plane_groups(case):
    """Faces of the merged mesh grouped by (mesh part, supporting plane).  A contact whose closest point lies on the
    edge shared by two coplanar triangles (the diagonal of a quad) is an exact tie between them: the reference
    resolves it by BVH traversal order, a brute-force scan by face index, and 1-ulp differences in the distance
    decide it either way -- so per-face forces are compared after summing over coplanar neighbours."""
    m = phys_cases.merged_mesh(case)
    v, f = m["verts"].astype(np.float64), m["faces"]
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = (n * v[f[:, 0]]).sum(1)
    key = np.concatenate([m["mesh_map"][:, None], np.round(n * 1e4), np.round(d * 1e6)[:, None]], 1)
    _, grp = np.unique(key, axis=0, return_inverse=True)
    return grp.reshape(-1)


def _check_forces(got, want, case):
    grp = _plane_groups(case)
    ng = grp.max() + 1
    sg, sw = np.zeros((ng, 3)), np.zeros((ng, 3))
    np.add.at(sg, grp, got)
    np.add.at(sw, grp, want)
    scale = np.abs(want).max()
    assert np.abs(sg - sw).max() <= 2e-3 * scale + 1e-3, "per-plane contact forces differ from the reference's"
    same_face = np.abs(got - want).max(1) <= 2e-3 * scale + 1e-3
    assert same_face.mean() >= 0.9, "more than a few faces lost their force to a coplanar neighbour"


def _check_candidates(num, idx, g, k):
    want_num, rows = util.golden_coll_rows(g, k)
    assert np.array_equal(num, want_num), f"frame {k}: candidate counts differ from the reference's"
    for i in np.nonzero(want_num)[0]:
        assert np.array_equal(idx[i, :num[i]], rows[i]), f"frame {k}: candidate row {i} differs"


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("name", list(phys_cases.CASES))
def test_batched_kernels_match_the_reference_source(name, precise):
    """BatchedSpringMass (3 identical environments in one launch) on every golden case."""
    case, g = util.load_phys_golden(name)
    c = util.cuda_from_case(case, E=3, precise=precise, coll_row_cap=500 if name == "two_ropes_collide" else 0)
    p = phys_cases.params_of(case)
    for k, tables in enumerate(case["frames"]):
        if p["self_collision"]:
            c.update_collision_graph()
            assert int(c.status[:, 1].sum()) == 0, "candidate row overflow"
            for e in (0, 2):
                _check_candidates(c.coll_num[e].cpu().numpy(), c.coll_idx[e].cpu().numpy(), g, k)
        if tables is not None:
            c.set_mesh_motion(*tables)
        c.step()
        x, v = c.get_state()
        x, v = x.cpu().numpy(), v.cpu().numpy()
        assert np.array_equal(x[0], x[1]) and np.array_equal(x[0], x[2])
        _check_state(x[0], v[0], g, k, name)
        if case["meshes"] is not None:
            want = g[f"f{k}_collision_forces"]
            got = c.collision_forces[0].cpu().numpy()
            if HOLD[name][2] == 0.0:                        # no particle may have changed branch: forces agree per plane
                _check_forces(got, want, case)
            else:                                           # a few contacts flipped: the total impulse still agrees
                assert np.abs(got.sum(0) - want.sum(0)).max() <= 0.05 * np.abs(want).sum(0).max() + 1.0


@pytest.mark.parametrize("name", ["rope_s10", "gripper_graze", "static_and_gripper", "pusher_tblock", "chain_reverse_z",
                                  "two_ropes_collide"])
def test_dropin_class_driven_like_the_reference(name):
    """real2sim_eval_b200.physics.SpringMassSystemWarp constructed and driven by the SAME code that constructed and
    drove the reference's class when the goldens were made (phys_cases.build_sim / drive_frame: the calls of
    sim/physics/phystwin.py:336-357, 365-366, 455-460, 515-519, with wp.capture_launch of `simulator.graph`)."""
    from real2sim_eval_b200.physics import SpringMassSystemWarp
    from real2sim_eval_b200.compat import warp as wp
    case, g = util.load_phys_golden(name)
    sim = phys_cases.build_sim(case, SpringMassSystemWarp, "cuda:0", use_graph=True)
    if case["meshes"] is not None:
        assert np.array_equal(sim.mesh_map.numpy(), g["mesh_map"])
        assert np.array_equal(sim.face_map.numpy(), g["face_map"])
        assert sim.collision_forces.numpy().shape == (len(g["face_map"]), 3)
    for k in range(len(case["frames"])):
        phys_cases.drive_frame(sim, case, k, wp, "cuda:0", use_graph=True)
        if phys_cases.params_of(case)["self_collision"]:
            assert sim.wp_collision_indices.numpy().shape[1] == 500 and sim.collision_row_overflow == 0
            _check_candidates(sim.wp_collision_number.numpy(), sim.wp_collision_indices.numpy(), g, k)
        x = wp.to_torch(sim.wp_state.wp_x).cpu().numpy()
        v = wp.to_torch(sim.wp_state.wp_v).cpu().numpy()
        _check_state(x, v, g, k, name)
        if case["meshes"] is not None and HOLD[name][2] == 0.0:
            _check_forces(sim.collision_forces.numpy(), g[f"f{k}_collision_forces"], case)


@pytest.mark.parametrize("mesh_accel", [-1, 1])
def test_tool_beside_a_static_obstacle_brute_force_and_grid(mesh_accel):
    """VERDICT r1 missing #7: a large rigid tool TOGETHER with static meshes.  mesh_accel = 1 searches the tool through
    its rest-frame grid and scans the static faces beside it (nearer hit wins, inside either mesh is inside);
    mesh_accel = -1 scans every face.  Both against the reference's own kernel source on the merged mesh."""
    case, g = util.load_phys_golden("pusher_static_tblock")
    c = util.cuda_from_case(case, mesh_accel=mesh_accel)
    free, _ = util.load_phys_golden("pusher_tblock")
    for k, tables in enumerate(case["frames"]):
        c.update_collision_graph(); c.set_mesh_motion(*tables); c.step()
        x, v = c.get_state()
        _check_state(x[0].cpu().numpy(), v[0].cpu().numpy(), g, k, "pusher_static_tblock")
    gf = util.load_phys_golden("pusher_tblock")[1]
    assert np.abs(g["f0_x"] - gf["f0_x"]).max() > 1e-4, "the obstacle must have changed the outcome"
    want = g["f0_collision_forces"]
    got = c.collision_forces[0].cpu().numpy()
    n_tool = int((g["mesh_map"] >= 0).sum())
    assert np.abs(want[n_tool:]).max() > 0, "the static faces must carry contact force"
    assert np.abs(got[n_tool:].sum(0) - want[n_tool:].sum(0)).max() <= 0.05 * np.abs(want[n_tool:]).sum(0).max() + 1.0


def test_grasp_force_faces_follow_the_requery():
    """SMW:397 re-assigns `query`, so a finger contact books its force on the face of the re-query (SMW:414);
    phystwin.py:386-391 reads faces [18], [19], [1] of each finger from that array.  The golden's per-face
    forces come from the reference source: agreement per supporting plane in every frame (exact ties between
    coplanar triangles aside, see _plane_groups) and per face in the tie-free first frame pins the attribution."""
    case, g = util.load_phys_golden("gripper_graze")
    c = util.cuda_from_case(case, precise=True)
    for k, tables in enumerate(case["frames"]):
        c.update_collision_graph(); c.set_mesh_motion(*tables); c.step()
        want = g[f"f{k}_collision_forces"]
        got = c.collision_forces[0].cpu().numpy()
        assert (np.abs(want).max(1) > 0).sum() >= 4
        _check_forces(got, want, case)
        if k == 0:   # no exact tie in this frame: every face carries the reference's force (first-query booking
            #          moves the force of faces 4/5 and 60/61 to their coplanar neighbour and fails here)
            assert np.abs(got - want).max() <= 2e-3 * np.abs(want).max()


def test_dense_contact_exceeds_the_compact_row_capacity():
    """the reference's candidate rows hold 500 entries (SMW:544-549); the batched default is 64.
    Four interleaved copies of a 500-particle blob produce rows of up to ~240 entries: the compact handle must REPORT the overflow
    (status[:,1] > 0) and a handle with coll_row_cap=500 must reproduce the oracle's rows."""
    from real2sim_eval_b200 import synth
    rng = np.random.default_rng(11)
    n, copies = 500, 4
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    blob = (d * (0.010 * rng.uniform(0, 1, (n, 1)) ** (1 / 3)) + np.array([0.3, 0.0, 0.05])).astype(np.float32)
    springs1, rest1 = synth.build_springs(blob.astype(np.float64), blob.astype(np.float64), 0.004, 10)
    far = np.concatenate([blob + np.float32([0.1 * c, 0, 0]) for c in range(copies)], 0).astype(np.float32)
    near = np.concatenate([blob + np.float32([0.0004 * c, 0.0003 * c, 0.0002 * c]) for c in range(copies)], 0).astype(np.float32)
    springs = np.concatenate([springs1 + n * c for c in range(copies)], 0)
    cat = lambda a: np.concatenate([a] * copies, 0)
    sc = synth.Scene("dense", far, np.zeros_like(far), springs, cat(rest1),
                     np.full(len(springs), np.log(np.float32(3e4)), np.float32), np.ones(n * copies, np.float32),
                     dict(synth.DEFAULT_PARAMS))
    o = util.oracle_from_scene(sc, 2)
    o.x[:] = near
    o.update_collision_graph()
    assert o.coll_num.max() > 64 and o.coll_num.max() <= 500, o.coll_num.max()
    small = util.cuda_from_scenes([sc], 2)
    small.set_state(near[None], np.zeros_like(near)[None])
    small.update_collision_graph()
    assert int(small.status[0, 1]) > 0, "the compact handle must report dropped candidates"
    big = util.cuda_from_scenes([sc], 2, coll_row_cap=500)
    big.set_state(near[None], np.zeros_like(near)[None])
    big.update_collision_graph()
    assert int(big.status[0, 1]) == 0
    num = big.coll_num[0].cpu().numpy()
    assert np.array_equal(num, o.coll_num)
    idx = big.coll_idx[0].cpu().numpy()
    for i in np.argsort(-num)[:100]:
        assert np.array_equal(idx[i, :num[i]], o.coll_idx[i, :num[i]])
    o.step(); big.step()
    x, _ = big.get_state()
    dx = np.abs(x[0].cpu().numpy() - o.x).max(1)
    assert (dx > 5e-6).mean() <= 0.01
