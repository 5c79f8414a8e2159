"""CPU tests of the physics oracle (oracle/physics_ref.c): known-answer tests derived from the
formulas of sim/physics/spring_mass_warp.py, and the shipped T-block graph.  The reference holds
no golden vectors for this path, so these KATs are what pins the restatement."""
import numpy as np
import pytest

import r2s_testutil as _util
from real2sim_eval_b200 import synth
from oracle.physics_ref import SpringMassOracle, build_csr, step_batch

F32 = np.float32


def _two(dist, rest, Y=3e4, v=(0, 0, 0), z=1.0, **kw):
    x = np.array([[0, 0, z], [dist, 0, z]], F32)
    vv = np.array([v, [0, 0, 0]], F32)
    return SpringMassOracle(x, vv, np.array([[0, 1]], np.int32), np.array([rest], F32), np.log(np.array([Y], F32)),
                            np.ones(2, F32), self_collision=False, **kw)


def test_one_substep_spring_force_closed_form():
    """SMW:87-129: F = clamp(Y)*(len/rest - 1)*d + damp*dot(v2-v1, d)*d; v' = (v + (F/m + g) dt)*exp(-dt*drag)."""
    dt, damp, drag, Y = 5e-5, 100.0, 3.0, 3e4
    o = _two(0.012, 0.010, Y=Y, v=(0.3, 0, 0), n_substeps=1, dt=dt, dashpot_damping=damp, drag_damping=drag)
    o.step()
    F = Y * (0.012 / 0.010 - 1.0) + damp * (0.0 - 0.3)          # along +x on particle 0
    fac = np.exp(-dt * drag)
    v0x = (0.3 + F * dt) * fac
    v1x = (0.0 - F * dt) * fac
    vz = (-9.8 * dt) * fac
    assert o.v[0, 0] == pytest.approx(v0x, rel=2e-6) and o.v[1, 0] == pytest.approx(v1x, rel=2e-6)
    assert o.v[0, 2] == pytest.approx(vz, rel=1e-6)
    assert o.x[0, 0] == pytest.approx(0.0 + v0x * dt, abs=1e-9)   # semi-implicit: x advances with the NEW velocity


def test_inactive_springs_are_skipped_and_stiffness_is_clamped():
    """exp(Y) > Y_min guard (SMW:75) drops the whole spring (dashpot too); clamp to Y_max (SMW:93)."""
    o = _two(0.012, 0.010, Y=50.0, v=(0.3, 0, 0), n_substeps=1, spring_Y_min=100.0)
    o.step()
    assert o.v[0, 0] == pytest.approx(0.3 * np.exp(-5e-5 * 3.0), rel=1e-6)
    a = _two(0.012, 0.010, Y=1e7, n_substeps=1)                 # above Y_max = 1e5 -> clamped
    b = _two(0.012, 0.010, Y=1e5, n_substeps=1)
    a.step(); b.step()
    assert np.allclose(a.v, b.v, rtol=1e-6)                     # exp(log(1e5)) is 1e5 only to rounding
    c = _two(0.012, 0.010, Y=2e5, n_substeps=1)
    c.step()
    assert np.array_equal(a.v, c.v)                             # both clamped to exactly Y_max


def test_free_fall_and_ground_bounce_with_time_of_impact():
    """SMW:424-474: bounce when the advanced z < 0 and v_z < -1e-4; x = x + v0*toi + v1*(dt - toi)."""
    dt, e, mu = 5e-5, 0.5, 0.3
    x = np.array([[0, 0, 1e-5]], F32)
    v = np.array([[0.2, 0, -1.0]], F32)
    o = SpringMassOracle(x, v, np.zeros((0, 2), np.int32), np.zeros(0, F32), np.zeros(0, F32), np.ones(1, F32),
                         n_substeps=1, self_collision=False, collide_elas=e, collide_fric=mu, drag_damping=0.0)
    o.step()
    vz0 = -1.0 - 9.8 * dt
    toi = 1e-5 / -vz0
    a = max(0.0, 1.0 - mu * (1 + e) * abs(vz0) / 0.2)
    assert o.v[0, 2] == pytest.approx(-e * vz0, rel=1e-5)
    assert o.v[0, 0] == pytest.approx(a * 0.2, rel=1e-5)
    assert o.x[0, 2] == pytest.approx(1e-5 + vz0 * toi + (-e * vz0) * (dt - toi), abs=1e-9)
    assert o.x[0, 2] >= 0.0


def test_two_particle_oscillation_period_and_damping():
    """Strain-form Hooke spring: relative motion obeys m_red*u'' = -(Y/rest)*u - damp*u' (m_red = 1/2)."""
    rest, Y, damp = 0.01, 3e4, 1.0
    n = 4000
    o = _two(0.0101, rest, Y=Y, z=5.0, n_substeps=1, dashpot_damping=damp, drag_damping=0.0)
    sep = []
    for _ in range(n):
        o.step()
        sep.append(o.x[1, 0] - o.x[0, 0] - rest)
    sep = np.asarray(sep, np.float64)
    zc = np.nonzero(np.diff(np.sign(sep)) != 0)[0]
    period = 2 * np.mean(np.diff(zc)) * 5e-5
    omega = np.sqrt(2 * Y / rest - damp ** 2)                    # underdamped pair, reduced mass 1/2
    assert period == pytest.approx(2 * np.pi / omega, rel=0.02)
    assert np.abs(sep[-200:]).max() < np.abs(sep[:200]).max()     # the dashpot removes energy


def test_gather_order_equals_scatter_order_to_rounding():
    sc = synth.make_rope(v_scale=0.05)
    a = _util.oracle_from_scene(sc, 5, gather_order=True, self_collision=False)
    b = _util.oracle_from_scene(sc, 5, gather_order=False, self_collision=False)
    a.step(); b.step()
    assert np.abs(a.x - b.x).max() < 2e-7 and np.abs(a.v - b.v).max() < 2e-3
    row_ptr, nbr, sid = build_csr(sc.N, sc.springs)
    assert row_ptr[-1] == 2 * sc.S and (np.diff(row_ptr) == np.bincount(sc.springs.reshape(-1), minlength=sc.N)).all()
    for i in (0, 17, sc.N - 1):
        seg = slice(row_ptr[i], row_ptr[i + 1])
        assert (np.diff(sid[seg]) > 0).all()                      # ascending spring index per particle
        assert all(i in sc.springs[t] and nbr_ in sc.springs[t] for t, nbr_ in zip(sid[seg], nbr[seg]))


def test_tblock_shipped_state_stays_at_rest():
    """The real T-block graph at its shipped rest state (max strain 1.35e-3): bounded motion."""
    sc = synth.load_tblock()
    assert (sc.N, sc.S) == (2229, 63100)
    o = _util.oracle_from_scene(sc, 200)
    o.update_collision_graph()
    assert o.coll_num.sum() == 0, "all close pairs are resting pairs at reset"
    o.step()
    assert np.isfinite(o.x).all() and np.abs(o.x - sc.x).max() < 2e-3 and o.x[:, 2].min() > -1e-6


def test_hash_grid_candidates_match_brute_force():
    """update_potential_collision (SMW:196-227): candidates of i = non-resting j != i closer than
    collision_dist, whatever the grid iteration order."""
    rng = np.random.default_rng(0)
    n = 400
    x = rng.uniform(-0.03, 0.03, (n, 3)).astype(F32)              # straddles the origin (cell truncation)
    o = SpringMassOracle(x + F32(1.0), np.zeros_like(x), np.zeros((0, 2), np.int32), np.zeros(0, F32), np.zeros(0, F32),
                         np.ones(n, F32), n_substeps=1)
    assert o.resting.sum() > 0 and np.array_equal(o.resting, o.resting.T)
    o.x[:] = x                                                    # move everything: new neighbours appear
    o.update_collision_graph()
    d = np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1, dtype=F32))
    for i in range(n):
        want = {j for j in range(n) if j != i and not o.resting[i, j] and d[i, j] < F32(0.005)}
        got = set(o.coll_idx[i, :o.coll_num[i]].tolist())
        assert got == want, i


def test_resting_pairs_need_no_distance_test():
    """build_resting_collision_pairs (SMW:272-291) marks EVERY grid-query neighbour j < i."""
    x = np.array([[0.001, 0.001, 0.001], [0.024, 0.024, 0.024], [0.2, 0.2, 0.2]], F32)  # 0,1 share a cell, 40 mm apart
    o = SpringMassOracle(x, np.zeros_like(x), np.zeros((0, 2), np.int32), np.zeros(0, F32), np.zeros(0, F32),
                         np.ones(3, F32), n_substeps=1)
    assert o.resting[0, 1] and o.resting[1, 0] and not o.resting[0, 2]


def test_self_collision_impulse_two_particles():
    """object_collision (SMW:132-268): head-on approach, equal masses: v_i -= J/m with
    J = -(1+e) v_rel_n / (1/m1 + 1/m2)."""
    x = np.array([[0, 0, 1.0], [0.004, 0, 1.0]], F32)
    v = np.array([[0.5, 0, 0], [-0.5, 0, 0]], F32)
    o = SpringMassOracle(x, v, np.zeros((0, 2), np.int32), np.zeros(0, F32), np.zeros(0, F32), np.ones(2, F32),
                         n_substeps=1, collide_self_elas=0.5, collide_self_fric=0.3, drag_damping=0.0)
    o.resting[:] = 0                                               # not a resting pair
    o.update_collision_graph()
    assert o.coll_num.tolist() == [1, 1]
    o.step()
    # relative normal velocity -1.0 -> impulse (1+e)*1.0/2 on each: 0.5 - 0.75 = -0.25
    assert o.v[0, 0] == pytest.approx(-0.25, abs=1e-6) and o.v[1, 0] == pytest.approx(0.25, abs=1e-6)


def test_mesh_query_sign_and_push_out():
    """mesh_collision (SMW:295-421): a particle 3 mm outside a finger face (margin 5 mm) is stopped and
    projected to the margin; a particle inside the finger gets sign -1 from the winding number."""
    g = synth.make_gripper(center=(0.0, 0.0, 0.0), gap=0.03)
    mesh = _util.gripper_mesh_dict(g)
    left_face_y = g.verts[:24, 1].max()                            # inner face of the left finger
    x = np.array([[0.0, left_face_y + 0.003, 0.02], [0.0, g.verts[:24, 1].mean(), 0.02]], F32)
    v = np.array([[0, -0.5, 0], [0, 0, 0]], F32)
    kw = dict(n_substeps=1, self_collision=False, drag_damping=0.0)
    o = SpringMassOracle(x, v, np.zeros((0, 2), np.int32), np.zeros(0, F32), np.zeros(0, F32), np.ones(2, F32),
                         mesh=mesh, **kw)
    o.step()
    assert o.v[0, 1] > -1e-3, "eef elasticity 0: the normal velocity is removed"
    assert o.x[0, 1] >= left_face_y + 0.005 - 2e-5, "projected out to the 5 mm margin"
    assert np.abs(o.collision_forces).max() > 0
    o1 = SpringMassOracle(x, v, np.zeros((0, 2), np.int32), np.zeros(0, F32), np.zeros(0, F32), np.ones(2, F32),
                          mesh=mesh, sign_mode=1, **kw)
    o1.step()
    assert not np.allclose(o.x[1], o1.x[1]), "inside the finger the exact winding sign differs from always-outside"


def test_openmp_batch_equals_serial():
    sc = synth.make_rope(v_scale=0.02)
    scenes = [synth.pose_scene(sc, 5 + e) for e in range(6)]
    a = [_util.oracle_from_scene(s, 4) for s in scenes]
    b = [_util.oracle_from_scene(s, 4) for s in scenes]
    step_batch(a)
    for o in b:
        o.step()
    assert all(np.array_equal(p.x, q.x) and np.array_equal(p.v, q.v) for p, q in zip(a, b))
