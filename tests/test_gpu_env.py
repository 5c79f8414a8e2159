"""GPU tests of the batched env step (physics -> re-bind -> render) at the shapes of BASELINE configs 3
and 4: two cameras per env sharing one Gaussian set, the sloth and T-block particle counts."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sloth_two_cameras_640x480():
    """configs[2] shape (sloth soft body, 640x480, two-camera render), reduced env / Gaussian counts."""
    import torch
    from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig
    cfg = EnvBatchConfig(scene="sloth", E=3, W=640, H=480, cameras=2, n_substeps=10, P=30000)
    env = BatchedEnv(cfg, "cuda")
    assert env.phys.smem_state, "3500 particles x 44 B still fit the 227 KB shared memory"
    acts = [tuple(torch.tensor(a).cuda() for a in env.make_actions(f)) for f in range(3)]
    x0 = env.phys.get_state()[0].clone()
    rest_rows = env.means3D[:, env.n_obj:env.n_obj + env.n_robot].clone()
    for f, m in enumerate(acts):
        lp = torch.tensor(env.make_link_poses(f + 1)).cuda()
        color, depth = env.step(m, link_pose=lp)
    # robot rows follow this frame's link poses (N2) and match the oracle for env 2
    from oracle import links_ref
    sc = env.scan
    p, q = links_ref.transform_gs(sc.points, sc.quats, sc.link_id, lp[2].cpu().numpy(), sc.base_pose, sc.link_offset)
    rows = slice(env.n_obj, env.n_obj + env.n_robot)
    assert np.abs(env.means3D[2, rows].cpu().numpy() - p).max() < 2e-6
    assert np.abs(env.rotations[2, rows].cpu().numpy() - q).max() < 2e-6
    assert not torch.equal(env.means3D[:, rows], rest_rows)
    total, overflow = env.raster.status()
    assert not overflow and total > 0
    assert tuple(color.shape) == (6, 3, 480, 640) and tuple(depth.shape) == (6, 1, 480, 640)
    assert torch.isfinite(color).all() and float(color.min()) >= 0.0 and float(depth.max()) <= 15.0
    assert not torch.equal(color[0], color[1]), "the two cameras of one env see different images"
    assert env.cams[0].tanfovx != env.cams[1].tanfovx, "and differ in intrinsics: one (tanfovx, tanfovy) per view"
    env.check()
    x1 = env.phys.get_state()[0]
    assert torch.isfinite(x1).all() and float((x1 - x0).abs().max()) > 1e-6 and float(x1[..., 2].min()) > -1e-6
    # object Gaussians follow the particles (LBS): they stay within a few mm of their bound particles
    idx0 = env.lbs.weights_indices[:, 0].long()
    gap = (env.means3D[:, :env.n_obj] - x1[:, idx0]).norm(dim=-1)
    assert float(gap.max()) < 0.02 and int(env.lbs.rank_flags.min()) == 1
    # a view rendered alone equals the same view inside the batch
    from real2sim_eval_b200.rasterizer import BatchedRasterizer
    r = BatchedRasterizer("cuda")
    c1, _, d1 = r.forward(env.means3D[1:2], env.opacities[1:2], viewmatrix=env.view[3:4], projmatrix=env.proj[3:4],
                          campos=env.campos[3:4], bg=env.bg, W=640, H=480, tanfovx=env.cams[3].tanfovx,
                          tanfovy=env.cams[3].tanfovy, shs=env.shs[1:2], scales=env.scales[1:2],
                          rotations=env.rotations[1:2], max_instances=2_000_000)
    assert torch.equal(c1[0], color[3]) and torch.equal(d1[0], depth[3])


def test_tblock_1024_envs_physics_only():
    """configs[3] shape: the real T-block graph x 1024 envs, physics only (no pusher mesh -- see DESIGN.md §8):
    identical envs stay bit-identical, the shipped rest state stays at rest, env 517 matches the oracle."""
    import torch
    import r2s_testutil as _util
    from real2sim_eval_b200 import synth
    sc = synth.load_tblock(v_scale=0.02)
    c = _util.cuda_from_scenes([sc] * 1024, 10, per_env_rest=False)
    c.update_collision_graph(); c.step()
    x, v = c.get_state()
    assert torch.equal(x[0].expand_as(x), x) and torch.isfinite(x).all()
    o = _util.oracle_from_scene(sc, 10)
    o.update_collision_graph(); o.step()
    assert np.abs(x[517].cpu().numpy() - o.x).max() <= 2e-6


def test_command_driven_env_closed_loop_matches_the_oracles():
    """N3 in the loop: the env is driven by end-effector COMMANDS (19 floats per env and frame); vertex tables and
    the grasp hysteresis are made on the device from last frame's finger forces.  The CPU side closes the same
    loop with oracle/eef_ref.py + oracle/physics_ref.c."""
    import torch
    import r2s_testutil as util
    from oracle import eef_ref
    from real2sim_eval_b200 import synth
    from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig
    cfg = EnvBatchConfig(scene="rope", E=3, W=64, H=64, cameras=1, n_substeps=10, P=3000)
    env = BatchedEnv(cfg, "cuda")
    e = 1
    x0, v0 = env.phys.get_state()
    sc = synth.pose_scene(env.base, cfg.seed + e)
    o = util.oracle_from_scene(sc, cfg.n_substeps, mesh=util.gripper_mesh_dict(env.gripper))
    assert np.array_equal(o.x, x0[e].cpu().numpy())
    o.create_resting_case() if env.phys.self_collision else None
    table = env.eef.table.cpu().numpy()
    faces = eef_ref.force_faces(env.gripper.mesh_map)
    cur, grasped = None, False
    for f in range(3):
        cmd = env.make_commands(f)
        env.step(command=tuple(torch.tensor(a).cuda().contiguous() for a in cmd))
        xyz, vel, rot, rvel, opn = (a[e] for a in cmd)
        r = eef_ref.eef_step(table, env.eef_init, xyz, vel, rot, rvel, opn, dt=env.dt, n_substeps=cfg.n_substeps,
                             current_openness=cur, grasped=grasped, forces=o.collision_forces, faces=faces)
        cur, grasped = r["current_openness"], r["grasped"]
        o.set_mesh_interactive(r["interp_pts"], r["interp_center"], r["dyn_vel"], r["dyn_omega"])
        if env.phys.self_collision:
            o.update_collision_graph()
        o.step()
        assert float(env.eef.current_openness[e]) == cur and bool(env.eef.grasped[e]) == grasped
        x = env.phys.get_state()[0][e].cpu().numpy()
        off = np.abs(x - o.x).max(1) > 1e-5
        assert off.mean() <= 0.01, f"frame {f}: {off.sum()} particles beyond 1e-5 m"   # contact ties, see DESIGN.md §2
    assert torch.isfinite(env.color).all()


def test_pusher_env_driven_by_commands_matches_the_oracles():
    """The push-T tool in the batched env: a rigid rod (use_pusher) whose per-substep vertex table comes from the
    device-side end-effector step (opening fixed at 1.0, one velocity row); the CPU side runs oracle/eef_ref.py
    + oracle/physics_ref.c with the same commands."""
    import torch
    import r2s_testutil as util
    from oracle import eef_ref
    from real2sim_eval_b200 import synth
    from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig
    cfg = EnvBatchConfig(scene="tblock", E=2, W=64, H=64, n_substeps=20, P=3000, gripper=False, pusher=True,
                         pusher_res=(24, 16))
    env = BatchedEnv(cfg, "cuda")
    assert env.phys.use_pusher and len(env.gripper.faces) == 816
    e = 1
    x0 = env.phys.get_state()[0][e].cpu().numpy()
    sc = synth.pose_scene(env.base, cfg.seed + e)
    o = util.oracle_from_scene(sc, cfg.n_substeps, mesh=util.gripper_mesh_dict(env.gripper), use_pusher=True,
                               collide_eef_fric=0.2)
    assert np.array_equal(o.x, x0)
    if env.phys.self_collision:
        o.create_resting_case()
    o_free = util.oracle_from_scene(sc, cfg.n_substeps)
    table = env.eef.table.cpu().numpy()
    for f in range(2):
        cmd = env.make_commands(f)
        env.step(command=tuple(None if a is None else torch.tensor(a).cuda().contiguous() for a in cmd))
        xyz, vel, rot, rvel = (a[e] for a in cmd[:4])
        r = eef_ref.eef_step(table, env.eef_init, xyz, vel, rot, rvel, 1.0, dt=env.dt, n_substeps=cfg.n_substeps,
                             use_pusher=True)
        o.set_mesh_interactive(r["interp_pts"], r["interp_center"], r["dyn_vel"], r["dyn_omega"])
        if env.phys.self_collision:
            o.update_collision_graph(); o_free.update_collision_graph()
        o.step(); o_free.step()
        x = env.phys.get_state()[0][e].cpu().numpy()
        off = np.abs(x - o.x).max(1) > 1e-5
        assert off.mean() <= 0.01, f"frame {f}: {off.sum()} particles beyond 1e-5 m"
    assert np.abs(o.x - o_free.x).max() > 1e-4, "the rod must push the block"
    assert float(env.eef.current_openness[e]) == 1.0
    assert torch.isfinite(env.color).all()


def test_success_window_numbers_frames_like_the_reference_pickles():
    """the reference saves state/{cnt:06d}.pkl BEFORE env.step (experiments/eval_policy.py:209-225), so file
    k holds the state after k steps and the success scripts count files k >= start.  BatchedEnv must book the state a
    step produces as frame (steps so far), not (steps so far - 1): with start_frame 3 and a T-block resting on its
    target, five steps give 3 hits (files 3, 4, 5) -- the episode rule of oracle/metrics_ref.py."""
    from oracle import metrics_ref
    from real2sim_eval_b200.envs import BatchedEnv, EnvBatchConfig
    cfg = EnvBatchConfig(scene="tblock", E=1, W=64, H=64, n_substeps=4, P=2000, gripper=False, pusher=False,
                         success_start_frame=3)
    env = BatchedEnv(cfg, "cuda")
    env.success.target.copy_(env.x_init[0])            # the block rests on its target: every frame passes
    for _ in range(5):
        env.step()
    succ, hits = env.success.result()
    want = metrics_ref.episode_rule([True] * 6, start_frame=3)[5][0]      # files 0..5, file 0 is the pre-step state
    assert want == 3 and int(hits[0]) == want and not succ.any()
    env.check()


@pytest.mark.parametrize("tool", ["gripper", "pusher"])
def test_device_loop_matches_the_complete_reference_stack(tool):
    """tests/golden/stack_*.npz: positions returned by the reference's unmodified phystwin.py + spring_mass_warp.py
    (CPU, under oracle/warp_exec.py; tests/golden/make_stack_golden.py) after each of three closed-loop frames.
    Here the same commands go through the DEVICE path -- r2s_eef_forward (tables + grasp hysteresis from last
    frame's finger forces) writing into the physics handle, then r2s_phys_step -- and must land on the reference's
    positions and grasp state."""
    import hashlib
    import torch
    import phys_cases
    import ref_harness
    from real2sim_eval_b200.eef import BatchedEefMotion
    from real2sim_eval_b200.physics import BatchedSpringMass
    use_pusher = tool == "pusher"
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"stack_{tool}.npz"))
    sn = ref_harness.stack_scenario(use_pusher)
    sc = sn["scene"]
    assert hashlib.sha256(np.ascontiguousarray(sc.springs).tobytes()).hexdigest() == str(g["springs_sha"])
    # rest lengths as the reference module makes them (torch.linalg.norm on the float32 cloud, phystwin.py:285-286)
    pts = torch.tensor(sc.x.astype(np.float64), dtype=torch.float32)
    spr = torch.tensor(sc.springs.astype(np.int64))
    rest = torch.linalg.norm(pts[spr[:, 0]] - pts[spr[:, 1]], dim=1).numpy()
    # (bitwise equal to the golden's on the build box -- rest_sha -- but a host with another SIMD width may round the
    # last place differently; 1 ulp of a rest length moves positions by ~1e-7 m, far inside the tolerance below)
    assert np.allclose(rest, sc.rest, rtol=3e-7, atol=0)
    p = dict(sc.params)
    if use_pusher:
        p["collide_eef_fric"] = 0.2
    phys = BatchedSpringMass(1, sc.springs, rest, num_particles=sc.N, n_substeps=sn["S"], log_spring_Y=sc.log_Y,
                             masses=sc.mass, use_pusher=use_pusher, precise=True, coll_row_cap=500, **p)
    phys.set_state(sc.x[None], sc.v[None])
    m = phys_cases.merged_mesh(dict(meshes=dict(dynamic=sn["meshes"], static=[])))
    phys.set_mesh(**m)
    phys.create_resting_case()
    eef = BatchedEefMotion(1, sn["table"], sn["center"], dt=sn["dt"], n_substeps=sn["S"], use_pusher=use_pusher,
                           mesh_map=None if use_pusher else m["mesh_map"], phys=phys)
    t = lambda a: torch.tensor(np.asarray(a, np.float32)[None]).cuda().contiguous()
    budget = 0.03 if use_pusher else 0.01
    for f, (xyz, vel, rot, rvel, opn) in enumerate(sn["commands"]):
        phys.update_collision_graph()
        eef.forward(t(xyz), t(vel), t(rot), t(rvel), None if use_pusher else t([opn]).reshape(1))
        phys.step()
        x = phys.get_state()[0][0].cpu().numpy()
        dx = np.abs(x - g["x"][f]).max(1)
        assert (dx > 1e-5).mean() <= budget and dx.max() < 2e-3, f"frame {f}: |dx|max {dx.max()}, {(dx > 1e-5).sum()} particles"
    if not use_pusher:
        assert float(eef.current_openness[0]) == pytest.approx(float(g["current_openness"]), abs=1e-7)
        assert bool(eef.grasped[0]) == bool(g["grasped"])
