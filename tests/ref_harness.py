"""TEST INFRASTRUCTURE: runs the reference's OWN sim/physics/phystwin.py against this repository's CUDA
SpringMassSystemWarp through real2sim_eval_b200.compat.install() -- the "sim/envs and experiments/eval_policy* run
unchanged" claim, executed.

The reference tree is not part of this repository and is not present on the GPU box, so the root is looked up in
$R2S_REFERENCE_ROOT, then /root/reference; tests that need it skip when neither holds sim/physics/phystwin.py.
(To execute on a GPU box: stage a throw-away, untracked copy of the two reference files under a gpurun-visible
path, point R2S_REFERENCE_ROOT at it for that one run, delete it afterwards -- profiles/r02_dropin_run.txt records
such a run.)

What is faked, and only because the packages are not installable here: kornia's
axis_angle_to_rotation_matrix (restated, as in oracle/eef_ref.py), open3d's PointCloud / KDTreeFlann hybrid search
(scipy cKDTree: nearest first, radius- and count-bounded), and the sapien/urdfpy-backed RobotPcSampler / KinHelper
(stand-ins that hand out the synthetic finger meshes).  The PhysTwin checkpoint files the constructor unpickles are
written to a temporary directory with synthetic contents of the documented structure (phystwin.py:231-298).
"""
import os
import pickle
import sys
import types

import numpy as np

F32 = np.float32


def reference_root():
    for root in (os.environ.get("R2S_REFERENCE_ROOT"), "/root/reference"):
        if root and os.path.exists(os.path.join(root, "sim", "physics", "phystwin.py")):
            return root
    return None


def _aa2rm_torch(aa):
    """kornia.geometry.conversions.axis_angle_to_rotation_matrix on torch tensors (any device), float32."""
    import torch
    theta2 = (aa * aa).sum(1)
    theta = torch.sqrt(theta2)
    w = aa / (theta + 1e-6)[:, None]
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s = torch.cos(theta), torch.sin(theta)
    k = 1.0 - c
    normal = torch.stack([c + wx * wx * k, wx * wy * k - wz * s, wy * s + wx * wz * k,
                          wz * s + wx * wy * k, c + wy * wy * k, -wx * s + wy * wz * k,
                          -wy * s + wx * wz * k, wx * s + wy * wz * k, c + wz * wz * k], -1)
    one = torch.ones_like(theta2)
    rx, ry, rz = aa[:, 0], aa[:, 1], aa[:, 2]
    taylor = torch.stack([one, -rz, ry, rz, one, -rx, -ry, rx, one], -1)
    return torch.where((theta2 > 1e-6)[:, None], normal, taylor).reshape(-1, 3, 3)


class _Mesh:
    def __init__(self, vertices, triangles):
        self.vertices, self.triangles = np.asarray(vertices), np.asarray(triangles)


class FakeRobotPcSampler:
    """Stand-in for sim/utils/robot/robot_pc_sampler.py:RobotPcSampler (sapien + urdfpy): the two mesh getters
    SpringMassDynamicsModule.__init__ calls (phystwin.py:318-325)."""

    def __init__(self, gripper=None, pusher=None):
        self._g, self._p = gripper, pusher

    def get_xarm_gripper_meshes(self, gripper_openness=1.0):
        g = self._g
        half, fh = len(g.verts) // 2, len(g.faces) // 2
        return [_Mesh(g.verts[:half], g.faces[:fh]), _Mesh(g.verts[half:], g.faces[fh:] - half)]

    def get_xarm_pusher_meshes(self):
        return [_Mesh(self._p.verts, self._p.faces)]


def _fake_open3d():
    from scipy.spatial import cKDTree
    o3d = types.ModuleType("open3d")
    o3d.geometry, o3d.utility = types.SimpleNamespace(), types.SimpleNamespace()

    class PointCloud:
        points = None

    class KDTreeFlann:
        def __init__(self, pcd):
            self.pts = np.asarray(pcd.points, np.float64)
            self.tree = cKDTree(self.pts)

        def search_hybrid_vector_3d(self, query, radius, max_nn):
            d, idx = self.tree.query(np.asarray(query, np.float64), k=max_nn, distance_upper_bound=radius)
            keep = idx < len(self.pts)
            return [int(keep.sum()), idx[keep].tolist(), (d[keep] ** 2).tolist()]

    o3d.geometry.PointCloud, o3d.geometry.KDTreeFlann = PointCloud, KDTreeFlann
    o3d.utility.Vector3dVector = lambda a: np.asarray(a)
    return o3d


def load_phystwin(root, backend="cuda"):
    """Import <root>/sim/physics/phystwin.py, UNMODIFIED.
    backend "cuda":   `warp`, `diff_gaussian_rasterization` and `sim.physics.spring_mass_warp` resolve to this
                      repository through compat.install() -- the drop-in configuration.
    backend "interp": `warp` is oracle/warp_exec.py and `sim.physics.spring_mass_warp` is the reference's own file
                      executed under it -- the complete reference stack on the CPU (pins the composition of the
                      end-effector step and the simulator, and checks this harness where there is no GPU)."""
    if backend == "cuda":
        from real2sim_eval_b200 import compat
        compat.install(root)
    else:
        from oracle import warp_exec
        smw = warp_exec.load_reference(os.path.join(root, "sim", "physics", "spring_mass_warp.py"))
        sys.modules["warp"] = warp_exec
        sys.modules["sim.physics.spring_mass_warp"] = smw
    pkg = lambda name: (lambda m: (setattr(m, "__path__", []), m)[1])(types.ModuleType(name))
    kornia = pkg("kornia")
    kornia.geometry = pkg("kornia.geometry")
    kornia.geometry.conversions = types.ModuleType("kornia.geometry.conversions")
    kornia.geometry.conversions.axis_angle_to_rotation_matrix = _aa2rm_torch
    kin, rps = types.ModuleType("sim.utils.robot.kinematics_utils"), types.ModuleType("sim.utils.robot.robot_pc_sampler")
    kin.KinHelper = type("KinHelper", (), {})
    rps.RobotPcSampler = FakeRobotPcSampler
    sys.modules.update({"kornia": kornia, "kornia.geometry": kornia.geometry,
                        "kornia.geometry.conversions": kornia.geometry.conversions, "open3d": _fake_open3d(),
                        "sim.utils.robot.kinematics_utils": kin, "sim.utils.robot.robot_pc_sampler": rps})
    for name in [m for m in sys.modules if m == "sim" or m.startswith("sim.") and "spring_mass_warp" not in m
                 and "kinematics_utils" not in m and "robot_pc_sampler" not in m]:
        del sys.modules[name]
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    return importlib.import_module("sim.physics.phystwin")


def physics_cfg(**over):
    """cfg/physics/default.yaml:7-56 as the namespace hydra would hand over."""
    d = dict(use_graph=True, fps=30, dt=5e-5, num_substeps=2000, duration=30, dashpot_damping=100.0, drag_damping=3.0,
             init_spring_Y=3e4, spring_Y_min=0.0, spring_Y_max=1e5, object_radius=0.02, object_max_neighbours=30,
             controller_radius=0.04, controller_max_neighbours=50, collide_elas=0.5, collide_fric=0.3,
             collide_self_elas=0.5, collide_self_fric=0.3, collide_eef_elas=0.0, collide_eef_fric=1.0,
             collision_requires_grad=True, self_collision=True, collision_dist=0.005, reverse_z=False,
             table_height=0.0, grasp_force_threshold=3e4)
    d.update(over)
    return types.SimpleNamespace(**d)


def write_phystwin_assets(tmp, case, scene):
    """The three files SpringMassDynamicsModule.__init__ reads (phystwin.py:231-298) for a synthetic object."""
    import torch
    os.makedirs(f"{tmp}/data/{case}", exist_ok=True)
    os.makedirs(f"{tmp}/experiments_optimization/{case}", exist_ok=True)
    os.makedirs(f"{tmp}/experiments/{case}/train", exist_ok=True)
    x = scene.x.astype(np.float64)
    with open(f"{tmp}/data/{case}/final_data.pkl", "wb") as f:
        pickle.dump(dict(object_points=x[None], object_colors=np.zeros_like(x)[None],
                         surface_points=np.zeros((0, 3)), interior_points=np.zeros((0, 3))), f)
    with open(f"{tmp}/experiments_optimization/{case}/optimal_params.pkl", "wb") as f:
        pickle.dump(dict(global_spring_Y=3e4, collide_elas=0.5, collide_fric=0.3, collide_object_elas=0.5,
                         collide_object_fric=0.3), f)
    torch.save(dict(spring_Y=torch.tensor(np.exp(scene.log_Y)), collide_elas=torch.tensor([0.5]),
                    collide_fric=torch.tensor([0.3]), collide_object_elas=torch.tensor([0.5]),
                    collide_object_fric=torch.tensor([0.3]), num_object_springs=scene.S),
               f"{tmp}/experiments/{case}/train/best_0.pth")


def stack_scenario(use_pusher, frames=3):
    """The closed-loop scenario of drive_and_compare as plain arrays: what tests/golden/make_stack_golden.py feeds the
    full reference stack and tests/test_gpu_env.py feeds the device path.  Returns dict(scene, table, center, meshes,
    commands=[(xyz, vel, rot, rvel, openness)], dt, S)."""
    from real2sim_eval_b200 import synth
    sc = synth.make_rope()
    if use_pusher:
        center = (0.5, 0.03, 0.004)
        tool = synth.make_pusher(center, n_circ=12, n_len=6)
        table = np.repeat(tool.verts[None], 101, 0).astype(np.float32)
        meshes = [(tool.verts, tool.faces)]
    else:
        center = (0.5, 0.0, 0.004)
        table = synth.gripper_opening_table(center)
        g = synth.make_gripper(center, gap=0.08)
        half, fh = len(g.verts) // 2, len(g.faces) // 2
        meshes = [(g.verts[:half], g.faces[:fh]), (g.verts[half:], g.faces[fh:] - half)]
    dt, S = 5e-5, 20
    xyz = np.asarray(center, np.float32)
    cmds = []
    for f in range(frames):
        vel = np.float32([0.0, -0.6, 0.0]) if use_pusher else np.float32([0.0, 0.0, -0.25])
        rvel = np.float32([0.0, 0.0, 0.3])
        cmds.append((xyz.copy(), vel, synth.EEF_ROT_DOWN.copy(), rvel, np.float32(0.25 - 0.1 * f)))
        xyz = (xyz + vel * np.float32(dt * S)).astype(np.float32)
    return dict(scene=sc, table=table, center=np.asarray(center, np.float32), meshes=meshes, commands=cmds, dt=dt, S=S)


def drive_and_compare(pt, dev, use_pusher, tmp, frames=3):
    """Construct the reference's SpringMassDynamicsModule on the synthetic rope (whatever simulator class the loaded
    module is bound to) and step it `frames` times; the same commands go through oracle/eef_ref.py + the C physics
    oracle.  Returns (module, per-frame max |dx| over particles arrays, oracle)."""
    import torch
    import phys_cases
    import r2s_testutil as util
    from oracle import eef_ref
    from real2sim_eval_b200 import synth
    sc = synth.make_rope()
    write_phystwin_assets(str(tmp), "rope", sc)
    cfg = physics_cfg(fps=1000)                              # num_substeps = round(1 / fps / dt) = 20 per frame
    if use_pusher:
        center = (0.5, 0.03, 0.004)
        tool = synth.make_pusher(center, n_circ=12, n_len=6)
        table = np.repeat(tool.verts[None], 101, 0).astype(np.float32)
        robot = FakeRobotPcSampler(pusher=tool)
        meshes = robot.get_xarm_pusher_meshes()
    else:
        center = (0.5, 0.0, 0.004)                           # finger tips 4 mm above the table, straddling the rope
        table = synth.gripper_opening_table(center)
        robot = FakeRobotPcSampler(gripper=synth.make_gripper(center, gap=0.08))
        meshes = robot.get_xarm_gripper_meshes()
    mod = pt.SpringMassDynamicsModule(
        phystwin_cfg=cfg, device=dev, wp_device=dev, case_name="rope", data_path=f"{tmp}/data",
        zeroth_order_ckpt_path=f"{tmp}/experiments_optimization", first_order_ckpt_path=f"{tmp}/experiments",
        init_pts=torch.tensor(sc.x), init_pose=torch.eye(4), static_meshes=[], robot=robot, robot_type="xarm7",
        use_pusher=use_pusher)
    S = cfg.num_substeps
    assert S == 20
    # the reference's KD-tree spring rule (phystwin.py:264-286) on the same cloud gives the synthetic scene's graph
    assert np.array_equal(mod.init_springs.cpu().numpy(), sc.springs)
    rest = mod.init_rest_lengths.cpu().numpy()               # torch.linalg.norm rounds the last place differently
    assert np.allclose(rest, sc.rest, rtol=3e-7, atol=0)
    sc.rest = rest.copy()                                    # the oracle gets the rest lengths the module made
    case = dict(meshes=dict(dynamic=[(m.vertices, m.triangles) for m in meshes], static=[]))
    o = util.oracle_from_scene(sc, S, mesh=phys_cases.merged_mesh(case), use_pusher=use_pusher,
                               collide_eef_fric=0.2 if use_pusher else 1.0, gather_order=False)
    func = eef_ref.make_eef_pts_func(table)
    t = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
    xyz = np.asarray(center, np.float32)
    cur, grasped = None, False
    faces = None if use_pusher else eef_ref.force_faces(o.mesh_map)
    errs, xs = [], []
    o_free = util.oracle_from_scene(sc, S, gather_order=False)   # the same rope without any tool
    for f in range(frames):
        vel = np.float32([0.0, -0.6, 0.0]) if use_pusher else np.float32([0.0, 0.0, -0.25])
        rvel = np.float32([0.0, 0.0, 0.3])
        openness = np.float32(0.25 - 0.1 * f)                # gap 26 mm -> 19 mm -> 12 mm: the fingers close on the rope
        x_ref = mod.step(eef_xyz=t(xyz[None]), eef_vel=t(vel[None]), eef_rot=t(synth.EEF_ROT_DOWN[None]),
                         eef_rot_vel=t(rvel[None]), gripper_openness=t([[openness]]), eef_pts_func=func,
                         init_eef_xyz=t(center))
        assert tuple(x_ref.shape) == (sc.N, 3)
        o.update_collision_graph()
        e = eef_ref.eef_step(table, center, xyz, vel, synth.EEF_ROT_DOWN, rvel, openness, dt=cfg.dt, n_substeps=S,
                             current_openness=cur, grasped=grasped, forces=None if use_pusher else o.collision_forces,
                             faces=faces, use_pusher=use_pusher)
        cur, grasped = e["current_openness"], e["grasped"]
        o.set_mesh_interactive(e["interp_pts"], e["interp_center"], e["dyn_vel"], e["dyn_omega"])
        o.step()
        o_free.update_collision_graph(); o_free.step()
        xs.append(x_ref.detach().cpu().numpy().copy())
        errs.append(np.abs(xs[-1] - o.x).max(1))
        if not use_pusher:
            assert abs(float(mod.current_openness) - cur) < 1e-7
        xyz = (xyz + vel * np.float32(cfg.dt * S)).astype(np.float32)
    assert np.abs(o.x - o_free.x).max() > 1e-4, "the tool must have moved the rope"
    mod._r2s_x_frames = xs
    return mod, errs, o
