"""Batched env.step(): E independent environments advanced by one frame each --
physics substeps (one persistent launch), object Gaussians re-bound to the
particles, one render per (env, camera) -- without a host synchronisation.

This is the per-frame sequence of the reference's BaseEnv.step + get_obs
(sim/envs/env.py:86-94, 53-74: physics.step -> renderer.update_state -> render)
restricted to the two hot paths this repository implements plus the LBS step between
them (r2s_lbs_forward: `interpolate_motions`, sim/utils/gs/transform_utils.py:58-212, as
sim/renderer/gs_renderer.py:732-749 calls it -- SURVEY.md §8f N1).
All device work goes to torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, synth
from .eef import BatchedEefMotion
from .lbs import BatchedLBS
from .metrics import BatchedSuccess
from .links import BatchedLinkTransform
from .physics import BatchedSpringMass
from .rasterizer import BatchedRasterizer


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@dataclass
class EnvBatchConfig:
    scene: str = "rope"          # rope | sloth | tblock
    E: int = 256                 # environments on this device
    W: int = 512
    H: int = 512
    cameras: int = 1             # views per environment
    n_substeps: int = 10
    P: int = 200_000             # Gaussians per environment
    obj_frac: float = 0.1        # fraction bound to the particles
    robot_frac: float = 0.15     # fraction that is the robot scan, re-posed per frame from link poses (N2)
    knn: int = 16                # bones per Gaussian (gs_renderer.py:35 k_wgt)
    k_rel: int = 8               # neighbours per bone (gs_renderer.py:34 k_rel)
    seed: int = 1234
    env_offset: int = 0          # global index of this shard's first env (multi-GPU sharding)
    gripper: bool = True
    pusher: bool = False         # the push-T tool instead of the gripper: one rigid rod (use_pusher, phystwin.py:462-510)
    pusher_res: tuple = (112, 112)   # (n_circ, n_len) of the rod mesh: 112 x 112 = 25,312 triangles (the shipped STL has 25,368)
    success_start_frame: int | None = None   # None: the task's own (1700 / 800 / 350); frames before it do not count
    state_ring: int = 0          # frames of packed particle positions kept on the device (0 = none)
    instances_per_gaussian: float = 8.0
    sort_object_gaussians: bool = True   # lay the object Gaussians out along a Morton curve (a warp shares its bones)
    fast_composite: bool = False  # ex2.approx compositing variant (1e-4 relative contract, not bit-identical)


def _scene(name: str) -> synth.Scene:
    if name == "rope":
        return synth.make_rope()
    if name == "sloth":
        return synth.make_sloth()
    if name == "tblock":
        return synth.load_tblock()
    raise ValueError(name)


class BatchedEnv:
    """E environments of one scene type on one device, synthetic data."""

    def __init__(self, cfg: EnvBatchConfig, device="cuda"):
        self.cfg = cfg
        self.device = dev = torch.device(device)
        self.lib = _lib.load()
        E = cfg.E
        base = _scene(cfg.scene)
        self.base = base
        pose_tf = [synth.pose_transform(base, cfg.seed + cfg.env_offset + e) for e in range(E)]
        poses = [synth.pose_scene(base, cfg.seed + cfg.env_offset + e) for e in range(E)]
        rest = np.stack([p.rest for p in poses])
        pr = dict(base.params)
        if cfg.pusher:
            pr["use_pusher"], pr["collide_eef_fric"] = True, 0.2     # phystwin.py:305-306
        self.phys = BatchedSpringMass(E, base.springs, rest, num_particles=base.N, n_substeps=cfg.n_substeps,
                                      log_spring_Y=base.log_Y, masses=base.mass, device=dev, **pr)
        self.x_init = torch.tensor(np.stack([p.x for p in poses]), device=dev)
        self.phys.set_state(self.x_init, torch.tensor(np.stack([p.v for p in poses]), device=dev))
        if self.phys.self_collision:
            self.phys.create_resting_case()
        self.dt = pr["dt"]
        # N4: per-frame task-success test + episode counters on the device (calculate_success_{T,rope,sloth}.py)
        task = {"rope": "rope", "sloth": "sloth", "tblock": "pusht"}[cfg.scene]
        kw = dict(start_frame=cfg.success_start_frame, ring_slots=cfg.state_ring, device=dev)
        if task == "rope":
            self.success = BatchedSuccess(task, E, base.N, springs=base.springs, **kw)
        elif task == "pusht":
            self.success = BatchedSuccess(task, E, base.N, target=base.x, **kw)
        else:   # container box of the packing task: 0.2 x 0.13 x 0.27 m, enlarged by 1.05 (calculate_success_sloth.py:155-158)
            obb = (np.array([0.0, 0.0, 0.135]), np.eye(3), np.array([0.2, 0.13, 0.27]) * 1.05)
            self.success = BatchedSuccess(task, E, base.N, obb=obb, **kw)
        # gripper: two fingers straddling the object near its centre, per-env motion tables
        self.gripper = None
        if cfg.pusher:
            # the rod stands just outside the object's -x face, tip 4 mm above the table; every env pushes along +x
            tip = np.array([float(base.x[:, 0].min()) - 0.0375 + 0.0006, float(base.x[:, 1].mean()), 0.004])
            self.gripper = g = synth.make_pusher(center=tip, n_circ=cfg.pusher_res[0], n_len=cfg.pusher_res[1])
            self.phys.set_mesh(g.verts, g.faces, g.mesh_map, g.face_map, len(g.verts))
            self.eef_init = tip.astype(np.float32)
            table = np.repeat(g.verts[None], 2, 0)       # eef_pts_func(1.0) is all the pusher ever asks (phystwin.py:474-477)
            self.eef = BatchedEefMotion(E, table, self.eef_init, dt=pr["dt"], n_substeps=cfg.n_substeps,
                                        use_pusher=True, phys=self.phys, device=dev)
            self.eef_pose = np.zeros((E, 3), np.float32)
        elif cfg.gripper:
            ctr = base.x.mean(0)
            self.gripper = synth.make_gripper(center=(float(ctr[0]), float(ctr[1]), 0.004), gap=0.03)
            g = self.gripper
            self.phys.set_mesh(g.verts, g.faces, g.mesh_map, g.face_map, len(g.verts))
            self.finger_pose = np.zeros((E, 3), np.float32)  # accumulated eef translation per env
            # N3: the end-effector command -> per-substep tables + grasp hysteresis on the device, written in
            # place into the physics handle's motion tables (phystwin.py:362-460)
            self.eef_init = np.array([ctr[0], ctr[1], 0.004], np.float32)
            self.eef = BatchedEefMotion(E, synth.gripper_opening_table(self.eef_init), self.eef_init, dt=pr["dt"],
                                        n_substeps=cfg.n_substeps, mesh_map=g.mesh_map, phys=self.phys, device=dev)
            self.eef_pose = np.zeros((E, 3), np.float32)
        # Gaussians: per-env sets generated on the device from per-env seeds
        P = cfg.P
        n_obj = int(P * cfg.obj_frac)
        self.n_obj, self.K = n_obj, min(cfg.knn, base.N)
        gen = torch.Generator(device=dev)
        means = torch.empty((E, P, 3), device=dev)
        scales = torch.empty((E, P, 3), device=dev)
        rots = torch.empty((E, P, 4), device=dev)
        opac = torch.empty((E, P, 1), device=dev)
        shs = torch.empty((E, P, 1, 3), device=dev)
        lo = torch.tensor([-0.1, -0.6, 0.0], device=dev)
        hi = torch.tensor([1.1, 0.6, 0.6], device=dev)
        # object Gaussians share their binding (relations, weights: invariant under each env's rigid pose)
        rng = np.random.default_rng(cfg.seed)
        from scipy.spatial import cKDTree
        src = base.x[rng.integers(0, base.N, n_obj)].astype(np.float64) + rng.normal(0, 0.002, (n_obj, 3))
        if cfg.sort_object_gaussians:   # storage order of the object Gaussians: spatial neighbours side by side
            src = src[synth.spatial_order(src)]
        tree = cKDTree(base.x)
        dist, idx = tree.query(src, k=self.K)                      # gs_renderer.py:202-211 knn_weights
        w = 1.0 / (dist.reshape(n_obj, self.K).astype(np.float32) + np.float32(1e-6))
        w = (w / w.sum(1, keepdims=True)).astype(np.float32)
        rel = tree.query(base.x, k=cfg.k_rel + 1)[1][:, 1:]        # gs_renderer.py:195-200 knn_relations
        self.lbs = BatchedLBS(E, base.N, P, n_obj, rel, w, idx.reshape(n_obj, self.K), device=dev,
                              bone_positions=base.x if cfg.sort_object_gaussians else None)
        self.x_prev4 = torch.empty_like(self.phys.x4)
        # robot scan: rows [n_obj, n_obj + n_robot) of every env, one shared scan, per-env link poses
        self.n_robot = int(P * cfg.robot_frac)
        self.links = None
        if self.n_robot:
            self.scan = synth.make_robot_scan(self.n_robot, cfg.seed, origin=(0.3, -0.3, 0.05),
                                              volume=([-0.1, -0.6, 0.0], [1.1, 0.6, 0.6]))
            self.links = BatchedLinkTransform(E, P, n_obj, self.scan.link_id, self.scan.points, self.scan.quats,
                                              self.scan.link_offset, self.scan.base_pose, device=dev)
            base = self.scan.base_pose
            prev = np.concatenate([np.eye(4)[None], base[:-1]])
            self._joint_rel = np.linalg.inv(prev) @ base                       # (L,4,4) joint frames at rest
            self._joint_phase = np.random.default_rng(cfg.seed + 5).uniform(0, 6.28, (E + cfg.env_offset, len(base), 3))[cfg.env_offset:]
            self.link_pose = torch.tensor(self.make_link_poses(0), device=dev)
        for e in range(E):
            gen.manual_seed(cfg.seed + 7 * (cfg.env_offset + e) + 1)
            means[e] = lo + (hi - lo) * torch.rand((P, 3), device=dev, generator=gen)
            scales[e] = torch.exp(np.log(0.006) + 0.5 * torch.randn((P, 3), device=dev, generator=gen))
            q = torch.randn((P, 4), device=dev, generator=gen)
            rots[e] = q / q.norm(dim=1, keepdim=True)
            opac[e] = torch.sigmoid(1.5 + 1.5 * torch.randn((P, 1), device=dev, generator=gen))
            shs[e] = (torch.rand((P, 1, 3), device=dev, generator=gen) - 0.5) / 0.28209479177387814
            means[e, :n_obj] = torch.tensor(synth.pose_points(src, pose_tf[e]), device=dev)   # posed with the env's object
        self.means3D, self.scales, self.rotations, self.opacities, self.shs = means, scales, rots, opac, shs
        if self.links is not None:
            self.links.forward(self.link_pose, self.means3D, self.rotations)
        # cameras: cfg.cameras fixed views per env (side, and a top-down second view), small per-env jitter
        cams = []
        for e in range(E):
            for c in range(cfg.cameras):
                cams.append(synth.make_camera(cfg.W, cfg.H, "side" if c == 0 else "top",
                                              jitter_seed=cfg.seed + 13 * (cfg.env_offset + e) + c))
        self.cams = cams
        self.view_h = torch.tensor(np.stack([c.view for c in cams])).pin_memory()
        self.proj_h = torch.tensor(np.stack([c.proj for c in cams])).pin_memory()
        self.campos_h = torch.tensor(np.stack([c.campos for c in cams])).pin_memory()
        self.view, self.proj, self.campos = self.view_h.to(dev), self.proj_h.to(dev), self.campos_h.to(dev)
        # one (tanfovx, tanfovy) per view when the cameras differ in intrinsics (the reference builds one settings
        # tuple per camera, transform_utils.py:17-30)
        tfx = np.array([c.tanfovx for c in cams], np.float32)
        tfy = np.array([c.tanfovy for c in cams], np.float32)
        if np.all(tfx == tfx[0]) and np.all(tfy == tfy[0]):
            self.tanfovx, self.tanfovy = float(tfx[0]), float(tfy[0])
        else:
            self.tanfovx, self.tanfovy = torch.tensor(tfx, device=dev), torch.tensor(tfy, device=dev)
        self.bg = torch.zeros(3, device=dev)
        self.raster = BatchedRasterizer(dev)
        self.B = E * cfg.cameras
        self.max_instances = int(cfg.instances_per_gaussian * self.B * P)
        self.color = torch.empty((self.B, 3, cfg.H, cfg.W), device=dev)
        self.depth = torch.empty((self.B, 1, cfg.H, cfg.W), device=dev)
        self.frame = 0

    # ---- per-frame host-side action -> gripper motion tables (numpy, pinned by the caller if needed)
    def make_actions(self, frame: int):
        """Seeded per-env eef velocities for this frame and the per-substep tables the physics takes
        (what sim/physics/phystwin.py:374-460 computes per frame from the policy's action)."""
        cfg, E, ns = self.cfg, self.cfg.E, self.cfg.n_substeps
        rng = np.random.default_rng(cfg.seed + 1000 * frame + cfg.env_offset)
        vel = rng.uniform(-0.1, 0.1, (E, 3)).astype(np.float32)
        vel[:, 2] = rng.uniform(-0.05, 0.02, E)
        g = self.gripper
        t = (np.arange(1, ns + 1, dtype=np.float32) * np.float32(self.dt))[None, :, None, None]
        base_pts = g.verts[None, None] + self.finger_pose[:, None, None, :]
        pts = (base_pts + vel[:, None, None, :] * t).astype(np.float32)            # (E, ns, 48, 3)
        ctr = (g.verts.mean(0)[None, None] + self.finger_pose[:, None, :] + vel[:, None, :] * t[:, :, 0]).astype(np.float32)
        dyn_vel = np.repeat((vel * 0.5)[:, None, :], 2, axis=1).astype(np.float32)  # (E, 2, 3)
        dyn_omega = np.zeros((E, 1, 3), np.float32)
        self.finger_pose = self.finger_pose + vel * np.float32(self.dt * ns)
        return pts, ctr, dyn_vel, dyn_omega

    def make_commands(self, frame: int):
        """Seeded per-env end-effector commands for this frame, as PhysTwinDynamics.step hands them to the
        dynamics module (phystwin.py:104-147): eef_xyz (E,3), eef_vel (E,3), eef_rot (E,3,3), eef_rot_vel (E,3),
        gripper_openness (E,) -- 19 floats per environment instead of the (S,48,3) vertex tables."""
        cfg, E = self.cfg, self.cfg.E
        rng = np.random.default_rng(cfg.seed + 1000 * frame + cfg.env_offset)
        if cfg.pusher:   # push along +x with a little sideways drift and yaw; the opening is not used (None)
            vel = np.stack([rng.uniform(0.05, 0.3, E), rng.uniform(-0.03, 0.03, E), np.zeros(E)], 1).astype(np.float32)
            rot_vel = np.stack([np.zeros(E), np.zeros(E), rng.normal(0.0, 0.5, E)], 1).astype(np.float32)
            xyz = (self.eef_init[None] + self.eef_pose).astype(np.float32)
            rot = np.repeat(synth.EEF_ROT_DOWN[None], E, 0).astype(np.float32)
            self.eef_pose = self.eef_pose + vel * np.float32(self.dt * cfg.n_substeps)
            return xyz, vel, rot, rot_vel, None
        vel = rng.uniform(-0.1, 0.1, (E, 3)).astype(np.float32)
        vel[:, 2] = rng.uniform(-0.05, 0.02, E)
        rot_vel = rng.normal(0.0, 0.2, (E, 3)).astype(np.float32)
        phase = np.random.default_rng(cfg.seed + 3).uniform(0, 6.28, E + cfg.env_offset)[cfg.env_offset:]
        openness = (0.3 + 0.12 * np.sin(phase + 0.25 * frame)).astype(np.float32)   # gap 2.1 .. 3.8 cm
        xyz = (self.eef_init[None] + self.eef_pose).astype(np.float32)
        rot = np.repeat(synth.EEF_ROT_DOWN[None], E, 0).astype(np.float32)
        self.eef_pose = self.eef_pose + vel * np.float32(self.dt * cfg.n_substeps)
        return xyz, vel, rot, rot_vel, openness

    def make_link_poses(self, frame: int) -> np.ndarray:
        """(E, L, 4, 4) float32 FK poses of every env's robot links for this frame: each joint of the chain
        swings smoothly about its rest frame (stand-in for sapien FK of the policy's qpos,
        robot_pc_sampler.py:131-136)."""
        ang = 0.05 * np.sin(self._joint_phase + 0.15 * frame)                  # (E,L,3) rotation vectors
        th = np.linalg.norm(ang, axis=-1)[..., None, None]
        k = ang / np.maximum(th[..., 0], 1e-12)
        K = np.zeros(ang.shape[:2] + (3, 3))
        K[..., 0, 1], K[..., 0, 2], K[..., 1, 0] = -k[..., 2], k[..., 1], k[..., 2]
        K[..., 1, 2], K[..., 2, 0], K[..., 2, 1] = -k[..., 0], -k[..., 1], k[..., 0]
        J = np.tile(np.eye(4), ang.shape[:2] + (1, 1))
        J[..., :3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
        out = np.empty_like(J)
        cur = np.tile(np.eye(4), (ang.shape[0], 1, 1))
        for s in range(ang.shape[1]):
            cur = cur @ J[:, s] @ self._joint_rel[s]
            out[:, s] = cur
        return out.astype(np.float32)

    def step(self, motion=None, out=None, link_pose=None, command=None, composite_stream=None):
        """One frame for every env: [end-effector step ->] collision graph -> substeps -> LBS -> robot links -> render.
        `command`: device tensors (eef_xyz, eef_vel, eef_rot, eef_rot_vel, gripper_openness) -- the per-substep
        tables and the grasp hysteresis are then made on the device (r2s_eef_forward); or
        `motion`: ready-made device tables (interp_pts, interp_center, dyn_vel, dyn_omega); or neither.
        `link_pose`: [E,L,4,4] device tensor of the robot's FK link poses for this frame, or None (robot kept).
        `composite_stream`: a torch stream for the compositing kernel (see r2s_raster_args.composite_stream): the images
        are complete on THAT stream; everything else of the step runs on the current stream.
        `out`: optional (color, depth[, rgb8]) device tensors to render into (double buffering); rgb8 is
        the [B,H,W,3] uint8 image the reference's evaluation loop builds on the host
        (experiments/eval_policy.py:248), written here by the compositing kernel."""
        color, depth = (out[0], out[1]) if out is not None else (self.color, self.depth)
        rgb8 = out[2] if out is not None and len(out) > 2 else None
        if self.phys.self_collision:
            self.phys.update_collision_graph()       # once per frame (phystwin.py:365-366)
        if command is not None:
            self.eef.forward(*command)               # reads last frame's collision_forces, fills the tables in place
        elif motion is not None:
            self.phys.set_mesh_motion(*motion)
        self.x_prev4.copy_(self.phys.x4)             # state['x'] before the frame (gs_renderer.py:727)
        self.phys.step()
        # the reference pickles state/{cnt:06d}.pkl BEFORE env.step (eval_policy.py:209-225): file k is the state after
        # k steps, so the state this frame produced is file number frame + 1
        self.success.update(self.phys.x4, frame=self.frame + 1)
        self.lbs.forward(self.x_prev4, self.phys.x4, self.means3D)
        if self.links is not None and link_pose is not None:   # robot Gaussians follow this frame's FK poses
            self.links.forward(link_pose, self.means3D, self.rotations)
        c = self.cfg
        self.raster.forward(self.means3D, self.opacities, viewmatrix=self.view, projmatrix=self.proj,
                            campos=self.campos, bg=self.bg, W=c.W, H=c.H, tanfovx=self.tanfovx,
                            tanfovy=self.tanfovy, shs=self.shs, scales=self.scales, rotations=self.rotations,
                            sh_degree=0, z_threshold=0.05, views_per_scene=c.cameras,
                            max_instances=self.max_instances, out_color=color, out_depth=depth,
                            want_radii=False, out_rgb8=rgb8, fast=c.fast_composite, composite_stream=composite_stream)
        self.frame += 1
        return color, depth

    def check(self):
        """Raise if anything was silently dropped since the last call: a render whose instance lists overflowed the
        workspace (that frame is background only), or self-collision candidates beyond the row capacity.  Both are
        counted on the device without a per-frame synchronisation; this call synchronises -- use it at the end
        of an episode (or of a benchmark's timed region)."""
        lost = self.raster.overflows(reset=True)
        if lost:
            raise _lib.R2SError(f"{lost} frame(s) overflowed the rasterizer workspace (max_instances={self.max_instances}): "
                                "raise EnvBatchConfig.instances_per_gaussian")
        if self.phys.self_collision:
            dropped = int(self.phys.status[:, 1].sum())
            if dropped:
                raise _lib.R2SError(f"{dropped} self-collision candidates exceeded the row capacity "
                                    f"{self.phys.coll_row_cap} in the last frame (the reference keeps 500 per particle, "
                                    "spring_mass_warp.py:544-549): construct the physics with a larger coll_row_cap")
