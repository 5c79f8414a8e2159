"""Generates tests/golden/links_*.npz by running the REFERENCE's own RobotPcSampler.transform_gs_torch
(sim/utils/robot/robot_pc_sampler.py:119-162) + quat_mult_torch on CPU, wrapped in the mask gather/scatter of
transform_gs_xarm_gripper (robot_pc_transformations.py:29,44-52) and the renderer's final normalisation
(gs_renderer.py:905).  Needs /root/reference; see oracle/links_ref.py for what is stubbed (asset loading, FK,
and kornia's rotation_matrix_to_quaternion, which is restated).

    python tests/golden/make_links_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import links_ref  # noqa: E402
from real2sim_eval_b200 import synth  # noqa: E402

CASES = [("a", 4000, 11, 0.3), ("b", 1500, 12, 1.5), ("c", 700, 13, 0.0)]


def main():
    mod = links_ref.load_reference()
    assert mod is not None, "/root/reference is not mounted"
    for name, n, seed, amount in CASES:
        scan = synth.make_robot_scan(n, seed)
        pose = synth.robot_link_poses(scan, seed + 100, amount)
        if name == "c":           # rest pose: every matrix is the identity up to rounding
            pose = scan.base_pose.copy()
        p, q = links_ref.reference_transform_gs(mod, scan.points, scan.quats, scan.total_mask, list(synth.XARM_LINK_IDS),
                                                scan.link_names, pose, scan.base_pose, scan.link_offset)
        out = os.path.join(HERE, f"links_{name}.npz")
        np.savez_compressed(out, points=scan.points, quats=scan.quats, total_mask=scan.total_mask, link_id=scan.link_id,
                            link_pose=pose, base_pose=scan.base_pose, link_offset=scan.link_offset,
                            out_points=p, out_quats=q)
        print(out, p.shape, os.path.getsize(out))


if __name__ == "__main__":
    main()
