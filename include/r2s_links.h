/*
 * r2s_links.h -- C ABI of the per-frame rigid re-posing of the robot's Gaussians.
 *
 * Replaces, for E environments that share one robot scan, the Gaussian arithmetic of
 *   transform_gs_xarm_gripper / transform_gs_xarm_pusher   sim/utils/robot/robot_pc_transformations.py:12-55
 *   RobotPcSampler.transform_gs_torch                      sim/utils/robot/robot_pc_sampler.py:119-162
 *   the final quaternion normalisation                     sim/renderer/gs_renderer.py:905
 * as sim/renderer/gs_renderer.py:887-893 calls them once per frame: 15 boolean-mask gathers, 15 small
 * matmuls / quaternion products and 15 mask scatters become one launch per link table plus one launch
 * over the Gaussians, indexed by a precomputed link-slot array.
 *
 * Not replaced (host side, out of scope): forward kinematics (sapien `compute_forward_kinematics`,
 * robot_pc_sampler.py:131-136).  The caller passes the FK pose of every link for this frame.
 *
 * Reference semantics kept, per link l (robot_pc_sampler.py:138-156):
 *   mat  = (pose_l @ offset_l) @ inverse(base_pose_l @ offset_l)      `rest_inv` = that inverse, made once
 *   p'   = p @ mat[:3,:3]^T + mat[:3,3]
 *   q'   = quat_mult(rotation_matrix_to_quaternion(mat[:3,:3]), normalize(q))     (w, x, y, z)
 * Gaussians whose mask value is not in the link list keep p and get normalize(q)
 * (robot_pc_transformations.py:29, 44-52); every quaternion is normalised once more at the end
 * (gs_renderer.py:905).  normalize(v) = v / max(|v|, 1e-12) (torch.nn.functional.normalize).
 * rotation_matrix_to_quaternion is kornia's (not vendored by the reference, not installed here): its
 * published four-branch algorithm is restated; see oracle/links_ref.py for the pinning status.
 */
#ifndef R2S_LINKS_H_
#define R2S_LINKS_H_

#include "r2s_common.h"

#ifdef __cplusplus
extern "C" {
#endif

#define R2S_LINKS_MAX 64 /* links per table (the xArm scan uses 15, robot_pc_transformations.py:35) */

typedef struct r2s_links_args {
    int32_t E;       /* environments                                                         */
    int32_t L;       /* links in the per-environment pose table (<= R2S_LINKS_MAX)            */
    int32_t P;       /* Gaussians per environment in `means3D` / `rotations` (row stride)     */
    int32_t first;   /* row of the first robot Gaussian inside each environment's P rows      */
    int32_t n_robot; /* robot-scan Gaussians: rows first .. first + n_robot - 1               */
    int32_t pad0_;
    const int32_t* link_id;   /* [n_robot] shared: slot in [0, L) or -1 = not attached to a moving link */
    const float* rest_means;  /* [n_robot, 3] shared: scan positions at base_qpos                        */
    const float* rest_quats;  /* [n_robot, 4] shared: scan rotations (w,x,y,z), un-normalised           */
    const float* link_pose;   /* [E, L, 16] row-major 4x4 FK pose of each link at this frame's qpos     */
    const float* link_offset; /* [L, 16] tf_obj_to_link (robot_pc_sampler.py:138)                       */
    const float* rest_inv;    /* [L, 16] inverse(base_pose @ offset) (robot_pc_sampler.py:145-147)      */
    float* means3D;           /* [E, P, 3] out: rows first ..                                            */
    float* rotations;         /* [E, P, 4] out: rows first ..                                            */
    float* link_scratch;      /* [E, L, 16] caller-owned: per link [R | t] (12) + quaternion (4)        */
} r2s_links_args;

/* Two launches on `stream`: compose the E*L link transforms, then re-pose the Gaussians. */
int r2s_links_forward(const r2s_links_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2S_LINKS_H_ */
