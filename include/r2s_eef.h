/*
 * r2s_eef.h -- C ABI of the per-frame end-effector step: the policy's end-effector
 * command of E environments -> the per-substep collision-mesh vertex tables the physics consumes, plus the
 * grasp hysteresis that decides the commanded finger opening, all on the device.
 *
 * Replaces the host arithmetic of SpringMassDynamicsModule.step, sim/physics/phystwin.py:362-510, per
 * environment and frame:
 *   :370-381  dts = linspace(1,S,S)*dt; eef_xyz_next = xyz + vel*dts; eef_rot_next = axis_angle(rot_vel*dts)^T @ rot
 *   :383-412  grasp hysteresis from the finger contact forces of the previous frame
 *             (`collision_forces.numpy()` / `mesh_map.numpy()` / `gripper_openness.item()`: three host round
 *             trips per frame in the reference) on faces [18], [19], [1] of each finger
 *   :415-431  finger vertices at the new and the previous opening (scipy interp1d over 101 samples,
 *             robot_pc_transformations.py:183-192), their per-substep interpolation in the gripper frame
 *   :433-452  interpolated_dynamic_points (S,V,3), interpolated_center (S,3), dynamic_velocity (2,3) incl. the
 *             mean closing velocity of each finger, dynamic_omega (1,3)
 *   :462-503  the same for a pusher (opening fixed at 1.0, one velocity row)
 * and hands the tables to `set_mesh_interactive` (:455-460, :505-510) -- here by writing them in place into
 * the buffers r2s_phys_motion_ptrs() exposes, so no table crosses PCIe and nothing synchronises.
 *
 * Not replaced: the pose bookkeeping of PhysTwinDynamics.step (:104-135, kornia quaternion/axis-angle
 * conversions of the action) -- the caller passes eef_xyz / eef_vel / eef_rot / eef_rot_vel as that code
 * computes them -- and inverse kinematics for the 101 finger-vertex samples (setup time).
 * kornia's axis_angle_to_rotation_matrix (dependency not vendored by the reference, not installed here) is
 * restated from its published algorithm: Rodrigues with w = aa/(theta + 1e-6) when theta^2 > 1e-6, else the
 * first-order matrix; see oracle/eef_ref.py for the pinning status.
 */
#ifndef R2S_EEF_H_
#define R2S_EEF_H_

#include "r2s_common.h"

#ifdef __cplusplus
extern "C" {
#endif

#define R2S_EEF_MAX_SUBSTEPS 2048 /* per-substep rotations are staged in shared memory */

typedef struct r2s_eef_args {
    int32_t E;          /* environments                                                               */
    int32_t n_substeps; /* S = phystwin_cfg.num_substeps                                              */
    int32_t n_pts;      /* V = dynamic collision-mesh vertices (both fingers: left half first)        */
    int32_t n_table;    /* samples of the opening -> vertices table (101, robot_pc_transformations.py:183) */
    int32_t use_pusher; /* phystwin.py:462: opening fixed at 1.0, no hysteresis, one velocity row     */
    int32_t F;          /* faces per environment in `collision_forces`                                */
    int32_t force_faces[6]; /* rows of collision_forces summed per finger: left {[18],[19],[1]} of the faces
                               with mesh_map == 0, right the same of mesh_map == 1 (phystwin.py:386-391) */
    int32_t dyn_vel_rows;   /* rows per environment in `dyn_vel` (2: the physics handle's layout)     */
    int32_t pad0_;
    float grasp_force_threshold; /* phystwin_cfg.grasp_force_threshold (3e4, cfg/physics/default.yaml:51) */
    float pad1_;
    double dt;                   /* phystwin_cfg.dt as the Python float it is: float32(dt) scales the substep
                                    times, float32(dt * S) and float32(2 * dt * S) the divisions (:429, :442) */
    const float* table;          /* [n_table, V, 3] finger vertices at opening k/(n_table-1), float32 as
                                    eef_pts_list holds them (robot_pc_transformations.py:185-189)       */
    const float* init_eef_xyz;   /* [3]                                                               */
    const float* eef_xyz;        /* [E, 3]    first gripper (phystwin.py:434 takes [:, 0])             */
    const float* eef_vel;        /* [E, 3]                                                            */
    const float* eef_rot;        /* [E, 3, 3] row-major                                               */
    const float* eef_rot_vel;    /* [E, 3]                                                            */
    const float* openness_cmd;   /* [E] gripper_openness of the action (ignored for a pusher)         */
    const float* collision_forces; /* [E, F, 3] of the previous frame, or NULL (all zero)             */
    double* current_openness;    /* [E] in/out: SpringMassDynamicsModule.current_openness; NaN = None  */
    int32_t* grasped;            /* [E] in/out: SpringMassDynamicsModule.grasped                       */
    float* interp_pts;           /* [E, S, V, 3] out                                                  */
    float* interp_center;        /* [E, S, 3]    out                                                  */
    float* dyn_vel;              /* [E, dyn_vel_rows, 3] out (row 1 untouched for a pusher)            */
    float* dyn_omega;            /* [E, 3]       out                                                  */
} r2s_eef_args;

/* One launch on `stream`, one CTA per environment. */
int r2s_eef_forward(const r2s_eef_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2S_EEF_H_ */
