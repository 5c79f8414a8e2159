/*
 * r2s_phys.h -- C ABI of the batched PhysTwin spring-mass substep loop.
 *
 * Replaces, for E independent environments that share one spring topology, the
 * reference's single-environment Warp driver
 *   sim/physics/spring_mass_warp.py  ("SMW", file:line relative to /root/reference)
 * behind the Python class of the same surface (real2sim_eval_b200/physics.py).
 * Each entry point names the reference interface it stands in for.
 *
 * One frame = n_substeps x { clear f, eval_springs, update_vel_from_force,
 * object_collision, set_mesh_points(+refit), mesh_collision,
 * integrate_ground_collision } (SMW:823-943) runs as ONE persistent launch:
 * one CTA per environment, particle state resident in shared memory.
 *
 * State layout in HBM: x4[E][N], v4[E][N] as float4 (xyz + one pad lane), so the
 * spring-force gather reads one 16-byte vector per neighbour.
 */
#ifndef R2S_PHYS_H_
#define R2S_PHYS_H_

#include "r2s_common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r2s_phys r2s_phys; /* opaque handle */

/* Construction parameters: the fields of phystwin_cfg that SMW:501-512 consumes
 * plus the constructor tensors of SMW:478-500. */
typedef struct r2s_phys_desc {
    int32_t E;              /* environments                                        */
    int32_t N;              /* particles per environment (num_object_points)       */
    int32_t S;              /* springs (shared topology)                           */
    int32_t n_substeps;     /* phystwin_cfg.num_substeps (PT:223)                  */
    int32_t self_collision; /* phystwin_cfg.self_collision                         */
    int32_t reverse_z;      /* phystwin_cfg.reverse_z                              */
    int32_t use_pusher;     /* SMW:499                                             */
    int32_t sign_mode;      /* mesh inside/outside: 0 exact winding number (thr 0.6),
                               1 always outside.  See DESIGN.md (Warp builtin unpinned) */
    int32_t coll_row_cap;   /* candidate-list row capacity; 0 -> 64.
                               [reference: 500 with no bound check, SMW:226,544-549]  */
    int32_t threads;        /* CTA size of the frame kernel; 0 -> default            */
    int32_t precise;        /* 1: IEEE sqrt/divide in the reference's expression order (SMW:87-99);
                               0 (default): 1/len by rsqrt + one Newton step and 1/rest precomputed --
                               the same formula to a few ulp, ~2x faster          */
    int32_t mesh_accel;     /* rigid-tool mesh accelerator: 0 auto (pusher meshes of >= 256 tool faces), 1 force,
                               -1 off.  Uniform grid in the tool's rest frame + pseudonormal inside/outside; the tool
                               must move rigidly (PT:462-510) and its faces must come first; static obstacle faces
                               may follow (they are scanned beside the grid query, SMW:652-676)             */
    int32_t pad0_;
    float dt, dashpot_damping, drag_damping;
    float spring_Y_min, spring_Y_max, collision_dist;
    float collide_elas, collide_fric;           /* SMW:591-598 */
    float collide_eef_elas, collide_eef_fric;   /* SMW:599-606 */
    float collide_self_elas, collide_self_fric; /* SMW:607-618 */
    const int32_t* springs;    /* [S,2] device (init_springs)                       */
    const float* rest_lengths; /* [S] or [E,S] device (init_rest_lengths)           */
    int32_t rest_per_env;      /* 0: rest_lengths is [S]; 1: [E,S]                  */
    const float* log_spring_Y; /* [S] device: log stiffness (PT:344); NULL -> log(3e4) */
    const float* masses;       /* [N] device (init_masses); NULL -> 1               */
    const int32_t* collision_mask; /* [N] device; NULL -> arange(N) (SMW:529-533)   */
} r2s_phys_desc;

/* SpringMassSystemWarp.__init__ (SMW:478-726) minus state/mesh upload.  Builds the
 * per-particle adjacency (CSR, ascending spring index) once.  May synchronise.
 * Returns NULL on failure (see r2s_last_error). */
r2s_phys* r2s_phys_create(const r2s_phys_desc* desc);
int r2s_phys_destroy(r2s_phys* h);

/* set_init_state (SMW:742-767).  x, v: [E,N,3] packed float32 device; v may be NULL
 * (zeros).  stride_env_elems = N*3 normally, 0 to broadcast one [N,3] to all envs. */
int r2s_phys_set_state(r2s_phys* h, const float* x, const float* v, int64_t stride_env_elems,
                       void* stream);
/* wp.to_torch(wp_state.wp_x / wp_v) (PT:523-531): packed [E,N,3] copies. */
int r2s_phys_get_state(r2s_phys* h, float* x, float* v, void* stream);

/* set_spring_Y (SMW:946-953): log stiffness [S]. */
int r2s_phys_set_spring_Y(r2s_phys* h, const float* log_spring_Y, void* stream);
/* Rest lengths [S] (per_env = 0) or [E,S] (per_env = 1). */
int r2s_phys_set_rest_lengths(r2s_phys* h, const float* rest, int per_env, void* stream);
/* set_collide / set_collide_eef / set_collide_self (SMW:955-995); NaN keeps a value. */
int r2s_phys_set_collide(r2s_phys* h, float elas, float fric, float eef_elas, float eef_fric,
                         float self_elas, float self_fric);

/* Merged collision mesh (SMW:626-712): HOST pointers, dynamic vertices first.
 * verts [V,3], faces [F,3], mesh_map [F] (0/1 finger, >=0 pusher, <0 static),
 * face_map [F].  Vertex table defaults to the rest pose, velocities to zero. */
int r2s_phys_set_mesh(r2s_phys* h, const float* verts, const int32_t* faces,
                      const int32_t* mesh_map, const int32_t* face_map, int32_t V, int32_t F,
                      int32_t n_dyn_verts);

/* set_mesh_interactive (SMW:769-804): per-substep dynamic vertex table
 * [(E,) n_substeps, n_dyn_verts, 3], centres [(E,) n_substeps, 3], dynamic_velocity
 * [(E,) 2, 3] (row 1 ignored for a pusher), dynamic_omega [(E,) 1, 3]; device.
 * per_env = 0 shares one table between all environments. */
int r2s_phys_set_mesh_motion(r2s_phys* h, const float* interp_pts, const float* interp_center,
                             const float* dyn_vel, const float* dyn_omega, int per_env,
                             void* stream);

/* The handle's own motion tables, for a producer that fills them in place on the device each frame
 * (r2s_eef_forward) instead of copying through r2s_phys_set_mesh_motion.  Shapes as in
 * r2s_phys_set_mesh_motion with dyn_vel always [(E,) 2, 3].  Switching per_env re-allocates (contents
 * undefined until written); the pointers stay valid until the next r2s_phys_set_mesh / switch. */
typedef struct r2s_phys_motion {
    float* interp_pts;    /* [n_env, n_substeps, n_dyn_verts, 3] */
    float* interp_center; /* [n_env, n_substeps, 3]              */
    float* dyn_vel;       /* [n_env, 2, 3]                       */
    float* dyn_omega;     /* [n_env, 3]                          */
    int32_t n_env, n_substeps, n_dyn_verts, dyn_vel_rows;
} r2s_phys_motion;
int r2s_phys_motion_ptrs(r2s_phys* h, int per_env, r2s_phys_motion* out);

/* create_resting_case (SMW:729-740) from the current x. */
int r2s_phys_create_resting_case(r2s_phys* h, void* stream);
/* update_collision_graph (SMW:806-821): rebuild candidate lists from current x. */
int r2s_phys_update_collision_graph(r2s_phys* h, void* stream);

/* step() / wp.capture_launch(graph) (SMW:823-943, PT:515-519): n_substeps <= 0
 * uses the descriptor's value.  One kernel launch; no host synchronisation. */
int r2s_phys_step(r2s_phys* h, int32_t n_substeps, void* stream);

/* Zero-copy views for the host shim. */
typedef struct r2s_phys_ptrs {
    float* x4;               /* [E,N,4] */
    float* v4;               /* [E,N,4] */
    float* collision_forces; /* [E,F,3] (last substep only, SMW:900,414) */
    int32_t* mesh_map;       /* [F] */
    int32_t* coll_num;       /* [E,N] */
    int32_t* coll_idx;       /* [E,N,coll_row_cap] */
    int32_t* status;         /* [E,4]: {candidate total, row overflow count, 0, 0} */
    int32_t F, coll_row_cap, smem_state, smem_bytes;
} r2s_phys_ptrs;
int r2s_phys_get_ptrs(r2s_phys* h, r2s_phys_ptrs* out);

/* Algorithmic bytes per environment per substep: 52*N + 16*S. */
int64_t r2s_phys_algorithmic_bytes(const r2s_phys* h);

#ifdef __cplusplus
}
#endif
#endif /* R2S_PHYS_H_ */
