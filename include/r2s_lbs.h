/*
 * r2s_lbs.h -- C ABI of the linear-blend-skinning step that moves the object Gaussians with the
 * particles between the physics frame and the render.
 *
 * Replaces `interpolate_motions(bones, motions, relations, xyz, weights, weights_indices, quat=None)`
 * (sim/utils/gs/transform_utils.py:58-212) as called once per frame by
 * sim/renderer/gs_renderer.py:732-749, for E environments that share relations / weights
 * (one PhysTwin, E poses).  Only the transformed xyz is produced: the reference call site passes
 * quat=None and discards the other returns.
 *
 * Reference semantics kept: per bone F = sum_a (new_a - new_i)(old_a - old_i)^T over its k_rel
 * neighbours, R = U diag(1,1,+-1) V^T (the proper rotation of the two dominant singular pairs);
 * if ANY bone of an environment has rank(F) < 2 every bone of that environment gets the identity
 * rotation (transform_utils.py:159-167: the shape-mismatched assignment falls into `except`);
 * xyz' = sum_k w_k (R_b (xyz - bone_b) + motion_b + bone_b).
 */
#ifndef R2S_LBS_H_
#define R2S_LBS_H_

#include "r2s_common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r2s_lbs_args {
    int32_t E;      /* environments                                              */
    int32_t N;      /* bones (particles) per environment                         */
    int32_t P;      /* Gaussians per environment in `means3D` (row stride)       */
    int32_t n_obj;  /* leading Gaussians of every environment that are skinned   */
    int32_t k_rel;  /* neighbours per bone (gs_renderer.py:34: 8)                */
    int32_t k_wgt;  /* bones per Gaussian (gs_renderer.py:35: 16)                */
    const int32_t* relations;       /* [N, k_rel]   shared                       */
    const int32_t* weights_indices; /* [n_obj, k_wgt] shared                     */
    const float* weights;           /* [n_obj, k_wgt] shared                     */
    const float* bones4;            /* [E, N, 4] particle positions BEFORE the frame (state['x'])  */
    const float* bones_new4;        /* [E, N, 4] particle positions AFTER the frame (x_pred)       */
    float* means3D;                 /* [E, P, 3] in/out: rows < n_obj are transformed in place     */
    float* rot_scratch;             /* [E, N, 12] bone transforms [R | new - R old] as three float4 rows;
                                       caller-owned scratch, 16-byte aligned                       */
    int32_t* rank_flags;            /* [E] out: 1 if every bone had rank >= 2, else 0 (identity used) */
    /* Optional layout hints, all three or none (null = bone transforms stored in bone order).  They change where
       a bone's transform sits in `rot_scratch` (and in the blend kernel's shared-memory copy), not what is computed:
       bone i is stored in slot bone_slot[i], and the blend reads row g of `weights_slots` / `weights_by_slot` --
       the same (bone, weight) pairs as row g of weights_indices / weights, with the bone replaced by its slot and
       the pairs in any order the caller likes (ascending slot keeps the 32 Gaussians of a warp on neighbouring
       rows).  The sum over a Gaussian's bones then runs in that order.                                         */
    const int32_t* bone_slot;       /* [N] a permutation of 0..N-1, shared                          */
    const int32_t* weights_slots;   /* [n_obj, k_wgt] shared                                        */
    const float* weights_by_slot;   /* [n_obj, k_wgt] shared                                        */
} r2s_lbs_args;

/* Two launches on `stream`: per-bone rotations, then the per-Gaussian blend. */
int r2s_lbs_forward(const r2s_lbs_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2S_LBS_H_ */
