/*
 * r2s_metrics.h -- C ABI of the on-device task-success metrics and the particle-state ring buffer
 * (SURVEY.md §8f row N4).
 *
 * The reference evaluates an episode offline: every env.step() pickles the full state to
 * <episode>/state/%06d.pkl (experiments/eval_policy.py:157-213) and three scripts re-read the files:
 *   experiments/utils/calculate_success_T.py:17-29      is_pusht_success:  mean_i |x_i - target_i|^2 < 0.002
 *   experiments/utils/calculate_success_rope.py:38-145  is_rope_success: springs crossing the y_min and the
 *                                                        y_max face of the routing box, each count >= 100
 *   experiments/utils/calculate_success_sloth.py:140-172 is_sloth_success: particles inside the (1.05x) oriented
 *                                                        bounding box of the container >= 3050
 * each followed by the same episode rule: the per-frame test is counted from a start frame on (T: 1700,
 * rope: 800, sloth: 350) and the episode succeeds once 30 frames have passed it (calculate_success_T.py:63-73,
 * _rope.py:193-203, _sloth.py:194-204).
 * Here one launch per frame evaluates the test for E environments from the particle state in HBM, keeps the
 * per-environment counters on the device, and (optionally) appends the packed positions to a ring buffer, so
 * the pickle files are not needed as the wire between simulation and evaluation; the final `success` flags
 * are what the multi-GPU metrics all-gather carries.
 *
 * Arithmetic kept: T in float32 like the numpy expression on float32 arrays (the mean's summation order
 * differs; 1e-6 relative); rope in float64 with np.isclose(.., 0, atol=1e-12) for the parallel / coplanar
 * cases and the eps-widened interval tests, so the COUNTS are exact; sloth as Open3D's
 * OrientedBoundingBox::GetPointIndicesWithinBoundingBox (|d . axis_k| <= extent_k / 2 in float64; open3d is a
 * dependency not installed here -- restated from its published source, parity unpinned for that one test).
 */
#ifndef R2S_METRICS_H_
#define R2S_METRICS_H_

#include "r2s_common.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { R2S_TASK_PUSHT = 0, R2S_TASK_ROPE = 1, R2S_TASK_SLOTH = 2 };

typedef struct r2s_success_args {
    int32_t E;           /* environments                                                        */
    int32_t N;           /* particles per environment                                           */
    int32_t S;           /* springs (rope only)                                                 */
    int32_t task;        /* R2S_TASK_*                                                          */
    int32_t frame;       /* index of this frame within the episode (the pickle file number)     */
    int32_t start_frame; /* frames before this one are not counted (1700 / 800 / 350)           */
    int32_t need_frames; /* frames that must pass for the episode to succeed (30)               */
    int32_t ring_slots;  /* > 0: also store this frame's positions in slot frame % ring_slots    */
    const float* x4;     /* [E, N, 4] particle positions (physics layout, .w ignored)           */
    const float* shift;  /* [3] added to every position first (world = x - global_translation,
                            phystwin.py:153), or NULL                                           */
    const float* target;     /* PUSHT: [N, 3] target positions (T_final_state.pkl)               */
    const int32_t* springs;  /* ROPE:  [S, 2]                                                    */
    double box[15];      /* ROPE:  box[0..2] = min xyz, box[3..5] = max xyz of the routing box
                            SLOTH: box[0..2] = OBB centre, box[3..11] = R row-major (columns = axes),
                                   box[12..14] = extent (already scaled, calculate_success_sloth.py:158) */
    double threshold;    /* PUSHT: 0.002 (mse <);  ROPE: 100 (each count >=);  SLOTH: 3050 (count >=) */
    float* value;        /* [E, 2] out: PUSHT {mse, -}; ROPE {y_min_count, y_max_count}; SLOTH {count, -} */
    int32_t* passed;     /* [E] out: this frame's test (0/1)                                    */
    int32_t* hits;       /* [E] in/out: frames >= start_frame that passed                       */
    int32_t* success;    /* [E] in/out: hits >= need_frames has been reached                    */
    float* ring;         /* [ring_slots, E, N, 3] or NULL                                       */
} r2s_success_args;

/* One launch on `stream`, one CTA per environment. */
int r2s_success_forward(const r2s_success_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2S_METRICS_H_ */
