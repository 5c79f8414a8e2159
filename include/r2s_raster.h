/*
 * r2s_raster.h -- C ABI of the batched forward Gaussian-splat rasterizer with
 * median depth.
 *
 * Stands in for the reference's CUDA rasterizer
 *   third-party/diff-gaussian-rasterization-w-depth ("DGR", file:line relative to it)
 *     CudaRasterizer::Rasterizer::forward     cuda_rasterizer/rasterizer.h:31-54,
 *                                             cuda_rasterizer/rasterizer_impl.cu:198-341
 *     CudaRasterizer::Rasterizer::markVisible cuda_rasterizer/rasterizer.h:24,
 *                                             cuda_rasterizer/rasterizer_impl.cu:141-153
 *   and its torch glue RasterizeGaussiansCUDA  rasterize_points.cu:35-117
 * for B independent (scene, camera) views per call.  The Python class
 * GaussianRasterizer of real2sim_eval_b200/rasterizer.py (same signature as
 * diff_gaussian_rasterization/__init__.py:149-198) binds these entry points.
 *
 * Differences of mechanism (results are the reference's):
 *   - no host synchronisation: the instance count stays on the device; the
 *     caller sizes `max_instances` up front and may poll r2s_raster_status()
 *     [reference: blocking cudaMemcpy of num_rendered, rasterizer_impl.cu:283-284];
 *   - instances are binned per SUPER-TILE (4x4 tiles) and each super-tile list is sorted on the
 *     unique key (depth bits, Gaussian id); a tile's list is the order-preserving subsequence of
 *     its super-tile's list whose tile rectangle contains the tile -- the same members in the
 *     same order that cub::DeviceRadixSort::SortPairs over (tile | depth) keys yields for the
 *     reference's emission order [rasterizer_impl.cu:70-111, 303-311, 116-138], with ~8x fewer
 *     instances emitted and sorted.
 */
#ifndef R2S_RASTER_H_
#define R2S_RASTER_H_

#include "r2s_common.h"

#ifdef __cplusplus
extern "C" {
#endif

#define R2S_TILE 16 /* BLOCK_X = BLOCK_Y, cuda_rasterizer/config.h:15-17 */

typedef struct r2s_raster_args {
    int32_t B;               /* views in this call                                         */
    int32_t views_per_scene; /* consecutive views sharing one Gaussian set (>= 1)          */
    int32_t P;               /* Gaussians per scene                                        */
    int32_t D, M;            /* active SH degree, SH coefficients per Gaussian             */
    int32_t W, H;            /* image size (all views)                                     */
    int32_t prefiltered;     /* accepted for signature parity; culled Gaussians are skipped */
    float scale_modifier, tanfovx, tanfovy, z_threshold;
    /* per scene: n_scenes = B / views_per_scene, leading dimension n_scenes */
    const float* means3D;        /* [n_scenes,P,3]                              */
    const float* scales;         /* [n_scenes,P,3] or NULL with cov3D_precomp   */
    const float* rotations;      /* [n_scenes,P,4] wxyz, used un-normalised     */
    const float* opacities;      /* [n_scenes,P]                                */
    const float* shs;            /* [n_scenes,P,M,3] or NULL with colors_precomp */
    const float* colors_precomp; /* [n_scenes,P,3] or NULL                      */
    const float* cov3D_precomp;  /* [n_scenes,P,6] or NULL                      */
    /* per view */
    const float* viewmatrix; /* [B,16] column-major w2c (transform_utils.py:11)  */
    const float* projmatrix; /* [B,16]                                           */
    const float* campos;     /* [B,3]                                            */
    const float* bg;         /* [3]                                              */
    /* outputs (caller-owned) */
    float* out_color; /* [B,3,H,W] */
    float* out_depth; /* [B,1,H,W] */
    int32_t* radii;   /* [B,P] or NULL; with NULL the workspace's `radii` / `tiles_touched` arrays are not written
                         either (no kernel reads them; 8 B per Gaussian of HBM traffic) */
    uint8_t* out_rgb8; /* [B,H,W,3] or NULL: the host-side image format of the reference's evaluation loop,
                          (clamp(color, 0, 1) * 255) truncated to uint8, HWC (gs_renderer.py:949 +
                          experiments/eval_policy.py:248), written by the same kernel as out_color */
    /* scratch (caller-owned); size from r2s_raster_workspace_bytes */
    void* workspace;
    size_t workspace_bytes;
    int64_t max_instances; /* capacity for (Gaussian, super-tile) instances over the whole batch; a
                              super-tile is 4x4 tiles, so the (Gaussian, tile) count is always enough */
    const float* tanfov_views; /* [B,2] per-view (tanfovx, tanfovy) or NULL (every view uses tanfovx / tanfovy
                                  above).  The reference builds one settings tuple per camera
                                  (transform_utils.py:17-30); its fixed and wrist cameras differ in intrinsics */
    int32_t* overflow_count;   /* device counter or NULL: incremented by every forward on which the instance lists
                                  exceeded max_instances (that frame holds background only).  Sticky -- the caller
                                  zeroes it and reads it when convenient, e.g. at the end of an episode */
    int32_t composite_mode;    /* R2S_COMPOSITE_PRECISE (0): IEEE expf in the reference's expression order, images
                                  bit-identical to the reference build.  R2S_COMPOSITE_FAST (1): log2(e) folded into
                                  the staged conic and ex2.approx -- within the 1e-4 relative contract, not bitwise */
    int32_t pad0_;
    void* composite_stream;    /* NULL: every kernel runs on the stream passed to r2s_raster_forward.  Otherwise the
                                  compositing kernel is enqueued on THIS stream, ordered after the sort by an event:
                                  the binning kernels (memory / latency-bound) of the next batch of views can then run
                                  under the compositing (issue-bound) of this one.  The caller orders whatever reuses
                                  the workspace or reads the images after this stream */
} r2s_raster_args;
#define R2S_COMPOSITE_PRECISE 0
#define R2S_COMPOSITE_FAST 1

size_t r2s_raster_workspace_bytes(int32_t B, int32_t P, int32_t W, int32_t H, int64_t max_instances);

/* Enqueue preprocess -> bin -> per-tile sort -> composite for all B views. */
int r2s_raster_forward(const r2s_raster_args* args, void* stream);

/* Blocking read-back of the device-side counters of the last forward on this
 * workspace: total (Gaussian, tile) instances (the reference's num_rendered summed over views)
 * and whether max_instances was exceeded (in which case the images hold only
 * the background).  This is the only call that synchronises. */
int r2s_raster_status(const void* workspace, void* stream, int64_t* num_rendered, int32_t* overflow);

/* markVisible: present[i] = (view-space z > 0.01). [P] bool (1 byte each). */
int r2s_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* Byte offsets of the intermediate arrays inside a workspace (for parity tests
 * and for the algorithmic-bytes accounting in bench.py). */
typedef struct r2s_raster_layout {
    size_t status;       /* int64 fine total, int32 overflow, int32 pad, int64 coarse total */
    size_t depths;       /* float  [B*P]  view-space z                                */
    size_t radii;        /* int32  [B*P]                                             */
    size_t tiles_touched; /* uint32 [B*P]                                            */
    size_t rec_a;        /* float4 [B*P][2]: one 32-byte record per Gaussian, {x, y, conic.x, conic.y} then
                            {conic.z, opacity, r, g} -- a visible Gaussian writes exactly one DRAM sector  */
    size_t rec_b;        /* = rec_a + 16: the second float4 of record 0 (stride 32 bytes)               */
    size_t rec_c;        /* float  [B*P] {b} (0 for culled Gaussians: written densely)                  */
    size_t rects;        /* uint32 [B*P] tile rectangle minx | miny<<8 | maxx<<16 | maxy<<24 (0 = culled) */
    size_t tile_count;   /* uint32 [B*ST] instances per super-tile (4x4 tiles)          */
    size_t tile_offset;  /* uint32 [B*ST+1] exclusive scan; list[s] = keys[off[s] .. off[s+1]) */
    size_t tile_fill;    /* uint32 [B*ST]                                            */
    size_t keys;         /* uint64 [max_instances] (depth bits << 32 | id), ascending inside each super-tile */
    size_t keys_alt;     /* uint64 [max_instances] merge scratch                      */
    size_t sorted_rect;  /* uint32 [max_instances] tile rectangle of each sorted entry */
    size_t total;        /* total bytes                                              */
    int32_t tiles_x, tiles_y, super_x, super_y;
} r2s_raster_layout;
int r2s_raster_workspace_layout(int32_t B, int32_t P, int32_t W, int32_t H, int64_t max_instances,
                                r2s_raster_layout* out);

/* Per-stage device timing of r2s_raster_forward for the roofline report: when enabled,
 * forward() records CUDA events between its kernels on the caller's stream (no sync);
 * r2s_raster_get_profile synchronises on the last event and returns the milliseconds of
 * {preprocess, scan, emit, tile_sort, composite} of the most recent forward. */
#define R2S_RASTER_STAGES 5
int r2s_raster_set_profile(int32_t enable);
int r2s_raster_get_profile(float ms[R2S_RASTER_STAGES]);

#ifdef __cplusplus
}
#endif
#endif /* R2S_RASTER_H_ */
