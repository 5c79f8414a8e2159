/*
 * r2s_common.h -- conventions shared by the two C-ABI entry-point families of
 * libr2s.so (r2s_phys.h: spring-mass substep loop, r2s_raster.h: Gaussian-splat
 * forward rasterizer).
 *
 * Conventions (they replace the reference's behaviour noted in brackets):
 *   - plain C ABI, no C++/torch types cross the boundary;
 *   - every function returns 0 on success or a negative r2s_status; the message
 *     is available from r2s_last_error() (thread-local).  Nothing throws.
 *     [reference: AT_ERROR / std::runtime_error / unchecked launches,
 *      rasterize_points.cu:57-59, rasterizer_impl.cu:243-246]
 *   - all device work is enqueued on the cudaStream_t passed in (as void*); the
 *     hot-path calls never synchronise the host.
 *     [reference: legacy default stream + one blocking cudaMemcpy per render,
 *      rasterizer_impl.cu:283-284]
 *   - the caller owns every buffer, including scratch ("workspace").
 *     [reference: three torch byte tensors grown through std::function
 *      callbacks, rasterize_points.cu:27-33,74-79]
 *   - pointers are DEVICE pointers unless a parameter is documented as host.
 */
#ifndef R2S_COMMON_H_
#define R2S_COMMON_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum r2s_status {
    R2S_OK = 0,
    R2S_ERR_INVALID = -1,   /* bad argument / shape                      */
    R2S_ERR_CUDA = -2,      /* a CUDA runtime call failed                */
    R2S_ERR_WORKSPACE = -3, /* workspace too small for the request       */
    R2S_ERR_UNSUPPORTED = -4
} r2s_status;

/* Message of the last failing call on this host thread ("" if none). */
const char* r2s_last_error(void);

/* ABI version of this library: major*10000 + minor*100 + patch. */
int r2s_version(void);

/* Number of kernel launches issued by this library in this process so far
 * (all entry points).  bench.py differences it around the timed region to
 * report "gpu_launches". */
int64_t r2s_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* R2S_COMMON_H_ */
