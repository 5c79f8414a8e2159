// Internal helpers shared by phys.cu and raster.cu (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "r2s_common.h"

namespace r2s {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define R2S_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            r2s::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,           \
                           cudaGetErrorString(e__));                                    \
            return R2S_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define R2S_LAUNCH_CHECK()                                                              \
    do {                                                                                \
        cudaError_t e__ = cudaPeekAtLastError();                                        \
        if (e__ != cudaSuccess) {                                                       \
            r2s::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,       \
                           cudaGetErrorString(e__));                                    \
            return R2S_ERR_CUDA;                                                        \
        }                                                                               \
        r2s::count_launch();                                                            \
    } while (0)

#define R2S_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            r2s::set_error(__VA_ARGS__);                                                \
            return R2S_ERR_INVALID;                                                     \
        }                                                                               \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace r2s
