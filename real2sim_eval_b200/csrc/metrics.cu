// metrics.cu -- per-frame task-success tests + particle-state ring buffer on the device,
// sm_100a.  One CTA per environment reads the particle state once (HBM-bound: 16 B per particle, plus 8 B
// per spring for the rope test) and leaves a handful of integers; see include/r2s_metrics.h for the
// reference scripts each test restates.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "r2s_internal.h"
#include "r2s_metrics.h"

namespace {

constexpr int kThreads = 256;
constexpr double kEps = 1e-12;   // calculate_success_rope.py:39 (eps default of the plane test)

struct P3 {
    double x, y, z;
};

__device__ __forceinline__ P3 load_point(const float4* __restrict__ x4, int i, float sx, float sy, float sz)
{
    const float4 p = x4[i];
    // the shift is applied in float32, as the reference's tensors hold world-frame positions in float32
    return {(double)__fadd_rn(p.x, sx), (double)__fadd_rn(p.y, sy), (double)__fadd_rn(p.z, sz)};
}

// _segment_plane_intersections_xz (calculate_success_rope.py:38-75) for one segment and one plane, float64
__device__ __forceinline__ bool crosses(const P3& p0, const P3& p1, double y_plane, double x_min, double x_max,
                                        double z_min, double z_max)
{
    const double dy = p1.y - p0.y;
    const bool parallel = fabs(dy) <= kEps;                          // np.isclose(dy, 0.0, atol=eps)
    auto in_rect = [&](double x, double z) {
        return x >= x_min - kEps && x <= x_max + kEps && z >= z_min - kEps && z <= z_max + kEps;
    };
    if (!parallel) {
        const double t = (y_plane - p0.y) / dy;
        if (!(t >= -kEps && t <= 1.0 + kEps)) return false;
        const double xi = __dadd_rn(p0.x, __dmul_rn(t, p1.x - p0.x));
        const double zi = __dadd_rn(p0.z, __dmul_rn(t, p1.z - p0.z));
        return in_rect(xi, zi);
    }
    const bool coplanar = fabs(p0.y - y_plane) <= kEps;             // np.isclose(y0 - y_plane, 0.0, atol=eps)
    return coplanar && (in_rect(p0.x, p0.z) || in_rect(p1.x, p1.z));
}

__global__ void __launch_bounds__(kThreads) success_kernel(const r2s_success_args a)
{
    __shared__ double s_sum[kThreads / 32];
    __shared__ int s_cnt[2];
    const int e = blockIdx.x, tid = threadIdx.x, N = a.N;
    const float4* x4 = reinterpret_cast<const float4*>(a.x4) + (size_t)e * N;
    const float sx = a.shift ? a.shift[0] : 0.0f, sy = a.shift ? a.shift[1] : 0.0f, sz = a.shift ? a.shift[2] : 0.0f;
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();

    if (a.ring_slots > 0 && a.ring) {   // the state the reference pickles every frame (eval_policy.py:207-213)
        float* dst = a.ring + ((size_t)(a.frame % a.ring_slots) * a.E + e) * (size_t)N * 3;
        for (int i = tid; i < N; i += kThreads) {
            const float4 p = x4[i];
            dst[3 * i] = __fadd_rn(p.x, sx);
            dst[3 * i + 1] = __fadd_rn(p.y, sy);
            dst[3 * i + 2] = __fadd_rn(p.z, sz);
        }
    }

    double v0 = 0.0, v1 = 0.0;
    bool pass = false;
    if (a.task == R2S_TASK_PUSHT) {
        // ((x - x_target) ** 2).sum(1).mean() (calculate_success_T.py:26): squares and the 3-term sum in float32
        double acc = 0.0;
        for (int i = tid; i < N; i += kThreads) {
            const float4 p = x4[i];
            const float dx = __fsub_rn(__fadd_rn(p.x, sx), a.target[3 * i]);
            const float dy = __fsub_rn(__fadd_rn(p.y, sy), a.target[3 * i + 1]);
            const float dz = __fsub_rn(__fadd_rn(p.z, sz), a.target[3 * i + 2]);
            acc += (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) s_sum[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) t += s_sum[w];
            const float mse = (float)(t / (double)N);
            v0 = mse;
            pass = (double)mse < a.threshold;
        }
    } else if (a.task == R2S_TASK_ROPE) {
        // count_xz_plane_intersections (calculate_success_rope.py:77-134) on the y_min and y_max faces
        int c0 = 0, c1 = 0;
        for (int s = tid; s < a.S; s += kThreads) {
            const int2 sp = reinterpret_cast<const int2*>(a.springs)[s];
            const P3 p0 = load_point(x4, sp.x, sx, sy, sz), p1 = load_point(x4, sp.y, sx, sy, sz);
            c0 += crosses(p0, p1, a.box[1], a.box[0], a.box[3], a.box[2], a.box[5]);
            c1 += crosses(p0, p1, a.box[4], a.box[0], a.box[3], a.box[2], a.box[5]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
        if ((tid & 31) == 0) {
            atomicAdd(&s_cnt[0], c0);
            atomicAdd(&s_cnt[1], c1);
        }
        __syncthreads();
        if (tid == 0) {
            v0 = s_cnt[0]; v1 = s_cnt[1];
            pass = v0 >= a.threshold && v1 >= a.threshold;          // calculate_success_rope.py:167
        }
    } else {
        // OrientedBoundingBox::GetPointIndicesWithinBoundingBox: |d . axis_k| <= extent_k / 2 (float64)
        int c = 0;
        const double* R = a.box + 3;
        for (int i = tid; i < N; i += kThreads) {
            const P3 p = load_point(x4, i, sx, sy, sz);
            const double dx = p.x - a.box[0], dy = p.y - a.box[1], dz = p.z - a.box[2];
            bool in = true;
#pragma unroll
            for (int k = 0; k < 3; ++k) {   // axis k = column k of R
                const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, R[k]), __dmul_rn(dy, R[3 + k])), __dmul_rn(dz, R[6 + k]));
                in = in && fabs(d) <= a.box[12 + k] / 2.0;
            }
            c += in;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0) atomicAdd(&s_cnt[0], c);
        __syncthreads();
        if (tid == 0) {
            v0 = s_cnt[0];
            pass = v0 >= a.threshold;                               // calculate_success_sloth.py:169
        }
    }
    if (tid == 0) {
        a.value[2 * e] = (float)v0;
        a.value[2 * e + 1] = (float)v1;
        a.passed[e] = pass;
        if (a.frame >= a.start_frame && pass) {
            const int h = a.hits[e] + 1;
            a.hits[e] = h;
            if (h >= a.need_frames) a.success[e] = 1;
        }
    }
}

}  // namespace

extern "C" int r2s_success_forward(const r2s_success_args* a, void* stream)
{
    R2S_REQUIRE(a, "r2s_success_forward: null args");
    R2S_REQUIRE(a->E > 0 && a->N > 0 && a->task >= R2S_TASK_PUSHT && a->task <= R2S_TASK_SLOTH,
                "r2s_success_forward: bad sizes E=%d N=%d task=%d", a->E, a->N, a->task);
    R2S_REQUIRE(a->x4 && a->value && a->passed && a->hits && a->success, "r2s_success_forward: null state or output");
    R2S_REQUIRE(((uintptr_t)a->x4 & 15) == 0, "r2s_success_forward: x4 must be 16-byte aligned");
    R2S_REQUIRE(a->task != R2S_TASK_PUSHT || a->target, "r2s_success_forward: the push-T test needs target positions");
    R2S_REQUIRE(a->task != R2S_TASK_ROPE || (a->springs && a->S > 0 && ((uintptr_t)a->springs & 7) == 0),
                "r2s_success_forward: the rope test needs the (8-byte aligned) spring list");
    R2S_REQUIRE(a->ring_slots <= 0 || a->ring, "r2s_success_forward: ring_slots > 0 without a ring buffer");
    R2S_REQUIRE(a->need_frames > 0, "r2s_success_forward: need_frames must be positive");
    success_kernel<<<a->E, kThreads, 0, (cudaStream_t)stream>>>(*a);
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}
