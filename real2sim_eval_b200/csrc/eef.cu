// eef.cu -- per-frame end-effector step on the device, sm_100a.
//
// One CTA per environment restates SpringMassDynamicsModule.step (sim/physics/phystwin.py:362-510) up to
// the call of set_mesh_interactive: grasp hysteresis (one thread, double precision as the reference's
// Python floats), per-substep end-effector poses (one thread per substep, staged in shared memory), then
// one thread per collision-mesh vertex walks the substeps and writes its row of the vertex table.  The
// float32 operations that the reference issues as separate torch kernels (mul, then add) are spelled
// with round-to-nearest intrinsics so nothing is contracted into an FMA the reference does not have.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "r2s_eef.h"
#include "r2s_internal.h"

namespace {

constexpr int kThreads = 256;

// scipy.interpolate.interp1d(np.arange(n)/(n-1), table, axis=0)(x), linear (_call_linear of scipy >= 1.15, the
// reference's pin): searchsorted-left index clipped to [1, n-1];
// y = ((x - x_lo)/(x_hi - x_lo)) * y_hi + ((x_hi - x)/(x_hi - x_lo)) * y_lo in float64; the caller rounds to
// float32 (phystwin.py:417).
struct Interp {
    int lo;
    double w_lo, w_hi;
};

__device__ __forceinline__ Interp interp_setup(double x, int n)
{
    const double den = (double)(n - 1);
    int idx = (int)ceil(x * den);
    idx = max(0, min(n, idx));
    while (idx > 0 && (double)(idx - 1) / den >= x) --idx;   // searchsorted(side='left'): #grid values < x
    while (idx < n && (double)idx / den < x) ++idx;
    idx = max(1, min(n - 1, idx));
    const double x_lo = (double)(idx - 1) / den, x_hi = (double)idx / den;
    Interp it;
    it.lo = idx - 1;
    it.w_hi = (x - x_lo) / (x_hi - x_lo);
    it.w_lo = (x_hi - x) / (x_hi - x_lo);
    return it;
}

__device__ __forceinline__ float interp_eval(const float* __restrict__ table, size_t stride, size_t off, const Interp& it)
{
    const double y_lo = (double)table[(size_t)it.lo * stride + off], y_hi = (double)table[(size_t)(it.lo + 1) * stride + off];
    return (float)__dadd_rn(__dmul_rn(it.w_hi, y_hi), __dmul_rn(it.w_lo, y_lo));
}

// kornia.geometry.conversions.axis_angle_to_rotation_matrix (restated, see r2s_eef.h), row-major 3x3
__device__ void axis_angle_to_matrix(const float aa[3], float R[9])
{
    const float theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > 1e-6f) {
        const float theta = sqrtf(theta2);
        const float d = theta + 1e-6f;
        const float wx = aa[0] / d, wy = aa[1] / d, wz = aa[2] / d;
        const float c = cosf(theta), s = sinf(theta), k = 1.0f - c;
        R[0] = c + wx * wx * k;        R[1] = wx * wy * k - wz * s;   R[2] = wy * s + wx * wz * k;
        R[3] = wz * s + wx * wy * k;   R[4] = c + wy * wy * k;        R[5] = -wx * s + wy * wz * k;
        R[6] = -wy * s + wx * wz * k;  R[7] = wx * s + wy * wz * k;   R[8] = c + wz * wz * k;
    } else {
        R[0] = 1.0f;   R[1] = -aa[2]; R[2] = aa[1];
        R[3] = aa[2];  R[4] = 1.0f;   R[5] = -aa[0];
        R[6] = -aa[1]; R[7] = aa[0];  R[8] = 1.0f;
    }
}

__global__ void __launch_bounds__(kThreads) eef_kernel(const r2s_eef_args a)
{
    extern __shared__ float s_pose[];          // [S][12]: rot_next (9, row-major) + xyz_next (3)
    __shared__ double s_open[2];               // opening now / before (clipped)
    __shared__ float s_close[2][kThreads / 32][3];
    const int e = blockIdx.x, tid = threadIdx.x, S = a.n_substeps, V = a.n_pts;
    const float dtf = (float)a.dt;
    const float* rot = a.eef_rot + 9 * (size_t)e;
    const float* xyz = a.eef_xyz + 3 * (size_t)e;
    const float* vel = a.eef_vel + 3 * (size_t)e;
    const float* rvel = a.eef_rot_vel + 3 * (size_t)e;

    if (tid == 0) {
        if (a.use_pusher) {
            a.current_openness[e] = 1.0;       // phystwin.py:466 "just for placeholding"
            s_open[0] = s_open[1] = 1.0;
        } else {
            // phystwin.py:370-412
            double openness = (double)a.openness_cmd[e];
            double cur = a.current_openness[e];
            int grasped = a.grasped[e];
            if (isnan(cur)) cur = openness;
            float nrm[2] = {0.0f, 0.0f};
            if (a.collision_forces) {
                const float* f = a.collision_forces + 3 * (size_t)e * a.F;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float* f0 = f + 3 * a.force_faces[3 * k];
                    const float* f1 = f + 3 * a.force_faces[3 * k + 1];
                    const float* f2 = f + 3 * a.force_faces[3 * k + 2];
                    const float x = __fadd_rn(__fadd_rn(f0[0], f1[0]), f2[0]);
                    const float y = __fadd_rn(__fadd_rn(f0[1], f1[1]), f2[1]);
                    const float z = __fadd_rn(__fadd_rn(f0[2], f1[2]), f2[2]);
                    nrm[k] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
                }
            }
            const double before = cur;
            if (nrm[0] < 100.0f && nrm[1] < 100.0f) grasped = 0;                  // release
            if (openness < cur) {
                if (nrm[0] > a.grasp_force_threshold && nrm[1] > a.grasp_force_threshold) {
                    openness = cur;
                    grasped = 1;                                                 // establish grasp
                } else if (grasped) {
                    cur = fmax(openness, cur - 0.05);
                    openness = cur;
                } else {
                    cur = openness;
                }
            } else {
                cur = openness;
            }
            a.current_openness[e] = cur;
            a.grasped[e] = grasped;
            s_open[0] = fmin(fmax(openness, 0.0), 1.0);
            s_open[1] = fmin(fmax(before, 0.0), 1.0);
        }
    }
    // per-substep poses (phystwin.py:374-381)
    for (int s = tid; s < S; s += kThreads) {
        const float dts = __fmul_rn((float)(s + 1), dtf);
        float aa[3], D[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) aa[i] = __fmul_rn(rvel[i], dts);
        axis_angle_to_matrix(aa, D);
        float* o = s_pose + 12 * s;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)   // (D^T @ rot)[i][j]
                o[3 * i + j] = D[i] * rot[j] + D[3 + i] * rot[3 + j] + D[6 + i] * rot[6 + j];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float c = __fadd_rn(xyz[i], __fmul_rn(vel[i], dts));
            o[9 + i] = c;
            a.interp_center[((size_t)e * S + s) * 3 + i] = c;     // interpolated_center (phystwin.py:436)
        }
    }
    __syncthreads();

    const double x_now = s_open[0], x_bef = s_open[1];
    const Interp it_now = interp_setup(x_now, a.n_table), it_bef = interp_setup(x_bef, a.n_table);
    const float span = (float)(a.dt * (double)S);                 // dt * n_substeps (Python float -> f32 scalar)
    const float span2 = (float)(2.0 * a.dt * (double)S);   // phystwin.py:442
    const size_t tstride = (size_t)V * 3;
    const float init[3] = {a.init_eef_xyz[0], a.init_eef_xyz[1], a.init_eef_xyz[2]};
    float cl[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    const int half = V / 2;

    for (int pnt = tid; pnt < V; pnt += kThreads) {
        float rel[3], dlt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float now = interp_eval(a.table, tstride, 3 * (size_t)pnt + i, it_now);
            const float bef = interp_eval(a.table, tstride, 3 * (size_t)pnt + i, it_bef);
            const float sg = i == 0 ? 1.0f : -1.0f;                 // flip y, z (phystwin.py:423-428)
            dlt[i] = sg * __fsub_rn(now, bef);
            rel[i] = sg * __fsub_rn(bef, init[i]);
        }
        if (!a.use_pusher) {   // closing velocity (phystwin.py:441-448): (delta @ rot^T) / (2 dt S), mean per finger
            const int k = pnt < half ? 0 : 1;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                cl[k][i] += (dlt[0] * rot[3 * i] + dlt[1] * rot[3 * i + 1] + dlt[2] * rot[3 * i + 2]) / span2;
        }
        float step[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) step[i] = __fdiv_rn(dlt[i], span);
        float* out = a.interp_pts + ((size_t)e * S * V + pnt) * 3;
        for (int s = 0; s < S; ++s, out += tstride) {
            const float* ps = s_pose + 12 * s;
            const float dts = __fmul_rn((float)(s + 1), dtf);
            const float r0 = __fadd_rn(rel[0], __fmul_rn(step[0], dts));
            const float r1 = __fadd_rn(rel[1], __fmul_rn(step[1], dts));
            const float r2 = __fadd_rn(rel[2], __fmul_rn(step[2], dts));
            // eef_xyz_next + relative @ rot_next^T (phystwin.py:432)
            out[0] = __fadd_rn(ps[9], ps[0] * r0 + ps[1] * r1 + ps[2] * r2);
            out[1] = __fadd_rn(ps[10], ps[3] * r0 + ps[4] * r1 + ps[5] * r2);
            out[2] = __fadd_rn(ps[11], ps[6] * r0 + ps[7] * r1 + ps[8] * r2);
        }
    }

    // dynamic_velocity / dynamic_omega (phystwin.py:439-452, 495-501)
    if (!a.use_pusher) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float v = cl[k][i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((tid & 31) == 0) s_close[k][tid >> 5][i] = v;
            }
    }
    __syncthreads();
    if (tid < 3) {
        const float half_v = __fmul_rn(vel[tid], 0.5f);
        float* dv = a.dyn_vel + (size_t)e * a.dyn_vel_rows * 3;
        if (a.use_pusher) {
            dv[tid] = half_v;
        } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float sum = 0.0f;
                for (int w = 0; w < kThreads / 32; ++w) sum += s_close[k][w][tid];
                const int cnt = k == 0 ? half : V - half;
                dv[3 * k + tid] = __fadd_rn(half_v, sum / (float)cnt);
            }
        }
        a.dyn_omega[3 * (size_t)e + tid] = __fmul_rn(-rvel[tid], 0.5f);
    }
}

}  // namespace

extern "C" int r2s_eef_forward(const r2s_eef_args* a, void* stream)
{
    R2S_REQUIRE(a, "r2s_eef_forward: null args");
    R2S_REQUIRE(a->E > 0 && a->n_substeps > 0 && a->n_substeps <= R2S_EEF_MAX_SUBSTEPS && a->n_pts > 0 && a->n_table >= 2,
                "r2s_eef_forward: bad sizes E=%d S=%d V=%d table=%d", a->E, a->n_substeps, a->n_pts, a->n_table);
    R2S_REQUIRE(a->table && a->init_eef_xyz && a->eef_xyz && a->eef_vel && a->eef_rot && a->eef_rot_vel,
                "r2s_eef_forward: null input");
    R2S_REQUIRE(a->current_openness && a->interp_pts && a->interp_center && a->dyn_vel && a->dyn_omega,
                "r2s_eef_forward: null output");
    R2S_REQUIRE(a->dyn_vel_rows >= (a->use_pusher ? 1 : 2), "r2s_eef_forward: dyn_vel_rows too small");
    if (!a->use_pusher) {
        R2S_REQUIRE(a->openness_cmd && a->grasped, "r2s_eef_forward: a gripper needs openness_cmd and grasped");
        R2S_REQUIRE(a->n_pts % 2 == 0, "r2s_eef_forward: a gripper has two fingers of equal vertex count");
        if (a->collision_forces)
            for (int k = 0; k < 6; ++k)
                R2S_REQUIRE(a->force_faces[k] >= 0 && a->force_faces[k] < a->F,
                            "r2s_eef_forward: force_faces[%d]=%d outside [0,%d)", k, a->force_faces[k], a->F);
    }
    const size_t smem = sizeof(float) * 12 * (size_t)a->n_substeps;
    if (smem > 48 * 1024)
        R2S_CUDA_TRY(cudaFuncSetAttribute(eef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eef_kernel<<<a->E, kThreads, smem, (cudaStream_t)stream>>>(*a);
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}
