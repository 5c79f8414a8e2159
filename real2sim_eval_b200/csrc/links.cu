// links.cu -- per-frame rigid re-posing of the robot's Gaussians for E environments:
// sim/utils/robot/robot_pc_transformations.py:12-55 + robot_pc_sampler.py:119-162 + gs_renderer.py:905.
// One thread per (env, link) composes the link transform and its quaternion; one thread per
// (env, Gaussian) applies it.  Streaming HBM work (28 B read from L2-resident shared scan arrays,
// 28 B written per Gaussian); no tensor cores.
#include <cuda_runtime.h>
#include <math.h>

#include "r2s_internal.h"
#include "r2s_links.h"

namespace {

__device__ __forceinline__ void matmul4(const float* a, const float* b, float* c)  // row-major 4x4
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = a[4 * i] * b[j];
#pragma unroll
            for (int k = 1; k < 4; ++k) s += a[4 * i + k] * b[4 * k + j];
            c[4 * i + j] = s;
        }
}

// kornia.geometry.conversions.rotation_matrix_to_quaternion (w,x,y,z), eps = 1e-8: four branches selected
// by the trace / the largest diagonal entry, divisions guarded by clamp(denominator, min=FLT_MIN).
__device__ __forceinline__ float safe_div(float n, float d) { return n / fmaxf(d, 1.17549435e-38f); }

__device__ void rotmat_to_quat(const float* m /* row-major 3x3 with row stride 4 */, float q[4])
{
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[4], m11 = m[5], m12 = m[6], m20 = m[8], m21 = m[9],
                m22 = m[10];
    const float eps = 1e-8f;
    const float trace = m00 + m11 + m22;
    if (trace > 0.0f) {
        const float sq = sqrtf(trace + 1.0f + eps) * 2.0f;  // 4 qw
        q[0] = 0.25f * sq; q[1] = safe_div(m21 - m12, sq); q[2] = safe_div(m02 - m20, sq); q[3] = safe_div(m10 - m01, sq);
    } else if (m00 > m11 && m00 > m22) {
        const float sq = sqrtf(1.0f + m00 - m11 - m22 + eps) * 2.0f;  // 4 qx
        q[0] = safe_div(m21 - m12, sq); q[1] = 0.25f * sq; q[2] = safe_div(m01 + m10, sq); q[3] = safe_div(m02 + m20, sq);
    } else if (m11 > m22) {
        const float sq = sqrtf(1.0f + m11 - m00 - m22 + eps) * 2.0f;  // 4 qy
        q[0] = safe_div(m02 - m20, sq); q[1] = safe_div(m01 + m10, sq); q[2] = 0.25f * sq; q[3] = safe_div(m12 + m21, sq);
    } else {
        const float sq = sqrtf(1.0f + m22 - m00 - m11 + eps) * 2.0f;  // 4 qz
        q[0] = safe_div(m10 - m01, sq); q[1] = safe_div(m02 + m20, sq); q[2] = safe_div(m12 + m21, sq); q[3] = 0.25f * sq;
    }
}

// robot_pc_sampler.py:138-150: mat = (pose @ offset) @ rest_inv, quat = rotation_matrix_to_quaternion(mat[:3,:3])
__global__ void link_compose_kernel(const r2s_links_args a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.E * a.L) return;
    const int l = i % a.L;
    float pose[16], off[16], inv[16], t1[16], m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        pose[k] = a.link_pose[(size_t)i * 16 + k];
        off[k] = a.link_offset[l * 16 + k];
        inv[k] = a.rest_inv[l * 16 + k];
    }
    matmul4(pose, off, t1);
    matmul4(t1, inv, m);
    float q[4];
    rotmat_to_quat(m, q);
    float* out = a.link_scratch + (size_t)i * 16;
#pragma unroll
    for (int k = 0; k < 12; ++k) out[k] = m[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) out[12 + k] = q[k];
}

__device__ __forceinline__ void normalize4(float q[4])  // torch.nn.functional.normalize(dim=-1), eps 1e-12
{
    const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// grid = (ceil(n_robot / 256), E)
__global__ void __launch_bounds__(256) link_apply_kernel(const r2s_links_args a)
{
    __shared__ float s_T[R2S_LINKS_MAX * 16];
    const int e = blockIdx.y;
    for (int k = threadIdx.x; k < a.L * 16; k += blockDim.x) s_T[k] = a.link_scratch[(size_t)e * a.L * 16 + k];
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_robot) return;
    const int l = __ldg(a.link_id + g);
    float p[3] = {__ldg(a.rest_means + 3 * g), __ldg(a.rest_means + 3 * g + 1), __ldg(a.rest_means + 3 * g + 2)};
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(a.rest_quats) + g);
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    normalize4(q);                                        // robot_pc_transformations.py:29
    if (l >= 0 && l < a.L) {
        const float* T = s_T + 16 * l;
        // p @ R^T + t  (robot_pc_sampler.py:151)
        const float x = p[0] * T[0] + p[1] * T[1] + p[2] * T[2] + T[3];
        const float y = p[0] * T[4] + p[1] * T[5] + p[2] * T[6] + T[7];
        const float z = p[0] * T[8] + p[1] * T[9] + p[2] * T[10] + T[11];
        p[0] = x; p[1] = y; p[2] = z;
        // quat_mult_torch(q_link, q)  (robot_pc_sampler.py:17-24)
        const float w1 = T[12], x1 = T[13], y1 = T[14], z1 = T[15];
        const float w2 = q[0], x2 = q[1], y2 = q[2], z2 = q[3];
        q[0] = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2;
        q[1] = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2;
        q[2] = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2;
        q[3] = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2;
    }
    normalize4(q);                                        // gs_renderer.py:905
    const size_t row = (size_t)e * a.P + a.first + g;
    float* om = a.means3D + 3 * row;
    om[0] = p[0]; om[1] = p[1]; om[2] = p[2];
    reinterpret_cast<float4*>(a.rotations)[row] = make_float4(q[0], q[1], q[2], q[3]);
}

}  // namespace

extern "C" int r2s_links_forward(const r2s_links_args* a, void* stream)
{
    R2S_REQUIRE(a, "r2s_links_forward: null args");
    R2S_REQUIRE(a->E > 0 && a->L > 0 && a->L <= R2S_LINKS_MAX && a->n_robot >= 0 && a->first >= 0 &&
                    (long long)a->first + a->n_robot <= a->P,
                "r2s_links_forward: bad sizes E=%d L=%d P=%d first=%d n_robot=%d", a->E, a->L, a->P, a->first, a->n_robot);
    R2S_REQUIRE(a->link_pose && a->link_offset && a->rest_inv && a->link_scratch, "r2s_links_forward: null link table");
    R2S_REQUIRE(a->n_robot == 0 || (a->link_id && a->rest_means && a->rest_quats && a->means3D && a->rotations),
                "r2s_links_forward: null Gaussian array");
    R2S_REQUIRE(((uintptr_t)a->rest_quats & 15) == 0 && ((uintptr_t)a->rotations & 15) == 0,
                "r2s_links_forward: quaternion arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    link_compose_kernel<<<r2s::ceil_div((long long)a->E * a->L, 128), 128, 0, st>>>(*a);
    R2S_LAUNCH_CHECK();
    if (a->n_robot > 0) {
        link_apply_kernel<<<dim3(r2s::ceil_div(a->n_robot, 256), a->E), 256, 0, st>>>(*a);
        R2S_LAUNCH_CHECK();
    }
    return R2S_OK;
}
