// lbs.cu -- linear-blend skinning of the object Gaussians (sim/utils/gs/transform_utils.py:58-212,
// quat=None path) for E environments: one thread per (env, bone) fits the bone's rotation, one thread
// per (env, Gaussian) blends its k_wgt bone transforms.  HBM-bound gather work; no tensor cores.
#include <cuda_runtime.h>
#include <math.h>

#include "r2s_internal.h"
#include "r2s_lbs.h"

namespace {

// Eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi sweeps (fp32).  Returns the eigenvalues
// in w and the eigenvectors as the COLUMNS of V.
__device__ void jacobi3(float a[3][3], float w[3], float V[3][3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0f : 0.0f;
    for (int sweep = 0; sweep < 8; ++sweep) {
        const float off = fabsf(a[0][1]) + fabsf(a[0][2]) + fabsf(a[1][2]);
        const float diag = fabsf(a[0][0]) + fabsf(a[1][1]) + fabsf(a[2][2]);
        if (off <= 1e-12f * diag || off == 0.0f) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const float apq = a[p][q];
            if (fabsf(apq) < 1e-30f) continue;
            const float theta = (a[q][q] - a[p][p]) / (2.0f * apq);
            const float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
            const float c = 1.0f / sqrtf(t * t + 1.0f), s = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- A J
                const float akp = a[k][p], akq = a[k][q];
                a[k][p] = c * akp - s * akq;
                a[k][q] = s * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- J^T A
                const float apk = a[p][k], aqk = a[q][k];
                a[p][k] = c * apk - s * aqk;
                a[q][k] = s * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq;
                V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
}

__device__ __forceinline__ void cross3(const float* a, const float* b, float* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// One thread per (env, bone): F, its two dominant singular pairs, and the proper rotation
//   R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T
// which equals U diag(1, 1, det(U) det(V)) V^T -- what the reference's SVD + determinant fix-ups
// (transform_utils.py:93-118) produce for every rank >= 2 case, reflections included.
__global__ void lbs_rotation_kernel(const r2s_lbs_args a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)a.E * a.N) return;
    const int e = (int)(t / a.N), i = (int)(t % a.N);
    const float4* b0 = reinterpret_cast<const float4*>(a.bones4) + (size_t)e * a.N;
    const float4* b1 = reinterpret_cast<const float4*>(a.bones_new4) + (size_t)e * a.N;
    const float4 o0 = b0[i], n0 = b1[i];
    float F[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (int k = 0; k < a.k_rel; ++k) {
        const int j = a.relations[(size_t)i * a.k_rel + k];
        const float4 oj = b0[j], nj = b1[j];
        const float ad[3] = {oj.x - o0.x, oj.y - o0.y, oj.z - o0.z};   // adj_bones
        const float an[3] = {nj.x - n0.x, nj.y - n0.y, nj.z - n0.z};   // adj_bones_new
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) F[r][c] += an[r] * ad[c];
    }
    float S[3][3], w[3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) S[r][c] = F[0][r] * F[0][c] + F[1][r] * F[1][c] + F[2][r] * F[2][c];  // F^T F
    jacobi3(S, w, V);
    // order the eigenvalues descending
    int i0 = 0, i1 = 1, i2 = 2;
    if (w[i0] < w[i1]) { int s = i0; i0 = i1; i1 = s; }
    if (w[i0] < w[i2]) { int s = i0; i0 = i2; i2 = s; }
    if (w[i1] < w[i2]) { int s = i1; i1 = i2; i2 = s; }
    // singular values as |F v|, not as roots of the eigenvalues of F^T F: the latter carry absolute noise of
    // ~sqrt(eps) * s1, which would count an exactly rank-1 F (collinear bone neighbourhoods) as rank 2, while
    // |F v2| is good to ~eps * s1 -- the accuracy class of the SVD torch.linalg.matrix_rank runs
    float v1[3] = {V[0][i0], V[1][i0], V[2][i0]}, v2[3] = {V[0][i1], V[1][i1], V[2][i1]}, v3[3];
    float u1[3], u2[3], u3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        u1[r] = F[r][0] * v1[0] + F[r][1] * v1[1] + F[r][2] * v1[2];
        u2[r] = F[r][0] * v2[0] + F[r][1] * v2[1] + F[r][2] * v2[2];
    }
    const float s1 = sqrtf(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    const float s2 = sqrtf(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    // torch.linalg.matrix_rank: singular values above sigma_max * max(m, n) * eps count
    const float tol = s1 * 3.0f * 1.1920929e-07f;
    const bool rank_ok = s2 > tol && s1 > 0.0f;
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (rank_ok) {
        float l = rsqrtf(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        u1[0] *= l; u1[1] *= l; u1[2] *= l;
        const float d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
        u2[0] -= d * u1[0]; u2[1] -= d * u1[1]; u2[2] -= d * u1[2];
        l = rsqrtf(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        u2[0] *= l; u2[1] *= l; u2[2] *= l;
        cross3(u1, u2, u3);
        cross3(v1, v2, v3);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[3 * r + c] = u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c];
    } else {
        atomicAnd(a.rank_flags + e, 0);
    }
    // bone transform as three float4 rows [R | c], c = new - R old, so that the blend is sum_k w_k (R x + c):
    // algebraically the reference's R (x - old) + motion + old
    float4* out = reinterpret_cast<float4*>(a.rot_scratch) + ((size_t)e * a.N + (a.bone_slot ? a.bone_slot[i] : i)) * 3;
    out[0] = make_float4(R[0], R[1], R[2], n0.x - (R[0] * o0.x + R[1] * o0.y + R[2] * o0.z));
    out[1] = make_float4(R[3], R[4], R[5], n0.y - (R[3] * o0.x + R[4] * o0.y + R[5] * o0.z));
    out[2] = make_float4(R[6], R[7], R[8], n0.z - (R[6] * o0.x + R[7] * o0.y + R[8] * o0.z));
}

__global__ void lbs_flag_reset_kernel(int* flags, int E)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) flags[e] = 1;
}

// xyz' = sum_k w_k (R_b (xyz - bone_b) + motion_b + bone_b), evaluated as sum_k w_k (R_b xyz + c_b) with
// c_b = new_b - R_b old_b from the rotation kernel.  grid = (chunks of Gaussians, E): a block first stages its
// environment's N bone transforms (48 B each) in shared memory with coalesced 16-byte loads, then every thread
// gathers its k_wgt bones from there; kStage = false reads them from global memory (N too large to stage).
template <bool kStage>
__global__ void lbs_blend_kernel(const r2s_lbs_args a, int per_block)
{
    extern __shared__ float4 s_T[];  // [N][3]
    const int e = blockIdx.y;
    const float4* Rm = reinterpret_cast<const float4*>(a.rot_scratch) + (size_t)e * a.N * 3;
    const bool use_R = a.rank_flags[e] != 0;   // reference quirk: one deficient bone -> identity for all
    if (kStage && use_R) {
        for (int k = threadIdx.x; k < 3 * a.N; k += blockDim.x) s_T[k] = Rm[k];
        __syncthreads();
    }
    const float4* b0 = reinterpret_cast<const float4*>(a.bones4) + (size_t)e * a.N;
    const float4* b1 = reinterpret_cast<const float4*>(a.bones_new4) + (size_t)e * a.N;
    const int g_end = min(a.n_obj, (int)(blockIdx.x + 1) * per_block);
    const bool vec4 = (a.k_wgt & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.weights_indices) |
                                              reinterpret_cast<uintptr_t>(a.weights) |
                                              reinterpret_cast<uintptr_t>(a.weights_slots) |
                                              reinterpret_cast<uintptr_t>(a.weights_by_slot)) & 15) == 0;
    // `T` is either the shared-memory copy or the global table; the loop is instantiated once per address space
    // (a pointer selected at run time would compile to generic loads, which cost several times an LDS here)
    auto run = [&](const float4* __restrict__ T) {
        for (int g = blockIdx.x * per_block + threadIdx.x; g < g_end; g += blockDim.x) {
            float* x = a.means3D + ((size_t)e * a.P + g) * 3;
            const float px = x[0], py = x[1], pz = x[2];
            float ox = 0.f, oy = 0.f, oz = 0.f;
            // with R: rows of T are addressed by slot (the caller's layout hint) when one is given
            const int* wi = (use_R && a.weights_slots ? a.weights_slots : a.weights_indices) + (size_t)g * a.k_wgt;
            const float* ww = (use_R && a.weights_slots ? a.weights_by_slot : a.weights) + (size_t)g * a.k_wgt;
            auto bone = [&](int b, float wk) {
                float tx, ty, tz;
                if (use_R) {
                    const float4 r0 = T[3 * b], r1 = T[3 * b + 1], r2 = T[3 * b + 2];
                    tx = r0.x * px + r0.y * py + r0.z * pz + r0.w;
                    ty = r1.x * px + r1.y * py + r1.z * pz + r1.w;
                    tz = r2.x * px + r2.y * py + r2.z * pz + r2.w;
                } else {
                    const float4 o = b0[b], n = b1[b];
                    tx = (px - o.x) + (n.x - o.x) + o.x;
                    ty = (py - o.y) + (n.y - o.y) + o.y;
                    tz = (pz - o.z) + (n.z - o.z) + o.z;
                }
                ox += tx * wk; oy += ty * wk; oz += tz * wk;
            };
            if (vec4) {   // rows of k_wgt entries start 16-byte aligned: four bones per 128-bit load
                for (int k = 0; k < a.k_wgt; k += 4) {
                    const int4 b = __ldg(reinterpret_cast<const int4*>(wi + k));
                    const float4 w = __ldg(reinterpret_cast<const float4*>(ww + k));
                    bone(b.x, w.x); bone(b.y, w.y); bone(b.z, w.z); bone(b.w, w.w);
                }
            } else {
                for (int k = 0; k < a.k_wgt; ++k) bone(__ldg(wi + k), __ldg(ww + k));
            }
            x[0] = ox; x[1] = oy; x[2] = oz;
        }
    };
    if (kStage && use_R) run(s_T);
    else run(Rm);
}

}  // namespace

extern "C" int r2s_lbs_forward(const r2s_lbs_args* a, void* stream)
{
    R2S_REQUIRE(a, "r2s_lbs_forward: null args");
    R2S_REQUIRE(a->E > 0 && a->N > 0 && a->P >= a->n_obj && a->n_obj >= 0 && a->k_rel > 0 && a->k_wgt > 0,
                "r2s_lbs_forward: bad sizes E=%d N=%d P=%d n_obj=%d k_rel=%d k_wgt=%d", a->E, a->N, a->P, a->n_obj,
                a->k_rel, a->k_wgt);
    R2S_REQUIRE(a->relations && a->weights_indices && a->weights && a->bones4 && a->bones_new4 && a->means3D &&
                    a->rot_scratch && a->rank_flags,
                "r2s_lbs_forward: null pointer");
    R2S_REQUIRE((a->bone_slot != nullptr) == (a->weights_slots != nullptr) &&
                    (a->bone_slot != nullptr) == (a->weights_by_slot != nullptr),
                "r2s_lbs_forward: bone_slot, weights_slots and weights_by_slot go together");
    cudaStream_t st = (cudaStream_t)stream;
    lbs_flag_reset_kernel<<<r2s::ceil_div(a->E, 256), 256, 0, st>>>(a->rank_flags, a->E);
    R2S_LAUNCH_CHECK();
    lbs_rotation_kernel<<<r2s::ceil_div((long long)a->E * a->N, 128), 128, 0, st>>>(*a);
    R2S_LAUNCH_CHECK();
    if (a->n_obj > 0) {
        const size_t smem = sizeof(float4) * 3 * (size_t)a->N;
        const int per_block = 2048;  // Gaussians per block: amortises the staging of N transforms
        const dim3 grid(r2s::ceil_div(a->n_obj, per_block), a->E);
        if (smem <= 200 * 1024) {
            R2S_CUDA_TRY(cudaFuncSetAttribute(lbs_blend_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            lbs_blend_kernel<true><<<grid, 512, smem, st>>>(*a, per_block);
        } else {
            lbs_blend_kernel<false><<<grid, 512, 0, st>>>(*a, per_block);
        }
        R2S_LAUNCH_CHECK();
    }
    return R2S_OK;
}
