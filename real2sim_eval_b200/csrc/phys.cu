// phys.cu -- batched PhysTwin spring-mass substep loop for sm_100a.
//
// What the reference runs as ~9 Warp kernels x num_substeps inside a CUDA graph
// (sim/physics/spring_mass_warp.py:823-943, "SMW") runs here as ONE persistent
// launch per frame: one CTA per environment, particle positions/velocities
// resident in shared memory as float4 for every substep of the frame, spring
// topology (shared by all environments) streamed from L2, per-environment rest
// lengths streamed from HBM.  The force scatter with float atomics (SMW:103-104)
// becomes an atomic-free gather over a per-particle adjacency list: the spring
// force is exactly antisymmetric under endpoint swap, so the force on particle i
// is the sum over its incident springs of F(x_i -> x_j).
//
// Kernel map (reference kernel -> here):
//   eval_springs + update_vel_from_force   SMW:61-129   -> frame_kernel phase A
//   object_collision (+loop)               SMW:132-268  -> frame_kernel phase B
//   set_mesh_points / refit / zero forces  SMW:889-900  -> frame_kernel phase A prologue (warp 0/1)
//   mesh_collision                         SMW:295-421  -> frame_kernel phase C
//   integrate_ground_collision             SMW:424-474  -> frame_kernel phase C
//   HashGrid.build + update_potential_collision / build_resting_collision_pairs
//                                          SMW:196-227, 272-291 -> grid_kernel<false/true>
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <map>
#include <vector>

#include "r2s_internal.h"
#include "r2s_phys.h"

namespace {

// ------------------------------------------------------------------ vec3 helpers
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float len3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize3(float3 a)
{
    float l = len3(a);
    return l > 0.0f ? a / l : f3(0.f, 0.f, 0.f);
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float3 xyz(float4 a) { return f3(a.x, a.y, a.z); }

// ------------------------------------------------------------------ kernel params
struct FrameParams {
    int E, N, n_sub, has_self_collision, use_pusher, sign_mode, precise;
    int V, F, n_dyn, coll_cap, stage_dyn, smem_forces;
    float dt, dashpot, drag_damping, rf, coll_dist;
    float c_elas, c_fric, ce_elas, ce_fric, cs_elas, cs_fric;
    const int* row_ptr;     // [N+1]
    const int2* nbr_k;      // [nd] {neighbour, float bits of clamped stiffness (<0: inactive)}
    const float* rest;      // [E or 1][nd]: rest length (precise) or its reciprocal (fast)
    long long rest_stride;  // nd or 0
    const float* mass;      // [N]
    int unit_mass;          // all masses are exactly 1.0f: x / m == x, the three IEEE divisions per particle are skipped
    const int* mask;        // [N]
    float4* x4;             // [E][N]
    float4* v4;             // [E][N]
    float* vb_scratch;      // [E][3][N]  (only when state does not fit shared memory)
    const int* coll_num;    // [E][N]
    const int* coll_idx;    // [E][N][cap]
    const int* status;      // [E][4]
    const float* stat_verts;  // [V][3]  rest-pose vertices (dynamic ones are overridden per substep)
    const int* faces;         // [F][3]
    const int* mesh_map;      // [F]
    const int* face_map;      // [F]
    const int* dyn_part;      // [n_dyn] part (0..kMaxParts-1) of each dynamic vertex: one bounding box per part
    int grp[15];              // face groups {first, end, closed} x (kMaxParts + 1); n_grp = 0 disables grouping
    int n_grp;
    const float* interp_pts;  // [(E)][n_sub_table][n_dyn][3]
    long long interp_stride;  // per env (0 = shared)
    const float* interp_center;  // [(E)][n_sub_table][3]
    long long center_stride;
    const float* dyn_vel;  // [(E)][2][3]
    long long dynvel_stride;
    const float* dyn_omega;  // [(E)][3]
    long long omega_stride;
    float* coll_forces;  // [E][F][3]
    // rigid dynamic mesh accelerator (set when the whole mesh is one rigidly moving tool, e.g. the pusher)
    int accel;                  // 0: brute force over the faces; 1: uniform grid + pseudonormal sign
    int F_dyn;                  // faces [0, F_dyn) are the rigid tool (gridded); [F_dyn, F) static obstacles (scanned)
    const float* frec;          // [F][24] rest-frame face record: v0 v1 v2 | n | e01 e12 e20 | pad
    const float* vnorm;         // [V][3] angle-weighted vertex pseudonormals
    const int* cell_start;      // [ncell + 1]
    const unsigned char* cell_skip;  // [ncell] Chebyshev distance in cells to the nearest non-empty cell
    const int* cell_tris;       // triangle ids per cell
    float gx0, gy0, gz0, gh;    // grid origin, cell edge
    int gnx, gny, gnz;
    int a0, a1, a2;             // anchor vertices (non-collinear) that define the rigid frame
    float rest_frame[12];       // rest basis f1 f2 f3 (rows) and rest anchor r0
    float rest_box[6];          // rest-pose bounding box lo / hi
};

// ------------------------------------------------------------------ mesh query
constexpr int kMaxParts = 4;                               // separately boxed parts of the dynamic mesh
constexpr float kBoxGrow = 0.005f * 1.0001f + 1e-6f;       // the largest contact margin (+ slack)

struct MeshView {
    const float* dyn;   // staged (shared) or table row (global): n_dyn x 3
    const float* stat;  // global rest-pose vertices
    const int* faces;
    int n_dyn, F;
    // face groups (one per dynamic part + one for the static faces): [k] = {first face, end face, closed?}, and
    // their bounding boxes of this substep (6 floats each, lo then hi); n_grp = 0: no grouping, scan every face
    const int* grp;
    const float* boxes;
    int n_grp;
    __device__ __forceinline__ float3 vert(int idx) const
    {
        const float* p = idx < n_dyn ? dyn + 3 * idx : stat + 3 * idx;
        return f3(p[0], p[1], p[2]);
    }
};

// Closest point on triangle (Ericson 5.1.5) as barycentrics (u, v):
// point = u*a + v*b + (1-u-v)*c, the form wp.mesh_eval_position evaluates.
__device__ __forceinline__ void closest_bary(float3 a, float3 b, float3 c, float3 p, float& u, float& v)
{
    float3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { u = 1.0f; v = 0.0f; return; }
    float3 bp = p - b;
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { u = 0.0f; v = 1.0f; return; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float t = d1 / (d1 - d3);
        u = 1.0f - t; v = t; return;
    }
    float3 cp = p - c;
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { u = 0.0f; v = 0.0f; return; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        u = 1.0f - w; v = 0.0f; return;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        u = 0.0f; v = 1.0f - w; return;
    }
    float denom = 1.0f / (va + vb + vc);
    float vv = vb * denom, ww = vc * denom;
    u = 1.0f - vv - ww; v = vv;
}

// As closest_bary, also reporting the Voronoi region of the closest point:
// 0 face interior, 1/2/3 vertex a/b/c, 4 edge ab, 5 edge ac, 6 edge bc.
__device__ __forceinline__ int closest_bary_region(float3 a, float3 b, float3 c, float3 p, float& u, float& v)
{
    float3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { u = 1.0f; v = 0.0f; return 1; }
    float3 bp = p - b;
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { u = 0.0f; v = 1.0f; return 2; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float t = d1 / (d1 - d3);
        u = 1.0f - t; v = t; return 4;
    }
    float3 cp = p - c;
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { u = 0.0f; v = 0.0f; return 3; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        u = 1.0f - w; v = 0.0f; return 5;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        u = 0.0f; v = 1.0f - w; return 6;
    }
    float denom = 1.0f / (va + vb + vc);
    float vv = vb * denom, ww = vc * denom;
    u = 1.0f - vv - ww; v = vv;
    return 0;
}

// Closest point of a RIGID mesh through a uniform grid built once in its rest frame (one warp per query,
// p already transformed into the rest frame).  Shells of cells around p's cell are visited in growing
// Chebyshev radius r; every triangle closer than r*h has been seen after shell r, so the search stops when the
// best distance is <= r*h (or r*h exceeds max_dist).  Same result as the brute-force scan: strictly smaller
// squared distance from max_dist^2, ties to the lower face index.  The sign is the angle-weighted pseudonormal
// test at the closest feature (Baerentzen & Aanaes): for a closed mesh it is the inside/outside the winding
// number > 0.6 test yields, without a sum over all faces.  Returns the closest point in the rest frame.
struct GridView {  // the fields of FrameParams the grid query needs, passed by value (kernel parameters live in
                   // the constant bank; taking a reference to the whole struct would copy it to local memory)
    const float* frec; const float* vnorm; const int* cell_start; const int* cell_tris; const int* faces;
    const unsigned char* cell_skip;
    float gx0, gy0, gz0, gh;
    int gnx, gny, gnz, sign_mode;
};

constexpr int kTriQueue = 256;   // triangle ids a warp collects before testing them 32 at a time

__device__ __noinline__ bool mesh_query_grid(const GridView p, float3 q, float max_dist, int* queue, int& face,
                                             float3& c_rest, float& sign)
{
    const unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float inv_h = 1.0f / p.gh;
    const int cx = (int)floorf((q.x - p.gx0) * inv_h), cy = (int)floorf((q.y - p.gy0) * inv_h),
              cz = (int)floorf((q.z - p.gz0) * inv_h);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= p.gnx || cy >= p.gny || cz >= p.gnz) return false;  // farther than max_dist
    float best = max_dist * max_dist;
    int hit = 0x7fffffff, region = 0;
    float bu = 0.f, bv = 0.f;
    auto test = [&](int fc) {
        const float* fr = p.frec + (size_t)fc * 24;
        const float3 a = f3(fr[0], fr[1], fr[2]), b = f3(fr[3], fr[4], fr[5]), c = f3(fr[6], fr[7], fr[8]);
        float uu, vv;
        const int reg = closest_bary_region(a, b, c, q, uu, vv);
        const float3 cp = a * uu + b * vv + c * (1.0f - uu - vv);
        const float3 d = cp - q;
        const float d2 = dot3(d, d);
        if (d2 < best || (d2 == best && fc < hit)) { best = d2; hit = fc; bu = uu; bv = vv; region = reg; }
    };
    // Lanes enumerate the cells of a shell, but most cells are empty: the triangle ids of the non-empty ones are
    // first compacted into the warp's queue and then tested one per lane, so the (expensive) closest-point test
    // runs on full warps.  The result does not depend on the order (minimum distance, ties to the lower face).
    int qn = 0;   // warp-uniform fill of the queue
    auto drain = [&]() {
        __syncwarp();
        for (int j = lane; j < qn; j += 32) test(queue[j]);
        __syncwarp();
        qn = 0;
    };
    const int r_max = (int)ceilf(max_dist * inv_h) + 1;
    // shells closer than the nearest non-empty cell hold no triangle: start there (nothing in reach: no hit)
    const int r_first = p.cell_skip[(cz * p.gny + cy) * p.gnx + cx];
    for (int r = r_first; r <= r_max; ++r) {
        const int w = 2 * r + 1, n_cells = w * w * w;
        for (int k0 = 0; k0 < n_cells; k0 += 32) {
            const int k = k0 + lane;
            int t0 = 0, cnt = 0;
            if (k < n_cells) {
                const int dz = k / (w * w) - r, dy = (k / w) % w - r, dx = k % w - r;
                const int x = cx + dx, y = cy + dy, z = cz + dz;
                // interior cells were visited by earlier shells
                if (max(max(abs(dx), abs(dy)), abs(dz)) == r && x >= 0 && y >= 0 && z >= 0 && x < p.gnx && y < p.gny &&
                    z < p.gnz) {
                    const int cell = (z * p.gny + y) * p.gnx + x;
                    t0 = p.cell_start[cell];
                    cnt = p.cell_start[cell + 1] - t0;
                }
            }
            if (!__any_sync(kFull, cnt > 0)) continue;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += u;
            }
            const int total = __shfl_sync(kFull, incl, 31);
            if (total > kTriQueue) {   // a batch of very crowded cells: test them cell by cell
                for (int t = t0; t < t0 + cnt; ++t) test(p.cell_tris[t]);
                continue;
            }
            if (qn + total > kTriQueue) drain();
            int* dst = queue + qn + incl - cnt;
            for (int i = 0; i < cnt; ++i) dst[i] = p.cell_tris[t0 + i];
            qn += total;
        }
        drain();
        float wb = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wb = fminf(wb, __shfl_xor_sync(kFull, wb, o));
        const float reach = (float)r * p.gh;
        if (wb <= reach * reach || reach > max_dist) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(kFull, best, o);
        const int oh = __shfl_xor_sync(kFull, hit, o);
        const float ou = __shfl_xor_sync(kFull, bu, o), ov = __shfl_xor_sync(kFull, bv, o);
        const int oreg = __shfl_xor_sync(kFull, region, o);
        if (ob < best || (ob == best && oh < hit)) { best = ob; hit = oh; bu = ou; bv = ov; region = oreg; }
    }
    if (hit == 0x7fffffff) return false;
    face = hit;
    const float* fr = p.frec + (size_t)hit * 24;
    const float3 a = f3(fr[0], fr[1], fr[2]), b = f3(fr[3], fr[4], fr[5]), c = f3(fr[6], fr[7], fr[8]);
    c_rest = a * bu + b * bv + c * (1.0f - bu - bv);
    float3 n;
    if (region == 0) n = f3(fr[9], fr[10], fr[11]);
    else if (region == 4) n = f3(fr[12], fr[13], fr[14]);       // edge ab
    else if (region == 6) n = f3(fr[15], fr[16], fr[17]);       // edge bc
    else if (region == 5) n = f3(fr[18], fr[19], fr[20]);       // edge ca
    else {
        const int vid = p.faces[3 * hit + (region - 1)];
        n = f3(p.vnorm[3 * vid], p.vnorm[3 * vid + 1], p.vnorm[3 * vid + 2]);
    }
    sign = (p.sign_mode == 1 || dot3(q - c_rest, n) >= 0.0f) ? 1.0f : -1.0f;
    return true;
}

__device__ __forceinline__ float solid_angle(float3 a, float3 b, float3 c, float3 p)
{
    a = a - p; b = b - p; c = c - p;
    float la = len3(a), lb = len3(b), lc = len3(c);
    float det = dot3(a, cross3(b, c));
    float den = la * lb * lc + dot3(a, b) * lc + dot3(b, c) * la + dot3(c, a) * lb;
    return 2.0f * atan2f(det, den);
}

// wp.mesh_query_point_sign_winding_number restated brute force (SMW:322-324), evaluated by
// one WARP per query point: lanes stride the faces, then an arg-min / sum reduction.
// Semantics are those of a sequential scan: strictly smaller squared distance from
// max_dist^2 wins and ties go to the lower face index; sign from the exact winding number
// against `threshold`.  All lanes return the same result.
__device__ __noinline__ bool mesh_query_warp(const MeshView& m, float3 p, float max_dist, float threshold,
                                             int sign_mode, int& face, float& u, float& v, float& sign)
{
    const unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float best = max_dist * max_dist;
    int hit = 0x7fffffff;
    float bu = 0.f, bv = 0.f;
    auto scan = [&](int f0, int f1) {
        for (int fc = f0 + lane; fc < f1; fc += 32) {
            const int* t = m.faces + 3 * fc;
            float3 a = m.vert(t[0]), b = m.vert(t[1]), c = m.vert(t[2]);
            float uu, vv;
            closest_bary(a, b, c, p, uu, vv);
            float3 q = a * uu + b * vv + c * (1.0f - uu - vv);
            float3 d = q - p;
            float d2 = dot3(d, d);
            if (d2 < best || (d2 == best && fc < hit)) { best = d2; hit = fc; bu = uu; bv = vv; }
        }
    };
    // squared distance from p to each group's box (0 inside): no face of the group is closer than that, so groups
    // are scanned nearest box first and the scan stops at the first box farther than the best hit so far --
    // the same arg-min (ties to the lower face) as the scan over all faces
    float gd[kMaxParts + 1];
#pragma unroll
    for (int k = 0; k <= kMaxParts; ++k) {
        gd[k] = 3e38f;
        if (k < m.n_grp && m.grp[3 * k + 1] > m.grp[3 * k]) {
            const float* b = m.boxes + 6 * k;
            const float ex = fmaxf(fmaxf(b[0] - p.x, p.x - b[3]), 0.0f), ey = fmaxf(fmaxf(b[1] - p.y, p.y - b[4]), 0.0f),
                        ez = fmaxf(fmaxf(b[2] - p.z, p.z - b[5]), 0.0f);
            gd[k] = ex * ex + ey * ey + ez * ez;
        }
    }
    if (m.n_grp == 0) {
        scan(0, m.F);
    } else {
        unsigned seen = 0u;
        for (int it = 0; it < m.n_grp; ++it) {
            int kmin = -1;
            float dmin = 3e38f;
#pragma unroll
            for (int k = 0; k <= kMaxParts; ++k)
                if (!((seen >> k) & 1u) && gd[k] < dmin) { dmin = gd[k]; kmin = k; }
            if (kmin < 0) break;
            seen |= 1u << kmin;
            float wb = best;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wb = fminf(wb, __shfl_xor_sync(kFull, wb, o));
            if (dmin > wb) break;   // every remaining group is at least this far
            scan(m.grp[3 * kmin], m.grp[3 * kmin + 1]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(kFull, best, o);
        const int oh = __shfl_xor_sync(kFull, hit, o);
        const float ou = __shfl_xor_sync(kFull, bu, o);
        const float ov = __shfl_xor_sync(kFull, bv, o);
        // lanes that found nothing keep hit = INT_MAX and best = max_dist^2, so they never win a tie
        if (ob < best || (ob == best && oh < hit)) { best = ob; hit = oh; bu = ou; bv = ov; }
    }
    if (hit == 0x7fffffff) return false;
    face = hit; u = bu; v = bv;
    if (sign_mode == 1) { sign = 1.0f; return true; }
    float total = 0.0f;
    auto wind = [&](int f0, int f1) {
        for (int fc = f0 + lane; fc < f1; fc += 32) {
            const int* t = m.faces + 3 * fc;
            total += solid_angle(m.vert(t[0]), m.vert(t[1]), m.vert(t[2]), p);
        }
    };
    if (m.n_grp == 0) {
        wind(0, m.F);
    } else {
        // a CLOSED group (every edge shared by two opposite half-edges) subtends a total solid angle of exactly 0
        // from any point outside it, in particular from outside its box: only open groups and groups whose box
        // contains p are summed (the omitted terms are rounding noise ~1e-7 against the 0.6 threshold)
#pragma unroll
        for (int k = 0; k <= kMaxParts; ++k)
            if (k < m.n_grp && !(m.grp[3 * k + 2] && gd[k] > 0.0f)) wind(m.grp[3 * k], m.grp[3 * k + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
    float wn = total * 0.25f * 0.31830988618379067f;
    sign = (wn > threshold) ? -1.0f : 1.0f;
    return true;
}

__device__ __forceinline__ float3 mesh_eval(const MeshView& m, int face, float u, float v)
{
    const int* t = m.faces + 3 * face;
    return m.vert(t[0]) * u + m.vert(t[1]) * v + m.vert(t[2]) * (1.0f - u - v);
}

// ------------------------------------------------------------------ frame kernel
// G = lanes cooperating on one particle's adjacency row in phase A.
// kPrecise: IEEE sqrt / divisions in the reference's expression order (SMW:87-99).
// !kPrecise: one rsqrt + Newton step gives 1/len, the rest length is stored as its reciprocal;
// same formula, ~4x fewer instructions per spring, results within a few ulp of the precise path.
template <int G, bool kSmemState, bool kPrecise, bool kAccel>
__global__ void __launch_bounds__(1024, 1) frame_kernel(const FrameParams p)
{
    extern __shared__ float4 smem4[];
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int N = p.N;

    // ---- carve shared memory (pointers derived directly from the shared array so that the
    //      compiler emits LDS/STS rather than generic loads)
    const int Npad = (N + 3) & ~3;
    float4* const sx = kSmemState ? smem4 : p.x4 + (size_t)e * N;
    float4* const sv = kSmemState ? smem4 + N : p.v4 + (size_t)e * N;
    float* const svb = kSmemState ? reinterpret_cast<float*>(smem4 + 2 * N) : p.vb_scratch + (size_t)e * 3 * N;  // [3][N]
    unsigned char* sp = reinterpret_cast<unsigned char*>(smem4) +
                        (kSmemState ? sizeof(float4) * 2 * N + sizeof(float) * 3 * Npad : 0);
    // boxes: kMaxParts parts of the dynamic mesh (one per finger / tool) at [6k..6k+5], static vertices at [24..29]
    float* s_aabb = reinterpret_cast<float*>(sp); sp += sizeof(float) * 32;
    int* s_ncand = reinterpret_cast<int*>(s_aabb + 30);                      // candidates of this substep
    float* s_rigid = reinterpret_cast<float*>(sp); sp += sizeof(float) * 16; // world = R rest + t of this substep (accel)
    int* s_cand = reinterpret_cast<int*>(sp);                                // particles near the mesh
    if (p.F > 0) sp += sizeof(int) * ((N + 3) & ~3);
    float* s_dyn = reinterpret_cast<float*>(sp);
    if (p.stage_dyn) sp += sizeof(float) * 3 * ((p.n_dyn + 3) & ~3);
    float* s_forces = reinterpret_cast<float*>(sp);
    if (p.smem_forces) sp += sizeof(float) * 3 * p.F;
    int* s_triq = reinterpret_cast<int*>(sp);                                // [warps][kTriQueue] (grid accelerator only)
    __shared__ int s_grp[3 * (kMaxParts + 1)];
    if (tid < 3 * (kMaxParts + 1)) s_grp[tid] = p.grp[tid];

    const bool has_mesh = p.F > 0;
    const float dt = p.dt, rf = p.rf;
    const float drag = expf(-dt * p.drag_damping);  // SMW:123

    // ---- load state
    if (kSmemState) {
        const float4* gx = p.x4 + (size_t)e * N;
        const float4* gv = p.v4 + (size_t)e * N;
        for (int i = tid; i < N; i += nthreads) { sx[i] = gx[i]; sv[i] = gv[i]; }
    }
    // static part of the mesh bounding box (vertices n_dyn..V-1), once
    if (has_mesh && tid < 32) {
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        for (int k = p.n_dyn + tid; k < p.V; k += 32)
            for (int c = 0; c < 3; ++c) {
                float q = p.stat_verts[3 * k + c];
                lo[c] = fminf(lo[c], q); hi[c] = fmaxf(hi[c], q);
            }
        for (int c = 0; c < 3; ++c)
            for (int o = 16; o > 0; o >>= 1) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
        if (tid == 0)
            for (int c = 0; c < 3; ++c) { s_aabb[24 + c] = lo[c]; s_aabb[27 + c] = hi[c]; }
    }
    const bool run_B = p.has_self_collision && p.status[4 * e] > 0;
    const float* rest = p.rest + (size_t)e * p.rest_stride;
    const float* g_interp = has_mesh ? p.interp_pts + (size_t)e * p.interp_stride : nullptr;
    const float* g_center = has_mesh ? p.interp_center + (size_t)e * p.center_stride : nullptr;
    const float* g_dynvel = has_mesh ? p.dyn_vel + (size_t)e * p.dynvel_stride : nullptr;
    const float* g_omega = has_mesh ? p.dyn_omega + (size_t)e * p.omega_stride : nullptr;
    float* g_forces = has_mesh ? p.coll_forces + (size_t)e * p.F * 3 : nullptr;
    float* forces = p.smem_forces ? s_forces : g_forces;
    // collision_forces is zeroed before every substep (SMW:900), so only the LAST substep's contacts are ever
    // visible: clear once here and accumulate only in that substep
    if (has_mesh)
        for (int k = tid; k < 3 * p.F; k += nthreads) forces[k] = 0.0f;
    __syncthreads();

    constexpr unsigned kFull = 0xffffffffu;
    const int lane_g = tid % G;
    const int grp = tid / G;
    const int n_groups = nthreads / G;
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << ((tid & 31) / G * G));

    for (int step = 0; step < p.n_sub; ++step) {
        // ============ phase A prologue: mesh vertices of this substep, force clear
        if (has_mesh) {
            const float* row = g_interp + (size_t)step * p.n_dyn * 3;
            if (kAccel) {
                // rigid tool: fit world = R rest + t from three anchor vertices of this substep's row, and take
                // the bounding box of the transformed rest box (no per-vertex pass, no refit)
                if (tid == 0) {
                    const float3 c0 = f3(row[3 * p.a0], row[3 * p.a0 + 1], row[3 * p.a0 + 2]);
                    const float3 c1 = f3(row[3 * p.a1], row[3 * p.a1 + 1], row[3 * p.a1 + 2]);
                    const float3 c2 = f3(row[3 * p.a2], row[3 * p.a2 + 1], row[3 * p.a2 + 2]);
                    const float3 e1 = normalize3(c1 - c0);
                    const float3 e3 = normalize3(cross3(c1 - c0, c2 - c0));
                    const float3 e2 = cross3(e3, e1);
                    const float* f = p.rest_frame;  // rows f1, f2, f3, then r0
                    float R[9];
                    for (int r = 0; r < 3; ++r) {
                        const float er[3] = {r == 0 ? e1.x : (r == 1 ? e1.y : e1.z), r == 0 ? e2.x : (r == 1 ? e2.y : e2.z),
                                             r == 0 ? e3.x : (r == 1 ? e3.y : e3.z)};
                        for (int c = 0; c < 3; ++c) R[3 * r + c] = er[0] * f[c] + er[1] * f[3 + c] + er[2] * f[6 + c];
                    }
                    const float3 r0 = f3(f[9], f[10], f[11]);
                    const float3 t = c0 - f3(R[0] * r0.x + R[1] * r0.y + R[2] * r0.z, R[3] * r0.x + R[4] * r0.y + R[5] * r0.z,
                                             R[6] * r0.x + R[7] * r0.y + R[8] * r0.z);
                    for (int k = 0; k < 9; ++k) s_rigid[k] = R[k];
                    s_rigid[9] = t.x; s_rigid[10] = t.y; s_rigid[11] = t.z;
                    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
                    for (int corner = 0; corner < 8; ++corner) {
                        const float bx = p.rest_box[(corner & 1) ? 3 : 0], by = p.rest_box[(corner & 2) ? 4 : 1],
                                    bz = p.rest_box[(corner & 4) ? 5 : 2];
                        const float w[3] = {R[0] * bx + R[1] * by + R[2] * bz + t.x, R[3] * bx + R[4] * by + R[5] * bz + t.y,
                                            R[6] * bx + R[7] * by + R[8] * bz + t.z};
                        for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], w[c]); hi[c] = fmaxf(hi[c], w[c]); }
                    }
                    for (int c = 0; c < 3; ++c) { s_aabb[c] = lo[c]; s_aabb[3 + c] = hi[c]; }
                    for (int k = 1; k < kMaxParts; ++k)
                        for (int c = 0; c < 3; ++c) { s_aabb[6 * k + c] = 3e38f; s_aabb[6 * k + 3 + c] = -3e38f; }
                    *s_ncand = 0;
                }
            } else if (tid < 32 * kMaxParts) {  // SMW:889-899 set_mesh_points + refit (bounds): warp w = part w
                const int part = tid >> 5, ln = tid & 31;
                float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
                for (int k = ln; k < p.n_dyn; k += 32) {
                    if (p.dyn_part[k] != part) continue;
                    for (int c = 0; c < 3; ++c) {
                        float q = row[3 * k + c];
                        if (p.stage_dyn) s_dyn[3 * k + c] = q;
                        lo[c] = fminf(lo[c], q); hi[c] = fmaxf(hi[c], q);
                    }
                }
                for (int c = 0; c < 3; ++c)
                    for (int o = 16; o > 0; o >>= 1) {
                        lo[c] = fminf(lo[c], __shfl_xor_sync(kFull, lo[c], o));
                        hi[c] = fmaxf(hi[c], __shfl_xor_sync(kFull, hi[c], o));
                    }
                if (ln == 0)
                    for (int c = 0; c < 3; ++c) { s_aabb[6 * part + c] = lo[c]; s_aabb[6 * part + 3 + c] = hi[c]; }
                if (tid == 0) *s_ncand = 0;
            }
        }
        // ============ phase A: spring forces (gather) + velocity update -> svb
        for (int i = grp; i < N; i += n_groups) {
            const float4 xi4 = sx[i], vi4 = sv[i];
            const float3 xi = xyz(xi4), vi = xyz(vi4);
            const int beg = __ldg(p.row_ptr + i), end = __ldg(p.row_ptr + i + 1);
            float3 acc = f3(0.f, 0.f, 0.f);
            // software pipeline: the {neighbour, stiffness} / rest-length loads of trip t+1 (L2 / HBM) are
            // issued before the arithmetic of trip t
            int k = beg + lane_g;
            bool have = k < end;
            int2 nk_next = make_int2(0, 0);
            float r_next = 1.0f;
            if (have) { nk_next = __ldg(p.nbr_k + k); r_next = __ldg(rest + k); }
            while (have) {
                const int2 nk = nk_next;
                const float r = r_next;
                k += G;
                have = k < end;
                if (have) { nk_next = __ldg(p.nbr_k + k); r_next = __ldg(rest + k); }
                const float kk = __int_as_float(nk.y);
                if (kk >= 0.0f) {  // exp(Y) > Y_min guard (SMW:75), resolved at set_spring_Y time
                    const float3 xj = xyz(sx[nk.x]), vj = xyz(sv[nk.x]);
                    const float3 dis = xj - xi;
                    if (kPrecise) {
                        const float dis_len = len3(dis);
                        const float3 d = dis / fmaxf(dis_len, 1e-6f);
                        const float3 spring_force = d * (kk * (dis_len / r - 1.0f));
                        const float v_rel = dot3(vj - vi, d);
                        const float3 dashpot = d * (p.dashpot * v_rel);
                        acc = acc + (spring_force + dashpot);
                    } else {
                        const float len2 = dot3(dis, dis);
                        const float c2 = fmaxf(len2, 1e-12f);         // max(len, 1e-6)^2
                        float y = rsqrtf(c2);
                        y = y * (1.5f - 0.5f * c2 * y * y);           // Newton step: 1/max(len,1e-6) to ~1 ulp
                        const float dis_len = len2 >= 1e-12f ? len2 * y : sqrtf(len2);
                        // F = (k (len / rest - 1) + damp (dv . d)) d with d = dis * y, factored so that the unit vector is
                        // never formed: one scalar coefficient on dis (9 operations instead of 19)
                        const float s_coef = kk * fmaf(dis_len, r, -1.0f);               // r = 1/rest
                        const float v_rel = dot3(vj - vi, dis) * y;
                        const float coef = fmaf(p.dashpot, v_rel, s_coef) * y;
                        acc.x = fmaf(coef, dis.x, acc.x); acc.y = fmaf(coef, dis.y, acc.y); acc.z = fmaf(coef, dis.z, acc.z);
                    }
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(gmask, acc.x, o);
                acc.y += __shfl_xor_sync(gmask, acc.y, o);
                acc.z += __shfl_xor_sync(gmask, acc.z, o);
            }
            if (lane_g == 0) {  // SMW:107-129
                const float m0 = p.mass[i];
                const float3 g = f3(0.0f * m0, 0.0f * m0, -9.8f * m0) * rf;
                const float3 all_force = acc + g;
                const float3 a = p.unit_mass ? all_force : all_force / m0;   // x / 1.0f is x, bit for bit
                const float3 v1 = vi + a * dt;
                const float3 v2 = v1 * drag;
                svb[i] = v2.x; svb[N + i] = v2.y; svb[2 * N + i] = v2.z;
            }
        }
        __syncthreads();

        // ============ phase B: self collision (SMW:132-193, 230-268): svb -> sv
        if (run_B) {
            const float ce = clampf(p.cs_elas, 0.0f, 1.0f);
            const float cf = clampf(p.cs_fric, 0.0f, 2.0f);
            const int* cnum = p.coll_num + (size_t)e * N;
            const int* cidx = p.coll_idx + (size_t)e * N * p.coll_cap;
            for (int i = tid; i < N; i += nthreads) {
                const float3 x1 = xyz(sx[i]);
                const float3 v1 = f3(svb[i], svb[N + i], svb[2 * N + i]);
                const float m1 = p.mass[i];
                const int mask1 = p.mask[i];
                float valid = 0.0f;
                float3 J_sum = f3(0.f, 0.f, 0.f);
                const int cnt = cnum[i];
                for (int k = 0; k < cnt; ++k) {
                    const int j = cidx[(size_t)i * p.coll_cap + k];
                    const float3 x2 = xyz(sx[j]);
                    const float3 v2 = f3(svb[j], svb[N + j], svb[2 * N + j]);
                    const float m2 = p.mass[j];
                    const float3 dis = x2 - x1;
                    const float dis_len = len3(dis);
                    const float3 rel = v2 - v1;
                    if (mask1 != p.mask[j] && dis_len < p.coll_dist && dot3(dis, rel) < -1e-4f) {
                        valid += 1.0f;
                        const float3 n = dis / fmaxf(dis_len, 1e-6f);
                        const float3 v_rel_n = n * dot3(rel, n);
                        const float inv_m = 1.0f / m1 + 1.0f / m2;
                        const float3 impulse_n = (v_rel_n * (-(1.0f + ce))) / inv_m;
                        const float v_rel_n_len = len3(v_rel_n);
                        const float3 v_rel_t = rel - v_rel_n;
                        const float v_rel_t_len = fmaxf(len3(v_rel_t), 1e-6f);
                        const float a = fmaxf(0.0f, 1.0f - cf * (1.0f + ce) * v_rel_n_len / v_rel_t_len);
                        const float3 impulse_t = (v_rel_t * (a - 1.0f)) / inv_m;
                        J_sum = J_sum + (impulse_n + impulse_t);
                    }
                }
                float3 vout = v1;
                if (valid > 0.0f) {
                    const float3 J_avg = J_sum / valid;
                    vout = v1 - J_avg / m1;
                }
                sv[i] = make_float4(vout.x, vout.y, vout.z, 0.0f);
            }
            __syncthreads();
        }

        // ============ phase C: mesh collision (SMW:295-421) + ground (SMW:424-474)
        MeshView mesh;
        float3 box_lo = f3(0, 0, 0), box_hi = f3(0, 0, 0);
        float3 c0 = f3(0, 0, 0), omega = f3(0, 0, 0), dv0 = f3(0, 0, 0), dv1 = f3(0, 0, 0);
        if (has_mesh) {
            mesh.dyn = p.stage_dyn ? s_dyn : g_interp + (size_t)step * p.n_dyn * 3;
            mesh.stat = p.stat_verts; mesh.faces = p.faces; mesh.n_dyn = p.n_dyn; mesh.F = p.F;
            mesh.grp = s_grp; mesh.boxes = s_aabb; mesh.n_grp = p.n_grp;
            // merged box of dynamic + static vertices, grown by the largest contact margin (+ slack).
            // A query can only change a particle that is inside the mesh (then it is inside the box) or
            // closer than `margin` (5 mm fingers, 1 mm otherwise) to its surface: a hit between margin and
            // max_dist = 20 mm has err >= 0 and leaves v unchanged and x advanced, exactly like no hit
            // (SMW:349-351, 415-421), so those particles need no query at all.
            // The test is made per part of the mesh (each finger / the tool / the static vertices has its own
            // box): between two open fingers nothing is queued until a finger comes within the margin.
            box_lo = f3(3e38f, 3e38f, 3e38f);
            box_hi = f3(-3e38f, -3e38f, -3e38f);
            for (int k = 0; k <= kMaxParts; ++k) {
                box_lo = f3(fminf(box_lo.x, s_aabb[6 * k]), fminf(box_lo.y, s_aabb[6 * k + 1]), fminf(box_lo.z, s_aabb[6 * k + 2]));
                box_hi = f3(fmaxf(box_hi.x, s_aabb[6 * k + 3]), fmaxf(box_hi.y, s_aabb[6 * k + 4]), fmaxf(box_hi.z, s_aabb[6 * k + 5]));
            }
            box_lo = box_lo - f3(kBoxGrow, kBoxGrow, kBoxGrow);
            box_hi = box_hi + f3(kBoxGrow, kBoxGrow, kBoxGrow);
            c0 = f3(g_center[3 * step], g_center[3 * step + 1], g_center[3 * step + 2]);
            omega = f3(g_omega[0], g_omega[1], g_omega[2]);
            dv0 = f3(g_dynvel[0], g_dynvel[1], g_dynvel[2]);
            dv1 = f3(g_dynvel[3], g_dynvel[4], g_dynvel[5]);
        }
        // ground bounce + position update of one particle (SMW:424-474)
        auto integrate_ground = [&](int i, float3 x0, float3 v0) {
            const float x_z = x0.z, v_z = v0.z;
            const float next_x_z = (x_z + v_z * dt) * rf;
            float3 v1; float toi;
            if (next_x_z < 0.0f && v_z * rf < -1e-4f) {
                const float3 normal = f3(0.0f, 0.0f, 1.0f) * rf;
                const float3 v_normal = normal * dot3(v0, normal);
                const float3 v_tao = v0 - v_normal;
                const float v_normal_len = len3(v_normal);
                const float v_tao_len = fmaxf(len3(v_tao), 1e-6f);
                const float ce = clampf(p.c_elas, 0.0f, 1.0f);
                const float cf = clampf(p.c_fric, 0.0f, 2.0f);
                const float3 v_normal_new = v_normal * (-ce);
                const float a = fmaxf(0.0f, 1.0f - cf * (1.0f + ce) * v_normal_len / v_tao_len);
                v1 = v_normal_new + v_tao * a;
                toi = -(x_z - 0.0f) / v_z;
            } else {
                v1 = v0; toi = 0.0f;
            }
            const float3 xn = x0 + v0 * toi + v1 * (dt - toi);
            sx[i] = make_float4(xn.x, xn.y, xn.z, 0.0f);
            sv[i] = make_float4(v1.x, v1.y, v1.z, 0.0f);
        };
        // C1: one thread per particle.  Particles whose advanced position lies outside the mesh
        // bounding box (grown by max_dist) cannot hit: they finish here.  The rest are queued.
        for (int i = tid; i < N; i += nthreads) {
            const float3 x0 = xyz(sx[i]);
            const float3 v0 = run_B ? xyz(sv[i]) : f3(svb[i], svb[N + i], svb[2 * N + i]);
            if (has_mesh) {
                const float3 next_x = x0 + v0 * dt;
                bool near_box = next_x.x >= box_lo.x && next_x.x <= box_hi.x && next_x.y >= box_lo.y &&
                                next_x.y <= box_hi.y && next_x.z >= box_lo.z && next_x.z <= box_hi.z;
                if (near_box) {   // inside the union box: test the parts
                    near_box = false;
                    for (int k = 0; k <= kMaxParts; ++k) {
                        const float* b = s_aabb + 6 * k;
                        near_box = near_box || (next_x.x >= b[0] - kBoxGrow && next_x.x <= b[3] + kBoxGrow &&
                                                next_x.y >= b[1] - kBoxGrow && next_x.y <= b[4] + kBoxGrow &&
                                                next_x.z >= b[2] - kBoxGrow && next_x.z <= b[5] + kBoxGrow);
                    }
                }
                if (near_box) {
                    s_cand[atomicAdd(s_ncand, 1)] = i;
                    if (!run_B) sv[i] = make_float4(v0.x, v0.y, v0.z, 0.0f);  // C2 reads v_before_ground from sv
                    continue;
                }
                integrate_ground(i, next_x, v0);  // mesh_collision with no hit still advances x (SMW:417-421)
            } else {
                integrate_ground(i, x0, v0);
            }
        }
        // C2: one warp per queued particle (mesh_collision, SMW:295-421, then the ground step)
        if (has_mesh) {
            __syncthreads();
            const int n_cand = *s_ncand;
            const int lane = tid & 31;
            for (int c = tid >> 5; c < n_cand; c += nthreads >> 5) {
                const int i = s_cand[c];
                const float3 x0 = xyz(sx[i]);
                float3 v0 = xyz(sv[i]);
                float3 next_x = x0 + v0 * dt;
                float3 next_v = v0;
                // closest point on the mesh within max_dist of `q` and the inside/outside sign there
                auto query = [&](float3 q, int& qface, float3& qpc, float& qsign) -> bool {
                    if (!kAccel) {
                        float qu, qv;
                        if (!mesh_query_warp(mesh, q, 0.02f, 0.6f, p.sign_mode, qface, qu, qv, qsign)) return false;
                        qpc = mesh_eval(mesh, qface, qu, qv);
                        return true;
                    }
                    const float* R = s_rigid;  // world = R rest + t  =>  rest = R^T (world - t)
                    const float3 w = f3(q.x - R[9], q.y - R[10], q.z - R[11]);
                    const float3 qr = f3(R[0] * w.x + R[3] * w.y + R[6] * w.z, R[1] * w.x + R[4] * w.y + R[7] * w.z,
                                         R[2] * w.x + R[5] * w.y + R[8] * w.z);
                    float3 cr;
                    const GridView gv = {p.frec, p.vnorm, p.cell_start, p.cell_tris, p.faces, p.cell_skip, p.gx0, p.gy0, p.gz0, p.gh,
                                         p.gnx, p.gny, p.gnz, p.sign_mode};
                    const bool hit_t = mesh_query_grid(gv, qr, 0.02f, s_triq + (tid >> 5) * kTriQueue, qface, cr, qsign);
                    if (hit_t)
                        qpc = f3(R[0] * cr.x + R[1] * cr.y + R[2] * cr.z + R[9], R[3] * cr.x + R[4] * cr.y + R[5] * cr.z + R[10],
                                 R[6] * cr.x + R[7] * cr.y + R[8] * cr.z + R[11]);
                    if (p.F_dyn == p.F) return hit_t;
                    // static obstacles beside the tool (the reference keeps both in one BVH, SMW:652-676): their faces are
                    // scanned by the warp (group kMaxParts of the brute-force query: box pruning, exact winding number over
                    // the static faces).  The merged query is the nearer of the two hits (ties to the lower face index, i.e.
                    // the tool), and the merged winding number is the sum of the two meshes': inside either one is inside.
                    MeshView sm = mesh;
                    sm.grp = s_grp + 3 * kMaxParts; sm.boxes = s_aabb + 6 * kMaxParts; sm.n_grp = 1;
                    int sface; float su, sv, ssign;
                    const bool hit_s = mesh_query_warp(sm, q, 0.02f, 0.6f, p.sign_mode, sface, su, sv, ssign);
                    if (!hit_t && !hit_s) return false;
                    const bool inside = (hit_t && qsign < 0.0f) || (hit_s && ssign < 0.0f);
                    if (hit_s) {
                        const float3 spc = mesh_eval(sm, sface, su, sv);
                        const float3 ds = spc - q, dt_ = qpc - q;
                        if (!hit_t || dot3(ds, ds) < dot3(dt_, dt_)) { qface = sface; qpc = spc; }
                    }
                    qsign = inside ? -1.0f : 1.0f;
                    return true;
                };
                int face; float sign; float3 pc;
                if (query(next_x, face, pc, sign)) {
                    int is_gripper;
                    const int mm = p.mesh_map[face];
                    if (!p.use_pusher) is_gripper = (mm == 0) ? 1 : ((mm == 1) ? 2 : 0);
                    else is_gripper = (mm >= 0) ? 1 : 0;
                    const float3 delta = next_x - pc;
                    const float dist = len3(delta) * sign;
                    const float margin = (is_gripper >= 1 && !p.use_pusher) ? 0.005f : 0.001f;
                    const float err = dist - margin;
                    if (err < 0.0f) {  // warp-uniform: every lane holds the same query result
                        const float3 normal = normalize3(delta) * sign;
                        float3 real_dyn = f3(0.f, 0.f, 0.f);
                        float ce, cf;
                        if (is_gripper >= 1) {
                            const float3 rot = cross3(omega, x0 - c0);
                            real_dyn = (is_gripper == 1 ? dv0 : dv1) + rot;
                            v0 = v0 - real_dyn;
                            ce = clampf(p.ce_elas, 0.0f, 1.0f);
                            cf = clampf(p.ce_fric, 0.0f, 2.0f);
                        } else {
                            ce = clampf(p.c_elas, 0.0f, 1.0f);
                            cf = clampf(p.c_fric, 0.0f, 2.0f);
                        }
                        const float3 v_normal = normal * dot3(v0, normal);
                        const float3 v_tao = v0 - v_normal;
                        const float v_normal_len = len3(v_normal);
                        const float v_tao_len = fmaxf(len3(v_tao), 1e-6f);
                        const float3 v_normal_new = v_normal * (-ce);
                        const float a = fmaxf(0.0f, 1.0f - cf * (1.0f + ce) * v_normal_len / v_tao_len);
                        const float3 v_tao_new = v_tao * a;
                        next_v = v_normal_new + v_tao_new;
                        if (is_gripper >= 1) next_v = next_v + real_dyn;
                        // SMW:397 re-assigns `query`, so the force of a finger / tool contact is booked on the face
                        // of the RE-QUERY (SMW:414); a missed re-query leaves Warp's default-constructed result, face 0
                        int force_face = face;
                        if (is_gripper >= 1) {
                            next_x = x0 + next_v * dt;
                            int face2; float sign2; float3 p2;
                            force_face = 0;
                            if (query(next_x, face2, p2, sign2)) {
                                force_face = face2;
                                const float3 delta2 = next_x - p2;
                                const float err2 = len3(delta2) * sign2 - margin;
                                if (err2 < 0.0f) {
                                    const float3 n2 = normalize3(delta2) * sign2;
                                    next_x = next_x - n2 * err2;
                                }
                            }
                        } else {
                            next_x = next_x - normal * err;
                        }
                        if (lane == 0 && step == p.n_sub - 1) {
                            const float3 dvn = (v_normal_new - v_normal) / dt;
                            const int fm = p.face_map[force_face];
                            atomicAdd(forces + 3 * fm, dvn.x);
                            atomicAdd(forces + 3 * fm + 1, dvn.y);
                            atomicAdd(forces + 3 * fm + 2, dvn.z);
                        }
                    }
                }
                __syncwarp();   // every lane has read sx[i] / sv[i] (top of this iteration) before lane 0 overwrites them
                if (lane == 0) integrate_ground(i, next_x, next_v);
            }
        }
        __syncthreads();
    }

    // ---- write back
    if (kSmemState) {
        float4* gx = p.x4 + (size_t)e * N;
        float4* gv = p.v4 + (size_t)e * N;
        for (int i = tid; i < N; i += nthreads) { gx[i] = sx[i]; gv[i] = sv[i]; }
    }
    if (has_mesh && p.smem_forces)
        for (int k = tid; k < 3 * p.F; k += nthreads) g_forces[k] = s_forces[k];
}

// ------------------------------------------------------------------ hash grid (Warp HashGrid restated)
#define R2S_GRID_DIM 128
__device__ __forceinline__ int grid_cell(int x, int y, int z)
{
    const int origin = 1 << 20;
    x += origin; y += origin; z += origin;
    x = max(0, x); y = max(0, y); z = max(0, z);
    return (z % R2S_GRID_DIM) * (R2S_GRID_DIM * R2S_GRID_DIM) + (y % R2S_GRID_DIM) * R2S_GRID_DIM +
           (x % R2S_GRID_DIM);
}

struct GridParams {
    int E, N, npow2, cap, words;
    float radius, coll_dist;
    const float4* x4;
    const int* mask;
    unsigned* resting;              // [E or 1][N][words]
    long long resting_stride;       // N*words or 0
    int* coll_num;                  // [E][N]
    int* coll_idx;                  // [E][N][cap]
    int* status;                    // [E][4]
    unsigned long long* key_scratch;  // [E][npow2] when keys do not fit shared memory, else null
    int stage_x;                      // positions staged in shared memory in cell-sorted order (after the keys)
};

// One CTA per environment: cell-sort the particles (bitonic on (cell, id) keys),
// then run the Warp query loop per particle.  kResting: build_resting_collision_pairs
// (SMW:272-291); else update_potential_collision (SMW:196-227).
template <bool kResting>
__global__ void __launch_bounds__(1024, 1) grid_kernel(const GridParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, N = p.N;
    unsigned long long* keys = p.key_scratch ? p.key_scratch + (size_t)e * p.npow2
                                             : reinterpret_cast<unsigned long long*>(smem_raw);
    const float4* x4 = p.x4 + (size_t)e * N;
    const float inv_w = 1.0f / p.radius;
    for (int i = tid; i < p.npow2; i += nt) {
        unsigned long long k = ~0ull;
        if (i < N) {
            const float4 q = x4[i];
            const int c = grid_cell((int)(q.x * inv_w), (int)(q.y * inv_w), (int)(q.z * inv_w));
            k = ((unsigned long long)(unsigned)c << 32) | (unsigned)i;
        }
        keys[i] = k;
    }
    if (!kResting && tid == 0) { p.status[4 * e] = 0; p.status[4 * e + 1] = 0; }
    __syncthreads();
    // ---- cell sort.  Fast path: when the occupied cells span a box of at most N cells (<= 128 per axis), a
    // counting sort on the box-local (z, y, x) cell rank gives exactly the order of the Warp
    // cell id (ascending id inside a cell): histogram, scan, unordered scatter, then each entry finds its rank
    // inside its cell by counting the smaller ids there -- 6 barriers instead of the bitonic network's
    // log2(n)(log2(n)+1)/2.  Otherwise (huge or wrapping extents): bitonic sort of the (cell, id) keys.
    __shared__ int s_lo[3], s_hi[3], s_c0[3], s_fast;
    if (tid < 3) { s_lo[tid] = 0x7fffffff; s_hi[tid] = -0x7fffffff; }
    __syncthreads();
    {
        int lo3[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi3[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
        for (int i = tid; i < N; i += nt) {
            const float4 q = x4[i];
            const int c3[3] = {(int)(q.x * inv_w), (int)(q.y * inv_w), (int)(q.z * inv_w)};
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo3[a] = min(lo3[a], c3[a]); hi3[a] = max(hi3[a], c3[a]); }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            for (int o = 16; o > 0; o >>= 1) {
                lo3[a] = min(lo3[a], __shfl_xor_sync(0xffffffffu, lo3[a], o));
                hi3[a] = max(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
            }
            if ((tid & 31) == 0) { atomicMin(&s_lo[a], lo3[a]); atomicMax(&s_hi[a], hi3[a]); }
        }
    }
    __syncthreads();
    if (tid == 0) {
        const long long ex = (long long)s_hi[0] - s_lo[0] + 1, ey = (long long)s_hi[1] - s_lo[1] + 1,
                        ez = (long long)s_hi[2] - s_lo[2] + 1;
        const int origin = 1 << 20;
        bool ok = p.stage_x && !p.key_scratch && ex * ey * ez <= (long long)N && ex <= R2S_GRID_DIM &&
                  ey <= R2S_GRID_DIM && ez <= R2S_GRID_DIM;
        // Per axis the hashed coordinate (c + 2^20) % 128 orders the box's cells; it wraps at most once inside a
        // box of <= 128 cells, at c0 = the first coordinate whose hash is 0: cells c >= c0 come first.
        // s_c0[a] = c0 when lo < c0 <= hi, else lo (no wrap).  Shifted coordinates must be non-negative
        // (grid_cell clamps at 0).
        for (int a = 0; a < 3 && ok; ++a) {
            ok = s_lo[a] + origin >= 0;
            const int c0 = (s_lo[a] + origin + R2S_GRID_DIM - 1) / R2S_GRID_DIM * R2S_GRID_DIM - origin;
            s_c0[a] = (c0 > s_lo[a] && c0 <= s_hi[a]) ? c0 : s_lo[a];
        }
        s_fast = ok ? 1 : 0;
    }
    __syncthreads();
    if (s_fast) {
        // scratch in the region that holds the cell-sorted positions afterwards (16 B per particle):
        // cnt[nc] | start[nc] | tmp_id[N] | cidx[N]  with nc <= N
        const int ex = s_hi[0] - s_lo[0] + 1, ey = s_hi[1] - s_lo[1] + 1, ez = s_hi[2] - s_lo[2] + 1;
        const int nc = ex * ey * ez;
        int* cnt = reinterpret_cast<int*>(smem_raw + sizeof(unsigned long long) * (size_t)p.npow2);
        int* start = cnt + nc;
        int* tmp_id = start + nc;
        int* cidx = tmp_id + N;
        __shared__ int s_wsum[32];
        for (int c = tid; c < nc; c += nt) cnt[c] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += nt) {
            const float4 q = x4[i];
            // rank of a coordinate in hash order: the part at or above the wrap point first
            auto rank = [&](int c, int a) { return c >= s_c0[a] ? c - s_c0[a] : (s_hi[a] - s_c0[a] + 1) + (c - s_lo[a]); };
            const int c = (rank((int)(q.z * inv_w), 2) * ey + rank((int)(q.y * inv_w), 1)) * ex + rank((int)(q.x * inv_w), 0);
            cidx[i] = c;
            atomicAdd(cnt + c, 1);
        }
        __syncthreads();
        // exclusive scan of cnt -> start (block-wide, chunks of nt)
        int carry = 0;
        for (int c0 = 0; c0 < nc; c0 += nt) {
            const int c = c0 + tid;
            const int v = c < nc ? cnt[c] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if ((tid & 31) >= o) incl += u;
            }
            if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
            __syncthreads();
            int base = carry;
            for (int w = 0; w < (tid >> 5); ++w) base += s_wsum[w];
            if (c < nc) { start[c] = base + incl - v; cnt[c] = 0; }
            int tot = 0;
            for (int w = 0; w < (nt >> 5); ++w) tot += s_wsum[w];
            carry += tot;
            __syncthreads();
        }
        for (int i = tid; i < N; i += nt) {
            const int c = cidx[i];
            tmp_id[start[c] + atomicAdd(cnt + c, 1)] = i;
        }
        __syncthreads();
        for (int t = tid; t < N; t += nt) {
            const int i = tmp_id[t], c = cidx[i];
            const int s0 = start[c], s1 = s0 + cnt[c];
            int rank = 0;
            for (int u = s0; u < s1; ++u) rank += tmp_id[u] < i;
            const float4 q = x4[i];
            const unsigned cell = (unsigned)grid_cell((int)(q.x * inv_w), (int)(q.y * inv_w), (int)(q.z * inv_w));
            keys[s0 + rank] = ((unsigned long long)cell << 32) | (unsigned)i;
        }
        __syncthreads();
    } else {
    for (int k = 2; k <= p.npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < p.npow2; i += nt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    unsigned* resting = p.resting + (size_t)e * p.resting_stride;
    // positions in cell-sorted order next to the keys, so that walking a cell reads consecutive shared memory
    float4* sxs = reinterpret_cast<float4*>(smem_raw + sizeof(unsigned long long) * (size_t)p.npow2);
    if (p.stage_x) {
        for (int t = tid; t < N; t += nt) sxs[t] = x4[(int)(keys[t] & 0xffffffffu)];
        __syncthreads();
    }
    int total = 0, overflow = 0;
    float d2_lim = p.coll_dist * p.coll_dist;
    while (sqrtf(d2_lim) >= p.coll_dist && d2_lim > 0.0f) d2_lim = __uint_as_float(__float_as_uint(d2_lim) - 1u);
    while (sqrtf(d2_lim) < p.coll_dist) d2_lim = __uint_as_float(__float_as_uint(d2_lim) + 1u);
    for (int t = tid; t < N; t += nt) {
        const int i = (int)(keys[t] & 0xffffffffu);  // wp.hash_grid_point_id: cell-sorted order
        const float4 q = p.stage_x ? sxs[t] : x4[i];
        const float3 x1 = xyz(q);
        // Warp queries the cells overlapping x +- 5*dist (the build radius) and update_potential_collision keeps
        // a candidate only if |dx| < dist (SMW:216-225).  The cell map is monotone per axis, so every candidate
        // that can pass lies in the cells overlapping x +- dist, and visiting only those (same x/y/z order)
        // yields the same row in the same order; the range is widened by 1e-4 relative + 1e-6 m so that float
        // rounding of the distance test cannot admit a point from outside it.  build_resting_collision_pairs
        // does not filter by distance and keeps the full range.
        const float r = kResting ? p.radius : p.coll_dist * 1.0001f + 1e-6f;
        const int xs = (int)((x1.x - r) * inv_w), ys = (int)((x1.y - r) * inv_w), zs = (int)((x1.z - r) * inv_w);
        const int xe = min((int)((x1.x + r) * inv_w), xs + R2S_GRID_DIM - 1);
        const int ye = min((int)((x1.y + r) * inv_w), ys + R2S_GRID_DIM - 1);
        const int ze = min((int)((x1.z + r) * inv_w), zs + R2S_GRID_DIM - 1);
        const int mask1 = p.mask[i];
        int cnt = 0;
        int* row = kResting ? nullptr : p.coll_idx + ((size_t)e * N + i) * p.cap;
        // cells are visited x fastest, then y, then z (Warp's hash_grid_query order); the cells of one x-run
        // have consecutive ids unless the run wraps modulo 128, so one binary search serves the whole run
        const bool x_wraps = ((xs + (1 << 20)) % R2S_GRID_DIM) > ((xe + (1 << 20)) % R2S_GRID_DIM) || xs + (1 << 20) < 0;
        for (int z = zs; z <= ze; ++z)
            for (int y = ys; y <= ye; ++y)
                for (int xx = xs; xx <= xe; ++xx) {
                    const unsigned cell = (unsigned)grid_cell(xx, y, z);
                    const unsigned cell_hi = x_wraps ? cell : (unsigned)grid_cell(xe, y, z);
                    const unsigned long long lo_key = (unsigned long long)cell << 32;
                    int lo = 0, hi = N;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (keys[mid] < lo_key) lo = mid + 1; else hi = mid;
                    }
                    for (; lo < N && (unsigned)(keys[lo] >> 32) <= cell_hi; ++lo) {
                        const int j = (int)(keys[lo] & 0xffffffffu);
                        if (kResting) {
                            if (j < i) {
                                atomicOr(resting + (size_t)i * p.words + (j >> 5), 1u << (j & 31));
                                atomicOr(resting + (size_t)j * p.words + (i >> 5), 1u << (i & 31));
                            }
                        } else {
                            // the three conditions of SMW:216-225 are a pure conjunction: the (cheap, usually
                            // failing) distance test goes first, the resting-pair bits are fetched only for close pairs
                            if (j == i) continue;
                            const float3 dis = xyz(p.stage_x ? sxs[lo] : x4[j]) - x1;
                            // len(dis) < dist (SMW:219) without the square root: sqrtf is monotone and correctly
                            // rounded, so sqrtf(d2) < dist  <=>  d2 < d2_lim (the least float whose root is >= dist)
                            if (!(dot3(dis, dis) < d2_lim) || mask1 == p.mask[j]) continue;
                            // resting[i][j] OR resting[j][i] (SMW:216-218): the relation is only ever written
                            // symmetrically (both bits per pair, above), so row i -- one 4*words-byte row per
                            // particle, cache-resident across its candidates -- decides
                            if ((resting[(size_t)i * p.words + (j >> 5)] >> (j & 31)) & 1u) continue;
                            if (cnt < p.cap) row[cnt++] = j;
                            else overflow++;
                        }
                    }
                    if (!x_wraps) break;  // the whole x-run was walked
                }
        if (!kResting) { p.coll_num[(size_t)e * N + i] = cnt; total += cnt; }
    }
    if (!kResting) {
        if (total) atomicAdd(p.status + 4 * e, total);
        if (overflow) atomicAdd(p.status + 4 * e + 1, overflow);
    }
}

// ------------------------------------------------------------------ small utility kernels
__global__ void pack_state_kernel(const float* __restrict__ src, long long stride_env, float4* __restrict__ dst,
                                  int E, int N)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)E * N) return;
    const int e = (int)(idx / N), i = (int)(idx % N);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src) {
        const float* s = src + (size_t)e * stride_env + 3 * (size_t)i;
        o = make_float4(s[0], s[1], s[2], 0.f);
    }
    dst[idx] = o;
}

__global__ void unpack_state_kernel(const float4* __restrict__ src, float* __restrict__ dst, long long n)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float4 q = src[idx];
    dst[3 * idx] = q.x; dst[3 * idx + 1] = q.y; dst[3 * idx + 2] = q.z;
}

// per directed entry: neighbour + clamped stiffness bits (negative = inactive), SMW:75,93
__global__ void stiffness_kernel(const float* __restrict__ logY, const int* __restrict__ sid,
                                 const int* __restrict__ nbr, float y_min, float y_max, int2* __restrict__ out,
                                 int nd)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nd) return;
    const float y = expf(logY[sid[k]]);
    const float kk = (y > y_min) ? fminf(fmaxf(y, y_min), y_max) : -1.0f;
    out[k] = make_int2(nbr[k], __float_as_int(kk));
}

__global__ void fill_kernel(float* p, float v, long long n)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) p[idx] = v;
}
__global__ void iota_kernel(int* p, int n)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) p[idx] = idx;
}

__global__ void rest_gather_kernel(const float* __restrict__ rest, long long rest_stride_env,
                                   const int* __restrict__ sid, float* __restrict__ out, int n_env, int nd,
                                   int reciprocal)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_env * nd) return;
    const int e = (int)(idx / nd), k = (int)(idx % nd);
    const float r = rest[(size_t)e * rest_stride_env + sid[k]];
    out[idx] = reciprocal ? 1.0f / r : r;
}

}  // namespace

// ====================================================================== host side
struct r2s_phys {
    r2s_phys_desc d;
    int nd = 0;  // directed adjacency entries (2S)
    int threads = 1024;
    int coll_cap = 64;
    // device buffers
    int* row_ptr = nullptr;
    int* nbr = nullptr;
    int* sid = nullptr;
    int2* nbr_k = nullptr;
    float* rest_csr = nullptr;
    int rest_envs = 1;
    float* logY = nullptr;
    float* mass = nullptr;
    int unit_mass = 1;          // every mass is exactly 1: the division by the mass is the identity
    int* mask = nullptr;
    float4* x4 = nullptr;
    float4* v4 = nullptr;
    float* vb_scratch = nullptr;
    int* coll_num = nullptr;
    int* coll_idx = nullptr;
    int* status = nullptr;
    unsigned* resting = nullptr;
    int words = 0;
    unsigned long long* key_scratch = nullptr;
    int npow2 = 0;
    // mesh
    int V = 0, F = 0, n_dyn = 0;
    float* stat_verts = nullptr;
    int* faces = nullptr;
    int* mesh_map = nullptr;
    int* face_map = nullptr;
    int* dyn_part = nullptr;
    int grp[15] = {0};
    int n_grp = 0;
    float* coll_forces = nullptr;
    float* interp_pts = nullptr;
    float* interp_center = nullptr;
    float* dyn_vel = nullptr;
    float* dyn_omega = nullptr;
    int motion_per_env = 0;
    int motion_substeps = 0;
    // rigid-tool accelerator
    int accel = 0;
    int F_dyn = 0;
    float* frec = nullptr;
    float* vnorm = nullptr;
    int* cell_start = nullptr;
    unsigned char* cell_skip = nullptr;
    int* cell_tris = nullptr;
    float grid0[3] = {0, 0, 0}, grid_h = 0.f;
    int grid_n[3] = {0, 0, 0};
    int anchors[3] = {0, 0, 0};
    float rest_frame[12] = {0};
    float rest_box[6] = {0};
    // launch config
    bool smem_state = true;
    size_t smem_bytes = 0;
    int stage_dyn = 0, smem_forces = 0;
    int max_smem_optin = 0;
};

namespace {

template <typename T>
int dmalloc(T** p, size_t n)
{
    *p = nullptr;
    if (n == 0) n = 1;
    R2S_CUDA_TRY(cudaMalloc((void**)p, n * sizeof(T)));
    return R2S_OK;
}

int configure_launch(r2s_phys* h)
{
    const int N = h->d.N;
    size_t misc = sizeof(float) * 48;   // bounding boxes + rigid transform
    h->stage_dyn = 0;
    h->smem_forces = 0;
    if (h->F > 0) {
        size_t dynb = sizeof(float) * 3 * ((h->n_dyn + 3) & ~3);
        size_t fb = sizeof(float) * 3 * h->F;
        misc += sizeof(int) * ((N + 3) & ~3);  // queue of particles near the mesh
        if (dynb <= 24 * 1024 && !h->accel) { h->stage_dyn = 1; misc += dynb; }
        if (fb <= 24 * 1024) { h->smem_forces = 1; misc += fb; }
        if (h->accel) misc += sizeof(int) * 32 * 256;   // per-warp triangle queues of mesh_query_grid (kTriQueue)
    }
    size_t state = sizeof(float4) * 2 * (size_t)N + sizeof(float) * 3 * ((N + 3) & ~3);
    h->smem_state = state + misc + 256 <= (size_t)h->max_smem_optin;
    h->smem_bytes = (h->smem_state ? state : 0) + misc + 64;
    if (!h->smem_state && !h->vb_scratch)
        if (int rc = dmalloc(&h->vb_scratch, (size_t)h->d.E * 3 * N)) return rc;
    return R2S_OK;
}

// Rigid-tool accelerator (host, once per set_mesh): face / edge / vertex pseudonormals, three anchor vertices
// that define the tool's frame, and a uniform grid of triangle lists over the rest pose grown by max_dist.
int build_accel(r2s_phys* h, const float* verts, const int* faces, int V, int F)
{
    auto P = [&](int i) { return verts + 3 * (size_t)i; };
    auto sub = [](const float* a, const float* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; };
    auto crs = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto nrm = [](double* a) { double l = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); if (l > 0) { a[0] /= l; a[1] /= l; a[2] /= l; } return l; };
    std::vector<double> fn(3 * (size_t)F), vn(3 * (size_t)V, 0.0);
    for (int f = 0; f < F; ++f) {
        const int* t = faces + 3 * f;
        double e1[3], e2[3], n[3];
        sub(P(t[1]), P(t[0]), e1); sub(P(t[2]), P(t[0]), e2);
        crs(e1, e2, n); nrm(n);
        for (int c = 0; c < 3; ++c) fn[3 * (size_t)f + c] = n[c];
        for (int k = 0; k < 3; ++k) {  // angle-weighted vertex pseudonormals
            double a[3], b[3];
            sub(P(t[(k + 1) % 3]), P(t[k]), a); sub(P(t[(k + 2) % 3]), P(t[k]), b);
            nrm(a); nrm(b);
            double cs = std::max(-1.0, std::min(1.0, a[0] * b[0] + a[1] * b[1] + a[2] * b[2]));
            const double ang = acos(cs);
            for (int c = 0; c < 3; ++c) vn[3 * (size_t)t[k] + c] += ang * n[c];
        }
    }
    // edge -> adjacent faces
    std::vector<std::pair<long long, int>> edges;
    edges.reserve(3 * (size_t)F);
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) {
            long long a = faces[3 * f + k], b = faces[3 * f + (k + 1) % 3];
            edges.push_back({std::min(a, b) * (long long)V + std::max(a, b), f});
        }
    std::sort(edges.begin(), edges.end());
    auto other_face = [&](int f, int a, int b) {
        const long long key = (long long)std::min(a, b) * V + std::max(a, b);
        auto it = std::lower_bound(edges.begin(), edges.end(), std::make_pair(key, -1));
        for (; it != edges.end() && it->first == key; ++it)
            if (it->second != f) return it->second;
        return f;  // boundary edge of an open mesh
    };
    std::vector<float> frec(24 * (size_t)F, 0.f), vnf(3 * (size_t)V);
    for (int i = 0; i < 3 * V; ++i) vnf[i] = (float)vn[i];
    for (int f = 0; f < F; ++f) {
        const int* t = faces + 3 * f;
        float* r = frec.data() + 24 * (size_t)f;
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) r[3 * k + c] = P(t[k])[c];
        for (int c = 0; c < 3; ++c) r[9 + c] = (float)fn[3 * (size_t)f + c];
        const int pairs[3][2] = {{0, 1}, {1, 2}, {2, 0}};  // e01, e12, e20
        for (int k = 0; k < 3; ++k) {
            const int g = other_face(f, t[pairs[k][0]], t[pairs[k][1]]);
            for (int c = 0; c < 3; ++c) r[12 + 3 * k + c] = (float)(fn[3 * (size_t)f + c] + fn[3 * (size_t)g + c]);
        }
    }
    // anchors: vertex 0, the vertex farthest from it, the vertex farthest from their line
    int a0 = 0, a1 = 0, a2 = 0;
    double best = -1;
    for (int i = 0; i < V; ++i) {
        double d[3]; sub(P(i), P(a0), d);
        const double l = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (l > best) { best = l; a1 = i; }
    }
    double ax[3]; sub(P(a1), P(a0), ax); nrm(ax);
    best = -1;
    for (int i = 0; i < V; ++i) {
        double d[3], c[3]; sub(P(i), P(a0), d); crs(ax, d, c);
        const double l = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
        if (l > best) { best = l; a2 = i; }
    }
    if (best <= 1e-12) { r2s::set_error("r2s_phys_set_mesh: degenerate (collinear) mesh cannot be accelerated"); return R2S_ERR_INVALID; }
    double d2[3], e3[3], e2[3];
    sub(P(a2), P(a0), d2); crs(ax, d2, e3); nrm(e3); crs(e3, ax, e2);
    for (int c = 0; c < 3; ++c) { h->rest_frame[c] = (float)ax[c]; h->rest_frame[3 + c] = (float)e2[c]; h->rest_frame[6 + c] = (float)e3[c];
                                  h->rest_frame[9 + c] = P(a0)[c]; }
    h->anchors[0] = a0; h->anchors[1] = a1; h->anchors[2] = a2;
    // grid over the rest box grown by max_dist + one cell
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = 0; i < V; ++i)
        for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], P(i)[c]); hi[c] = std::max(hi[c], P(i)[c]); }
    for (int c = 0; c < 3; ++c) { h->rest_box[c] = lo[c]; h->rest_box[3 + c] = hi[c]; }
    float cell = 0.004f;
    long long ncell = 0;
    for (;;) {
        ncell = 1;
        for (int c = 0; c < 3; ++c) {
            h->grid0[c] = lo[c] - (0.02f + cell);
            h->grid_n[c] = (int)ceilf((hi[c] - lo[c] + 2.f * (0.02f + cell)) / cell) + 1;
            ncell *= h->grid_n[c];
        }
        if (ncell <= (1ll << 22)) break;
        cell *= 1.5f;
    }
    h->grid_h = cell;
    std::vector<int> start(ncell + 1, 0);
    auto cell_range = [&](int f, int* c0, int* c1) {
        const int* t = faces + 3 * f;
        for (int c = 0; c < 3; ++c) {
            const float mn = std::min(P(t[0])[c], std::min(P(t[1])[c], P(t[2])[c]));
            const float mx = std::max(P(t[0])[c], std::max(P(t[1])[c], P(t[2])[c]));
            c0[c] = std::max(0, (int)floorf((mn - h->grid0[c]) / cell));
            c1[c] = std::min(h->grid_n[c] - 1, (int)floorf((mx - h->grid0[c]) / cell));
        }
    };
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<int> cur;
        std::vector<int> tris;
        if (pass == 1) {
            for (long long i = 0; i < ncell; ++i) start[i + 1] += start[i];
            cur.assign(start.begin(), start.end() - 1);
            tris.resize(start[ncell]);
        }
        for (int f = 0; f < F; ++f) {
            int c0[3], c1[3];
            cell_range(f, c0, c1);
            for (int z = c0[2]; z <= c1[2]; ++z)
                for (int y = c0[1]; y <= c1[1]; ++y)
                    for (int x = c0[0]; x <= c1[0]; ++x) {
                        const long long id = ((long long)z * h->grid_n[1] + y) * h->grid_n[0] + x;
                        if (pass == 0) start[id + 1]++;
                        else tris[cur[id]++] = f;
                    }
        }
        if (pass == 1) {
            if (dmalloc(&h->frec, frec.size()) || dmalloc(&h->vnorm, vnf.size()) || dmalloc(&h->cell_start, start.size()) ||
                dmalloc(&h->cell_tris, tris.size()))
                return R2S_ERR_CUDA;
            R2S_CUDA_TRY(cudaMemcpy(h->frec, frec.data(), sizeof(float) * frec.size(), cudaMemcpyHostToDevice));
            R2S_CUDA_TRY(cudaMemcpy(h->vnorm, vnf.data(), sizeof(float) * vnf.size(), cudaMemcpyHostToDevice));
            R2S_CUDA_TRY(cudaMemcpy(h->cell_start, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice));
            if (!tris.empty())
                R2S_CUDA_TRY(cudaMemcpy(h->cell_tris, tris.data(), sizeof(int) * tris.size(), cudaMemcpyHostToDevice));
            // Chebyshev distance (in cells, capped) from every cell to the nearest non-empty one: two 3-D chamfer
            // sweeps, exact for the chessboard metric.  A query starts its shell search at that radius.
            const int nx = h->grid_n[0], ny = h->grid_n[1], nz = h->grid_n[2];
            std::vector<unsigned char> skip((size_t)ncell);
            for (long long i = 0; i < ncell; ++i) skip[i] = start[i + 1] > start[i] ? 0 : 255;
            auto at = [&](int x, int y, int z) -> int {
                return (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) ? 255 : skip[((size_t)z * ny + y) * nx + x];
            };
            for (int sweep = 0; sweep < 2; ++sweep) {
                const int dir = sweep ? -1 : 1;
                for (int zz = 0; zz < nz; ++zz)
                    for (int yy = 0; yy < ny; ++yy)
                        for (int xx = 0; xx < nx; ++xx) {
                            const int x = sweep ? nx - 1 - xx : xx, y = sweep ? ny - 1 - yy : yy, z = sweep ? nz - 1 - zz : zz;
                            int best_d = at(x, y, z);
                            for (int dz = -1; dz <= 1; ++dz)      // the 13 neighbours already swept in this direction
                                for (int dy = -1; dy <= 1; ++dy)
                                    for (int dx = -1; dx <= 1; ++dx) {
                                        const int lin = (dz * 3 + dy) * 3 + dx;
                                        if (lin * dir >= 0) continue;
                                        best_d = std::min(best_d, at(x + dx, y + dy, z + dz) + 1);
                                    }
                            skip[((size_t)z * ny + y) * nx + x] = (unsigned char)std::min(best_d, 255);
                        }
            }
            if (dmalloc(&h->cell_skip, skip.size())) return R2S_ERR_CUDA;
            R2S_CUDA_TRY(cudaMemcpy(h->cell_skip, skip.data(), skip.size(), cudaMemcpyHostToDevice));
        }
    }
    h->accel = 1;
    return R2S_OK;
}

template <int G, bool kSmem, bool kPrecise, bool kAccel>
int launch_frame_t(r2s_phys* h, const FrameParams& fp, cudaStream_t st)
{
    R2S_CUDA_TRY(cudaFuncSetAttribute(frame_kernel<G, kSmem, kPrecise, kAccel>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    frame_kernel<G, kSmem, kPrecise, kAccel><<<h->d.E, h->threads, h->smem_bytes, st>>>(fp);
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}

template <int G, bool kAccel>
int launch_frame(r2s_phys* h, const FrameParams& fp, cudaStream_t st)
{
    if (h->smem_state)
        return h->d.precise ? launch_frame_t<G, true, true, kAccel>(h, fp, st) : launch_frame_t<G, true, false, kAccel>(h, fp, st);
    return h->d.precise ? launch_frame_t<G, false, true, kAccel>(h, fp, st) : launch_frame_t<G, false, false, kAccel>(h, fp, st);
}

}  // namespace

extern "C" {

r2s_phys* r2s_phys_create(const r2s_phys_desc* desc)
{
    if (!desc || desc->E <= 0 || desc->N <= 0 || desc->S < 0 || !desc->springs || !desc->rest_lengths) {
        r2s::set_error("r2s_phys_create: bad descriptor (E=%d N=%d S=%d)", desc ? desc->E : -1,
                       desc ? desc->N : -1, desc ? desc->S : -1);
        return nullptr;
    }
    r2s_phys* h = new r2s_phys();
    h->d = *desc;
    const int E = desc->E, N = desc->N, S = desc->S;
    h->nd = 2 * S;
    h->coll_cap = desc->coll_row_cap > 0 ? desc->coll_row_cap : 64;
    h->threads = desc->threads > 0 ? desc->threads : 1024;
    if (h->threads % 32 || h->threads > 1024) {
        r2s::set_error("r2s_phys_create: threads must be a multiple of 32, <= 1024");
        delete h;
        return nullptr;
    }
    int dev = 0;
    auto fail = [&](const char* what) -> r2s_phys* {
        if (what) r2s::set_error("r2s_phys_create: %s", what);
        r2s_phys_destroy(h);
        return nullptr;
    };
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return fail("cannot query device");

    // ---- adjacency (host): per particle, incident springs in ascending spring index
    std::vector<int> springs(2 * (size_t)S);
    if (S && cudaMemcpy(springs.data(), desc->springs, sizeof(int) * 2 * S, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail("cannot read springs (must be a device pointer)");
    std::vector<int> row_ptr(N + 1, 0), nbr(h->nd ? h->nd : 1), sid(h->nd ? h->nd : 1);
    for (int t = 0; t < S; ++t) {
        const int a = springs[2 * t], b = springs[2 * t + 1];
        if (a < 0 || a >= N || b < 0 || b >= N) return fail("spring index out of range");
        row_ptr[a + 1]++; row_ptr[b + 1]++;
    }
    for (int i = 0; i < N; ++i) row_ptr[i + 1] += row_ptr[i];
    {
        std::vector<int> cur(row_ptr.begin(), row_ptr.end() - 1);
        for (int t = 0; t < S; ++t) {
            const int a = springs[2 * t], b = springs[2 * t + 1];
            nbr[cur[a]] = b; sid[cur[a]++] = t;
            nbr[cur[b]] = a; sid[cur[b]++] = t;
        }
    }
    bool ok = true;
    ok &= dmalloc(&h->row_ptr, N + 1) == 0 && dmalloc(&h->nbr, h->nd) == 0 && dmalloc(&h->sid, h->nd) == 0 &&
          dmalloc(&h->nbr_k, h->nd) == 0 && dmalloc(&h->logY, S) == 0 && dmalloc(&h->mass, N) == 0 &&
          dmalloc(&h->mask, N) == 0 && dmalloc(&h->x4, (size_t)E * N) == 0 &&
          dmalloc(&h->v4, (size_t)E * N) == 0 && dmalloc(&h->status, (size_t)E * 4) == 0;
    if (!ok) return fail(nullptr);
    cudaMemcpy(h->row_ptr, row_ptr.data(), sizeof(int) * (N + 1), cudaMemcpyHostToDevice);
    if (h->nd) {
        cudaMemcpy(h->nbr, nbr.data(), sizeof(int) * h->nd, cudaMemcpyHostToDevice);
        cudaMemcpy(h->sid, sid.data(), sizeof(int) * h->nd, cudaMemcpyHostToDevice);
    }
    cudaMemset(h->status, 0, sizeof(int) * 4 * E);
    cudaMemset(h->x4, 0, sizeof(float4) * (size_t)E * N);
    cudaMemset(h->v4, 0, sizeof(float4) * (size_t)E * N);
    h->unit_mass = 1;
    if (desc->masses) {
        cudaMemcpy(h->mass, desc->masses, sizeof(float) * N, cudaMemcpyDeviceToDevice);
        std::vector<float> hm(N);
        cudaMemcpy(hm.data(), desc->masses, sizeof(float) * N, cudaMemcpyDeviceToHost);
        for (float m : hm) if (m != 1.0f) { h->unit_mass = 0; break; }   // PhysTwin masses are all 1 (PT:335)
    } else { fill_kernel<<<r2s::ceil_div(N, 256), 256>>>(h->mass, 1.0f, N); r2s::count_launch(); }
    if (desc->collision_mask) cudaMemcpy(h->mask, desc->collision_mask, sizeof(int) * N, cudaMemcpyDeviceToDevice);
    else { iota_kernel<<<r2s::ceil_div(N, 256), 256>>>(h->mask, N); r2s::count_launch(); }
    if (desc->log_spring_Y) {
        if (r2s_phys_set_spring_Y(h, desc->log_spring_Y, nullptr)) return fail(nullptr);
    } else if (S) {
        fill_kernel<<<r2s::ceil_div(S, 256), 256>>>(h->logY, logf(3e4f), S);
        r2s::count_launch();
        if (r2s_phys_set_spring_Y(h, h->logY, nullptr)) return fail(nullptr);
    }
    if (r2s_phys_set_rest_lengths(h, desc->rest_lengths, desc->rest_per_env, nullptr)) return fail(nullptr);

    if (desc->self_collision) {
        h->words = (N + 31) / 32;
        h->npow2 = 1;
        while (h->npow2 < N) h->npow2 <<= 1;
        ok = dmalloc(&h->coll_num, (size_t)E * N) == 0 && dmalloc(&h->coll_idx, (size_t)E * N * h->coll_cap) == 0 &&
             dmalloc(&h->resting, (size_t)E * N * h->words) == 0;
        if (!ok) return fail(nullptr);
        cudaMemset(h->coll_num, 0, sizeof(int) * (size_t)E * N);
        cudaMemset(h->resting, 0, sizeof(unsigned) * (size_t)E * N * h->words);
        if (sizeof(unsigned long long) * (size_t)h->npow2 > (size_t)h->max_smem_optin - 1024)
            if (dmalloc(&h->key_scratch, (size_t)E * h->npow2)) return fail(nullptr);
    }
    if (configure_launch(h)) return fail(nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("device error during setup");
    return h;
}

int r2s_phys_destroy(r2s_phys* h)
{
    if (!h) return R2S_OK;
    void* ptrs[] = {h->row_ptr, h->nbr, h->sid, h->nbr_k, h->rest_csr, h->logY, h->mass, h->mask, h->x4, h->v4,
                    h->vb_scratch, h->coll_num, h->coll_idx, h->status, h->resting, h->key_scratch,
                    h->stat_verts, h->faces, h->mesh_map, h->face_map, h->dyn_part, h->coll_forces, h->interp_pts,
                    h->interp_center, h->dyn_vel, h->dyn_omega, h->frec, h->vnorm, h->cell_start, h->cell_tris, h->cell_skip};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete h;
    return R2S_OK;
}

int r2s_phys_set_state(r2s_phys* h, const float* x, const float* v, int64_t stride, void* stream)
{
    R2S_REQUIRE(h && x, "r2s_phys_set_state: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)h->d.E * h->d.N;
    pack_state_kernel<<<r2s::ceil_div(n, 256), 256, 0, st>>>(x, stride, h->x4, h->d.E, h->d.N);
    R2S_LAUNCH_CHECK();
    pack_state_kernel<<<r2s::ceil_div(n, 256), 256, 0, st>>>(v, stride, h->v4, h->d.E, h->d.N);
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}

int r2s_phys_get_state(r2s_phys* h, float* x, float* v, void* stream)
{
    R2S_REQUIRE(h, "r2s_phys_get_state: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)h->d.E * h->d.N;
    if (x) { unpack_state_kernel<<<r2s::ceil_div(n, 256), 256, 0, st>>>(h->x4, x, n); R2S_LAUNCH_CHECK(); }
    if (v) { unpack_state_kernel<<<r2s::ceil_div(n, 256), 256, 0, st>>>(h->v4, v, n); R2S_LAUNCH_CHECK(); }
    return R2S_OK;
}

int r2s_phys_set_spring_Y(r2s_phys* h, const float* logY, void* stream)
{
    R2S_REQUIRE(h && logY, "r2s_phys_set_spring_Y: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (h->d.S == 0) return R2S_OK;
    if (logY != h->logY)
        R2S_CUDA_TRY(cudaMemcpyAsync(h->logY, logY, sizeof(float) * h->d.S, cudaMemcpyDeviceToDevice, st));
    stiffness_kernel<<<r2s::ceil_div(h->nd, 256), 256, 0, st>>>(h->logY, h->sid, h->nbr, h->d.spring_Y_min,
                                                                h->d.spring_Y_max, h->nbr_k, h->nd);
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}

int r2s_phys_set_rest_lengths(r2s_phys* h, const float* rest, int per_env, void* stream)
{
    R2S_REQUIRE(h && rest, "r2s_phys_set_rest_lengths: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_env = per_env ? h->d.E : 1;
    if (!h->rest_csr || h->rest_envs != n_env) {
        if (h->rest_csr) { cudaFree(h->rest_csr); h->rest_csr = nullptr; }
        if (int rc = dmalloc(&h->rest_csr, (size_t)n_env * (h->nd ? h->nd : 1))) return rc;
        h->rest_envs = n_env;
    }
    if (h->nd == 0) return R2S_OK;
    const long long n = (long long)n_env * h->nd;
    rest_gather_kernel<<<r2s::ceil_div(n, 256), 256, 0, st>>>(rest, h->d.S, h->sid, h->rest_csr, n_env, h->nd,
                                                                 h->d.precise ? 0 : 1);
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}

int r2s_phys_set_collide(r2s_phys* h, float elas, float fric, float eef_elas, float eef_fric, float self_elas,
                         float self_fric)
{
    R2S_REQUIRE(h, "r2s_phys_set_collide: null handle");
    if (elas == elas) h->d.collide_elas = elas;
    if (fric == fric) h->d.collide_fric = fric;
    if (eef_elas == eef_elas) h->d.collide_eef_elas = eef_elas;
    if (eef_fric == eef_fric) h->d.collide_eef_fric = eef_fric;
    if (self_elas == self_elas) h->d.collide_self_elas = self_elas;
    if (self_fric == self_fric) h->d.collide_self_fric = self_fric;
    return R2S_OK;
}

int r2s_phys_set_mesh(r2s_phys* h, const float* verts, const int32_t* faces, const int32_t* mesh_map,
                      const int32_t* face_map, int32_t V, int32_t F, int32_t n_dyn)
{
    R2S_REQUIRE(h && verts && faces && mesh_map && face_map, "r2s_phys_set_mesh: null argument");
    R2S_REQUIRE(V > 0 && F > 0 && n_dyn >= 0 && n_dyn <= V, "r2s_phys_set_mesh: bad sizes V=%d F=%d n_dyn=%d", V, F,
                n_dyn);
    for (int k = 0; k < 3 * F; ++k)
        R2S_REQUIRE(faces[k] >= 0 && faces[k] < V, "r2s_phys_set_mesh: face index out of range");
    for (int k = 0; k < F; ++k)
        R2S_REQUIRE(face_map[k] >= 0 && face_map[k] < F, "r2s_phys_set_mesh: face_map out of range");
    void* old[] = {h->stat_verts, h->faces, h->mesh_map, h->face_map, h->dyn_part, h->coll_forces, h->interp_pts,
                   h->interp_center, h->dyn_vel, h->dyn_omega};
    for (void* p : old)
        if (p) cudaFree(p);
    h->interp_pts = h->interp_center = h->dyn_vel = h->dyn_omega = nullptr;
    for (void* q : {(void*)h->frec, (void*)h->vnorm, (void*)h->cell_start, (void*)h->cell_tris, (void*)h->cell_skip})
        if (q) cudaFree(q);
    h->frec = h->vnorm = nullptr; h->cell_start = h->cell_tris = nullptr; h->cell_skip = nullptr; h->accel = 0;
    h->V = V; h->F = F; h->n_dyn = n_dyn;
    const int E = h->d.E, ns = h->d.n_substeps;
    if (dmalloc(&h->stat_verts, (size_t)3 * V) || dmalloc(&h->faces, (size_t)3 * F) || dmalloc(&h->mesh_map, F) ||
        dmalloc(&h->face_map, F) || dmalloc(&h->coll_forces, (size_t)E * F * 3))
        return R2S_ERR_CUDA;
    R2S_CUDA_TRY(cudaMemcpy(h->stat_verts, verts, sizeof(float) * 3 * V, cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemcpy(h->faces, faces, sizeof(int) * 3 * F, cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemcpy(h->mesh_map, mesh_map, sizeof(int) * F, cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemcpy(h->face_map, face_map, sizeof(int) * F, cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemset(h->coll_forces, 0, sizeof(float) * (size_t)E * F * 3));
    {   // one broad-phase box per dynamic mesh (finger / tool): part = rank of the mesh_map value of a face that
        // uses the vertex, among the distinct values on dynamic faces; more than four meshes share the last box
        std::vector<int> part(n_dyn > 0 ? n_dyn : 1, 0), ids;
        for (int f = 0; f < F; ++f)
            for (int c = 0; c < 3; ++c) {
                const int v = faces[3 * f + c];
                if (v >= n_dyn) continue;
                int k = 0;
                while (k < (int)ids.size() && ids[k] != mesh_map[f]) ++k;
                if (k == (int)ids.size()) ids.push_back(mesh_map[f]);
                part[v] = k < 4 ? k : 3;
            }
        h->dyn_part = nullptr;
        if (dmalloc(&h->dyn_part, part.size())) return R2S_ERR_CUDA;
        R2S_CUDA_TRY(cudaMemcpy(h->dyn_part, part.data(), sizeof(int) * part.size(), cudaMemcpyHostToDevice));
        // face groups for the brute-force query: group = part of the face's (dynamic) vertices, 4 = static faces.
        // Usable only if every face lies in one group and each group is one contiguous face range; a group is
        // closed when every directed edge a->b has exactly one partner b->a inside the group.
        bool ok = true;
        std::vector<int> fg(F, 0);
        for (int f = 0; f < F && ok; ++f) {
            const int* t = faces + 3 * f;
            const bool d0 = t[0] < n_dyn, d1 = t[1] < n_dyn, d2 = t[2] < n_dyn;
            if (d0 && d1 && d2) {
                fg[f] = part[t[0]];
                ok = part[t[1]] == fg[f] && part[t[2]] == fg[f];
            } else if (!d0 && !d1 && !d2) {
                fg[f] = 4;
            } else {
                ok = false;
            }
        }
        for (int k = 0; k < 15; ++k) h->grp[k] = 0;
        for (int g = 0; g < 5 && ok; ++g) {
            int first = -1, last = -1, count = 0;
            for (int f = 0; f < F; ++f)
                if (fg[f] == g) { if (first < 0) first = f; last = f; ++count; }
            if (count == 0) continue;
            ok = last - first + 1 == count;
            std::map<std::pair<int, int>, int> half;   // directed edge -> multiplicity
            for (int f = first; f <= last && ok; ++f)
                for (int c = 0; c < 3; ++c) half[{faces[3 * f + c], faces[3 * f + (c + 1) % 3]}]++;
            bool closed = true;
            for (const auto& kv : half) {
                const auto it = half.find({kv.first.second, kv.first.first});
                if (kv.second != 1 || it == half.end() || it->second != 1) { closed = false; break; }
            }
            h->grp[3 * g] = first; h->grp[3 * g + 1] = last + 1; h->grp[3 * g + 2] = closed ? 1 : 0;
        }
        h->n_grp = ok ? 5 : 0;
    }
    // SMW:699-711 defaults: rest pose repeated over the substeps, centre = mean, zero velocities
    std::vector<float> tbl((size_t)ns * n_dyn * 3 + 1), ctr((size_t)ns * 3 + 1), zero(6, 0.f);
    double c[3] = {0, 0, 0};
    for (int k = 0; k < n_dyn; ++k)
        for (int a = 0; a < 3; ++a) c[a] += verts[3 * k + a];
    for (int s = 0; s < ns; ++s) {
        if (n_dyn) memcpy(&tbl[(size_t)s * n_dyn * 3], verts, sizeof(float) * 3 * n_dyn);
        for (int a = 0; a < 3; ++a) ctr[3 * s + a] = n_dyn ? (float)(c[a] / n_dyn) : 0.f;
    }
    if (dmalloc(&h->interp_pts, tbl.size()) || dmalloc(&h->interp_center, ctr.size()) || dmalloc(&h->dyn_vel, 6) ||
        dmalloc(&h->dyn_omega, 3))
        return R2S_ERR_CUDA;
    R2S_CUDA_TRY(cudaMemcpy(h->interp_pts, tbl.data(), sizeof(float) * tbl.size(), cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemcpy(h->interp_center, ctr.data(), sizeof(float) * ctr.size(), cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemcpy(h->dyn_vel, zero.data(), sizeof(float) * 6, cudaMemcpyHostToDevice));
    R2S_CUDA_TRY(cudaMemcpy(h->dyn_omega, zero.data(), sizeof(float) * 3, cudaMemcpyHostToDevice));
    h->motion_per_env = 0;
    h->motion_substeps = ns;
    // a mesh that is entirely one rigidly moving tool (the pusher: every vertex dynamic) gets the grid accelerator
    // (static obstacles may follow it: their faces are scanned beside the grid query)
    const int want = h->d.mesh_accel;
    int F_dyn = F;
    bool tool_first = true;   // dynamic faces are the first F_dyn faces and use only dynamic vertices; the rest are static
    if (n_dyn < V) {
        tool_first = h->n_grp == 5 && h->grp[0] == 0 && h->grp[3 * 4 + 1] == F && h->grp[1] == h->grp[3 * 4] &&
                     h->grp[3 * 1 + 1] == 0 && h->grp[3 * 2 + 1] == 0 && h->grp[3 * 3 + 1] == 0;
        F_dyn = tool_first ? h->grp[1] : F;
    }
    h->F_dyn = F;
    if (want > 0 || (want == 0 && h->d.use_pusher && tool_first && n_dyn > 0 && F_dyn >= 256)) {
        R2S_REQUIRE(tool_first && n_dyn > 0, "r2s_phys_set_mesh: mesh_accel needs ONE rigid tool whose faces come first (static faces may follow)");
        if (int rc = build_accel(h, verts, faces, n_dyn, F_dyn)) return rc;
        h->F_dyn = F_dyn;
    }
    return configure_launch(h);
}

int r2s_phys_set_mesh_motion(r2s_phys* h, const float* interp_pts, const float* interp_center, const float* dyn_vel,
                             const float* dyn_omega, int per_env, void* stream)
{
    R2S_REQUIRE(h && h->F > 0, "r2s_phys_set_mesh_motion: no mesh set");
    R2S_REQUIRE(interp_pts && interp_center && dyn_vel && dyn_omega, "r2s_phys_set_mesh_motion: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int ne = per_env ? h->d.E : 1, ns = h->d.n_substeps;
    if (h->motion_per_env != (per_env ? 1 : 0)) {
        cudaFree(h->interp_pts); cudaFree(h->interp_center); cudaFree(h->dyn_vel); cudaFree(h->dyn_omega);
        h->interp_pts = h->interp_center = h->dyn_vel = h->dyn_omega = nullptr;
        if (dmalloc(&h->interp_pts, (size_t)ne * ns * h->n_dyn * 3) || dmalloc(&h->interp_center, (size_t)ne * ns * 3) ||
            dmalloc(&h->dyn_vel, (size_t)ne * 6) || dmalloc(&h->dyn_omega, (size_t)ne * 3))
            return R2S_ERR_CUDA;
        h->motion_per_env = per_env ? 1 : 0;
    }
    const int nv = h->d.use_pusher ? 1 : 2;
    R2S_CUDA_TRY(cudaMemcpyAsync(h->interp_pts, interp_pts, sizeof(float) * (size_t)ne * ns * h->n_dyn * 3,
                                 cudaMemcpyDeviceToDevice, st));
    R2S_CUDA_TRY(cudaMemcpyAsync(h->interp_center, interp_center, sizeof(float) * (size_t)ne * ns * 3,
                                 cudaMemcpyDeviceToDevice, st));
    if (nv == 2) {
        R2S_CUDA_TRY(cudaMemcpyAsync(h->dyn_vel, dyn_vel, sizeof(float) * (size_t)ne * 6, cudaMemcpyDeviceToDevice, st));
    } else {
        R2S_CUDA_TRY(cudaMemcpy2DAsync(h->dyn_vel, sizeof(float) * 6, dyn_vel, sizeof(float) * 3, sizeof(float) * 3, ne,
                                       cudaMemcpyDeviceToDevice, st));
    }
    R2S_CUDA_TRY(cudaMemcpyAsync(h->dyn_omega, dyn_omega, sizeof(float) * (size_t)ne * 3, cudaMemcpyDeviceToDevice, st));
    return R2S_OK;
}

int r2s_phys_motion_ptrs(r2s_phys* h, int per_env, r2s_phys_motion* out)
{
    R2S_REQUIRE(h && h->F > 0, "r2s_phys_motion_ptrs: no mesh set");
    R2S_REQUIRE(out, "r2s_phys_motion_ptrs: null output");
    const int ne = per_env ? h->d.E : 1, ns = h->d.n_substeps;
    if (h->motion_per_env != (per_env ? 1 : 0)) {   // re-shape the tables; contents start undefined
        cudaFree(h->interp_pts); cudaFree(h->interp_center); cudaFree(h->dyn_vel); cudaFree(h->dyn_omega);
        h->interp_pts = h->interp_center = h->dyn_vel = h->dyn_omega = nullptr;
        if (dmalloc(&h->interp_pts, (size_t)ne * ns * h->n_dyn * 3) || dmalloc(&h->interp_center, (size_t)ne * ns * 3) ||
            dmalloc(&h->dyn_vel, (size_t)ne * 6) || dmalloc(&h->dyn_omega, (size_t)ne * 3))
            return R2S_ERR_CUDA;
        R2S_CUDA_TRY(cudaMemset(h->dyn_vel, 0, sizeof(float) * (size_t)ne * 6));
        R2S_CUDA_TRY(cudaMemset(h->dyn_omega, 0, sizeof(float) * (size_t)ne * 3));
        h->motion_per_env = per_env ? 1 : 0;
    }
    out->interp_pts = h->interp_pts; out->interp_center = h->interp_center;
    out->dyn_vel = h->dyn_vel; out->dyn_omega = h->dyn_omega;
    out->n_env = ne; out->n_substeps = ns; out->n_dyn_verts = h->n_dyn; out->dyn_vel_rows = 2;
    return R2S_OK;
}

static int run_grid(r2s_phys* h, bool resting, cudaStream_t st)
{
    R2S_REQUIRE(h && h->d.self_collision, "self collision is disabled for this system");
    GridParams g{};
    g.E = h->d.E; g.N = h->d.N; g.npow2 = h->npow2; g.cap = h->coll_cap; g.words = h->words;
    g.radius = h->d.collision_dist * 5.0f;
    g.coll_dist = h->d.collision_dist;
    g.x4 = h->x4; g.mask = h->mask; g.resting = h->resting; g.resting_stride = (long long)h->d.N * h->words;
    g.coll_num = h->coll_num; g.coll_idx = h->coll_idx; g.status = h->status; g.key_scratch = h->key_scratch;
    size_t smem = h->key_scratch ? 0 : sizeof(unsigned long long) * (size_t)h->npow2;
    g.stage_x = 0;
    if (!h->key_scratch && smem + sizeof(float4) * (size_t)h->d.N + 1024 <= (size_t)h->max_smem_optin) {
        g.stage_x = 1;
        smem += sizeof(float4) * (size_t)h->d.N;
    }
    if (resting) {
        R2S_CUDA_TRY(cudaMemsetAsync(h->resting, 0, sizeof(unsigned) * (size_t)h->d.E * h->d.N * h->words, st));
        R2S_CUDA_TRY(cudaFuncSetAttribute(grid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grid_kernel<true><<<h->d.E, 1024, smem, st>>>(g);
    } else {
        R2S_CUDA_TRY(cudaFuncSetAttribute(grid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grid_kernel<false><<<h->d.E, 1024, smem, st>>>(g);
    }
    R2S_LAUNCH_CHECK();
    return R2S_OK;
}

int r2s_phys_create_resting_case(r2s_phys* h, void* stream) { return run_grid(h, true, (cudaStream_t)stream); }
int r2s_phys_update_collision_graph(r2s_phys* h, void* stream) { return run_grid(h, false, (cudaStream_t)stream); }

int r2s_phys_step(r2s_phys* h, int32_t n_substeps, void* stream)
{
    R2S_REQUIRE(h, "r2s_phys_step: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = n_substeps > 0 ? n_substeps : h->d.n_substeps;
    R2S_REQUIRE(h->F == 0 || ns <= h->d.n_substeps,
                "r2s_phys_step: %d substeps exceed the mesh motion table (%d rows)", ns, h->d.n_substeps);
    FrameParams p{};
    p.E = h->d.E; p.N = h->d.N; p.n_sub = ns;
    p.has_self_collision = h->d.self_collision; p.use_pusher = h->d.use_pusher; p.sign_mode = h->d.sign_mode;
    p.precise = h->d.precise;
    p.V = h->V; p.F = h->F; p.n_dyn = h->n_dyn; p.coll_cap = h->coll_cap;
    p.stage_dyn = h->stage_dyn; p.smem_forces = h->smem_forces;
    p.dt = h->d.dt; p.dashpot = h->d.dashpot_damping; p.drag_damping = h->d.drag_damping;
    p.rf = h->d.reverse_z ? -1.0f : 1.0f; p.coll_dist = h->d.collision_dist;
    p.c_elas = h->d.collide_elas; p.c_fric = h->d.collide_fric;
    p.ce_elas = h->d.collide_eef_elas; p.ce_fric = h->d.collide_eef_fric;
    p.cs_elas = h->d.collide_self_elas; p.cs_fric = h->d.collide_self_fric;
    p.row_ptr = h->row_ptr; p.nbr_k = h->nbr_k; p.rest = h->rest_csr;
    p.rest_stride = h->rest_envs > 1 ? h->nd : 0;
    p.mass = h->mass; p.unit_mass = h->unit_mass; p.mask = h->mask; p.x4 = h->x4; p.v4 = h->v4; p.vb_scratch = h->vb_scratch;
    p.coll_num = h->coll_num; p.coll_idx = h->coll_idx; p.status = h->status;
    p.stat_verts = h->stat_verts; p.faces = h->faces; p.mesh_map = h->mesh_map; p.face_map = h->face_map; p.dyn_part = h->dyn_part;
    for (int k = 0; k < 15; ++k) p.grp[k] = h->grp[k];
    p.n_grp = h->n_grp;
    p.interp_pts = h->interp_pts; p.interp_center = h->interp_center; p.dyn_vel = h->dyn_vel; p.dyn_omega = h->dyn_omega;
    const long long pe = h->motion_per_env ? 1 : 0;
    p.interp_stride = pe * h->d.n_substeps * h->n_dyn * 3;
    p.center_stride = pe * h->d.n_substeps * 3;
    p.dynvel_stride = pe * 6;
    p.omega_stride = pe * 3;
    p.coll_forces = h->coll_forces;
    p.accel = h->accel; p.F_dyn = h->F_dyn;
    p.frec = h->frec; p.vnorm = h->vnorm; p.cell_start = h->cell_start; p.cell_tris = h->cell_tris; p.cell_skip = h->cell_skip;
    p.gx0 = h->grid0[0]; p.gy0 = h->grid0[1]; p.gz0 = h->grid0[2]; p.gh = h->grid_h;
    p.gnx = h->grid_n[0]; p.gny = h->grid_n[1]; p.gnz = h->grid_n[2];
    p.a0 = h->anchors[0]; p.a1 = h->anchors[1]; p.a2 = h->anchors[2];
    for (int k = 0; k < 12; ++k) p.rest_frame[k] = h->rest_frame[k];
    for (int k = 0; k < 6; ++k) p.rest_box[k] = h->rest_box[k];
    // lanes per adjacency row: 8 suits degrees ~30-60; R2S_PHYS_G overrides it for tuning
    static const int g_lanes = [] { const char* e = getenv("R2S_PHYS_G"); return e ? atoi(e) : 8; }();
    if (p.accel) return launch_frame<8, true>(h, p, st);
    if (g_lanes == 4) return launch_frame<4, false>(h, p, st);
    if (g_lanes == 16) return launch_frame<16, false>(h, p, st);
    if (g_lanes == 32) return launch_frame<32, false>(h, p, st);
    return launch_frame<8, false>(h, p, st);
}

int r2s_phys_get_ptrs(r2s_phys* h, r2s_phys_ptrs* out)
{
    R2S_REQUIRE(h && out, "r2s_phys_get_ptrs: null argument");
    out->x4 = reinterpret_cast<float*>(h->x4);
    out->v4 = reinterpret_cast<float*>(h->v4);
    out->collision_forces = h->coll_forces;
    out->mesh_map = h->mesh_map;
    out->coll_num = h->coll_num;
    out->coll_idx = h->coll_idx;
    out->status = h->status;
    out->F = h->F;
    out->coll_row_cap = h->coll_cap;
    out->smem_state = h->smem_state ? 1 : 0;
    out->smem_bytes = (int32_t)h->smem_bytes;
    return R2S_OK;
}

int64_t r2s_phys_algorithmic_bytes(const r2s_phys* h)
{
    return h ? 52ll * h->d.N + 16ll * h->d.S : 0;
}

}  // extern "C"
