/*
 * oracle/physics_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU float32 restatement of the PhysTwin spring-mass substep loop of
 * kywind/real2sim-eval, kernel by kernel.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file.
 *
 * PINNING.  The reference physics executes only through warp-lang 1.7.0 (not installable offline), but its
 * kernel bodies are plain Python: oracle/warp_exec.py interprets the warp surface they use and runs the
 * reference's UNMODIFIED sim/physics/spring_mass_warp.py (and, in tests/test_oracle_reference_stack.py, the
 * unmodified phystwin.py on top of it) on the CPU.  tests/golden/phys_*.npz hold its outputs on nine seeded
 * scenarios (tests/golden/make_physics_golden.py); tests/test_oracle_physics_golden.py requires this file to
 * reproduce them BIT FOR BIT (positions, velocities, per-substep intermediates, candidate rows, resting pairs,
 * per-face contact forces).  Pinned that way: every kernel's arithmetic and step()'s ordering / aliasing.
 * STILL UNPINNED: the three Warp built-ins whose code lives in warp-lang's native library (HashGrid cell
 * arithmetic / iteration order, the mesh closest-point query and its face ties, the winding-number sign) --
 * restated here and in warp_exec.py from their published algorithm; see the notes at each function.
 * Analytic known-answer tests and the shipped T-block rest state stay in tests/test_oracle_physics.py.
 *
 * Reference (all file:line relative to /root/reference/):
 *   sim/physics/spring_mass_warp.py   ("SMW")
 *     eval_springs                :61-104
 *     update_vel_from_force       :107-129
 *     loop / object_collision     :132-193, 230-268
 *     update_potential_collision  :196-227
 *     build_resting_collision_pairs :272-291
 *     mesh_collision              :295-421
 *     integrate_ground_collision  :424-474
 *     SpringMassSystemWarp.step   :823-943 (order + buffer aliasing)
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * All arithmetic is IEEE float32 in the reference's expression order with FMA
 * contraction disabled.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float len(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 cross(v3 a, v3 b)
{
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 normalize(v3 a)
{
    float l = len(a);
    if (l > 0.0f) return divs(a, l);
    return V(0, 0, 0);
}
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline v3 ld(const float *p, int i) { return V(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static inline void st(float *p, int i, v3 a) { p[3 * i] = a.x; p[3 * i + 1] = a.y; p[3 * i + 2] = a.z; }

#define COLL_CAP 500 /* SMW:544-549 row capacity of collision_indices */

/* One environment.  Field order is mirrored by oracle/physics_ref.py (ctypes). */
typedef struct {
    int32_t N, S, n_substeps, self_collision;
    float dt, dashpot_damping, drag_damping, reverse_factor;
    float spring_Y_min, spring_Y_max, collision_dist;
    float collide_elas, collide_fric, collide_eef_elas, collide_eef_fric;
    float collide_self_elas, collide_self_fric;
    /* state (SMW:8-17) */
    float *x, *v, *v_bc, *v_bg, *f;
    /* topology (SMW:555-563, 586-590) */
    const int32_t *springs; /* S x 2 */
    const float *rest;      /* S */
    const float *logY;      /* S (log stiffness, PT:344) */
    const float *mass;      /* N */
    const int32_t *mask;    /* N */
    /* self-collision tables */
    int32_t *coll_idx; /* N x COLL_CAP */
    int32_t *coll_num; /* N */
    uint8_t *resting;  /* N x N bool (SMW:715-720) */
    /* merged mesh (SMW:626-712); n_faces == 0 means "no mesh" */
    int32_t n_verts, n_faces, n_dyn_verts, use_pusher;
    float *mesh_pts;        /* n_verts x 3, dynamic vertices first */
    const int32_t *faces;   /* n_faces x 3 */
    const int32_t *mesh_map; /* n_faces */
    const int32_t *face_map; /* n_faces */
    float *collision_forces; /* n_faces x 3 */
    const float *interp_pts;    /* n_substeps x n_dyn_verts x 3 */
    const float *interp_center; /* n_substeps x 3 (num_eefs == 1) */
    const float *dyn_vel;       /* 2 x 3 (gripper) or 1 x 3 (pusher) */
    const float *dyn_omega;     /* 1 x 3 */
    int32_t sign_mode; /* 0: exact winding number, 1: always outside (+1) */
    int32_t pad_;
} oracle_phys;

/* ------------------------------------------------------------------ P1 */
/* SMW:61-104.  The reference scatters with float atomics (unordered); here
 * springs are applied in index order, which is one of its legal outcomes. */
static void eval_springs(oracle_phys *s)
{
    for (int t = 0; t < s->S; ++t) {
        float k = expf(s->logY[t]);
        if (!(k > s->spring_Y_min)) continue;
        int i1 = s->springs[2 * t], i2 = s->springs[2 * t + 1];
        v3 x1 = ld(s->x, i1), v1 = ld(s->v, i1), x2 = ld(s->x, i2), v2 = ld(s->v, i2);
        float rest = s->rest[t];
        v3 dis = sub(x2, x1);
        float dis_len = len(dis);
        v3 d = divs(dis, fmaxf(dis_len, 1e-6f));
        float kk = clampf(k, s->spring_Y_min, s->spring_Y_max);
        v3 spring_force = muls(d, kk * (dis_len / rest - 1.0f));
        float v_rel = dot(sub(v2, v1), d);
        v3 dashpot = muls(d, s->dashpot_damping * v_rel);
        v3 F = add(spring_force, dashpot);
        st(s->f, i1, add(ld(s->f, i1), F));
        st(s->f, i2, sub(ld(s->f, i2), F));
    }
}

/* Gather form used by the CUDA kernel: force on particle i is the sum over its
 * incident springs, in spring-index order, of F(x_i -> x_j).  F is exactly
 * antisymmetric under endpoint swap, so each term is bit-identical to the
 * scatter form; only the summation order differs (per-particle, spring order).
 * Exposed so tests can pin the CUDA summation order bit-for-bit. */
static void eval_springs_gather(oracle_phys *s, const int32_t *row_ptr, const int32_t *nbr,
                                const int32_t *sid)
{
    for (int i = 0; i < s->N; ++i) {
        v3 acc = V(0, 0, 0);
        v3 x1 = ld(s->x, i), v1 = ld(s->v, i);
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
            int t = sid[e], j = nbr[e];
            float k = expf(s->logY[t]);
            if (!(k > s->spring_Y_min)) continue;
            v3 x2 = ld(s->x, j), v2 = ld(s->v, j);
            v3 dis = sub(x2, x1);
            float dis_len = len(dis);
            v3 d = divs(dis, fmaxf(dis_len, 1e-6f));
            float kk = clampf(k, s->spring_Y_min, s->spring_Y_max);
            v3 spring_force = muls(d, kk * (dis_len / s->rest[t] - 1.0f));
            float v_rel = dot(sub(v2, v1), d);
            v3 dashpot = muls(d, s->dashpot_damping * v_rel);
            acc = add(acc, add(spring_force, dashpot));
        }
        st(s->f, i, acc);
    }
}

/* ------------------------------------------------------------------ P2 */
/* SMW:107-129 */
static void update_vel_from_force(oracle_phys *s, float *v_new)
{
    float drag = expf(-s->dt * s->drag_damping);
    for (int i = 0; i < s->N; ++i) {
        v3 v0 = ld(s->v, i), f0 = ld(s->f, i);
        float m0 = s->mass[i];
        v3 g = muls(muls(V(0.0f, 0.0f, -9.8f), m0), s->reverse_factor);
        v3 all_force = add(f0, g);
        v3 a = divs(all_force, m0);
        v3 v1 = add(v0, muls(a, s->dt));
        st(v_new, i, muls(v1, drag));
    }
}

/* ------------------------------------------------------------------ P3 */
/* SMW:132-193 (loop) + SMW:230-268 (object_collision) */
static void object_collision(oracle_phys *s)
{
    float e = clampf(s->collide_self_elas, 0.0f, 1.0f);
    float mu = clampf(s->collide_self_fric, 0.0f, 2.0f);
    for (int i = 0; i < s->N; ++i) {
        v3 x1 = ld(s->x, i), v1 = ld(s->v_bc, i);
        float m1 = s->mass[i];
        int mask1 = s->mask[i];
        float valid = 0.0f;
        v3 J_sum = V(0, 0, 0);
        int cnt = s->coll_num[i];
        for (int k = 0; k < cnt; ++k) {
            int j = s->coll_idx[(size_t)i * COLL_CAP + k];
            v3 x2 = ld(s->x, j), v2 = ld(s->v_bc, j);
            float m2 = s->mass[j];
            v3 dis = sub(x2, x1);
            float dis_len = len(dis);
            v3 rel = sub(v2, v1);
            if (mask1 != s->mask[j] && dis_len < s->collision_dist && dot(dis, rel) < -1e-4f) {
                valid += 1.0f;
                v3 n = divs(dis, fmaxf(dis_len, 1e-6f));
                v3 v_rel_n = muls(n, dot(rel, n));
                float inv_m = 1.0f / m1 + 1.0f / m2;
                v3 impulse_n = divs(muls(v_rel_n, -(1.0f + e)), inv_m);
                float v_rel_n_len = len(v_rel_n);
                v3 v_rel_t = sub(rel, v_rel_n);
                float v_rel_t_len = fmaxf(len(v_rel_t), 1e-6f);
                float a = fmaxf(0.0f, 1.0f - mu * (1.0f + e) * v_rel_n_len / v_rel_t_len);
                v3 impulse_t = divs(muls(v_rel_t, a - 1.0f), inv_m);
                J_sum = add(J_sum, add(impulse_n, impulse_t));
            }
        }
        if (valid > 0.0f) {
            v3 J_avg = divs(J_sum, valid);
            st(s->v_bg, i, sub(v1, divs(J_avg, m1)));
        } else {
            st(s->v_bg, i, v1);
        }
    }
}

/* ------------------------------------------------------- HashGrid (Warp) */
/* Restated from the published warp/native/hashgrid.h algorithm (Warp 1.x):
 * cell = int(p * inv_w) (C truncation), + 2^20, clamp >= 0, mod 128 per axis;
 * points sorted by cell id, ascending point id inside a cell; a query visits
 * cells [int((p-r)*inv_w), min(int((p+r)*inv_w), start+127)] with x fastest,
 * then y, then z and yields every point stored in each visited cell (no
 * distance filter).  UNPINNED: cannot be checked against warp-lang offline. */
#define GRID_DIM 128
typedef struct {
    int n;
    float inv_w;
    int32_t *cell_of;   /* n: cell id of each sorted entry */
    int32_t *point_ids; /* n: point ids sorted by (cell, id) */
} hashgrid;

static inline int grid_cell(int x, int y, int z)
{
    const int origin = 1 << 20;
    x += origin; y += origin; z += origin;
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    if (z < 0) z = 0;
    int cx = x % GRID_DIM, cy = y % GRID_DIM, cz = z % GRID_DIM;
    return cz * (GRID_DIM * GRID_DIM) + cy * GRID_DIM + cx;
}

typedef struct { int32_t cell, id; } cell_entry;
static int cmp_cell_entry(const void *a, const void *b)
{
    const cell_entry *p = (const cell_entry *)a, *q = (const cell_entry *)b;
    if (p->cell != q->cell) return p->cell < q->cell ? -1 : 1;
    return p->id < q->id ? -1 : (p->id > q->id);
}

static void grid_build(hashgrid *g, const float *x, int n, float radius)
{
    g->n = n;
    g->inv_w = 1.0f / radius;
    cell_entry *tmp = (cell_entry *)malloc(sizeof(cell_entry) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        tmp[i].cell = grid_cell((int)(x[3 * i] * g->inv_w), (int)(x[3 * i + 1] * g->inv_w),
                                (int)(x[3 * i + 2] * g->inv_w));
        tmp[i].id = i;
    }
    qsort(tmp, (size_t)n, sizeof(cell_entry), cmp_cell_entry);
    g->cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    g->point_ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) { g->cell_of[i] = tmp[i].cell; g->point_ids[i] = tmp[i].id; }
    free(tmp);
}
static void grid_free(hashgrid *g) { free(g->cell_of); free(g->point_ids); }

/* first sorted slot whose cell id >= cell */
static int grid_lower(const hashgrid *g, int cell)
{
    int lo = 0, hi = g->n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (g->cell_of[mid] < cell) lo = mid + 1; else hi = mid;
    }
    return lo;
}

typedef void (*grid_visit)(void *ctx, int i, int index);
static void grid_query(const hashgrid *g, v3 p, float radius, void *ctx, int i, grid_visit fn)
{
    int xs = (int)((p.x - radius) * g->inv_w), ys = (int)((p.y - radius) * g->inv_w),
        zs = (int)((p.z - radius) * g->inv_w);
    int xe = (int)((p.x + radius) * g->inv_w), ye = (int)((p.y + radius) * g->inv_w),
        ze = (int)((p.z + radius) * g->inv_w);
    if (xe > xs + GRID_DIM - 1) xe = xs + GRID_DIM - 1;
    if (ye > ys + GRID_DIM - 1) ye = ys + GRID_DIM - 1;
    if (ze > zs + GRID_DIM - 1) ze = zs + GRID_DIM - 1;
    for (int z = zs; z <= ze; ++z)
        for (int y = ys; y <= ye; ++y)
            for (int xx = xs; xx <= xe; ++xx) {
                int cell = grid_cell(xx, y, z);
                for (int k = grid_lower(g, cell); k < g->n && g->cell_of[k] == cell; ++k)
                    fn(ctx, i, g->point_ids[k]);
            }
}

/* ------------------------------------------------------------------ P5 */
/* SMW:272-291 / 729-740: every grid-query neighbour j < i (no distance test)
 * becomes a symmetric "resting" pair. */
static void visit_resting(void *ctx, int i, int index)
{
    oracle_phys *s = (oracle_phys *)ctx;
    if (index < i) {
        s->resting[(size_t)i * s->N + index] = 1;
        s->resting[(size_t)index * s->N + i] = 1;
    }
}
void oracle_phys_create_resting_case(oracle_phys *s)
{
    hashgrid g;
    float radius = s->collision_dist * 5.0f;
    grid_build(&g, s->x, s->N, radius);
    for (int t = 0; t < s->N; ++t) {
        int i = g.point_ids[t]; /* wp.hash_grid_point_id: cell-sorted order */
        grid_query(&g, ld(s->x, i), radius, s, i, visit_resting);
    }
    grid_free(&g);
}

/* ------------------------------------------------------------------ P4 */
/* SMW:196-227 / 806-821.  The reference has no bound check against the row
 * capacity of 500 (out-of-bounds write); here rows stop growing at 500. */
static void visit_potential(void *ctx, int i, int index)
{
    oracle_phys *s = (oracle_phys *)ctx;
    if (index == i) return;
    if (s->resting[(size_t)i * s->N + index] || s->resting[(size_t)index * s->N + i]) return;
    v3 dis = sub(ld(s->x, index), ld(s->x, i));
    float dis_len = len(dis);
    if (s->mask[i] != s->mask[index] && dis_len < s->collision_dist) {
        int c = s->coll_num[i];
        if (c < COLL_CAP) {
            s->coll_idx[(size_t)i * COLL_CAP + c] = index;
            s->coll_num[i] = c + 1;
        }
    }
}
void oracle_phys_update_collision_graph(oracle_phys *s)
{
    hashgrid g;
    float radius = s->collision_dist * 5.0f;
    grid_build(&g, s->x, s->N, radius);
    memset(s->coll_num, 0, sizeof(int32_t) * (size_t)s->N);
    for (int t = 0; t < s->N; ++t) {
        int i = g.point_ids[t];
        grid_query(&g, ld(s->x, i), radius, s, i, visit_potential);
    }
    grid_free(&g);
}

/* ------------------------------------------------------- Mesh (Warp) */
/* Closest point on triangle (Ericson, Real-Time Collision Detection 5.1.5),
 * returned as barycentrics (u, v) with point = u*p0 + v*p1 + (1-u-v)*p2 as
 * wp.mesh_eval_position evaluates it. */
static void closest_bary(v3 a, v3 b, v3 c, v3 p, float *u, float *v)
{
    v3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { *u = 1.0f; *v = 0.0f; return; }
    v3 bp = sub(p, b);
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { *u = 0.0f; *v = 1.0f; return; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float t = d1 / (d1 - d3);
        *u = 1.0f - t; *v = t; return;
    }
    v3 cp = sub(p, c);
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { *u = 0.0f; *v = 0.0f; return; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        *u = 1.0f - w; *v = 0.0f; return;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        *u = 0.0f; *v = 1.0f - w; return;
    }
    float denom = 1.0f / (va + vb + vc);
    float vv = vb * denom, ww = vc * denom;
    *u = 1.0f - vv - ww; *v = vv;
}

static inline v3 mesh_eval(const oracle_phys *s, int face, float u, float v)
{
    const int32_t *t = s->faces + 3 * face;
    v3 p = ld(s->mesh_pts, t[0]), q = ld(s->mesh_pts, t[1]), r = ld(s->mesh_pts, t[2]);
    return add(add(muls(p, u), muls(q, v)), muls(r, 1.0f - u - v));
}

/* Signed solid angle of a triangle seen from p (Van Oosterom & Strackee). */
static inline float solid_angle(v3 a, v3 b, v3 c, v3 p)
{
    a = sub(a, p); b = sub(b, p); c = sub(c, p);
    float la = len(a), lb = len(b), lc = len(c);
    float det = dot(a, cross(b, c));
    float den = la * lb * lc + dot(a, b) * lc + dot(b, c) * la + dot(c, a) * lb;
    return 2.0f * atan2f(det, den);
}

/* wp.mesh_query_point_sign_winding_number(mesh, p, max_dist, accuracy, threshold)
 * restated brute force: strictly-smaller squared distance starting from
 * max_dist^2, faces visited in ascending index (Warp visits in BVH order, so
 * ties between faces sharing an edge/vertex are UNPINNED); sign from the exact
 * winding number (Warp uses a dipole far-field approximation governed by
 * `accuracy`; for meshes of <100 triangles the exact sum is the natural
 * restatement). */
static int mesh_query(const oracle_phys *s, v3 p, float max_dist, float threshold, int *face,
                      float *u, float *v, float *sign)
{
    float best = max_dist * max_dist;
    int hit = -1;
    float bu = 0, bv = 0;
    for (int fc = 0; fc < s->n_faces; ++fc) {
        const int32_t *t = s->faces + 3 * fc;
        v3 a = ld(s->mesh_pts, t[0]), b = ld(s->mesh_pts, t[1]), c = ld(s->mesh_pts, t[2]);
        float uu, vv;
        closest_bary(a, b, c, p, &uu, &vv);
        v3 q = add(add(muls(a, uu), muls(b, vv)), muls(c, 1.0f - uu - vv));
        v3 d = sub(q, p);
        float d2 = dot(d, d);
        if (d2 < best) { best = d2; hit = fc; bu = uu; bv = vv; }
    }
    if (hit < 0) return 0;
    *face = hit; *u = bu; *v = bv;
    if (s->sign_mode == 1) { *sign = 1.0f; return 1; }
    float total = 0.0f;
    for (int fc = 0; fc < s->n_faces; ++fc) {
        const int32_t *t = s->faces + 3 * fc;
        total += solid_angle(ld(s->mesh_pts, t[0]), ld(s->mesh_pts, t[1]), ld(s->mesh_pts, t[2]), p);
    }
    float wn = total * 0.25f * 0.31830988618379067f; /* / (4 pi) */
    *sign = (wn > threshold) ? -1.0f : 1.0f;
    return 1;
}

/* ------------------------------------------------------------------ P7 */
/* SMW:20-29 + :889-900 (refit is implicit in a brute-force query) */
static void set_mesh_points(oracle_phys *s, int step)
{
    memcpy(s->mesh_pts, s->interp_pts + (size_t)step * s->n_dyn_verts * 3,
           sizeof(float) * 3 * (size_t)s->n_dyn_verts);
    memset(s->collision_forces, 0, sizeof(float) * 3 * (size_t)s->n_faces);
}

/* ------------------------------------------------------------------ P6 */
/* SMW:295-421; x and v_bg are read and written in place (SMW:906-907,923-924) */
static void mesh_collision(oracle_phys *s, int step)
{
    const float dt = s->dt;
    /* bounding box of the mesh grown by max_dist: a point outside it has every triangle farther than
     * max_dist, so its query cannot hit (what Warp's BVH traversal prunes; the result is unchanged) */
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int k = 0; k < s->n_verts; ++k)
        for (int c = 0; c < 3; ++c) {
            float q = s->mesh_pts[3 * k + c];
            if (q < lo[c]) lo[c] = q;
            if (q > hi[c]) hi[c] = q;
        }
    const float grow = 0.02f * 1.0001f + 1e-6f;
    for (int i = 0; i < s->N; ++i) {
        v3 x0 = ld(s->x, i), v0 = ld(s->v_bg, i);
        v3 next_x = add(x0, muls(v0, dt));
        v3 next_v = v0;
        int face; float u, v, sign;
        int near_box = next_x.x >= lo[0] - grow && next_x.x <= hi[0] + grow && next_x.y >= lo[1] - grow &&
                       next_x.y <= hi[1] + grow && next_x.z >= lo[2] - grow && next_x.z <= hi[2] + grow;
        if (near_box && mesh_query(s, next_x, 0.02f, 0.6f, &face, &u, &v, &sign)) {
            int is_gripper;
            if (!s->use_pusher) {
                if (s->mesh_map[face] == 0) is_gripper = 1;
                else if (s->mesh_map[face] == 1) is_gripper = 2;
                else is_gripper = 0;
            } else {
                is_gripper = s->mesh_map[face] >= 0 ? 1 : 0;
            }
            v3 p = mesh_eval(s, face, u, v);
            v3 delta = sub(next_x, p);
            float dist = len(delta) * sign;
            float margin = (is_gripper >= 1 && !s->use_pusher) ? 0.005f : 0.001f;
            float err = dist - margin;
            if (err < 0.0f) {
                v3 normal = muls(normalize(delta), sign);
                v3 real_dyn = V(0, 0, 0);
                float ce, cf;
                if (is_gripper >= 1) {
                    v3 c0 = ld(s->interp_center, step);
                    v3 rot = cross(ld(s->dyn_omega, 0), sub(x0, c0));
                    if (is_gripper == 1) real_dyn = add(ld(s->dyn_vel, 0), rot);
                    else real_dyn = add(ld(s->dyn_vel, 1), rot);
                    v0 = sub(v0, real_dyn);
                    ce = clampf(s->collide_eef_elas, 0.0f, 1.0f);
                    cf = clampf(s->collide_eef_fric, 0.0f, 2.0f);
                } else {
                    ce = clampf(s->collide_elas, 0.0f, 1.0f);
                    cf = clampf(s->collide_fric, 0.0f, 2.0f);
                }
                v3 v_normal = muls(normal, dot(v0, normal));
                v3 v_tao = sub(v0, v_normal);
                float v_normal_len = len(v_normal);
                float v_tao_len = fmaxf(len(v_tao), 1e-6f);
                v3 v_normal_new = muls(v_normal, -ce);
                float a = fmaxf(0.0f, 1.0f - cf * (1.0f + ce) * v_normal_len / v_tao_len);
                v3 v_tao_new = muls(v_tao, a);
                next_v = add(v_normal_new, v_tao_new);
                if (is_gripper >= 1) next_v = add(next_v, real_dyn);
                /* SMW:397 re-assigns `query`: for a finger / tool contact the force is booked on the face of the
                 * re-query (SMW:414); a missed re-query leaves Warp's default-constructed result (face 0) */
                int force_face = face;
                if (is_gripper >= 1) {
                    next_x = add(x0, muls(next_v, dt));
                    int face2; float u2, v2, sign2;
                    force_face = 0;
                    if (mesh_query(s, next_x, 0.02f, 0.6f, &face2, &u2, &v2, &sign2)) {
                        force_face = face2;
                        v3 p2 = mesh_eval(s, face2, u2, v2);
                        v3 delta2 = sub(next_x, p2);
                        float dist2 = len(delta2) * sign2;
                        float err2 = dist2 - margin;
                        if (err2 < 0.0f) {
                            v3 n2 = muls(normalize(delta2), sign2);
                            next_x = sub(next_x, muls(n2, err2));
                        }
                    }
                } else {
                    next_x = sub(next_x, muls(normal, err));
                }
                v3 delta_v_normal = sub(v_normal_new, v_normal);
                int fm = s->face_map[force_face];
                st(s->collision_forces, fm, add(ld(s->collision_forces, fm), divs(delta_v_normal, dt)));
            }
        }
        st(s->x, i, next_x);
        st(s->v_bg, i, next_v);
    }
}

/* ------------------------------------------------------------------ P8 */
/* SMW:424-474 */
static void integrate_ground_collision(oracle_phys *s)
{
    const float dt = s->dt, rf = s->reverse_factor;
    for (int i = 0; i < s->N; ++i) {
        v3 x0 = ld(s->x, i), v0 = ld(s->v_bg, i);
        v3 normal = muls(V(0.0f, 0.0f, 1.0f), rf);
        float x_z = x0.z, v_z = v0.z;
        float next_x_z = (x_z + v_z * dt) * rf;
        v3 v1; float toi;
        if (next_x_z < 0.0f && v_z * rf < -1e-4f) {
            v3 v_normal = muls(normal, dot(v0, normal));
            v3 v_tao = sub(v0, v_normal);
            float v_normal_len = len(v_normal);
            float v_tao_len = fmaxf(len(v_tao), 1e-6f);
            float ce = clampf(s->collide_elas, 0.0f, 1.0f);
            float cf = clampf(s->collide_fric, 0.0f, 2.0f);
            v3 v_normal_new = muls(v_normal, -ce);
            float a = fmaxf(0.0f, 1.0f - cf * (1.0f + ce) * v_normal_len / v_tao_len);
            v1 = add(v_normal_new, muls(v_tao, a));
            toi = -(x_z - 0.0f) / v_z;
        } else {
            v1 = v0; toi = 0.0f;
        }
        st(s->x, i, add(add(x0, muls(v0, toi)), muls(v1, dt - toi)));
        st(s->v, i, v1);
    }
}

/* ------------------------------------------------------------------ P9 */
/* SMW:823-943: one frame = n_substeps x { clear f, P1, P2, P3, P7, P6, P8 }.
 * csr != NULL selects the gather summation order (same terms, per-particle
 * spring-index order) used by the CUDA kernel. */
typedef struct { const int32_t *row_ptr, *nbr, *sid; } oracle_csr;

void oracle_phys_step(oracle_phys *s, const oracle_csr *csr)
{
    for (int i = 0; i < s->n_substeps; ++i) {
        memset(s->f, 0, sizeof(float) * 3 * (size_t)s->N);
        if (csr && csr->row_ptr) eval_springs_gather(s, csr->row_ptr, csr->nbr, csr->sid);
        else eval_springs(s);
        if (s->self_collision) {
            update_vel_from_force(s, s->v_bc);
            object_collision(s);
        } else {
            update_vel_from_force(s, s->v_bg);
        }
        if (s->n_faces > 0) {
            set_mesh_points(s, i);
            mesh_collision(s, i);
        }
        integrate_ground_collision(s);
    }
}

/* Batched driver for the CPU baseline / --impl reference arm: E independent
 * environments, one OpenMP thread per environment at a time. */
void oracle_phys_step_batch(oracle_phys *envs, int E, const oracle_csr *csr)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int e = 0; e < E; ++e) oracle_phys_step(&envs[e], csr);
}

/* torchrun exports OMP_NUM_THREADS=1; the baseline legs of bench.py ask for the host's cores explicitly. */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_get_max_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_get_max_threads(void) { return 1; }
#endif

int oracle_phys_struct_size(void) { return (int)sizeof(oracle_phys); }
int oracle_phys_coll_cap(void) { return COLL_CAP; }
