/*
 * oracle/raster_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU float32 restatement of the forward Gaussian-splat rasterizer-with-depth
 * of kywind/real2sim-eval (third-party/diff-gaussian-rasterization-w-depth,
 * "DGR").  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file.
 *
 * Pinning: this restatement is checked against the reference's own CUDA
 * rasterizer (compiled unmodified into oracle/_ref/ by oracle/Makefile and run
 * on a B200) through the committed fixtures tests/golden/raster_*.npz, made by
 * tests/golden/make_raster_golden.py.
 *
 * Reference (file:line relative to /root/reference/third-party/
 * diff-gaussian-rasterization-w-depth/):
 *   cuda_rasterizer/forward.cu         computeColorFromSH :20-71, computeCov2D :74-113,
 *                                      computeCov3D :118-152, preprocessCUDA :155-257,
 *                                      renderCUDA :262-394
 *   cuda_rasterizer/auxiliary.h        ndc2Pix :41-44, getRect :46-56,
 *                                      transformPoint4x3/4x4 :58-77, in_frustum :139-165
 *   cuda_rasterizer/rasterizer_impl.cu getHigherMsb :35-50, checkFrustum :54-66,
 *                                      duplicateWithKeys :70-111, identifyTileRanges :116-138,
 *                                      Rasterizer::forward :198-341
 *   cuda_rasterizer/config.h           BLOCK_X = BLOCK_Y = 16, NUM_CHANNELS = 3 :15-17
 * GLM (un-vendored submodule, DGR/.gitmodules:1-3) column-major mat3 products
 * are restated term by term in the order GLM evaluates them.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16
#define BLOCK_Y 16

typedef struct {
    int32_t P, D, M, W, H;
    int32_t prefiltered;
    float scale_modifier, tanfovx, tanfovy, z_threshold;
    const float *means3D;       /* P x 3 */
    const float *scales;        /* P x 3 or NULL */
    const float *rotations;     /* P x 4 (w,x,y,z) or NULL */
    const float *opacities;     /* P */
    const float *shs;           /* P x M x 3 or NULL */
    const float *colors_precomp; /* P x 3 or NULL */
    const float *cov3D_precomp; /* P x 6 or NULL */
    const float *viewmatrix;    /* 16, column-major (w2c transposed) */
    const float *projmatrix;    /* 16 */
    const float *campos;        /* 3 */
    const float *bg;            /* 3 */
    /* outputs */
    float *out_color;   /* 3 x H x W */
    float *out_depth;   /* H x W */
    int32_t *radii;     /* P */
    /* optional intermediate outputs (may be NULL) */
    float *depths;        /* P */
    float *means2D;       /* P x 2 */
    float *cov3D;         /* P x 6 */
    float *conic_opacity; /* P x 4 */
    float *rgb;           /* P x 3 */
    uint32_t *tiles_touched; /* P */
    float *final_T;       /* H x W */
    uint32_t *n_contrib;  /* H x W */
    uint32_t *ranges;     /* tiles x 2 */
    uint32_t *point_list; /* capacity point_list_cap */
    uint64_t *point_keys; /* capacity point_list_cap */
    int64_t point_list_cap;
} oracle_raster;

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* AUX:41-44 -- evaluated in double, as the reference's double literals force */
static inline float ndc2Pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

/* AUX:46-56 */
static inline void getRect(float px, float py, int max_radius, uint32_t gx, uint32_t gy,
                           uint32_t *minx, uint32_t *miny, uint32_t *maxx, uint32_t *maxy)
{
    int a;
    a = (int)((px - max_radius) / BLOCK_X); if (a < 0) a = 0; *minx = (uint32_t)a < gx ? (uint32_t)a : gx;
    a = (int)((py - max_radius) / BLOCK_Y); if (a < 0) a = 0; *miny = (uint32_t)a < gy ? (uint32_t)a : gy;
    a = (int)((px + max_radius + BLOCK_X - 1) / BLOCK_X); if (a < 0) a = 0; *maxx = (uint32_t)a < gx ? (uint32_t)a : gx;
    a = (int)((py + max_radius + BLOCK_Y - 1) / BLOCK_Y); if (a < 0) a = 0; *maxy = (uint32_t)a < gy ? (uint32_t)a : gy;
}

/* AUX:58-77 */
static inline void xform4x3(const float *p, const float *m, float *o)
{
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static inline void xform4x4(const float *p, const float *m, float *o)
{
    xform4x3(p, m, o);
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* FWD:118-152.  GLM column-major: M = S*R with M[j][k] = s_k*R[j][k];
 * Sigma[j][i] = sum_k M[i][k]*M[j][k] in k order. */
static void computeCov3D(const float *scale, float mod, const float *rot, float *cov3D)
{
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    float R[3][3] = {
        {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
        {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
        {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    float Mm[3][3];
    for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
            /* S is diagonal: the two zero products GLM adds do not change the value */
            Mm[j][k] = s[k] * R[j][k];
        }
#define SIG(j, i) (Mm[i][0] * Mm[j][0] + Mm[i][1] * Mm[j][1] + Mm[i][2] * Mm[j][2])
    cov3D[0] = SIG(0, 0); cov3D[1] = SIG(0, 1); cov3D[2] = SIG(0, 2);
    cov3D[3] = SIG(1, 1); cov3D[4] = SIG(1, 2); cov3D[5] = SIG(2, 2);
#undef SIG
}

/* FWD:74-113 */
static void computeCov2D(const float *mean, float focal_x, float focal_y, float tan_fovx,
                         float tan_fovy, const float *c, const float *view, float *cov)
{
    float t[3];
    xform4x3(mean, view, t);
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    /* J columns */
    float J[3][3] = {{focal_x / t[2], 0.0f, -(focal_x * t[0]) / (t[2] * t[2])},
                     {0.0f, focal_y / t[2], -(focal_y * t[1]) / (t[2] * t[2])},
                     {0.0f, 0.0f, 0.0f}};
    /* W columns: W[k][i] = view[4*i + k] */
    float Wm[3][3];
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) Wm[k][i] = view[4 * i + k];
    /* T = W*J : T[j][i] = sum_k W[k][i]*J[j][k] */
    float T[3][3];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            T[j][i] = Wm[0][i] * J[j][0] + Wm[1][i] * J[j][1] + Wm[2][i] * J[j][2];
    float Vrk[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    /* A = transpose(T)*transpose(Vrk): A[j][i] = sum_k T[i][k]*Vrk[k][j] */
    float A[3][3];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            A[j][i] = T[i][0] * Vrk[0][j] + T[i][1] * Vrk[1][j] + T[i][2] * Vrk[2][j];
    /* cov = A*T : cov[j][i] = sum_k A[k][i]*T[j][k] */
#define COV(j, i) (A[0][i] * T[j][0] + A[1][i] * T[j][1] + A[2][i] * T[j][2])
    cov[0] = COV(0, 0) + 0.3f;
    cov[1] = COV(0, 1);
    cov[2] = COV(1, 1) + 0.3f;
#undef COV
}

/* FWD:20-71 */
static void computeColorFromSH(int idx, int deg, int max_coeffs, const float *means,
                               const float *campos, const float *shs, float *out)
{
    const float *pos = means + 3 * idx;
    float dir[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    float l = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    dir[0] /= l; dir[1] /= l; dir[2] /= l;
    const float *sh = shs + (size_t)idx * max_coeffs * 3;
    for (int ch = 0; ch < 3; ++ch) {
#define S(k) sh[3 * (k) + ch]
        float result = SH_C0 * S(0);
        if (deg > 0) {
            float x = dir[0], y = dir[1], z = dir[2];
            result = result - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z;
                float xy = x * y, yz = y * z, xz = x * z;
                result = result + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) +
                         SH_C2[2] * (2.0f * zz - xx - yy) * S(6) + SH_C2[3] * xz * S(7) +
                         SH_C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    result = result + SH_C3[0] * y * (3.0f * xx - yy) * S(9) +
                             SH_C3[1] * xy * z * S(10) +
                             SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                             SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                             SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) +
                             SH_C3[5] * z * (xx - yy) * S(14) +
                             SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
#undef S
        result += 0.5f;
        out[ch] = fmaxf(result, 0.0f);
    }
}

/* IMPL:35-50 */
uint32_t oracle_get_higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* IMPL:54-66 (markVisible): z_threshold hard-coded 0.01f */
void oracle_mark_visible(int P, const float *means3D, const float *view, const float *proj,
                         uint8_t *present)
{
    (void)proj;
    for (int i = 0; i < P; ++i) {
        float pv[3];
        xform4x3(means3D + 3 * i, view, pv);
        present[i] = !(pv[2] <= 0.01f);
    }
}

typedef struct { uint64_t key; uint32_t val; } kv;

/* stable LSD radix sort on bits [0, nbits), 8 bits per pass: same result as
 * cub::DeviceRadixSort::SortPairs(..., 0, 32 + bit) (IMPL:306-311) */
static void radix_sort_pairs(kv *a, kv *tmp, size_t n, int nbits)
{
    for (int shift = 0; shift < nbits; shift += 8) {
        int bits = nbits - shift < 8 ? nbits - shift : 8;
        uint32_t maskv = (1u << bits) - 1u;
        size_t count[257];
        memset(count, 0, sizeof(count));
        for (size_t i = 0; i < n; ++i) count[((a[i].key >> shift) & maskv) + 1]++;
        for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
        for (size_t i = 0; i < n; ++i) tmp[count[(a[i].key >> shift) & maskv]++] = a[i];
        kv *t = a; a = tmp; tmp = t;
    }
    /* result may live in either buffer; caller passes an even number of passes
     * or copies back -- handled by the caller via the returned parity */
}

/* Forward pass.  Returns num_rendered (IMPL:340), or -1 if point_list_cap is
 * too small for the optional sorted-list export. */
int64_t oracle_raster_forward(oracle_raster *a)
{
    const int P = a->P, W = a->W, H = a->H;
    const float focal_y = H / (2.0f * a->tanfovy); /* IMPL:223-224 */
    const float focal_x = W / (2.0f * a->tanfovx);
    const uint32_t gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const uint32_t tiles = gx * gy;

    float *depths = (float *)calloc((size_t)P, sizeof(float));
    float *xy = (float *)calloc((size_t)P * 2, sizeof(float));
    float *conic_o = (float *)calloc((size_t)P * 4, sizeof(float));
    float *rgb = (float *)calloc((size_t)P * 3, sizeof(float));
    uint32_t *touched = (uint32_t *)calloc((size_t)P, sizeof(uint32_t));
    int32_t *radii = a->radii ? a->radii : (int32_t *)calloc((size_t)P, sizeof(int32_t));
    float *cov3Ds = (float *)calloc((size_t)P * 6, sizeof(float));

    /* ---- preprocessCUDA, FWD:155-257 ---- */
    for (int idx = 0; idx < P; ++idx) {
        radii[idx] = 0;
        touched[idx] = 0;
        const float *p_orig = a->means3D + 3 * idx;
        float p_view[3];
        xform4x3(p_orig, a->viewmatrix, p_view);
        if (p_view[2] <= a->z_threshold) continue; /* in_frustum, AUX:139-165 */
        float p_hom[4];
        xform4x4(p_orig, a->projmatrix, p_hom);
        float p_w = 1.0f / (p_hom[3] + 0.0000001f);
        float p_proj[2] = {p_hom[0] * p_w, p_hom[1] * p_w};
        const float *cov3D;
        if (a->cov3D_precomp) cov3D = a->cov3D_precomp + 6 * idx;
        else {
            computeCov3D(a->scales + 3 * idx, a->scale_modifier, a->rotations + 4 * idx,
                         cov3Ds + 6 * idx);
            cov3D = cov3Ds + 6 * idx;
        }
        float cov[3];
        computeCov2D(p_orig, focal_x, focal_y, a->tanfovx, a->tanfovy, cov3D, a->viewmatrix, cov);
        float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float pix[2] = {ndc2Pix(p_proj[0], W), ndc2Pix(p_proj[1], H)};
        uint32_t minx, miny, maxx, maxy;
        getRect(pix[0], pix[1], (int)my_radius, gx, gy, &minx, &miny, &maxx, &maxy);
        if ((maxx - minx) * (maxy - miny) == 0) continue;
        if (a->colors_precomp == NULL)
            computeColorFromSH(idx, a->D, a->M, a->means3D, a->campos, a->shs, rgb + 3 * idx);
        depths[idx] = p_view[2];
        radii[idx] = (int32_t)my_radius;
        xy[2 * idx] = pix[0]; xy[2 * idx + 1] = pix[1];
        conic_o[4 * idx] = conic[0]; conic_o[4 * idx + 1] = conic[1];
        conic_o[4 * idx + 2] = conic[2]; conic_o[4 * idx + 3] = a->opacities[idx];
        touched[idx] = (maxy - miny) * (maxx - minx);
    }

    /* ---- InclusiveSum + duplicateWithKeys, IMPL:279-301 ---- */
    size_t R = 0;
    for (int i = 0; i < P; ++i) R += touched[i];
    kv *keys = (kv *)malloc(sizeof(kv) * (R ? R : 1));
    kv *tmp = (kv *)malloc(sizeof(kv) * (R ? R : 1));
    size_t off = 0;
    for (int idx = 0; idx < P; ++idx) {
        if (radii[idx] > 0) {
            uint32_t minx, miny, maxx, maxy;
            getRect(xy[2 * idx], xy[2 * idx + 1], radii[idx], gx, gy, &minx, &miny, &maxx, &maxy);
            uint32_t dbits;
            memcpy(&dbits, &depths[idx], 4);
            for (uint32_t y = miny; y < maxy; ++y)
                for (uint32_t x = minx; x < maxx; ++x) {
                    keys[off].key = ((uint64_t)(y * gx + x) << 32) | dbits;
                    keys[off].val = (uint32_t)idx;
                    off++;
                }
        }
    }
    /* ---- SortPairs over bits [0, 32+bit), IMPL:303-311 ---- */
    int nbits = 32 + (int)oracle_get_higher_msb(tiles);
    int passes = (nbits + 7) / 8;
    radix_sort_pairs(keys, tmp, R, nbits);
    kv *sorted = (passes & 1) ? tmp : keys;

    /* ---- identifyTileRanges, IMPL:313-321 ---- */
    uint32_t *ranges = (uint32_t *)calloc((size_t)tiles * 2, sizeof(uint32_t));
    for (size_t i = 0; i < R; ++i) {
        uint32_t cur = (uint32_t)(sorted[i].key >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(sorted[i - 1].key >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }

    /* ---- renderCUDA, FWD:262-394 (per pixel; block-level early exit only
     *      skips work, it never changes a pixel's result) ---- */
    const float *feat = a->colors_precomp ? a->colors_precomp : rgb;
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            uint32_t tile = (uint32_t)(py / BLOCK_Y) * gx + (uint32_t)(px / BLOCK_X);
            uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
            float pixf[2] = {(float)px, (float)py};
            float T = 1.0f, C[3] = {0, 0, 0}, Dm = 15.0f;
            uint32_t contributor = 0, last_contributor = 0;
            for (uint32_t k = r0; k < r1; ++k) {
                contributor++;
                uint32_t id = sorted[k].val;
                float dx = xy[2 * id] - pixf[0], dy = xy[2 * id + 1] - pixf[1];
                const float *co = conic_o + 4 * id;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float alpha = fminf(0.99f, co[3] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) break; /* done = true */
                for (int ch = 0; ch < 3; ++ch) C[ch] += feat[3 * id + ch] * alpha * T;
                if (T > 0.5f && test_T < 0.5) Dm = depths[id];
                T = test_T;
                last_contributor = contributor;
            }
            size_t pid = (size_t)py * W + px;
            for (int ch = 0; ch < 3; ++ch) a->out_color[(size_t)ch * H * W + pid] = C[ch] + T * a->bg[ch];
            a->out_depth[pid] = Dm;
            if (a->final_T) a->final_T[pid] = T;
            if (a->n_contrib) a->n_contrib[pid] = last_contributor;
        }

    int64_t ret = (int64_t)R;
    if (a->depths) memcpy(a->depths, depths, sizeof(float) * (size_t)P);
    if (a->means2D) memcpy(a->means2D, xy, sizeof(float) * 2 * (size_t)P);
    if (a->cov3D) memcpy(a->cov3D, cov3Ds, sizeof(float) * 6 * (size_t)P);
    if (a->conic_opacity) memcpy(a->conic_opacity, conic_o, sizeof(float) * 4 * (size_t)P);
    if (a->rgb) memcpy(a->rgb, rgb, sizeof(float) * 3 * (size_t)P);
    if (a->tiles_touched) memcpy(a->tiles_touched, touched, sizeof(uint32_t) * (size_t)P);
    if (a->ranges) memcpy(a->ranges, ranges, sizeof(uint32_t) * 2 * (size_t)tiles);
    if (a->point_list || a->point_keys) {
        if ((int64_t)R > a->point_list_cap) ret = -1;
        else
            for (size_t i = 0; i < R; ++i) {
                if (a->point_list) a->point_list[i] = sorted[i].val;
                if (a->point_keys) a->point_keys[i] = sorted[i].key;
            }
    }
    free(depths); free(xy); free(conic_o); free(rgb); free(touched); free(cov3Ds);
    if (!a->radii) free(radii);
    free(keys); free(tmp); free(ranges);
    return ret;
}

int oracle_raster_struct_size(void) { return (int)sizeof(oracle_raster); }
