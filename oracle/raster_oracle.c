/*
 * raster_oracle.c — CPU ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the tile-based differentiable 3D-Gaussian splat rasterizer that GoMAvatar calls at
 * reference models/modules/renderer/gaussian.py:9,20,53-67,83-91.  The algorithm lives in the third-party
 * package `diff_gaussian_rasterization` (graphdeco-inria; UNPINNED `pip install git+https://...`, reference
 * README.md:36; API shape implies main ~ 59f5f77).  Its source is NOT under /root/reference and not on this
 * machine, and the reference holds no tests / golden vectors for it:  **PARITY UNPINNED**.  This file restates
 * the published algorithm as specified in SURVEY.md Appendix A (sections cited per function), for the one
 * branch GoMAvatar exercises: colors_precomp + cov3D_precomp, sh_degree 0, scale_modifier 1.
 *
 * Arithmetic contract (shared with the CUDA kernels so that every integer decision is bit-identical):
 * fp32, every operation individually rounded (compile with -ffp-contract=off; the kernels use __fmul_rn /
 * __fadd_rn / IEEE div+sqrt), association exactly as written here, ndc2Pix in fp64.  exp() is never part of an
 * integer decision.  Gradient sums over pixels are accumulated in fp64 (upstream's order is undefined:
 * atomics), everything else is fp32.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC raster_oracle.c -o _build/libraster_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define MAXC 8

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* App. A.2: row-vector product [p,1]·M with M read as float[16] row-major. */
static inline void xform4x3(const float *m, const float *p, float *o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static inline void xform4x4(const float *m, const float *p, float *o) {
    xform4x3(m, p, o);
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* App. A.3 step 6: literals are double upstream -> fp64 evaluation, rounded on return. */
static inline float ndc2pix(float v, int S) { return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

/* Shared by forward (A.3 step 3) and backward (A.7 i-iii): clamped view-space point, J·R rows, cov2D. */
typedef struct {
    float t[3];
    float xmul, ymul;      /* 0 if the tangent was clamped */
    float M0[3], M1[3];    /* rows of J·R */
    float a, b, c;         /* cov2D incl. the +0.3 low-pass */
    float v0[3], v1[3];    /* Sigma·M0^T, Sigma·M1^T */
} Cov2D;

static void cov2d(const float *mean, const float *cov6, const float *view, float fx, float fy,
                  float tanfovx, float tanfovy, Cov2D *o) {
    float t[3];
    xform4x3(view, mean, t);
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    o->xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    o->ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    t[0] = fminf_(limx, fmaxf_(-limx, txtz)) * t[2];
    t[1] = fminf_(limy, fmaxf_(-limy, tytz)) * t[2];
    o->t[0] = t[0]; o->t[1] = t[1]; o->t[2] = t[2];
    const float J00 = fx / t[2];
    const float J02 = -(fx * t[0]) / (t[2] * t[2]);
    const float J11 = fy / t[2];
    const float J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* R[r][k] = view[4k + r] */
    for (int k = 0; k < 3; k++) {
        o->M0[k] = J00 * view[4 * k + 0] + J02 * view[4 * k + 2];
        o->M1[k] = J11 * view[4 * k + 1] + J12 * view[4 * k + 2];
    }
    const float S[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
    for (int k = 0; k < 3; k++) {
        o->v0[k] = S[k][0] * o->M0[0] + S[k][1] * o->M0[1] + S[k][2] * o->M0[2];
        o->v1[k] = S[k][0] * o->M1[0] + S[k][1] * o->M1[1] + S[k][2] * o->M1[2];
    }
    o->a = (o->M0[0] * o->v0[0] + o->M0[1] * o->v0[1] + o->M0[2] * o->v0[2]) + 0.3f;
    o->b = o->M0[0] * o->v1[0] + o->M0[1] * o->v1[1] + o->M0[2] * o->v1[2];
    o->c = (o->M1[0] * o->v1[0] + o->M1[1] * o->v1[1] + o->M1[2] * o->v1[2]) + 0.3f;
}

/* ------------------------------------------------------------------ App. A.3 preprocess (per Gaussian) */
/* Returns N_dup = sum(tiles_touched).  rect = (minx, miny, maxx, maxy), max exclusive. */
int64_t gor_preprocess(int P, int H, int W, const float *means3D, const float *cov6, const float *opacity,
                       const float *view, const float *proj, float tanfovx, float tanfovy,
                       int32_t *radii, float *depth, float *xy, float *conic_opacity, int32_t *rect,
                       uint32_t *tiles_touched) {
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    int64_t total = 0;
#pragma omp parallel for reduction(+ : total) schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0; tiles_touched[i] = 0; depth[i] = 0.f;
        xy[2 * i] = xy[2 * i + 1] = 0.f;
        for (int k = 0; k < 4; k++) { conic_opacity[4 * i + k] = 0.f; rect[4 * i + k] = 0; }
        const float *p = means3D + 3 * i;
        float pv[3];
        xform4x3(view, p, pv);
        if (pv[2] <= 0.2f) continue;                                   /* near cull */
        float ph[4];
        xform4x4(proj, p, ph);
        const float pw = 1.0f / (ph[3] + 0.0000001f);
        const float ppx = ph[0] * pw, ppy = ph[1] * pw;
        Cov2D q;
        cov2d(p, cov6 + 6 * i, view, fx, fy, tanfovx, tanfovy, &q);
        const float det = q.a * q.c - q.b * q.b;
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float conx = q.c * det_inv, cony = -q.b * det_inv, conz = q.a * det_inv;
        const float mid = 0.5f * (q.a + q.c);
        const float disc = sqrtf(fmaxf_(0.1f, mid * mid - det));
        const float lam1 = mid + disc, lam2 = mid - disc;
        const float my_radius = ceilf(3.f * sqrtf(fmaxf_(lam1, lam2)));
        const float px = ndc2pix(ppx, W), py = ndc2pix(ppy, H);
        const int r = (int)my_radius;
        const int minx = imin(gx, imax(0, (int)((px - (float)r) / (float)TILE)));
        const int miny = imin(gy, imax(0, (int)((py - (float)r) / (float)TILE)));
        const int maxx = imin(gx, imax(0, (int)((px + (float)r + (float)TILE - 1.0f) / (float)TILE)));
        const int maxy = imin(gy, imax(0, (int)((py + (float)r + (float)TILE - 1.0f) / (float)TILE)));
        const int area = (maxx - minx) * (maxy - miny);
        if (area == 0) continue;
        depth[i] = pv[2];
        radii[i] = r;
        xy[2 * i] = px; xy[2 * i + 1] = py;
        conic_opacity[4 * i + 0] = conx; conic_opacity[4 * i + 1] = cony;
        conic_opacity[4 * i + 2] = conz; conic_opacity[4 * i + 3] = opacity[i];
        rect[4 * i + 0] = minx; rect[4 * i + 1] = miny; rect[4 * i + 2] = maxx; rect[4 * i + 3] = maxy;
        tiles_touched[i] = (uint32_t)area;
        total += area;
    }
    return total;
}

/* ------------------------------------------------------------------ App. A.4 binning */
typedef struct { uint64_t key; uint32_t id; } KV;
static int kv_cmp(const void *a, const void *b) {
    const KV *x = (const KV *)a, *y = (const KV *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return (x->id > y->id) - (x->id < y->id);        /* stable LSD sort == ties by emission order == id */
}

/* keys_sorted[N_dup] (tile<<32 | depth bits), point_list[N_dup], ranges[T*2] ((0,0) for empty tiles). */
int gor_bin(int P, int H, int W, const int32_t *radii, const float *depth, const int32_t *rect,
            int64_t n_dup, uint64_t *keys_sorted, uint32_t *point_list, uint32_t *ranges) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    KV *kv = (KV *)malloc(sizeof(KV) * (size_t)(n_dup > 0 ? n_dup : 1));
    if (!kv) return -1;
    int64_t off = 0;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint32_t dbits;
        memcpy(&dbits, depth + i, 4);
        for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; y++)
            for (int x = rect[4 * i + 0]; x < rect[4 * i + 2]; x++) {
                kv[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                kv[off].id = (uint32_t)i;
                off++;
            }
    }
    if (off != n_dup) { free(kv); return -2; }
    qsort(kv, (size_t)n_dup, sizeof(KV), kv_cmp);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy));
    for (int64_t k = 0; k < n_dup; k++) {
        keys_sorted[k] = kv[k].key;
        point_list[k] = kv[k].id;
        const uint32_t tile = (uint32_t)(kv[k].key >> 32);
        if (k == 0 || tile != (uint32_t)(kv[k - 1].key >> 32)) ranges[2 * tile] = (uint32_t)k;
        if (k == n_dup - 1 || tile != (uint32_t)(kv[k + 1].key >> 32)) ranges[2 * tile + 1] = (uint32_t)(k + 1);
    }
    free(kv);
    return 0;
}

/* ------------------------------------------------------------------ App. A.5 blend forward */
/* colors [P,C]; out_color [C,H,W]; final_T [H,W]; n_contrib [H,W]. */
int gor_blend_forward(int H, int W, int C, const uint32_t *point_list, const uint32_t *ranges,
                      const float *xy, const float *conic_opacity, const float *colors, const float *bg,
                      float *out_color, float *final_T, uint32_t *n_contrib) {
    if (C > MAXC) return -1;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const uint32_t start = ranges[2 * tile], end = ranges[2 * tile + 1];
        const int tx = tile % gx, ty = tile / gx;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                const float pxf = (float)x, pyf = (float)y;
                float T = 1.0f, acc[MAXC] = {0};
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = start; k < end; k++) {
                    contributor++;
                    const uint32_t g = point_list[k];
                    const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                    const float *co = conic_opacity + 4 * g;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    const float alpha = fminf_(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break;                     /* done: this Gaussian is NOT blended */
                    for (int ch = 0; ch < C; ch++) acc[ch] += colors[(size_t)g * C + ch] * alpha * T;
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)y * W + x;
                final_T[pix] = T;
                n_contrib[pix] = last;
                for (int ch = 0; ch < C; ch++) out_color[(size_t)ch * H * W + pix] = acc[ch] + T * bg[ch];
            }
    }
    return 0;
}

/* ------------------------------------------------------------------ decision margins of the forward blend */
/* Test infrastructure for the parity tests (no upstream counterpart).  The blend takes three kinds of discontinuous decisions
 * per (pixel, Gaussian): power > 0, alpha < 1/255, test_T < 1e-4.  An implementation whose exp() differs from libm's in the
 * last bits (the CUDA kernels use ex2.approx) takes a DIFFERENT decision only where alpha or test_T sits within a few 1e-6
 * (relative) of its threshold — and then the pixel legitimately moves by up to alpha * colour, and the gradients of every
 * Gaussian blended at that pixel move with it.  This function reports, per pixel, the smallest relative distance of any decision
 * taken for it to its threshold (margin [H,W]; +inf where nothing was decided), and flags every Gaussian that passes the
 * alpha test at a pixel whose margin is below `thr` (fragile [P]).  The tests then demand that EVERY value outside the
 * north-star tolerance belongs to such a pixel / Gaussian, instead of allowing a fraction of unexplained outliers. */
int gor_blend_margins(int P, int H, int W, const uint32_t *point_list, const uint32_t *ranges, const float *xy,
                      const float *conic_opacity, float thr, float *margin, uint8_t *fragile) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    (void)P;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const uint32_t start = ranges[2 * tile], end = ranges[2 * tile + 1];
        const int tx = tile % gx, ty = tile / gx;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                const float pxf = (float)x, pyf = (float)y;
                float T = 1.0f, m = INFINITY;
                for (uint32_t k = start; k < end; k++) {
                    const uint32_t g = point_list[k];
                    const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                    const float *co = conic_opacity + 4 * g;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) { m = fminf(m, fabsf(power)); continue; }
                    const float alpha = fminf_(0.99f, co[3] * expf(power));
                    m = fminf(m, fabsf(alpha - 1.0f / 255.0f) * 255.0f);
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    m = fminf(m, fabsf(test_T - 0.0001f) * 10000.0f);
                    if (test_T < 0.0001f) break;
                    T = test_T;
                }
                margin[(size_t)y * W + x] = m;
                if (m < thr)                 /* every Gaussian that can contribute here inherits the pixel's fragility */
                    for (uint32_t k = start; k < end; k++) {
                        const uint32_t g = point_list[k];
                        const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                        const float *co = conic_opacity + 4 * g;
                        const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > 0.0f) continue;
                        if (fminf_(0.99f, co[3] * expf(power)) >= 0.5f / 255.0f) fragile[g] = 1;
                    }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------ App. A.6 blend backward */
/* dL_dpix [C,H,W] -> dL_dmean2D [P,2], dL_dconic [P,3] (A,B,C), dL_dopacity [P], dL_dcolors [P,C]. */
int gor_blend_backward(int P, int H, int W, int C, const uint32_t *point_list, const uint32_t *ranges,
                       const float *xy, const float *conic_opacity, const float *colors, const float *bg,
                       const float *final_T, const uint32_t *n_contrib, const float *dL_dpix,
                       float *dL_dmean2D, float *dL_dconic, float *dL_dopacity, float *dL_dcolors) {
    if (C > MAXC) return -1;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int stride = 6 + C;   /* mean2D 2, conic 3, opacity 1, colors C */
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    double *accs = (double *)calloc((size_t)nthreads * P * stride, sizeof(double));
    if (!accs) return -2;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *A = accs + (size_t)tid * P * stride;
#pragma omp for schedule(dynamic, 4)
        for (int tile = 0; tile < gx * gy; tile++) {
            const uint32_t start = ranges[2 * tile], end = ranges[2 * tile + 1];
            const int tx = tile % gx, ty = tile / gx;
            for (int ly = 0; ly < TILE; ly++)
                for (int lx = 0; lx < TILE; lx++) {
                    const int x = tx * TILE + lx, y = ty * TILE + ly;
                    if (x >= W || y >= H) continue;
                    const size_t pix = (size_t)y * W + x;
                    const float pxf = (float)x, pyf = (float)y;
                    const float T_final = final_T[pix];
                    float T = T_final;
                    uint32_t contributor = end - start;
                    const uint32_t last_contributor = n_contrib[pix];
                    float accum_rec[MAXC] = {0}, last_color[MAXC] = {0}, dpix[MAXC];
                    float bg_dot = 0.f;
                    for (int ch = 0; ch < C; ch++) dpix[ch] = dL_dpix[(size_t)ch * H * W + pix];
                    for (int ch = 0; ch < C; ch++) bg_dot += bg[ch] * dpix[ch];
                    float last_alpha = 0.f;
                    for (uint32_t k = end; k-- > start;) {
                        contributor--;
                        if (contributor >= last_contributor) continue;
                        const uint32_t g = point_list[k];
                        const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                        const float *co = conic_opacity + 4 * g;
                        const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > 0.0f) continue;
                        const float G = expf(power);
                        const float alpha = fminf_(0.99f, co[3] * G);
                        if (alpha < 1.0f / 255.0f) continue;
                        T = T / (1.f - alpha);
                        const float dchannel_dcolor = alpha * T;
                        float dL_dalpha = 0.f;
                        double *Ag = A + (size_t)g * stride;
                        for (int ch = 0; ch < C; ch++) {
                            const float c = colors[(size_t)g * C + ch];
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                            last_color[ch] = c;
                            dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                            Ag[6 + ch] += (double)(dchannel_dcolor * dpix[ch]);
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                        const float dL_dG = co[3] * dL_dalpha;       /* 0.99 clamp ignored, as upstream */
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                        const float dG_ddely = -gdy * co[2] - gdx * co[1];
                        Ag[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                        Ag[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                        Ag[2] += (double)(-0.5f * gdx * dx * dL_dG);
                        Ag[3] += (double)(-0.5f * gdx * dy * dL_dG);
                        Ag[4] += (double)(-0.5f * gdy * dy * dL_dG);
                        Ag[5] += (double)(G * dL_dalpha);
                    }
                }
        }
    }
    for (int g = 0; g < P; g++) {
        double s[6 + MAXC] = {0};
        for (int t = 0; t < nthreads; t++)
            for (int k = 0; k < stride; k++) s[k] += accs[((size_t)t * P + g) * stride + k];
        dL_dmean2D[2 * g] = (float)s[0]; dL_dmean2D[2 * g + 1] = (float)s[1];
        dL_dconic[3 * g] = (float)s[2]; dL_dconic[3 * g + 1] = (float)s[3]; dL_dconic[3 * g + 2] = (float)s[4];
        dL_dopacity[g] = (float)s[5];
        for (int ch = 0; ch < C; ch++) dL_dcolors[(size_t)g * C + ch] = (float)s[6 + ch];
    }
    free(accs);
    return 0;
}

/* ------------------------------------------------------------------ App. A.7 preprocess backward */
/* -> dL_dmeans3D [P,3], dL_dcov6 [P,6] (xx,xy,xz,yy,yz,zz; off-diagonals carry both symmetric entries). */
int gor_preprocess_backward(int P, int H, int W, const float *means3D, const float *cov6, const int32_t *radii,
                            const float *view, const float *proj, float tanfovx, float tanfovy,
                            const float *dL_dmean2D, const float *dL_dconic,
                            float *dL_dmeans3D, float *dL_dcov6) {
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] = 0.f;
        for (int k = 0; k < 6; k++) dL_dcov6[6 * i + k] = 0.f;
        if (!(radii[i] > 0)) continue;
        const float *p = means3D + 3 * i;
        Cov2D q;
        cov2d(p, cov6 + 6 * i, view, fx, fy, tanfovx, tanfovy, &q);
        const float a = q.a, b = q.b, c = q.c;
        const float gA = dL_dconic[3 * i], gB = dL_dconic[3 * i + 1], gC = dL_dconic[3 * i + 2];
        const float den = a * c - b * b;
        const float k2 = 1.0f / ((den * den) + 0.0000001f);
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        if (k2 != 0) {
            dL_da = k2 * (-c * c * gA + 2 * b * c * gB + (den - a * c) * gC);
            dL_dc = k2 * (-a * a * gC + 2 * a * b * gB + (den - a * c) * gA);
            dL_db = k2 * 2 * (b * c * gA - (den + 2 * b * b) * gB + a * b * gC);
            const float *M0 = q.M0, *M1 = q.M1;
            float *o = dL_dcov6 + 6 * i;
            o[0] = M0[0] * M0[0] * dL_da + M0[0] * M1[0] * dL_db + M1[0] * M1[0] * dL_dc;
            o[3] = M0[1] * M0[1] * dL_da + M0[1] * M1[1] * dL_db + M1[1] * M1[1] * dL_dc;
            o[5] = M0[2] * M0[2] * dL_da + M0[2] * M1[2] * dL_db + M1[2] * M1[2] * dL_dc;
            o[1] = 2 * M0[0] * M0[1] * dL_da + (M0[0] * M1[1] + M0[1] * M1[0]) * dL_db + 2 * M1[0] * M1[1] * dL_dc;
            o[2] = 2 * M0[0] * M0[2] * dL_da + (M0[0] * M1[2] + M0[2] * M1[0]) * dL_db + 2 * M1[0] * M1[2] * dL_dc;
            o[4] = 2 * M0[2] * M0[1] * dL_da + (M0[1] * M1[2] + M0[2] * M1[1]) * dL_db + 2 * M1[1] * M1[2] * dL_dc;
        }
        /* (iii) through M = J(t)·R */
        float dM0[3], dM1[3];
        for (int k = 0; k < 3; k++) {
            dM0[k] = 2 * q.v0[k] * dL_da + q.v1[k] * dL_db;
            dM1[k] = 2 * q.v1[k] * dL_dc + q.v0[k] * dL_db;
        }
        float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; k++) {
            dJ00 += view[4 * k + 0] * dM0[k];
            dJ02 += view[4 * k + 2] * dM0[k];
            dJ11 += view[4 * k + 1] * dM1[k];
            dJ12 += view[4 * k + 2] * dM1[k];
        }
        const float tz = 1.f / q.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = q.xmul * -fx * tz2 * dJ02;
        const float dty = q.ymul * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * q.t[0]) * tz3 * dJ02 + (2 * fy * q.t[1]) * tz3 * dJ12;
        float dmean[3];
        for (int k = 0; k < 3; k++) dmean[k] = view[4 * k + 0] * dtx + view[4 * k + 1] * dty + view[4 * k + 2] * dtz;
        /* (iv) mean2D -> mean through the projection */
        float mh[4];
        xform4x4(proj, p, mh);
        const float mw = 1.0f / (mh[3] + 0.0000001f);
        const float mul1 = mh[0] * mw * mw, mul2 = mh[1] * mw * mw;
        const float gx_ = dL_dmean2D[2 * i], gy_ = dL_dmean2D[2 * i + 1];
        for (int k = 0; k < 3; k++)
            dmean[k] += (proj[4 * k + 0] * mw - proj[4 * k + 3] * mul1) * gx_ + (proj[4 * k + 1] * mw - proj[4 * k + 3] * mul2) * gy_;
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] = dmean[k];
    }
    return 0;
}

int gor_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
