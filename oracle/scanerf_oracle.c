/*
 * scanerf_oracle.c -- CPU restatement of the reference's native hot-path ops.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the CPU
 * baseline for bench.py.  Nothing in the product package may link, import or
 * call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Arithmetic follows the reference's fp32 operation
 * order.  Where nvcc (default -fmad=true) is expected to contract a*b+c into
 * one FMA in the reference build, fmaf() is written explicitly; this file is
 * compiled with -ffp-contract=off so nothing else is fused.
 *
 * Pinning status: the reference ships no tests or golden vectors (SURVEY.md
 * section 4).  The oracle is pinned against outputs of the reference CUDA
 * extensions rebuilt unmodified for sm_100a (oracle/build_ref.py ->
 * oracle/_ref/ *.so) and run on the B200 box; the vectors that run produced are
 * committed under tests/golden/ (tests/golden/make_ref_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* tiny pthread parallel-for (the image's gcc has no libgomp)                  */
/* ------------------------------------------------------------------------- */
typedef void (*range_fn)(int begin, int end, void *ctx);
typedef struct { range_fn fn; void *ctx; int begin, end; } pf_job_t;
static void *pf_tramp(void *a) { pf_job_t *j = (pf_job_t *)a; j->fn(j->begin, j->end, j->ctx); return NULL; }

ORACLE_API int oracle_num_threads(void)
{
    const char *e = getenv("ORACLE_THREADS");
    int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : n);
}

static void parallel_for(int n, range_fn fn, void *ctx)
{
    int nt = oracle_num_threads();
    if (nt > n) nt = n;
    if (nt <= 1) { fn(0, n, ctx); return; }
    pthread_t th[256]; pf_job_t jobs[256];
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].begin = (int)((long long)n * t / nt);
        jobs[t].end = (int)((long long)n * (t + 1) / nt);
        pthread_create(&th[t], NULL, pf_tramp, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------- */
/* helpers restating cutil_math.h                                             */
/* ------------------------------------------------------------------------- */

/* cuda/include/cutil_math.h:913-916  signf(0) = +1 */
static inline int signf_i(float a) { return a >= 0.0f ? 1 : -1; }
/* cuda/include/cutil_math.h:924-926 */
static inline float safe_divide(float a, float b) { return b != 0.0f ? a / b : 100000000.0f; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return imax(lo, imin(v, hi)); }

/* ------------------------------------------------------------------------- */
/* 1. multi-resolution hash-grid encode                                       */
/* ------------------------------------------------------------------------- */

/* hashgrid/src/hashgrid_bg_kernel.cu:14-24 (same in hashgrid_kernel.cu): only
 * three primes are used, uint32 wrap-around, mask with T-1. */
static inline uint32_t hash3(int x, int y, int z, uint32_t mask)
{
    uint32_t r = 0;
    r ^= (uint32_t)x * 1u;
    r ^= (uint32_t)y * 2654435761u;
    r ^= (uint32_t)z * 805459861u;
    return r & mask;
}

/* hashgrid_bg_kernel.cu:79-90: corner order c = 4*dx + 2*dy + dz */
static inline void corner_indices(uint32_t idx[8], int ix, int iy, int iz, uint32_t mask)
{
    for (int c = 0; c < 8; ++c)
        idx[c] = hash3(ix + ((c >> 2) & 1), iy + ((c >> 1) & 1), iz + (c & 1), mask);
}

/* hashgrid_bg_kernel.cu:26-38: products evaluated left to right */
static inline void linear_weight(float w[8], float ox, float oy, float oz)
{
    w[0] = (1 - ox) * (1 - oy) * (1 - oz);
    w[1] = (1 - ox) * (1 - oy) * oz;
    w[2] = (1 - ox) * oy * (1 - oz);
    w[3] = (1 - ox) * oy * oz;
    w[4] = ox * (1 - oy) * (1 - oz);
    w[5] = ox * (1 - oy) * oz;
    w[6] = ox * oy * (1 - oz);
    w[7] = ox * oy * oz;
}

/* hashgrid_bg_kernel.cu:40-77 */
static inline void dweights(float dx[8], float dy[8], float dz[8], float ox, float oy, float oz)
{
    dx[0] = (-1.0f) * (1 - oy) * (1 - oz); dx[1] = (-1.0f) * (1 - oy) * oz;
    dx[2] = (-1.0f) * oy * (1 - oz);       dx[3] = (-1.0f) * oy * oz;
    dx[4] = (1 - oy) * (1 - oz);           dx[5] = (1 - oy) * oz;
    dx[6] = oy * (1 - oz);                 dx[7] = oy * oz;
    dy[0] = (1 - ox) * (-1.0f) * (1 - oz); dy[1] = (1 - ox) * (-1.0f) * oz;
    dy[2] = (1 - ox) * (1 - oz);           dy[3] = (1 - ox) * oz;
    dy[4] = ox * (-1.0f) * (1 - oz);       dy[5] = ox * (-1.0f) * oz;
    dy[6] = ox * (1 - oz);                 dy[7] = ox * oz;
    dz[0] = (1 - ox) * (1 - oy) * (-1.0f); dz[1] = (1 - ox) * (1 - oy);
    dz[2] = (1 - ox) * oy * (-1.0f);       dz[3] = (1 - ox) * oy;
    dz[4] = ox * (1 - oy) * (-1.0f);       dz[5] = ox * (1 - oy);
    dz[6] = ox * oy * (-1.0f);             dz[7] = ox * oy;
}

typedef struct { int ix, iy, iz; float ox, oy, oz; float sx, sy, sz; } cellpos_t;

/* BG variant prologue, hashgrid_bg_kernel.cu:123-130 and :177-184.
 * (p + 2.0f) / 4.0f is add then multiply by 1.0f/4.0f (cutil_math.h:411-415). */
static inline cellpos_t locate_bg(const float *p, const int *res)
{
    cellpos_t c;
    const float inv4 = 1.0f / 4.0f;
    float ux = (p[0] + 2.0f) * inv4, uy = (p[1] + 2.0f) * inv4, uz = (p[2] + 2.0f) * inv4;
    float vx = ux * (float)(res[0] - 1), vy = uy * (float)(res[1] - 1), vz = uz * (float)(res[2] - 1);
    c.ix = (int)vx; c.iy = (int)vy; c.iz = (int)vz;
    c.ox = vx - (float)c.ix; c.oy = vy - (float)c.iy; c.oz = vz - (float)c.iz;
    /* doffset_dpts = (res-1)/4.0f  (:186) */
    c.sx = (float)(res[0] - 1) * inv4; c.sy = (float)(res[1] - 1) * inv4; c.sz = (float)(res[2] - 1) * inv4;
    return c;
}

static inline float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }

/* bbox variant prologue, hashgrid/src/hashgrid_kernel.cu:126-141; the product
 * float(i)*g + corner is one FMA under nvcc's default contraction. */
static inline cellpos_t locate_bbox(const float *p, const int *res, const float *corner, const float *size)
{
    cellpos_t c;
    float q[3], g[3], o[3]; int i[3];
    for (int a = 0; a < 3; ++a) {
        q[a] = clampf(p[a], corner[a], corner[a] + size[a]);
        g[a] = size[a] / (float)(res[a] - 1);
        i[a] = (int)((q[a] - corner[a]) / g[a]);
        float vmin = fmaf((float)i[a], g[a], corner[a]);
        o[a] = (q[a] - vmin) / g[a];
    }
    c.ix = i[0]; c.iy = i[1]; c.iz = i[2];
    c.ox = o[0]; c.oy = o[1]; c.oz = o[2];
    /* backward scale 1.0f / grid_size (hashgrid_kernel.cu:237-239) */
    c.sx = 1.0f / g[0]; c.sy = 1.0f / g[1]; c.sz = 1.0f / g[2];
    return c;
}

/* Forward.  hashgrid_bg_kernel.cu:106-150 (bbox==NULL) / hashgrid_kernel.cu:105-158.
 * points[B,3], table[L,T,2], res[L,3] -> out[B,L,2]; optional idx_out[B,L,8]
 * returns the hashed corner indices so tests can assert them bit-exactly. */
typedef struct {
    const float *points, *grad_in, *table; const int *res; const float *corner, *size;
    float *out; uint32_t *idx_out; float *grad_points, *grad_table, *gp_lvl; int B, L, T;
} enc_ctx_t;

static void enc_fwd_range(int b0, int b1, void *vctx)
{
    enc_ctx_t *x = (enc_ctx_t *)vctx;
    const float *points = x->points, *table = x->table, *corner = x->corner, *size = x->size;
    const int *res = x->res; float *out = x->out; uint32_t *idx_out = x->idx_out;
    const int L = x->L, T = x->T;
    const uint32_t mask = (uint32_t)T - 1u;
    for (int b = b0; b < b1; ++b) {
        for (int l = 0; l < L; ++l) {
            const float *lev = table + (size_t)l * T * 2;
            cellpos_t c = corner ? locate_bbox(points + 3 * b, res + 3 * l, corner, size)
                                 : locate_bg(points + 3 * b, res + 3 * l);
            uint32_t idx[8]; float w[8];
            corner_indices(idx, c.ix, c.iy, c.iz, mask);
            linear_weight(w, c.ox, c.oy, c.oz);
            float ax = 0.0f, ay = 0.0f;
            for (int k = 0; k < 8; ++k) { /* acc = acc + w*F contracts to FMA */
                ax = fmaf(w[k], lev[2 * (size_t)idx[k]], ax);
                ay = fmaf(w[k], lev[2 * (size_t)idx[k] + 1], ay);
            }
            out[((size_t)b * L + l) * 2] = ax;
            out[((size_t)b * L + l) * 2 + 1] = ay;
            if (idx_out) memcpy(idx_out + ((size_t)b * L + l) * 8, idx, sizeof(idx));
        }
    }
}

ORACLE_API void oracle_hash_encode_fwd(const float *points, const float *table, const int *res,
                                       const float *corner, const float *size,
                                       float *out, uint32_t *idx_out, int B, int L, int T)
{
    enc_ctx_t x; memset(&x, 0, sizeof(x));
    x.points = points; x.table = table; x.res = res; x.corner = corner; x.size = size;
    x.out = out; x.idx_out = idx_out; x.B = B; x.L = L; x.T = T;
    parallel_for(B, enc_fwd_range, &x);
}

/* Backward.  hashgrid_bg_kernel.cu:152-226 / hashgrid_kernel.cu:160-243.
 * grad_table and grad_points are ACCUMULATED into (the reference atomically adds
 * into caller-zeroed buffers).  Levels are independent in grad_table, so the
 * parallel loop is over levels and the summation order is deterministic
 * (point-ascending) -- the reference's own order is not. */
static void enc_bwd_range(int l0, int l1, void *vctx)
{
    enc_ctx_t *x = (enc_ctx_t *)vctx;
    const float *points = x->points, *grad_in = x->grad_in, *table = x->table, *corner = x->corner, *size = x->size;
    const int *res = x->res; float *grad_table = x->grad_table, *gp_lvl = x->gp_lvl;
    const int B = x->B, L = x->L, T = x->T;
    const uint32_t mask = (uint32_t)T - 1u;
    for (int l = l0; l < l1; ++l) {
        const float *lev = table + (size_t)l * T * 2;
        float *glev = grad_table + (size_t)l * T * 2;
        for (int b = 0; b < B; ++b) {
            cellpos_t c = corner ? locate_bbox(points + 3 * b, res + 3 * l, corner, size)
                                 : locate_bg(points + 3 * b, res + 3 * l);
            float gx = grad_in[((size_t)b * L + l) * 2], gy = grad_in[((size_t)b * L + l) * 2 + 1];
            uint32_t idx[8]; float w[8], dx[8], dy[8], dz[8];
            corner_indices(idx, c.ix, c.iy, c.iz, mask);
            linear_weight(w, c.ox, c.oy, c.oz);
            dweights(dx, dy, dz, c.ox, c.oy, c.oz);
            float ddx[2] = {0, 0}, ddy[2] = {0, 0}, ddz[2] = {0, 0};
            for (int k = 0; k < 8; ++k) {
                float fx = lev[2 * (size_t)idx[k]], fy = lev[2 * (size_t)idx[k] + 1];
                glev[2 * (size_t)idx[k]] += w[k] * gx;
                glev[2 * (size_t)idx[k] + 1] += w[k] * gy;
                ddx[0] = fmaf(fx, dx[k], ddx[0]); ddx[1] = fmaf(fy, dx[k], ddx[1]);
                ddy[0] = fmaf(fx, dy[k], ddy[0]); ddy[1] = fmaf(fy, dy[k], ddy[1]);
                ddz[0] = fmaf(fx, dz[k], ddz[0]); ddz[1] = fmaf(fy, dz[k], ddz[1]);
            }
            float *gp = gp_lvl + ((size_t)l * B + b) * 3;
            gp[0] = c.sx * fmaf(gx, ddx[0], gy * ddx[1]);
            gp[1] = c.sy * fmaf(gx, ddy[0], gy * ddy[1]);
            gp[2] = c.sz * fmaf(gx, ddz[0], gy * ddz[1]);
        }
    }
}

static void enc_gp_range(int b0, int b1, void *vctx)
{
    enc_ctx_t *x = (enc_ctx_t *)vctx;
    for (int b = b0; b < b1; ++b)
        for (int l = 0; l < x->L; ++l)
            for (int a = 0; a < 3; ++a)
                x->grad_points[(size_t)b * 3 + a] += x->gp_lvl[((size_t)l * x->B + b) * 3 + a];
}

ORACLE_API void oracle_hash_encode_bwd(const float *points, const float *grad_in, const float *table,
                                       const int *res, const float *corner, const float *size,
                                       float *grad_points, float *grad_table, int B, int L, int T)
{
    enc_ctx_t x; memset(&x, 0, sizeof(x));
    x.points = points; x.grad_in = grad_in; x.table = table; x.res = res; x.corner = corner; x.size = size;
    x.grad_points = grad_points; x.grad_table = grad_table; x.B = B; x.L = L; x.T = T;
    x.gp_lvl = (float *)calloc((size_t)L * B * 3, sizeof(float));
    parallel_for(L, enc_bwd_range, &x);
    parallel_for(B, enc_gp_range, &x);
    free(x.gp_lvl);
}
