/* oracle/solver_restated.c — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the jpeg2png solver hot path, written from the algorithm, not from the
 * reference text.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (libjpeg2png_b200.so) never does.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so the
 * pin is the reference itself: tests/test_oracle.py runs this file against oracle/_ref/
 * libref_compute.so (the unmodified reference sources compiled with the reference's numeric
 * flags) and requires BIT-IDENTICAL float planes; the vectors under tests/golden/ were produced
 * by that reference build (tests/golden/make_golden.py) and are checked on machines where
 * /root/reference does not exist.
 *
 * Shape of the restatement (deliberately the shape of the CUDA kernels, not of the reference):
 *   - the reference accumulates the TV / TGV sub-gradients by SCATTER in scan order
 *     (compute.c:97-105, :165-183); here every frame pixel GATHERS its eleven contributions in
 *     the one order that reproduces the reference's floating-point sum;
 *   - box/unbox (box.c:5-36) are pure permutations and become addressing;
 *   - FISTA extrapolation, step and projection are expressed per pixel / per 8x8 block with no
 *     in-place aliasing, so every loop is order-independent (and runs under OpenMP), EXCEPT the
 *     fp64 sum of squares of compute.c:200-206 which is kept sequential to match bit for bit.
 *
 * Floating-point contract (reference Makefile:21-22,41-45; compute.c:15-18): fp32 expressions
 * are evaluated in fp32, no contraction into FMA, round-to-nearest; wherever the reference
 * multiplies by a `double` literal the product/sum is evaluated in fp64 and rounded to fp32 once
 * on assignment (ooura/dct.c:24-31).  This file must be compiled with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/jpeg2png_b200.h"

#if defined(__FP_FAST_FMA) && !defined(ORACLE_ALLOW_FMA)
/* -ffp-contract=off is what keeps the compiler from fusing; nothing to check at compile time
 * beyond documenting it.  (FLT_EVAL_METHOD must be 0, as in compute.c:15.) */
#endif
#include <float.h>
_Static_assert(FLT_EVAL_METHOD == 0, "oracle needs fp32 evaluation of fp32 expressions");

/* ------------------------------------------------------------------------------------------
 * 8-point transforms.  Constants: ooura/dct.c:24-31 (CkR = cos(k*pi/16)/2, CkI = sin(k*pi/16)/2,
 * C4R = 1/sqrt(8), W = cos(pi/4)).  They are doubles on purpose: every expression that touches
 * one is evaluated in fp64 and narrowed once.
 * ---------------------------------------------------------------------------------------- */
static const double K1R = 0.49039264020161522456, K1I = 0.09754516100806413392;
static const double K2R = 0.46193976625564337806, K2I = 0.19134171618254488586;
static const double K3R = 0.41573480615127261854, K3I = 0.27778511650980111237;
static const double K4R = 0.35355339059327376220, KW = 0.70710678118654752440;

/* narrow(a*u + b*v) and narrow(a*u - b*v) with fp64 products and fp64 sum, no fusion */
static inline float rot_add(double a, float u, double b, float v) {
        double pu = a * (double)u;
        double pv = b * (double)v;
        return (float)(pu + pv);
}
static inline float rot_sub(double a, float u, double b, float v) {
        double pu = a * (double)u;
        double pv = b * (double)v;
        return (float)(pu - pv);
}
static inline float scale(double a, float u) {
        return (float)(a * (double)u);
}

/* forward 8-point DCT-II on v[0], v[s], ..., v[7s] in place — op graph of ooura/dct.c:104-129 */
static void fdct8(float *v, size_t s) {
        const float i0 = v[0], i1 = v[s], i2 = v[2 * s], i3 = v[3 * s];
        const float i4 = v[4 * s], i5 = v[5 * s], i6 = v[6 * s], i7 = v[7 * s];
        /* butterflies in fp32 */
        const float s07 = i0 + i7, d07 = i0 - i7;
        const float s25 = i2 + i5, d25 = i2 - i5;
        const float s43 = i4 + i3, d43 = i4 - i3;
        const float s61 = i6 + i1, d61 = i6 - i1;
        /* even half */
        float er = s07 + s43, ei = s25 + s61;
        v[0] = scale(K4R, er + ei);
        v[4 * s] = scale(K4R, er - ei);
        er = s07 - s43;
        ei = s25 - s61;
        v[2 * s] = rot_sub(K2R, er, K2I, ei);
        v[6 * s] = rot_add(K2R, ei, K2I, er);
        /* odd half */
        const float m = scale(KW, d25 - d61);
        const float q = scale(KW, d25 + d61);
        const float oi3 = q - d43;
        const float oi1 = q + d43;
        const float or3 = d07 - m;
        const float or1 = d07 + m;
        v[s] = rot_sub(K1R, or1, K1I, oi1);
        v[7 * s] = rot_add(K1R, oi1, K1I, or1);
        v[3 * s] = rot_sub(K3R, or3, K3I, oi3);
        v[5 * s] = rot_add(K3R, oi3, K3I, or3);
}

/* inverse 8-point transform in place — op graph of ooura/dct.c:40-65 */
static void idct8(float *v, size_t s) {
        const float c0 = v[0], c1 = v[s], c2 = v[2 * s], c3 = v[3 * s];
        const float c4 = v[4 * s], c5 = v[5 * s], c6 = v[6 * s], c7 = v[7 * s];
        /* odd half */
        float o1r = rot_add(K1R, c1, K1I, c7);
        const float o1i = rot_sub(K1R, c7, K1I, c1);
        const float o3r = rot_add(K3R, c3, K3I, c5);
        float o3i = rot_sub(K3R, c5, K3I, c3);
        const float dr = o1r - o3r;
        const float di = o1i + o3i;
        o1r = o1r + o3r;
        o3i = o3i - o1i;
        const float p = scale(KW, dr + di);
        const float m = scale(KW, dr - di);
        /* even half */
        const float er = rot_add(K2R, c2, K2I, c6);
        const float ei = rot_sub(K2R, c6, K2I, c2);
        const float zr = scale(K4R, c0 + c4);
        const float zi = scale(K4R, c0 - c4);
        const float t2r = zr - er, t2i = zi - ei;
        const float t0r = zr + er, t0i = zi + ei;
        v[0] = t0r + o1r;
        v[7 * s] = t0r - o1r;
        v[2 * s] = t0i + p;
        v[5 * s] = t0i - p;
        v[4 * s] = t2r - o3i;
        v[3 * s] = t2r + o3i;
        v[6 * s] = t2i - m;
        v[s] = t2i + m;
}

/* 8x8 block, row-major 64 floats.  Both directions run the vertical pass (stride 8) first and
 * the horizontal pass second (ooura/dct.c:39-66 then :67-94; :103-130 then :131-158). */
void oracle_idct8x8(float *b) {
        for (unsigned j = 0; j < 8; j++) idct8(b + j, 8);
        for (unsigned j = 0; j < 8; j++) idct8(b + 8 * j, 1);
}
void oracle_dct8x8(float *b) {
        for (unsigned j = 0; j < 8; j++) fdct8(b + j, 8);
        for (unsigned j = 0; j < 8; j++) fdct8(b + 8 * j, 1);
}

/* ------------------------------------------------------------------------------------------
 * Conventional decode of one plane: dequantise, inverse transform, lay out as a raster.
 * Restates reference jpeg.c:83-92 (decode_coefficients) + jpeg2png.c:131-139 (unbox).
 * Result: coef->fdata = freshly allocated raster h x w (16-byte aligned).
 * ---------------------------------------------------------------------------------------- */
static float *alloc_plane(size_t n) {
        size_t bytes = n * sizeof(float);
        bytes = (bytes + 15) & ~(size_t)15;
        float *p = aligned_alloc(16, bytes ? bytes : 16);
        if (!p) abort();
        return p;
}

void oracle_decode_coefficients(struct coef *coef) {
        const unsigned bw = coef->w / 8, bh = coef->h / 8;
        float *out = alloc_plane((size_t)coef->w * coef->h);
#pragma omp parallel for schedule(static)
        for (unsigned by = 0; by < bh; by++) {
                for (unsigned bx = 0; bx < bw; bx++) {
                        const size_t blk = (size_t)by * bw + bx;
                        float t[64];
                        for (unsigned j = 0; j < 64; j++) {
                                /* int product converted once (jpeg.c:88) */
                                t[j] = (float)((int)coef->data[blk * 64 + j] * (int)coef->quant_table[j]);
                        }
                        oracle_idct8x8(t);
                        for (unsigned r = 0; r < 8; r++)
                                for (unsigned c = 0; c < 8; c++)
                                        out[(size_t)(by * 8 + r) * coef->w + bx * 8 + c] = t[r * 8 + c];
                }
        }
        coef->fdata = out;
}

/* ------------------------------------------------------------------------------------------
 * Colour conversion + quantisation to integer samples, as the reference PNG writer does before
 * handing rows to libpng (png.c:39-62).  `y` must already carry the +128 luma shift
 * (jpeg2png.c:156-159).  Each plane is indexed with its own stride (png.c:39-41).
 * out: bits==8 -> 3 bytes/pixel; bits==16 -> 6 bytes/pixel big-endian.
 * ---------------------------------------------------------------------------------------- */
static inline float clamp255(double x) {
        /* png.c:15-17: CLAMP in double, narrowed to float by the return type */
        return (float)(x > 255. ? 255. : (x < 0. ? 0. : x));
}

void oracle_ycc_to_rgb(unsigned w, unsigned h, unsigned bits, const float *y, unsigned y_stride,
                       const float *cb, unsigned cb_stride, const float *cr, unsigned cr_stride,
                       uint8_t *out) {
        const unsigned depth = bits / 8;
        const float bitfactor = (float)((double)(1 << bits) / 256.);
        for (unsigned i = 0; i < h; i++) {
                for (unsigned j = 0; j < w; j++) {
                        const float yi = y[(size_t)i * y_stride + j];
                        const float cbi = cb[(size_t)i * cb_stride + j];
                        const float cri = cr[(size_t)i * cr_stride + j];
                        /* float + double*float -> double; clamp; float * float; truncate */
                        const unsigned r = (unsigned)(clamp255((double)yi + 1.402 * (double)cri) * bitfactor);
                        const unsigned g = (unsigned)(clamp255(((double)yi - 0.34414 * (double)cbi) - 0.71414 * (double)cri) * bitfactor);
                        const unsigned b = (unsigned)(clamp255((double)yi + 1.772 * (double)cbi) * bitfactor);
                        uint8_t *px = out + ((size_t)i * w + j) * 3 * depth;
                        if (bits == 8) {
                                px[0] = r & 0xFF;
                                px[1] = g & 0xFF;
                                px[2] = b & 0xFF;
                        } else {
                                px[0] = (r >> 8) & 0xFF; px[1] = r & 0xFF;
                                px[2] = (g >> 8) & 0xFF; px[3] = g & 0xFF;
                                px[4] = (b >> 8) & 0xFF; px[5] = b & 0xFF;
                        }
                }
        }
}

/* ------------------------------------------------------------------------------------------
 * The solver.
 * ---------------------------------------------------------------------------------------- */
struct plane {
        /* geometry */
        unsigned cw, ch;    /* coefficient grid, samples */
        unsigned sw, sh;    /* upsampling factors */
        int resample;       /* compute.c:338 */
        const int16_t *data;
        const uint16_t *qt;
        /* frame-sized state */
        float *x;     /* current iterate x_k                        (reference aux.fdata after projection) */
        float *xp;    /* previous iterate x_{k-1}                   (reference aux.fista after the swap)   */
        float *y;     /* FISTA point, then stepped point            (reference aux.fdata during the step)  */
        float *g;     /* objective sub-gradient                     (reference aux.obj_gradient)           */
        float *dx;    /* forward difference in x of y               (reference aux.temp[0])                */
        float *dy;    /* forward difference in y of y               (reference aux.temp[1])                */
        /* per-source TV / TGV contributions, already multiplied/divided as the reference does */
        float *tv_self, *tv_right, *tv_below;
        float *t2_self, *t2_lr, *t2_ud, *t2_diag;
        /* coefficient-grid state */
        float *cosv;  /* clamped DCT coefficients of the last projection, block-major (aux.cos)          */
        float *sub;   /* block means / projected means, raster ch x cw                                     */
};

static inline size_t at(unsigned x, unsigned y, unsigned w) { return (size_t)y * w + x; }

/* aux_init — compute.c:278-310 */
static void plane_init(struct plane *p, struct coef *c, unsigned w, unsigned h) {
        const size_t n = (size_t)w * h;
        p->cw = c->w; p->ch = c->h; p->sw = c->w_samp; p->sh = c->h_samp;
        p->resample = !(c->w == w && c->h == h);
        p->data = c->data; p->qt = c->quant_table;
        p->x = alloc_plane(n); p->xp = alloc_plane(n); p->y = alloc_plane(n); p->g = alloc_plane(n);
        p->dx = alloc_plane(n); p->dy = alloc_plane(n);
        p->tv_self = alloc_plane(n); p->tv_right = alloc_plane(n); p->tv_below = alloc_plane(n);
        p->t2_self = alloc_plane(n); p->t2_lr = alloc_plane(n); p->t2_ud = alloc_plane(n); p->t2_diag = alloc_plane(n);
        const size_t nc = (size_t)c->w * c->h;
        p->cosv = alloc_plane(nc);
        p->sub = alloc_plane(nc);
        for (size_t i = 0; i < nc; i++)
                p->cosv[i] = (float)((int)c->data[i] * (int)c->quant_table[i & 63]);   /* :283 */
        for (unsigned yy = 0; yy < h; yy++) {
                unsigned cy = yy / c->h_samp; if (cy > c->h - 1) cy = c->h - 1;         /* :298 */
                for (unsigned xx = 0; xx < w; xx++) {
                        unsigned cx = xx / c->w_samp; if (cx > c->w - 1) cx = c->w - 1; /* :299 */
                        p->x[at(xx, yy, w)] = c->fdata[at(cx, cy, c->w)];
                }
        }
        memcpy(p->xp, p->x, n * sizeof(float));                                        /* :307-309 */
        free(c->fdata);                                                                /* :304-305 */
        c->fdata = NULL;
}

static void plane_free(struct plane *p) {
        free(p->xp); free(p->y); free(p->g); free(p->dx); free(p->dy);
        free(p->tv_self); free(p->tv_right); free(p->tv_below);
        free(p->t2_self); free(p->t2_lr); free(p->t2_ud); free(p->t2_diag);
        free(p->cosv); free(p->sub);
}

/* DCT-distance term — compute.c:38-70 / compute_simd_step.c:7-62.  Writes g = 0 + alpha*idct(r)
 * over the footprint of every coefficient sample; pixels outside every footprint keep 0.
 * Returns sum of 0.5*(r/q)^2 in the reference's block/j order (the SIMD build's log value). */
static double prob_term(struct plane *p, unsigned w, float p_alpha) {
        const unsigned bw = p->cw / 8, bh = p->ch / 8;
        double dist = 0.;
        for (unsigned by = 0; by < bh; by++) {
                for (unsigned bx = 0; bx < bw; bx++) {
                        const size_t blk = (size_t)by * bw + bx;
                        float r[64];
                        for (unsigned j = 0; j < 64; j++) {
                                const float q = (float)p->qt[j];
                                float v = p->cosv[blk * 64 + j] - (float)p->data[blk * 64 + j] * q;
                                const float nrm = v / q;
                                dist += (double)(nrm * nrm);
                                v = v / (q * q);
                                r[j] = v;
                        }
                        oracle_idct8x8(r);
                        for (unsigned iy = 0; iy < 8; iy++)
                                for (unsigned ix = 0; ix < 8; ix++) {
                                        const float contrib = p_alpha * r[iy * 8 + ix];
                                        for (unsigned sy = 0; sy < p->sh; sy++)
                                                for (unsigned sx = 0; sx < p->sw; sx++) {
                                                        const unsigned fx = (bx * 8 + ix) * p->sw + sx;
                                                        const unsigned fy = (by * 8 + iy) * p->sh + sy;
                                                        float acc = p->g[at(fx, fy, w)]; /* is 0 here */
                                                        acc += contrib;
                                                        p->g[at(fx, fy, w)] = acc;
                                                }
                                }
                }
        }
        return 0.5 * dist;
}

/* projection onto the quantisation box — compute.c:334-404, clamp :323-331, box.c */
static void project(struct plane *p, unsigned w) {
        const unsigned bw = p->cw / 8, bh = p->ch / 8;
        float *f = p->y;   /* stepped point, updated in place exactly like aux.fdata */
        if (p->resample) {
                const float cnt = (float)(p->sw * p->sh);
#pragma omp parallel for schedule(static)
                for (unsigned cy = 0; cy < p->ch; cy++)
                        for (unsigned cx = 0; cx < p->cw; cx++) {
                                float mean = 0.f;
                                for (unsigned sy = 0; sy < p->sh; sy++)
                                        for (unsigned sx = 0; sx < p->sw; sx++)
                                                mean += f[at(cx * p->sw + sx, cy * p->sh + sy, w)];
                                mean /= cnt;
                                p->sub[at(cx, cy, p->cw)] = mean;
                                for (unsigned sy = 0; sy < p->sh; sy++)
                                        for (unsigned sx = 0; sx < p->sw; sx++)
                                                f[at(cx * p->sw + sx, cy * p->sh + sy, w)] -= mean;
                        }
        }
        const float *src = p->resample ? p->sub : f;
        float *dst = p->resample ? p->sub : f;
        const unsigned stride = p->resample ? p->cw : w;
#pragma omp parallel for schedule(static)
        for (unsigned by = 0; by < bh; by++)
                for (unsigned bx = 0; bx < bw; bx++) {
                        const size_t blk = (size_t)by * bw + bx;
                        float t[64];
                        for (unsigned r = 0; r < 8; r++)
                                for (unsigned c = 0; c < 8; c++)
                                        t[r * 8 + c] = src[at(bx * 8 + c, by * 8 + r, stride)];
                        oracle_dct8x8(t);
                        for (unsigned j = 0; j < 64; j++) {
                                const float d = (float)p->data[blk * 64 + j];
                                const float q = (float)p->qt[j];
                                const float lo = (d - 0.5f) * q, hi = (d + 0.5f) * q;
                                float v = t[j];
                                v = v > hi ? hi : (v < lo ? lo : v);
                                t[j] = v;
                                p->cosv[blk * 64 + j] = v;                       /* :381 */
                        }
                        oracle_idct8x8(t);
                        for (unsigned r = 0; r < 8; r++)
                                for (unsigned c = 0; c < 8; c++)
                                        dst[at(bx * 8 + c, by * 8 + r, stride)] = t[r * 8 + c];
                }
        if (p->resample) {
#pragma omp parallel for schedule(static)
                for (unsigned cy = 0; cy < p->ch; cy++)
                        for (unsigned cx = 0; cx < p->cw; cx++) {
                                const float mean = p->sub[at(cx, cy, p->cw)];
                                for (unsigned sy = 0; sy < p->sh; sy++)
                                        for (unsigned sx = 0; sx < p->sw; sx++)
                                                f[at(cx * p->sw + sx, cy * p->sh + sy, w)] += mean;
                        }
        }
}

/* One full solve.  objective_log: NULL or iterations*4 doubles
 * (objective, prob_dist, tv, tv2 per iteration, compute.c:271-272, SIMD-build flavour). */
void oracle_compute(unsigned nchannel, struct coef *coefs, float weight, const float *pweight,
                    unsigned iterations, double *objective_log) {
        unsigned w = 0, h = 0;
        for (unsigned c = 0; c < nchannel; c++) {                                 /* compute.c:410-416 */
                if (coefs[c].w * coefs[c].w_samp > w) w = coefs[c].w * coefs[c].w_samp;
                if (coefs[c].h * coefs[c].h_samp > h) h = coefs[c].h * coefs[c].h_samp;
        }
        const size_t n = (size_t)w * h;
        struct plane pl[3];
        for (unsigned c = 0; c < nchannel; c++) plane_init(&pl[c], &coefs[c], w, h);
        float *n1 = alloc_plane(n), *n2 = alloc_plane(n);

        const float radius = sqrtf((float)h * (float)w) / 2;                      /* :425 */
        const float step = radius / sqrtf((float)(1 + iterations));               /* :443 */
        const float a1 = (float)(1. / (double)sqrtf((float)nchannel));            /* :90 */
        const float tgv_alpha = weight / sqrtf((float)(4 / 2));                   /* :258 */
        const float a2 = (float)(((double)tgv_alpha * 1.) / (double)sqrtf((float)nchannel)); /* :154 */
        float t = 1;

        for (unsigned it = 0; it < iterations; it++) {
                /* FISTA extrapolation — compute.c:431-440 */
                const float tnext = (1 + sqrtf(1 + 4 * (t * t))) / 2;
                const float factor = (t - 1) / tnext;
                t = tnext;
                for (unsigned c = 0; c < nchannel; c++) {
                        struct plane *p = &pl[c];
#pragma omp parallel for schedule(static)
                        for (size_t i = 0; i < n; i++) {
                                const float d = p->x[i] - p->xp[i];
                                p->y[i] = p->x[i] + factor * d;
                        }
                }

                /* gradient, term 1: DCT distance — compute.c:239-248 */
                double prob_dist = 0.;
                float total_alpha = 0.f;
                for (unsigned c = 0; c < nchannel; c++) {
                        memset(pl[c].g, 0, n * sizeof(float));
                        if (pweight[c] != 0.f) {
                                const float p_alpha = pweight[c] * 2 * 255 * sqrtf(2);
                                total_alpha += p_alpha;
                                prob_dist += prob_term(&pl[c], w, p_alpha);
                        }
                }

                /* per-source TV quantities — compute.c:73-113 */
#pragma omp parallel for schedule(static)
                for (unsigned yy = 0; yy < h; yy++)
                        for (unsigned xx = 0; xx < w; xx++) {
                                const size_t i = at(xx, yy, w);
                                float gx[3], gy[3];
                                float nn = 0.f;
                                for (unsigned c = 0; c < nchannel; c++) {
                                        const float *f = pl[c].y;
                                        gx[c] = xx >= w - 1 ? 0.f : f[i + 1] - f[i];
                                        gy[c] = yy >= h - 1 ? 0.f : f[i + w] - f[i];
                                        nn += gx[c] * gx[c];
                                        nn += gy[c] * gy[c];
                                }
                                nn = sqrtf(nn);
                                n1[i] = nn;
                                for (unsigned c = 0; c < nchannel; c++) {
                                        pl[c].dx[i] = gx[c];
                                        pl[c].dy[i] = gy[c];
                                        if (nn != 0.f) {
                                                pl[c].tv_self[i] = (a1 * -(gx[c] + gy[c])) / nn;
                                                pl[c].tv_right[i] = (a1 * gx[c]) / nn;
                                                pl[c].tv_below[i] = (a1 * gy[c]) / nn;
                                        }
                                }
                        }

                /* per-source TGV quantities — compute.c:128-186 */
                const int use_tgv = weight != 0.f;
                if (use_tgv) {
#pragma omp parallel for schedule(static)
                        for (unsigned yy = 0; yy < h; yy++)
                                for (unsigned xx = 0; xx < w; xx++) {
                                        const size_t i = at(xx, yy, w);
                                        float gxx[3], gyy[3], sym[3];
                                        float nn = 0.f;
                                        for (unsigned c = 0; c < nchannel; c++) {
                                                const float *dx = pl[c].dx, *dy = pl[c].dy;
                                                gxx[c] = xx == 0 ? 0.f : dx[i] - dx[i - 1];
                                                const float gyx = xx == 0 ? 0.f : dy[i] - dy[i - 1];
                                                const float gxy = yy == 0 ? 0.f : dx[i] - dx[i - w];
                                                gyy[c] = yy == 0 ? 0.f : dy[i] - dy[i - w];
                                                sym[c] = (float)((double)(gxy + gyx) / 2.);
                                                nn += (gxx[c] * gxx[c] + 2 * (sym[c] * sym[c])) + gyy[c] * gyy[c];
                                        }
                                        nn = sqrtf(nn);
                                        n2[i] = nn;
                                        if (nn != 0.f)
                                                for (unsigned c = 0; c < nchannel; c++) {
                                                        pl[c].t2_self[i] = a2 * (-((2 * gxx[c] + 2 * sym[c]) + 2 * gyy[c]) / nn);
                                                        pl[c].t2_lr[i] = a2 * ((sym[c] + gxx[c]) / nn);
                                                        pl[c].t2_ud[i] = a2 * ((gyy[c] + sym[c]) / nn);
                                                        pl[c].t2_diag[i] = a2 * ((-sym[c]) / nn);
                                                }
                                }
                }

                /* gather — the order below is the order in which the reference's scan-order
                 * scatter reaches each pixel */
#pragma omp parallel for schedule(static)
                for (unsigned yy = 0; yy < h; yy++)
                        for (unsigned xx = 0; xx < w; xx++) {
                                const size_t i = at(xx, yy, w);
                                const int up = yy > 0, dn = yy < h - 1, lf = xx > 0, rt = xx < w - 1;
                                for (unsigned c = 0; c < nchannel; c++) {
                                        const struct plane *p = &pl[c];
                                        float acc = p->g[i];
                                        if (up && n1[i - w] != 0.f) acc += p->tv_below[i - w];
                                        if (lf && n1[i - 1] != 0.f) acc += p->tv_right[i - 1];
                                        if (n1[i] != 0.f) acc += p->tv_self[i];
                                        if (use_tgv) {
                                                if (up && n2[i - w] != 0.f) acc += p->t2_ud[i - w];
                                                if (up && rt && n2[i - w + 1] != 0.f) acc += p->t2_diag[i - w + 1];
                                                if (lf && n2[i - 1] != 0.f) acc += p->t2_lr[i - 1];
                                                if (n2[i] != 0.f) acc += p->t2_self[i];
                                                if (rt && n2[i + 1] != 0.f) acc += p->t2_lr[i + 1];
                                                if (dn && lf && n2[i + w - 1] != 0.f) acc += p->t2_diag[i + w - 1];
                                                if (dn && n2[i + w] != 0.f) acc += p->t2_ud[i + w];
                                        }
                                        pl[c].g[i] = acc;
                                }
                        }

                /* objective values for the log — sequential fp64 sums in scan order */
                if (objective_log) {
                        double tv = 0., tv2 = 0.;
                        for (size_t i = 0; i < n; i++) tv += (double)(a1 * n1[i]);           /* compute_simd_step.c:87-90 */
                        total_alpha += (float)nchannel;
                        if (use_tgv) {
                                for (size_t i = 0; i < n; i++) tv2 += (double)(a2 * n2[i]);  /* :208-212 */
                                total_alpha += tgv_alpha * (float)nchannel;
                        }
                        objective_log[it * 4 + 0] = (tv + tv2 + prob_dist) / (double)total_alpha;
                        objective_log[it * 4 + 1] = prob_dist;
                        objective_log[it * 4 + 2] = tv;
                        objective_log[it * 4 + 3] = tv2;
                }

                /* normalised step — compute.c:200-216; then projection; then rotate the buffers */
                for (unsigned c = 0; c < nchannel; c++) {
                        struct plane *p = &pl[c];
                        double ss = 0.;
                        for (size_t i = 0; i < n; i++) ss += (double)(p->g[i] * p->g[i]);   /* sequential on purpose */
                        const float norm = sqrtf((float)ss);
                        if (norm != 0.f) {
#pragma omp parallel for schedule(static)
                                for (size_t i = 0; i < n; i++) p->y[i] = p->y[i] - step * (p->g[i] / norm);
                        }
                        project(p, w);
                        /* x_{k-1} <- x_k, x_k <- projected point */
                        float *old = p->xp;
                        p->xp = p->x;
                        p->x = p->y;
                        p->y = old;
                }
        }

        for (unsigned c = 0; c < nchannel; c++) {                                  /* compute.c:455-463 */
                coefs[c].fdata = pl[c].x;
                coefs[c].w = w;
                coefs[c].h = h;
                plane_free(&pl[c]);
        }
        free(n1); free(n2);
}
