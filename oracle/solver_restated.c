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
 * The solver, in strip form.
 *
 * A strip = frame rows [row0, row0+rows) of the frame plus two halo rows on every side that has
 * a neighbour (the stencil reach).  The whole frame is the special case of one strip without
 * halos, and that is how oracle_compute() below runs — so the reference-pinned path and the
 * multi-rank path (tests/test_strips_gloo.py, world size 2 over gloo) are the same code.  The
 * product's strip sessions (session.cu: j2p_session_create_strip / _gradient / _project / _halo)
 * mirror this interface one to one.
 * ---------------------------------------------------------------------------------------- */
struct plane {
        /* geometry */
        unsigned cw, ch;    /* coefficient grid held here: cw samples wide, ch rows (the strip's own rows) */
        unsigned sw, sh;    /* upsampling factors */
        int resample;       /* compute.c:338, decided on the WHOLE frame */
        int use_prob;
        float p_alpha;
        int16_t *data;      /* local copy of the strip's coefficient rows */
        uint16_t qt[64];
        float *fdata0;      /* conventional decode of the strip's coefficient rows */
        /* frame-sized state, Hl x W */
        float *x, *xp, *y, *g, *dx, *dy;
        float *tv_self, *tv_right, *tv_below;
        float *t2_self, *t2_lr, *t2_ud, *t2_diag;
        /* coefficient-grid state */
        float *cosv, *sub;
};

struct oracle_strip {
        unsigned nchannel, W, Hg, Hl, y0g, t0, t1;
        float weight, step, a1, a2, tgv_alpha, t, factor;
        int use_tgv;
        struct plane pl[3];
        float *n1, *n2;
        double prob_dist;    /* log value of the last gradient call */
        float total_alpha;
};

static inline size_t at(unsigned x, unsigned y, unsigned w) { return (size_t)y * w + x; }

struct oracle_strip *oracle_strip_create(unsigned nchannel, const unsigned *plane_w, const unsigned *plane_h,
                                         const unsigned *w_samp, const unsigned *h_samp, float weight,
                                         const float *pweight, unsigned iterations, unsigned row0, unsigned rows) {
        struct oracle_strip *s = calloc(1, sizeof *s);
        unsigned W = 0, H = 0;
        for (unsigned c = 0; c < nchannel; c++) {                                 /* compute.c:410-416 */
                if (plane_w[c] * w_samp[c] > W) W = plane_w[c] * w_samp[c];
                if (plane_h[c] * h_samp[c] > H) H = plane_h[c] * h_samp[c];
        }
        if (rows == 0) { row0 = 0; rows = H; }
        const unsigned halo_top = row0 > 0 ? 2 : 0, halo_bot = row0 + rows < H ? 2 : 0;
        s->nchannel = nchannel; s->W = W; s->Hg = H;
        s->Hl = rows + halo_top + halo_bot;
        s->y0g = row0 - halo_top; s->t0 = halo_top; s->t1 = halo_top + rows;
        s->weight = weight;
        const float radius = sqrtf((float)H * (float)W) / 2;                      /* :425 */
        s->step = radius / sqrtf((float)(1 + iterations));                        /* :443 */
        s->a1 = (float)(1. / (double)sqrtf((float)nchannel));                     /* :90 */
        s->tgv_alpha = weight / sqrtf((float)(4 / 2));                            /* :258 */
        s->a2 = (float)(((double)s->tgv_alpha * 1.) / (double)sqrtf((float)nchannel)); /* :154 */
        s->use_tgv = weight != 0.f;
        s->t = 1;
        const size_t n = (size_t)W * s->Hl;
        for (unsigned c = 0; c < nchannel; c++) {
                struct plane *p = &s->pl[c];
                const unsigned cy0 = row0 / h_samp[c];
                unsigned cy1 = (row0 + rows + h_samp[c] - 1) / h_samp[c];
                if (cy1 > plane_h[c]) cy1 = plane_h[c];
                p->cw = plane_w[c]; p->ch = cy1 - cy0; p->sw = w_samp[c]; p->sh = h_samp[c];
                p->resample = !(plane_w[c] == W && plane_h[c] == H);
                p->use_prob = pweight[c] != 0.f;                                  /* :244 */
                p->p_alpha = pweight[c] * 2 * 255 * sqrtf(2);                     /* :245 */
                p->x = alloc_plane(n); p->xp = alloc_plane(n); p->y = alloc_plane(n); p->g = alloc_plane(n);
                p->dx = alloc_plane(n); p->dy = alloc_plane(n);
                p->tv_self = alloc_plane(n); p->tv_right = alloc_plane(n); p->tv_below = alloc_plane(n);
                p->t2_self = alloc_plane(n); p->t2_lr = alloc_plane(n); p->t2_ud = alloc_plane(n); p->t2_diag = alloc_plane(n);
                const size_t nc = (size_t)p->cw * p->ch;
                p->cosv = alloc_plane(nc); p->sub = alloc_plane(nc); p->fdata0 = alloc_plane(nc);
                p->data = malloc(nc * sizeof(int16_t) + 16);
                memset(p->x, 0, n * sizeof(float)); memset(p->xp, 0, n * sizeof(float)); memset(p->y, 0, n * sizeof(float));
        }
        s->n1 = alloc_plane(n); s->n2 = alloc_plane(n);
        return s;
}

void oracle_strip_destroy(struct oracle_strip *s) {
        for (unsigned c = 0; c < s->nchannel; c++) {
                struct plane *p = &s->pl[c];
                free(p->x); free(p->xp); free(p->y); free(p->g); free(p->dx); free(p->dy);
                free(p->tv_self); free(p->tv_right); free(p->tv_below);
                free(p->t2_self); free(p->t2_lr); free(p->t2_ud); free(p->t2_diag);
                free(p->cosv); free(p->sub); free(p->fdata0); free(p->data);
        }
        free(s->n1); free(s->n2);
        free(s);
}

unsigned oracle_strip_width(const struct oracle_strip *s) { return s->W; }
unsigned oracle_strip_owned_rows(const struct oracle_strip *s) { return s->t1 - s->t0; }

/* aux_init on the owned rows — compute.c:278-310.  `data` / `fdata`: the strip's coefficient rows. */
void oracle_strip_upload(struct oracle_strip *s, unsigned c, const int16_t *data, const uint16_t *quant, const float *fdata) {
        struct plane *p = &s->pl[c];
        const size_t nc = (size_t)p->cw * p->ch;
        memcpy(p->data, data, nc * sizeof(int16_t));
        memcpy(p->qt, quant, 64 * sizeof(uint16_t));
        memcpy(p->fdata0, fdata, nc * sizeof(float));
        for (size_t i = 0; i < nc; i++)
                p->cosv[i] = (float)((int)p->data[i] * (int)p->qt[i & 63]);          /* :283 */
        const unsigned W = s->W;
        for (unsigned r = s->t0; r < s->t1; r++) {
                unsigned cy = (r - s->t0) / p->sh; if (cy > p->ch - 1) cy = p->ch - 1;   /* :298 (the clamp only bites in the last strip) */
                for (unsigned xx = 0; xx < W; xx++) {
                        unsigned cx = xx / p->sw; if (cx > p->cw - 1) cx = p->cw - 1;     /* :299 */
                        p->x[at(xx, r, W)] = p->fdata0[at(cx, cy, p->cw)];
                }
        }
        memcpy(p->xp, p->x, (size_t)W * s->Hl * sizeof(float));                      /* :307-309 */
        s->t = 1;
}

/* side 0 = top, 1 = bottom; what 0 = rows to send, 1 = rows to receive into.  Returns NULL/0 if
 * the strip has no neighbour on that side. */
float *oracle_strip_halo(struct oracle_strip *s, unsigned c, int side, int what, size_t *count) {
        const int has = side == 0 ? s->t0 > 0 : s->t1 < s->Hl;
        *count = has ? (size_t)2 * s->W : 0;
        if (!has) return NULL;
        float *x = s->pl[c].x;
        if (side == 0) return what == 0 ? x + (size_t)s->t0 * s->W : x;
        return what == 0 ? x + (size_t)(s->t1 - 2) * s->W : x + (size_t)s->t1 * s->W;
}

void oracle_strip_copy_halo_to_prev(struct oracle_strip *s) {
        for (unsigned c = 0; c < s->nchannel; c++) {
                struct plane *p = &s->pl[c];
                memcpy(p->xp, p->x, (size_t)s->t0 * s->W * sizeof(float));
                memcpy(p->xp + (size_t)s->t1 * s->W, p->x + (size_t)s->t1 * s->W, (size_t)(s->Hl - s->t1) * s->W * sizeof(float));
        }
}

void oracle_strip_download(struct oracle_strip *s, unsigned c, float *out) {
        memcpy(out, s->pl[c].x + (size_t)s->t0 * s->W, (size_t)(s->t1 - s->t0) * s->W * sizeof(float));
}

/* DCT-distance term — compute.c:38-70 / compute_simd_step.c:7-62.  Writes g = 0 + alpha*idct(r)
 * over the footprint of every coefficient sample of the strip; pixels outside every footprint keep 0.
 * Returns sum of 0.5*(r/q)^2 in the reference's block/j order (the SIMD build's log value). */
static double prob_term(struct oracle_strip *s, struct plane *p) {
        const unsigned bw = p->cw / 8, bh = p->ch / 8, w = s->W;
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
                                        const float contrib = p->p_alpha * r[iy * 8 + ix];
                                        for (unsigned sy = 0; sy < p->sh; sy++)
                                                for (unsigned sx = 0; sx < p->sw; sx++) {
                                                        const unsigned fx = (bx * 8 + ix) * p->sw + sx;
                                                        const unsigned fy = s->t0 + (by * 8 + iy) * p->sh + sy;
                                                        float acc = p->g[at(fx, fy, w)]; /* is 0 here */
                                                        acc += contrib;
                                                        p->g[at(fx, fy, w)] = acc;
                                                }
                                }
                }
        }
        return 0.5 * dist;
}

/* projection of the owned rows onto the quantisation box — compute.c:334-404, clamp :323-331, box.c */
static void project(struct oracle_strip *s, struct plane *p) {
        const unsigned bw = p->cw / 8, bh = p->ch / 8, w = s->W;
        float *f = p->y + (size_t)s->t0 * w;   /* stepped point, owned rows, updated in place like aux.fdata */
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

/* First half of an iteration: FISTA point, sub-gradient on the owned rows, this strip's fp64
 * sums of g^2 (sequential over the owned rows in raster order: for the whole frame that is the
 * reference's own order, compute.c:200-206).  objective: NULL or 4 doubles (valid for a
 * whole-frame strip): objective, prob_dist, tv, tv2 as the SIMD build logs them. */
void oracle_strip_gradient(struct oracle_strip *s, double *sums, double *objective) {
        const unsigned w = s->W, Hl = s->Hl, nchannel = s->nchannel;
        const size_t n = (size_t)w * Hl;
        const long last = (long)s->Hg - 1 - (long)s->y0g;   /* local index of the frame's last row  */
        const long first = -(long)s->y0g;                   /* local index of the frame's first row */
        struct plane *pl = s->pl;
        float *n1 = s->n1, *n2 = s->n2;
        const float a1 = s->a1, a2 = s->a2;

        /* FISTA extrapolation — compute.c:431-440 */
        const float tnext = (1 + sqrtf(1 + 4 * (s->t * s->t))) / 2;
        const float factor = (s->t - 1) / tnext;
        s->t = tnext;
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
                if (pl[c].use_prob) {
                        total_alpha += pl[c].p_alpha;
                        prob_dist += prob_term(s, &pl[c]);
                }
        }

        /* per-source TV quantities — compute.c:73-113.  Sources: every local row whose lower
         * neighbour is available (the last halo row is never a needed source). */
#pragma omp parallel for schedule(static)
        for (unsigned yy = 0; yy < Hl; yy++) {
                const int has_d = (long)yy < last;
                if (has_d && yy + 1 >= Hl) continue;
                for (unsigned xx = 0; xx < w; xx++) {
                        const size_t i = at(xx, yy, w);
                        float gx[3], gy[3];
                        float nn = 0.f;
                        for (unsigned c = 0; c < nchannel; c++) {
                                const float *f = pl[c].y;
                                gx[c] = xx >= w - 1 ? 0.f : f[i + 1] - f[i];
                                gy[c] = !has_d ? 0.f : f[i + w] - f[i];
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
        }

        /* per-source TGV quantities — compute.c:128-186.  Needs the row above unless it is the frame's first row. */
        if (s->use_tgv) {
#pragma omp parallel for schedule(static)
                for (unsigned yy = 0; yy < Hl; yy++) {
                        const int has_u = (long)yy > first;
                        if (has_u && yy == 0) continue;
                        if ((long)yy < last && yy + 1 >= Hl) continue;
                        for (unsigned xx = 0; xx < w; xx++) {
                                const size_t i = at(xx, yy, w);
                                float gxx[3], gyy[3], sym[3];
                                float nn = 0.f;
                                for (unsigned c = 0; c < nchannel; c++) {
                                        const float *dx = pl[c].dx, *dy = pl[c].dy;
                                        gxx[c] = xx == 0 ? 0.f : dx[i] - dx[i - 1];
                                        const float gyx = xx == 0 ? 0.f : dy[i] - dy[i - 1];
                                        const float gxy = !has_u ? 0.f : dx[i] - dx[i - w];
                                        gyy[c] = !has_u ? 0.f : dy[i] - dy[i - w];
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
        }

        /* gather on the owned rows — the order below is the order in which the reference's
         * scan-order scatter reaches each pixel */
#pragma omp parallel for schedule(static)
        for (unsigned yy = s->t0; yy < s->t1; yy++)
                for (unsigned xx = 0; xx < w; xx++) {
                        const size_t i = at(xx, yy, w);
                        const int up = (long)yy > first, dn = (long)yy < last, lf = xx > 0, rt = xx < w - 1;
                        for (unsigned c = 0; c < nchannel; c++) {
                                const struct plane *p = &pl[c];
                                float acc = p->g[i];
                                if (up && n1[i - w] != 0.f) acc += p->tv_below[i - w];
                                if (lf && n1[i - 1] != 0.f) acc += p->tv_right[i - 1];
                                if (n1[i] != 0.f) acc += p->tv_self[i];
                                if (s->use_tgv) {
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

        /* objective values for the log — sequential fp64 sums in scan order (owned rows) */
        if (objective) {
                double tv = 0., tv2 = 0.;
                const size_t i0 = (size_t)s->t0 * w, i1 = (size_t)s->t1 * w;
                for (size_t i = i0; i < i1; i++) tv += (double)(a1 * n1[i]);           /* compute_simd_step.c:87-90 */
                total_alpha += (float)nchannel;
                if (s->use_tgv) {
                        for (size_t i = i0; i < i1; i++) tv2 += (double)(a2 * n2[i]);  /* :208-212 */
                        total_alpha += s->tgv_alpha * (float)nchannel;
                }
                objective[0] = (tv + tv2 + prob_dist) / (double)total_alpha;
                objective[1] = prob_dist;
                objective[2] = tv;
                objective[3] = tv2;
        }

        /* this strip's sums of squares — sequential on purpose (compute.c:200-206) */
        for (unsigned c = 0; c < nchannel; c++) {
                const float *g = pl[c].g;
                double ss = 0.;
                const size_t i0 = (size_t)s->t0 * w, i1 = (size_t)s->t1 * w;
                for (size_t i = i0; i < i1; i++) ss += (double)(g[i] * g[i]);
                sums[c] = ss;
        }
        for (unsigned c = nchannel; c < 3; c++) sums[c] = 0.;
}

/* Second half: fold the per-rank sums in rank order, normalised step (compute.c:209-216),
 * projection, rotate the buffers.  The halo rows of the new iterate are stale until the driver
 * exchanges them. */
void oracle_strip_project(struct oracle_strip *s, const double *sums_by_rank, unsigned nranks) {
        const unsigned w = s->W;
        const size_t i0 = (size_t)s->t0 * w, i1 = (size_t)s->t1 * w;
        for (unsigned c = 0; c < s->nchannel; c++) {
                struct plane *p = &s->pl[c];
                double ss = 0.;
                for (unsigned r = 0; r < nranks; r++) ss += sums_by_rank[r * 3 + c];
                const float norm = sqrtf((float)ss);
                if (norm != 0.f) {
#pragma omp parallel for schedule(static)
                        for (size_t i = i0; i < i1; i++) p->y[i] = p->y[i] - s->step * (p->g[i] / norm);
                }
                project(s, p);
                /* x_{k-1} <- x_k, x_k <- projected point */
                float *old = p->xp;
                p->xp = p->x;
                p->x = p->y;
                p->y = old;
        }
}

/* One full solve of a whole frame, with the reference's ownership rules for coefs[c].fdata.
 * objective_log: NULL or iterations*4 doubles (compute.c:271-272, SIMD-build flavour). */
void oracle_compute(unsigned nchannel, struct coef *coefs, float weight, const float *pweight,
                    unsigned iterations, double *objective_log) {
        unsigned pw[3], ph[3], sw[3], sh[3];
        for (unsigned c = 0; c < nchannel; c++) {
                pw[c] = coefs[c].w; ph[c] = coefs[c].h; sw[c] = coefs[c].w_samp; sh[c] = coefs[c].h_samp;
        }
        struct oracle_strip *s = oracle_strip_create(nchannel, pw, ph, sw, sh, weight, pweight, iterations, 0, 0);
        for (unsigned c = 0; c < nchannel; c++) {
                oracle_strip_upload(s, c, coefs[c].data, coefs[c].quant_table, coefs[c].fdata);
                free(coefs[c].fdata);                                             /* compute.c:304-305 */
                coefs[c].fdata = NULL;
        }
        for (unsigned it = 0; it < iterations; it++) {
                double sums[3];
                oracle_strip_gradient(s, sums, objective_log ? objective_log + 4 * it : NULL);
                oracle_strip_project(s, sums, 1);
        }
        for (unsigned c = 0; c < nchannel; c++) {                                  /* compute.c:455-463 */
                float *out = alloc_plane((size_t)s->W * s->Hg);
                oracle_strip_download(s, c, out);
                coefs[c].fdata = out;
                coefs[c].w = s->W;
                coefs[c].h = s->Hg;
        }
        oracle_strip_destroy(s);
}
