/* jpeg_reader.c — JPEG file -> quantised DCT coefficients (see jpeg_reader.h).
 *
 * A from-scratch ITU T.81 entropy decoder that stops where libjpeg's jpeg_read_coefficients
 * stops: after Huffman decoding, before dequantisation.  Section references are to T.81.
 */
#include "jpeg_reader.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* zigzag position -> natural (row-major) index, T.81 figure A.6 */
static const uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                               41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                               30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct huff {
        int present;
        uint8_t bits[17];
        uint8_t vals[256];
        int mincode[17], maxcode[18], valptr[17];     /* T.81 F.2.2.3 */
};

struct comp {
        int id, h, v, tq;
        unsigned wb, hb;          /* real block grid (jpeg.c:52-53) */
        unsigned pwb, phb;        /* MCU-padded block grid used while decoding */
        int16_t *blk;             /* [phb][pwb][64], natural order */
        int dc_pred;
        int td, ta;               /* tables of the current scan */
};

struct dec {
        const uint8_t *p, *end;
        uint32_t acc;
        int nbits;
        int hit_marker;           /* a marker was met inside entropy data (stop feeding bits, give zeros) */
        uint16_t qt[4][64];       /* natural order */
        int qt_present[4];
        struct huff dc[4], ac[4];
        struct comp c[3];
        int ncomp, progressive;
        unsigned W, H, maxh, maxv, mcux, mcuy;
        unsigned restart_interval;
        char *err;
        size_t errlen;
        int failed;
};

static int fail(struct dec *d, const char *fmt, ...) {
        if (!d->failed && d->err && d->errlen) {
                va_list ap;
                va_start(ap, fmt);
                vsnprintf(d->err, d->errlen, fmt, ap);
                va_end(ap);
        }
        d->failed = 1;
        return -1;
}

/* ---- bit reader over entropy-coded data with byte stuffing (B.1.1.5) ----------------------- */
static void fill(struct dec *d) {
        while (d->nbits <= 24) {
                unsigned byte = 0;
                if (!d->hit_marker && d->p < d->end) {
                        byte = *d->p;
                        if (byte == 0xFF) {
                                if (d->p + 1 < d->end && d->p[1] == 0x00) {
                                        d->p += 2;
                                } else {
                                        d->hit_marker = 1;   /* leave the marker in place */
                                        byte = 0;
                                }
                        } else {
                                d->p++;
                        }
                }
                d->acc |= (uint32_t)byte << (24 - d->nbits);
                d->nbits += 8;
        }
}
static inline int getbits(struct dec *d, int n) {
        if (n <= 0) return 0;
        if (n > 16) {                                    /* only a corrupt Huffman table can ask for more (F.1.2.1.1: at most 15/16) */
                fail(d, "corrupt jpeg: bad magnitude category");
                return 0;
        }
        if (d->nbits < n) fill(d);
        const int v = (int)(d->acc >> (32 - n));
        d->acc <<= n;
        d->nbits -= n;
        return v;
}
static inline int getbit(struct dec *d) { return getbits(d, 1); }

static int build_huff(struct dec *d, struct huff *h) {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
                h->valptr[l] = k;
                h->mincode[l] = code;
                code += h->bits[l];
                k += h->bits[l];
                h->maxcode[l] = h->bits[l] ? code - 1 : -1;
                if (code > (1 << l)) return fail(d, "corrupt jpeg: bad huffman table");
                code <<= 1;
        }
        h->maxcode[17] = 0x7fffffff;
        h->present = 1;
        return 0;
}
static int decode_huff(struct dec *d, const struct huff *h) {
        int code = 0;
        for (int l = 1; l <= 16; l++) {
                code = (code << 1) | getbit(d);
                if (h->maxcode[l] >= 0 && code <= h->maxcode[l] && code >= h->mincode[l])
                        return h->vals[h->valptr[l] + code - h->mincode[l]];
        }
        fail(d, "corrupt jpeg: bad huffman code");
        return 0;
}
static inline int extend(int v, int s) { return (s < 1 || s > 16) ? 0 : (v < (1 << (s - 1)) ? v - (1 << s) + 1 : v); }   /* F.2.2.1 */

/* ---- block decoders ------------------------------------------------------------------------ */
static void block_sequential(struct dec *d, struct comp *c, int16_t *b) {
        int s = decode_huff(d, &d->dc[c->td]);
        int diff = s ? extend(getbits(d, s), s) : 0;
        c->dc_pred += diff;
        b[0] = (int16_t)c->dc_pred;
        for (int k = 1; k < 64; k++) {
                const int rs = decode_huff(d, &d->ac[c->ta]);
                const int r = rs >> 4;
                s = rs & 15;
                if (s) {
                        k += r;
                        if (k > 63) { fail(d, "corrupt jpeg: coefficient index out of range"); return; }
                        b[ZZ[k]] = (int16_t)extend(getbits(d, s), s);
                } else {
                        if (r != 15) break;     /* EOB */
                        k += 15;
                }
                if (d->failed) return;
        }
}
static void block_dc_first(struct dec *d, struct comp *c, int16_t *b, int al) {
        const int s = decode_huff(d, &d->dc[c->td]);
        const int diff = s ? extend(getbits(d, s), s) : 0;
        c->dc_pred += diff;
        b[0] = (int16_t)(c->dc_pred * (1 << al));
}
static void block_dc_refine(struct dec *d, int16_t *b, int al) {
        if (getbit(d)) b[0] |= (int16_t)(1 << al);
}
static void block_ac_first(struct dec *d, struct comp *c, int16_t *b, int ss, int se, int al, unsigned *eobrun) {
        if (*eobrun > 0) { (*eobrun)--; return; }
        for (int k = ss; k <= se; k++) {
                const int rs = decode_huff(d, &d->ac[c->ta]);
                const int r = rs >> 4, s = rs & 15;
                if (s) {
                        k += r;
                        if (k > 63) { fail(d, "corrupt jpeg: coefficient index out of range"); return; }
                        b[ZZ[k]] = (int16_t)(extend(getbits(d, s), s) * (1 << al));
                } else {
                        if (r == 15) { k += 15; }
                        else {
                                *eobrun = 1u << r;
                                if (r) *eobrun += (unsigned)getbits(d, r);
                                (*eobrun)--;
                                break;
                        }
                }
                if (d->failed) return;
        }
}
static void block_ac_refine(struct dec *d, struct comp *c, int16_t *b, int ss, int se, int al, unsigned *eobrun) {   /* G.1.2.3 */
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        if (*eobrun == 0) {
                for (; k <= se; k++) {
                        const int rs = decode_huff(d, &d->ac[c->ta]);
                        int r = rs >> 4, s = rs & 15;
                        if (d->failed) return;
                        if (s) {
                                s = getbit(d) ? p1 : m1;
                        } else if (r != 15) {
                                *eobrun = 1u << r;
                                if (r) *eobrun += (unsigned)getbits(d, r);
                                break;
                        }
                        /* skip over already non-zero coefficients (each takes a correction bit) and r zero ones */
                        do {
                                int16_t *coef = &b[ZZ[k]];
                                if (*coef != 0) {
                                        if (getbit(d) && (*coef & p1) == 0) *coef = (int16_t)(*coef + (*coef >= 0 ? p1 : m1));
                                } else {
                                        if (--r < 0) break;
                                }
                                k++;
                        } while (k <= se);
                        if (s && k <= se) b[ZZ[k]] = (int16_t)s;
                }
        }
        if (*eobrun > 0) {
                for (; k <= se; k++) {
                        int16_t *coef = &b[ZZ[k]];
                        if (*coef != 0 && getbit(d) && (*coef & p1) == 0) *coef = (int16_t)(*coef + (*coef >= 0 ? p1 : m1));
                }
                (*eobrun)--;
        }
}

/* ---- one scan ------------------------------------------------------------------------------ */
static int restart(struct dec *d, struct comp **sc, int ns, unsigned *eobrun, int expect) {
        /* byte-align, then RSTn (E.2.4) */
        d->acc = 0;
        d->nbits = 0;
        d->hit_marker = 0;
        while (d->p + 1 < d->end && !(d->p[0] == 0xFF && d->p[1] >= 0xD0 && d->p[1] <= 0xD7)) {
                if (d->p[0] == 0xFF && d->p[1] != 0x00 && d->p[1] != 0xFF) return fail(d, "corrupt jpeg: missing restart marker");
                d->p++;
        }
        if (d->p + 1 >= d->end) return fail(d, "corrupt jpeg: truncated at restart marker");
        if ((d->p[1] & 7) != (expect & 7)) return fail(d, "corrupt jpeg: restart marker out of sequence");
        d->p += 2;
        for (int i = 0; i < ns; i++) sc[i]->dc_pred = 0;
        *eobrun = 0;
        return 0;
}

static int decode_scan(struct dec *d, struct comp **sc, int ns, int ss, int se, int ah, int al) {
        d->acc = 0;
        d->nbits = 0;
        d->hit_marker = 0;
        for (int i = 0; i < ns; i++) sc[i]->dc_pred = 0;
        unsigned eobrun = 0, since_restart = 0;
        int rst = 0;
        const int interleaved = ns > 1;
        unsigned nmx, nmy;
        if (interleaved) { nmx = d->mcux; nmy = d->mcuy; }
        else { nmx = sc[0]->wb; nmy = sc[0]->hb; }           /* A.2.3: non-interleaved MCU = one block of the real grid */
        for (unsigned my = 0; my < nmy; my++)
                for (unsigned mx = 0; mx < nmx; mx++) {
                        if (d->restart_interval && since_restart == d->restart_interval) {
                                if (restart(d, sc, ns, &eobrun, rst++) != 0) return -1;
                                since_restart = 0;
                        }
                        for (int i = 0; i < ns; i++) {
                                struct comp *c = sc[i];
                                const int bh = interleaved ? c->h : 1, bv = interleaved ? c->v : 1;
                                for (int y = 0; y < bv; y++)
                                        for (int x = 0; x < bh; x++) {
                                                const unsigned bx = interleaved ? mx * c->h + x : mx;
                                                const unsigned by = interleaved ? my * c->v + y : my;
                                                int16_t *b = c->blk + ((size_t)by * c->pwb + bx) * 64;
                                                if (!d->progressive) block_sequential(d, c, b);
                                                else if (ss == 0) { if (ah == 0) block_dc_first(d, c, b, al); else block_dc_refine(d, b, al); }
                                                else if (ah == 0) block_ac_first(d, c, b, ss, se, al, &eobrun);
                                                else block_ac_refine(d, c, b, ss, se, al, &eobrun);
                                                if (d->failed) return -1;
                                        }
                        }
                        since_restart++;
                }
        /* leave d->p at the next marker */
        while (d->p + 1 < d->end && !(d->p[0] == 0xFF && d->p[1] != 0x00 && !(d->p[1] >= 0xD0 && d->p[1] <= 0xD7) && d->p[1] != 0xFF)) d->p++;
        return 0;
}

/* ---- marker segments ----------------------------------------------------------------------- */
static unsigned be16(const uint8_t *p) { return ((unsigned)p[0] << 8) | p[1]; }

static int parse_dqt(struct dec *d, const uint8_t *s, unsigned len) {
        while (len > 0) {
                const int pq = s[0] >> 4, tq = s[0] & 15;
                if (tq > 3 || pq > 1) return fail(d, "corrupt jpeg: bad DQT");
                const unsigned need = 1 + 64 * (pq + 1);
                if (len < need) return fail(d, "corrupt jpeg: short DQT");
                for (int k = 0; k < 64; k++) d->qt[tq][ZZ[k]] = pq ? (uint16_t)be16(s + 1 + 2 * k) : s[1 + k];
                d->qt_present[tq] = 1;
                s += need;
                len -= need;
        }
        return 0;
}
static int parse_dht(struct dec *d, const uint8_t *s, unsigned len) {
        while (len > 0) {
                if (len < 17) return fail(d, "corrupt jpeg: short DHT");
                const int tc = s[0] >> 4, th = s[0] & 15;
                if (tc > 1 || th > 3) return fail(d, "corrupt jpeg: bad DHT");
                struct huff *h = tc ? &d->ac[th] : &d->dc[th];
                unsigned n = 0;
                h->bits[0] = 0;
                for (int l = 1; l <= 16; l++) { h->bits[l] = s[l]; n += s[l]; }
                if (n > 256 || len < 17 + n) return fail(d, "corrupt jpeg: bad DHT");
                memcpy(h->vals, s + 17, n);
                if (build_huff(d, h) != 0) return -1;
                s += 17 + n;
                len -= 17 + n;
        }
        return 0;
}
static int parse_sof(struct dec *d, const uint8_t *s, unsigned len) {
        if (len < 6) return fail(d, "corrupt jpeg: short SOF");
        if (s[0] != 8) return fail(d, "unsupported jpeg: %d-bit samples", s[0]);
        d->H = be16(s + 1);
        d->W = be16(s + 3);
        d->ncomp = s[5];
        if (d->ncomp != 3) return fail(d, "only 3 component jpegs are supported");            /* jpeg.c:34 */
        if (d->W == 0 || d->H == 0) return fail(d, "unsupported jpeg: empty image or DNL-defined height");
        if (len < 6 + 3u * d->ncomp) return fail(d, "corrupt jpeg: short SOF");
        d->maxh = d->maxv = 1;
        for (int i = 0; i < 3; i++) {
                struct comp *c = &d->c[i];
                c->id = s[6 + 3 * i];
                c->h = s[7 + 3 * i] >> 4;
                c->v = s[7 + 3 * i] & 15;
                c->tq = s[8 + 3 * i];
                if (c->h < 1 || c->h > 4 || c->v < 1 || c->v > 4 || c->tq > 3) return fail(d, "corrupt jpeg: bad component spec");
                if ((unsigned)c->h > d->maxh) d->maxh = c->h;
                if ((unsigned)c->v > d->maxv) d->maxv = c->v;
        }
        d->mcux = (d->W + 8 * d->maxh - 1) / (8 * d->maxh);
        d->mcuy = (d->H + 8 * d->maxv - 1) / (8 * d->maxv);
        for (int i = 0; i < 3; i++) {
                struct comp *c = &d->c[i];
                const unsigned cw = (d->W * c->h + d->maxh - 1) / d->maxh, ch = (d->H * c->v + d->maxv - 1) / d->maxv;   /* A.1.1 */
                c->wb = (cw + 7) / 8;
                c->hb = (ch + 7) / 8;
                c->pwb = d->mcux * c->h;
                c->phb = d->mcuy * c->v;
                c->blk = calloc((size_t)c->pwb * c->phb * 64, sizeof(int16_t));
                if (!c->blk) return fail(d, "could not allocate memory for coefs");                /* jpeg.c:69 */
        }
        return 0;
}

int j2p_read_jpeg_mem(const uint8_t *buf, size_t len, struct j2p_jpeg *out, char *err, size_t errlen) {
        struct dec *d = calloc(1, sizeof *d);
        if (!d) return -1;
        d->p = buf; d->end = buf + len; d->err = err; d->errlen = errlen;
        if (err && errlen) err[0] = 0;
        memset(out, 0, sizeof *out);
        int have_sof = 0, done = 0;
        if (len < 4 || buf[0] != 0xFF || buf[1] != 0xD8) { fail(d, "not a jpeg file (no SOI marker)"); goto out; }
        d->p += 2;
        while (!done && !d->failed) {
                /* find next marker */
                while (d->p < d->end && *d->p != 0xFF) d->p++;
                while (d->p < d->end && *d->p == 0xFF) d->p++;
                if (d->p >= d->end) { fail(d, "corrupt jpeg: premature end of file"); break; }
                const unsigned m = *d->p++;
                if (m == 0xD9) { done = 1; break; }                       /* EOI */
                if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;      /* TEM, stray RSTn */
                if (d->p + 2 > d->end) { fail(d, "corrupt jpeg: truncated marker"); break; }
                const unsigned seglen = be16(d->p);
                if (seglen < 2 || d->p + seglen > d->end) { fail(d, "corrupt jpeg: bad segment length"); break; }
                const uint8_t *s = d->p + 2;
                const unsigned sl = seglen - 2;
                d->p += seglen;
                if (m == 0xDB) parse_dqt(d, s, sl);
                else if (m == 0xC4) parse_dht(d, s, sl);
                else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                        if (have_sof) { fail(d, "unsupported jpeg: multiple frames"); break; }
                        d->progressive = m == 0xC2;
                        if (parse_sof(d, s, sl) == 0) have_sof = 1;
                } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
                        fail(d, "unsupported jpeg: SOF%u (arithmetic, lossless or hierarchical coding)", m - 0xC0);
                } else if (m == 0xDD) {
                        if (sl < 2) fail(d, "corrupt jpeg: short DRI"); else d->restart_interval = be16(s);
                } else if (m == 0xDA) {
                        if (!have_sof) { fail(d, "corrupt jpeg: scan before frame header"); break; }
                        if (sl < 1) { fail(d, "corrupt jpeg: short SOS"); break; }
                        const int ns = s[0];
                        if (ns < 1 || ns > 3 || sl < 1 + 2u * ns + 3) { fail(d, "corrupt jpeg: bad SOS"); break; }
                        struct comp *sc[3];
                        for (int i = 0; i < ns; i++) {
                                sc[i] = NULL;
                                for (int k = 0; k < 3; k++) if (d->c[k].id == s[1 + 2 * i]) sc[i] = &d->c[k];
                                if (!sc[i]) { fail(d, "corrupt jpeg: scan names an unknown component"); break; }
                                sc[i]->td = s[2 + 2 * i] >> 4;
                                sc[i]->ta = s[2 + 2 * i] & 15;
                                if (sc[i]->td > 3 || sc[i]->ta > 3) fail(d, "corrupt jpeg: bad table selector");
                        }
                        if (d->failed) break;
                        int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
                        if (!d->progressive) { ss = 0; se = 63; ah = al = 0; }
                        else if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) { fail(d, "corrupt jpeg: bad progressive scan parameters"); break; }
                        for (int i = 0; i < ns; i++) {
                                if ((ss == 0 && !(d->progressive && ah) && !d->dc[sc[i]->td].present) ||
                                    ((!d->progressive || ss > 0) && !d->ac[sc[i]->ta].present)) { fail(d, "corrupt jpeg: scan uses an undefined huffman table"); break; }
                        }
                        if (d->failed) break;
                        decode_scan(d, sc, ns, ss, se, ah, al);
                }
                /* everything else (APPn, COM, DNL, ...) is skipped */
        }
        if (!d->failed && !have_sof) fail(d, "corrupt jpeg: no frame header");
        if (!d->failed) {
                out->w = d->W;
                out->h = d->H;
                for (int i = 0; i < 3 && !d->failed; i++) {
                        struct comp *c = &d->c[i];
                        struct coef *o = &out->coefs[i];
                        if (!d->qt_present[c->tq]) { fail(d, "weird jpeg: no quant table pointer"); break; }          /* jpeg.c:40 */
                        for (int j = 0; j < 64; j++) {
                                if (d->qt[c->tq][j] == 0) { fail(d, "invalid quantization table"); break; }            /* jpeg.c:43 */
                                o->quant_table[j] = d->qt[c->tq][j];
                        }
                        if (d->failed) break;
                        o->w = c->wb * 8;
                        o->h = c->hb * 8;
                        o->w_samp = d->maxh / c->h;                                                                    /* jpeg.c:57-58 */
                        o->h_samp = d->maxv / c->v;
                        if (o->h / 8 != (d->H / o->h_samp + 7) / 8) { fail(d, "jpeg invalid coef h size"); break; }    /* jpeg.c:59-61 */
                        if (o->w / 8 != (d->W / o->w_samp + 7) / 8) { fail(d, "jpeg invalid coef w size"); break; }    /* jpeg.c:62-64 */
                        o->data = malloc((size_t)o->w * o->h * sizeof(int16_t));
                        if (!o->data) { fail(d, "could not allocate memory for coefs"); break; }
                        for (unsigned by = 0; by < c->hb; by++)
                                memcpy(o->data + (size_t)by * c->wb * 64, c->blk + (size_t)by * c->pwb * 64, (size_t)c->wb * 64 * sizeof(int16_t));
                }
        }
out:
        for (int i = 0; i < 3; i++) free(d->c[i].blk);
        const int rc = d->failed ? -1 : 0;
        if (rc != 0) for (int i = 0; i < 3; i++) { free(out->coefs[i].data); out->coefs[i].data = NULL; }
        free(d);
        return rc;
}
