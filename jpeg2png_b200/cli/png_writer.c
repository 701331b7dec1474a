/* png_writer.c — YCbCr float planes -> RGB PNG, without libpng (zlib only).
 *
 * Replaces reference png.c:20-78.  The colour conversion and quantisation to integer samples
 * restate png.c:39-62 exactly (that is where "bit-identical PNG" is decided): double-precision
 * YCbCr->RGB, clamp to [0,255] in double narrowed to float, scale by (1<<bits)/256 in float,
 * TRUNCATE to unsigned; 16-bit samples big-endian.  Each plane is indexed with its own stride
 * (png.c:39-41).  The container is a plain non-interlaced truecolour PNG: filter type 0 on every
 * row, one zlib stream; pixels, not file bytes, are what must match a libpng-written file.
 */
#include "png_writer.h"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* png.c:15-17: the argument is narrowed to float by the call, compared against double bounds */
static float clamp255(float x) { return (float)((double)x > 255. ? 255. : ((double)x < 0. ? 0. : (double)x)); }

void j2p_ycc_to_rgb(unsigned w, unsigned h, unsigned bits, const float *y, unsigned ys, const float *cb, unsigned cbs,
                    const float *cr, unsigned crs, uint8_t *out, size_t out_stride) {
        const unsigned depth = bits / 8;
        const float bitfactor = (float)((double)(1 << bits) / 256.);                             /* png.c:43 */
#pragma omp parallel for schedule(static)
        for (unsigned i = 0; i < h; i++) {
                uint8_t *row = out + (size_t)i * out_stride;
                for (unsigned j = 0; j < w; j++) {
                        const float yi = y[(size_t)i * ys + j], cbi = cb[(size_t)i * cbs + j], cri = cr[(size_t)i * crs + j];
                        const unsigned r = (unsigned)(clamp255((double)yi + 1.402 * (double)cri) * bitfactor);                              /* png.c:44 */
                        const unsigned g = (unsigned)(clamp255(((double)yi - 0.34414 * (double)cbi) - 0.71414 * (double)cri) * bitfactor);  /* png.c:45 */
                        const unsigned b = (unsigned)(clamp255((double)yi + 1.772 * (double)cbi) * bitfactor);                              /* png.c:46 */
                        uint8_t *px = row + (size_t)j * 3 * depth;
                        if (bits == 8) {
                                px[0] = r & 0xFF; px[1] = g & 0xFF; px[2] = b & 0xFF;
                        } else {
                                px[0] = (r >> 8) & 0xFF; px[1] = r & 0xFF;
                                px[2] = (g >> 8) & 0xFF; px[3] = g & 0xFF;
                                px[4] = (b >> 8) & 0xFF; px[5] = b & 0xFF;
                        }
                }
        }
}

static int put_chunk(FILE *f, const char *type, const uint8_t *data, size_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1],
                          (uint8_t)type[2], (uint8_t)type[3]};
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data, (uInt)len);
        const uint8_t tail[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        return fwrite(hdr, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4 ? 0 : -1;
}

/* One zlib stream made of independently deflated pieces (the pigz construction): every piece but
 * the last ends with Z_SYNC_FLUSH — an empty stored block that pads to a byte boundary and leaves
 * the stream open — the last with Z_FINISH; the 2-byte zlib header goes in front and the Adler-32
 * of the whole input (combined from the per-piece checksums) behind.  Pieces cost ~1 % of
 * compression ratio (no history across them) and make an 8K frame's 100 MB of RGB a job for all
 * host threads instead of several seconds on one.  Inside the file-parallel loop of the command
 * line OpenMP runs the inner loop serially (nested parallelism is off), which is what we want. */
#define J2P_PIECE ((size_t)1 << 20)
#define J2P_FAIL() do { _Pragma("omp atomic write") failed = 1; } while (0)

/* Compression effort: the reference leaves libpng at zlib level 6 and its default filter heuristics.
 * What must match is the pixels, not the file bytes, and on a GPU the deflate of a multi-megapixel
 * image costs several times the solve it follows (profiles/r01_cli_batch.txt: 170-290 ms per 1080p
 * file against 30 ms).  Images of a megapixel or more are therefore written with the Up filter
 * (type 2: every byte minus the byte above it — smooth images leave small residuals) and deflated at
 * level 1 with the run-length strategy, which suits such residuals: on a 1080p frame 72 ms and
 * 2.9 MB against 166 ms and 4.9 MB for unfiltered scanlines at level 1 (183 ms, 5.0 MB at level 6).
 * Small images keep unfiltered scanlines at level 6. */
static int big_image(size_t raw_bytes) { return raw_bytes >= (size_t)3 << 20; }
static int deflate_level(size_t raw_bytes) { return big_image(raw_bytes) ? 1 : 6; }
static int deflate_strategy(size_t raw_bytes) { return big_image(raw_bytes) ? Z_RLE : Z_DEFAULT_STRATEGY; }

/* Up filter in place.  Bottom-up, so the row above is still unfiltered when it is subtracted; row 0
 * keeps filter type 0 (its "row above" is all zeros, Up would leave it unchanged anyway). */
static void filter_up_in_place(uint8_t *raw, size_t stride, unsigned h) {
        for (unsigned y = h; y-- > 1;) {
                uint8_t *cur = raw + (size_t)y * stride;
                const uint8_t *up = cur - stride;
                cur[0] = 2;
                for (size_t i = 1; i < stride; i++) cur[i] = (uint8_t)(cur[i] - up[i]);
        }
}

static uint8_t *deflate_pieces(const uint8_t *raw, size_t len, size_t *zlen) {
        const int level = deflate_level(len);
        const size_t np = (len + J2P_PIECE - 1) / J2P_PIECE;
        uint8_t **buf = calloc(np, sizeof *buf);
        size_t *blen = calloc(np, sizeof *blen);
        uLong *adl = calloc(np, sizeof *adl);
        if (!buf || !blen || !adl) { free(buf); free(blen); free(adl); return NULL; }
        int failed = 0;
#pragma omp parallel for schedule(dynamic)
        for (long i = 0; i < (long)np; i++) {
                const size_t off = (size_t)i * J2P_PIECE, n = len - off < J2P_PIECE ? len - off : J2P_PIECE;
                z_stream zs;
                memset(&zs, 0, sizeof zs);
                if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, deflate_strategy(len)) != Z_OK) { J2P_FAIL(); continue; }
                const size_t cap = deflateBound(&zs, (uLong)n) + 64;
                buf[i] = malloc(cap);
                if (buf[i]) {
                        zs.next_in = (Bytef *)(raw + off);
                        zs.avail_in = (uInt)n;
                        zs.next_out = buf[i];
                        zs.avail_out = (uInt)cap;
                        const int last = (size_t)i + 1 == np;
                        const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
                        if ((last ? rc != Z_STREAM_END : rc != Z_OK) || zs.avail_in != 0 || zs.avail_out == 0) J2P_FAIL();
                        blen[i] = cap - zs.avail_out;
                        adl[i] = adler32(adler32(0L, Z_NULL, 0), raw + off, (uInt)n);
                } else {
                        J2P_FAIL();
                }
                deflateEnd(&zs);
        }
        uint8_t *z = NULL;
        if (!failed) {
                size_t total = 2 + 4;
                for (size_t i = 0; i < np; i++) total += blen[i];
                z = malloc(total);
                if (z) {
                        size_t pos = 0;
                        z[pos++] = 0x78;                                 /* deflate, 32 KB window */
                        z[pos++] = 0x9C;                                 /* default compression, check bits */
                        uLong a = adler32(0L, Z_NULL, 0);
                        for (size_t i = 0; i < np; i++) {
                                memcpy(z + pos, buf[i], blen[i]);
                                pos += blen[i];
                                const size_t off = i * J2P_PIECE, n = len - off < J2P_PIECE ? len - off : J2P_PIECE;
                                a = adler32_combine(a, adl[i], (z_off_t)n);
                        }
                        z[pos++] = (uint8_t)(a >> 24); z[pos++] = (uint8_t)(a >> 16); z[pos++] = (uint8_t)(a >> 8); z[pos++] = (uint8_t)a;
                        *zlen = pos;
                }
        }
        for (size_t i = 0; i < np; i++) free(buf[i]);
        free(buf); free(blen); free(adl);
        return z;
}

/* `raw`: h UNFILTERED scanlines of 1 + w*3*bits/8 bytes, each starting with filter-type byte 0.
 * Large images are filtered in place (the buffer is the caller's scratch). */
int j2p_write_png_scanlines(FILE *out, unsigned w, unsigned h, unsigned bits, uint8_t *raw) {
        if (bits != 8 && bits != 16) return -1;
        const size_t stride = (size_t)w * 3 * (bits / 8) + 1;
        size_t zlen = 0;
        uint8_t *z = NULL;
        if (big_image(stride * h)) filter_up_in_place(raw, stride, h);
        if (stride * h > 2 * J2P_PIECE) {
                z = deflate_pieces(raw, stride * h, &zlen);
        } else {
                uLongf zl = compressBound((uLong)(stride * h));
                z = malloc(zl);
                if (z && compress2(z, &zl, raw, (uLong)(stride * h), deflate_level(stride * h)) == Z_OK) zlen = zl;
                else { free(z); z = NULL; }
        }
        int rc = -1;
        if (z) {
                static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
                const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16),
                                          (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)bits, 2 /* truecolour */, 0, 0, 0};
                if (fwrite(sig, 1, 8, out) == 8 && put_chunk(out, "IHDR", ihdr, 13) == 0 && put_chunk(out, "IDAT", z, zlen) == 0 &&
                    put_chunk(out, "IEND", NULL, 0) == 0)
                        rc = 0;
        }
        free(z);
        return rc;
}

int j2p_write_png(FILE *out, unsigned w, unsigned h, unsigned bits, const float *y, unsigned ys, const float *cb, unsigned cbs,
                  const float *cr, unsigned crs) {
        if (bits != 8 && bits != 16) return -1;
        const size_t row = (size_t)w * 3 * (bits / 8), stride = row + 1;      /* +1: filter-type byte */
        uint8_t *raw = malloc(stride * h);
        if (!raw) return -1;
        for (unsigned i = 0; i < h; i++) raw[(size_t)i * stride] = 0;          /* filter 0 (None) */
        j2p_ycc_to_rgb(w, h, bits, y, ys, cb, cbs, cr, crs, raw + 1, stride);
        const int rc = j2p_write_png_scanlines(out, w, h, bits, raw);
        free(raw);
        return rc;
}
