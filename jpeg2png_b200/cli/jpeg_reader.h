/* jpeg_reader.h — JPEG file -> quantised DCT coefficients, without libjpeg.
 *
 * Replaces reference jpeg.c:22-80 (read_jpeg, built on libjpeg's jpeg_read_coefficients, which is
 * not available on the build box): same output contract — three `struct coef` with
 *   w,h            = width_in_blocks*8, height_in_blocks*8 of the component, NOT padded to whole
 *                    MCUs (jpeg.c:52-53)
 *   w_samp,h_samp  = max_h/h_i, max_v/v_i (jpeg.c:57-58)
 *   data           = int16 [blocks][64], natural (row-major) order, blocks in raster order
 *   quant_table    = natural order (jpeg.c:46)
 * and the same rejections: not exactly 3 components (jpeg.c:34), a zero quantisation entry
 * (jpeg.c:41-45), component sizes that do not match the sampling factors (jpeg.c:59-64).
 *
 * Supported: baseline and extended sequential Huffman (SOF0, SOF1), progressive Huffman (SOF2),
 * 8-bit precision, restart intervals, interleaved and non-interleaved scans.  Not supported:
 * arithmetic coding, lossless, hierarchical, 12-bit.
 */
#ifndef J2P_JPEG_READER_H
#define J2P_JPEG_READER_H

#include <stddef.h>
#include <stdint.h>

#include "../../include/jpeg2png_b200.h"

struct j2p_jpeg {
        unsigned w, h;              /* image size in pixels */
        struct coef coefs[3];       /* data malloc'd (caller frees), fdata NULL */
};

/* Returns 0 on success; on failure returns non-zero and writes a message into err (if non-NULL).
 * The messages for the reference's own rejections are the reference's (jpeg.c:34,43,60,63). */
int j2p_read_jpeg_mem(const uint8_t *buf, size_t len, struct j2p_jpeg *out, char *err, size_t errlen);

#endif
