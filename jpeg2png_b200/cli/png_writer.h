/* png_writer.h — see png_writer.c.  Replaces reference png.h:7 / png.c:20-78. */
#ifndef J2P_PNG_WRITER_H
#define J2P_PNG_WRITER_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

/* YCbCr (y already carries the +128 of jpeg2png.c:156-159) -> interleaved RGB samples,
 * 3 bytes (bits == 8) or 6 bytes big-endian (bits == 16) per pixel; png.c:39-62. */
void j2p_ycc_to_rgb(unsigned w, unsigned h, unsigned bits, const float *y, unsigned y_stride, const float *cb,
                    unsigned cb_stride, const float *cr, unsigned cr_stride, uint8_t *out, size_t out_stride);

/* Writes a w x h truecolour PNG of `bits` bits per sample.  Returns 0 on success. */
int j2p_write_png(FILE *out, unsigned w, unsigned h, unsigned bits, const float *y, unsigned y_stride, const float *cb,
                  unsigned cb_stride, const float *cr, unsigned cr_stride);

/* Same container from unfiltered scanlines (h rows of 1 + w*3*bits/8 bytes, filter byte 0 first):
 * what j2p_session_download_scanlines() delivers from the device.  `raw` is scratch: large images
 * are filtered in place before they are deflated. */
int j2p_write_png_scanlines(FILE *out, unsigned w, unsigned h, unsigned bits, uint8_t *raw);

#endif
