// kernels_epilogue.cu — what the reference does to the solver's result before libpng sees it,
// as one device pass over a joint (3-plane) session:
//   luma += 128                                   (jpeg2png.c:156-159)
//   YCbCr -> RGB in double, clamp to [0, 255], scale by (1 << bits) / 256, TRUNCATE
//                                                 (png.c:39-47, clamp: utils.h CLAMP via png.c:15-17)
//   8-bit samples, or 16-bit big-endian           (png.c:51-62)
// The output is the image as PNG scanlines — every row prefixed with filter type 0 — so the host
// only has to deflate it: 3 (or 6) bytes per pixel cross PCIe instead of 12.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"

namespace j2p {

constexpr int EP_NT = 256;

// png.c:15-17 + :44-46: the double expression is narrowed to float by the call to clamp(), compared
// with the double bounds 0. and 255., the (float) result multiplied by the float bitfactor and
// converted to unsigned (truncation)
__device__ __forceinline__ unsigned to_sample(double v, float bitfactor) {
    float x = __double2float_rn(v);
    x = (double)x > 255. ? 255.f : ((double)x < 0. ? 0.f : x);
    return __float2uint_rz(__fmul_rn(x, bitfactor));
}

// one CTA = EP_NT consecutive pixels of one row; samples are staged in shared memory so that the
// byte stream of the row (which starts at an odd address: filter byte first) is written with
// consecutive threads on consecutive bytes
__global__ void __launch_bounds__(EP_NT) k_scanlines(const float *Y, const float *Cb, const float *Cr, int W, int w, int h, int bits, uint8_t *out) {
    __shared__ uint8_t sm[EP_NT * 6];
    const int row = blockIdx.y, x0 = blockIdx.x * EP_NT, tid = threadIdx.x;
    const int depth = bits >> 3, px = x0 + tid;
    const float bitfactor = bits == 8 ? 1.0f : 256.0f;                        // (1 << bits) / 256.
    if (px < w) {
        const size_t gi = (size_t)row * W + px;
        const float yi = __fadd_rn(Y[gi], 128.f);                             // jpeg2png.c:158
        const double dy = (double)yi, dcb = (double)Cb[gi], dcr = (double)Cr[gi];
        const unsigned r = to_sample(__dadd_rn(dy, __dmul_rn(1.402, dcr)), bitfactor);                                        // png.c:44
        const unsigned g = to_sample(__dsub_rn(__dsub_rn(dy, __dmul_rn(0.34414, dcb)), __dmul_rn(0.71414, dcr)), bitfactor);  // png.c:45
        const unsigned b = to_sample(__dadd_rn(dy, __dmul_rn(1.772, dcb)), bitfactor);                                        // png.c:46
        uint8_t *p = sm + tid * 3 * depth;
        if (depth == 1) {
            p[0] = (uint8_t)r; p[1] = (uint8_t)g; p[2] = (uint8_t)b;
        } else {
            p[0] = (uint8_t)(r >> 8); p[1] = (uint8_t)r; p[2] = (uint8_t)(g >> 8); p[3] = (uint8_t)g; p[4] = (uint8_t)(b >> 8); p[5] = (uint8_t)b;
        }
    }
    __syncthreads();
    const size_t stride = (size_t)w * 3 * depth + 1;
    uint8_t *dst = out + (size_t)row * stride;
    if (x0 == 0 && tid == 0) dst[0] = 0;                                      // PNG filter type 0 (None)
    const int npx = min(EP_NT, w - x0), nbytes = npx * 3 * depth;
    dst += 1 + (size_t)x0 * 3 * depth;
    for (int i = tid; i < nbytes; i += EP_NT) dst[i] = sm[i];
}

cudaError_t launch_scanlines(const float *Y, const float *Cb, const float *Cr, int W, int w, int h, int bits, uint8_t *out, cudaStream_t s) {
    const dim3 grid((w + EP_NT - 1) / EP_NT, h);
    k_scanlines<<<grid, EP_NT, 0, s>>>(Y, Cb, Cr, W, w, h, bits, out);
    return cudaGetLastError();
}

}  // namespace j2p
