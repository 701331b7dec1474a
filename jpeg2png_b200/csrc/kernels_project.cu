// kernels_project.cu — step + projection kernel, conventional decode, aux_init.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"

namespace j2p {

// ------------------------------------------------------------------------------------------
// 8x8 block transposes among the 8 lanes that own one block (lane j holds row j).
// tile: 8 rows x 8 floats; element (r, c) lives at r*8 + (c ^ (((r>>2)&1)<<2)) — the two float4
// halves of rows 4..7 are swapped, which makes both the 128-bit row accesses and the scalar
// column accesses bank-conflict free (tiles of the four blocks of a warp are 72 floats apart).
// ------------------------------------------------------------------------------------------
constexpr int TILE_STRIDE = 72;

__device__ __forceinline__ void rows_to_cols(float (&v)[8], float *tile, int j) {
    const int h = (j >> 2) & 1;
    float4 *row = reinterpret_cast<float4 *>(tile + j * 8);
    row[h] = make_float4(v[0], v[1], v[2], v[3]);
    row[h ^ 1] = make_float4(v[4], v[5], v[6], v[7]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))];
    __syncwarp();
}
__device__ __forceinline__ void cols_to_rows(float (&v)[8], float *tile, int j) {
#pragma unroll
    for (int i = 0; i < 8; i++) tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))] = v[i];
    __syncwarp();
    const int h = (j >> 2) & 1;
    const float4 *row = reinterpret_cast<const float4 *>(tile + j * 8);
    const float4 lo = row[h], hi = row[h ^ 1];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    __syncwarp();
}

// 2-D transforms for a thread that holds row j of the block and ends holding row j.
// Vertical pass first, horizontal second (ooura/dct.c:39-94, :103-158).
__device__ __forceinline__ void fdct8x8_rows(float (&v)[8], float *tile, int j) {
    rows_to_cols(v, tile, j);
    fdct8(v);
    cols_to_rows(v, tile, j);
    fdct8(v);
}
__device__ __forceinline__ void idct8x8_rows(float (&v)[8], float *tile, int j) {
    rows_to_cols(v, tile, j);
    idct8(v);
    cols_to_rows(v, tile, j);
    idct8(v);
}

// ------------------------------------------------------------------------------------------
// k_project — 8 threads per coefficient block (thread j owns row j), 32 blocks per CTA.
// Template <SW, SH>: compile-time sampling factors of the plane (float4 I/O, stepped values of
// the whole footprint kept in registers); SW == 0 selects the run-time generic path.
// ------------------------------------------------------------------------------------------
constexpr int P_NT = 256, P_BW = 8, P_BH = 4;   // CTA tile: 8 x 4 coefficient blocks

struct ProjPlane {
    int c;          // plane index
    int gx;         // CTAs per row
};

// the stepped point at one frame pixel: y = x + f (x - xp), then y - step * (g / norm)
// (compute.c:436, :213).  `rn` = RN(1/norm) from k_gradient's last CTA.
struct Stepper {
    float factor, step, norm, rn;
    bool stepping;
    // IEEE division (generic / fallback paths)
    __device__ __forceinline__ float operator()(float x, float xp, float g) const {
        float y = fadd(x, fmul(factor, fsub(x, xp)));
        if (stepping) y = fsub(y, fmul(step, fdiv(g, norm)));
        return y;
    }
    // shared-reciprocal division; `key` collects the guard of numerics.cuh (smallest non-zero |g|)
    __device__ __forceinline__ float fast(float x, float xp, float g, unsigned &key) const {
        float y = fadd(x, fmul(factor, fsub(x, xp)));
        if (stepping) {
            key = min(key, qdiv_key(g));
            y = fsub(y, fmul(step, qdiv_core(g, norm, rn)));
        }
        return y;
    }
};

template <int SW, int SH>
__global__ void __launch_bounds__(P_NT, (SW * SH <= 1) ? 4 : 2) k_project(const __grid_constant__ FrameDev F, const ProjPlane G, const float factor) {
    __shared__ __align__(16) float tiles[2][P_NT / 8][TILE_STRIDE];
    __shared__ __align__(16) float sq[3][64];          // q, q*q, RN(1/(q*q)) of this plane
    const int tid = threadIdx.x;
    const int c = G.c;
    const PlaneDev &P = F.pl[c];
    if (tid < 64) {
        sq[0][tid] = F.q[c][tid];
        sq[1][tid] = F.qq[c][tid];
        sq[2][tid] = F.rqq[c][tid];
    }
    __syncthreads();
    const int ctay = blockIdx.x / G.gx, ctax = blockIdx.x - ctay * G.gx;
    const int W = F.W, H = F.H;
    const int b = tid >> 3, j = tid & 7;
    const int bx = ctax * P_BW + (b & (P_BW - 1)), by = ctay * P_BH + (b >> 3);
    const bool real = bx < (P.cw >> 3) && by < (P.ch >> 3);
    const int sw = SW ? SW : P.sw, sh = SW ? SH : P.sh;
    Stepper stepper;
    stepper.factor = factor;
    stepper.step = F.step;
    stepper.norm = F.norms[c];
    stepper.rn = F.norms[4 + c];
    stepper.stepping = stepper.norm != 0.f;                        // compute.c:211
    const bool norm_ok = qdiv_divisor_ok(stepper.norm);
    float *tileA = tiles[0][b], *tileB = tiles[1][b];
    const int cy = by * 8 + j;

    if (!real) {
        // pixels of the frame that no coefficient block covers: step only (compute.c:349-350 never visits them)
        for (int i = 0; i < 8; i++)
            for (int sy = 0; sy < sh; sy++)
                for (int sx = 0; sx < sw; sx++) {
                    const int px = (bx * 8 + i) * sw + sx, py = cy * sh + sy;
                    if (px < W && py < H) {
                        const size_t gi = (size_t)py * W + px;
                        P.xp[gi] = stepper(P.x[gi], P.xp[gi], P.g[gi]);
                    }
                }
        return;   // whole 8-lane groups leave together; the remaining lanes still __syncwarp among themselves
    }

    // quantised coefficients of this row: issued first so the latency hides behind the pixel loads
    const int4 draw = __ldg(reinterpret_cast<const int4 *>(P.data + ((size_t)(by * (P.cw >> 3) + bx) * 64 + j * 8)));

    // ---- stepped point of the footprint, block-row means (compute.c:348-370) ------------------
    constexpr int ZW = SW ? SW * 8 : 1, ZH = SW ? SH : 1;
    float z[ZH][ZW];
    float v[8], mean[8];
    if constexpr (SW > 0) {
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int sy = 0; sy < SH; sy++) {
            const size_t gi = (size_t)(cy * SH + sy) * W + (size_t)bx * 8 * SW;
            const float4 *xr = reinterpret_cast<const float4 *>(P.x + gi);
            const float4 *pr = reinterpret_cast<const float4 *>(P.xp + gi);
            const float4 *gr = reinterpret_cast<const float4 *>(P.g + gi);
#pragma unroll
            for (int k = 0; k < SW * 2; k++) {
                const float4 a = xr[k], p = pr[k], g = gr[k];
                z[sy][k * 4 + 0] = stepper.fast(a.x, p.x, g.x, key);
                z[sy][k * 4 + 1] = stepper.fast(a.y, p.y, g.y, key);
                z[sy][k * 4 + 2] = stepper.fast(a.z, p.z, g.z, key);
                z[sy][k * 4 + 3] = stepper.fast(a.w, p.w, g.w, key);
            }
        }
        if (stepper.stepping && !(norm_ok && key >= QDIV_KEY_MIN)) {   // outside the proven range: IEEE division
#pragma unroll
            for (int sy = 0; sy < SH; sy++) {
                const size_t gi = (size_t)(cy * SH + sy) * W + (size_t)bx * 8 * SW;
#pragma unroll
                for (int k = 0; k < SW * 8; k++) z[sy][k] = stepper(P.x[gi + k], P.xp[gi + k], P.g[gi + k]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (SW * SH > 1 || P.resample) {
                float m = 0.f;                                               // compute.c:351
#pragma unroll
                for (int sy = 0; sy < SH; sy++)
#pragma unroll
                    for (int sx = 0; sx < SW; sx++) m = fadd(m, z[sy][i * SW + sx]);
                constexpr int CNT = SW * SH;
                if constexpr ((CNT & (CNT - 1)) == 0) m = fmul(m, 1.0f / CNT);   // exact: power-of-two divisor
                else m = fdiv(m, (float)CNT);                                // compute.c:359
                mean[i] = m;
                v[i] = m;
            } else {
                mean[i] = 0.f;
                v[i] = z[0][i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int cx = bx * 8 + i;
            if (P.resample) {
                float m = 0.f;
                for (int sy = 0; sy < sh; sy++)
                    for (int sx = 0; sx < sw; sx++) {
                        const size_t gi = (size_t)(cy * sh + sy) * W + cx * sw + sx;
                        m = fadd(m, stepper(P.x[gi], P.xp[gi], P.g[gi]));
                    }
                m = fdiv(m, P.cnt);
                mean[i] = m;
                v[i] = m;
            } else {
                const size_t gi = (size_t)cy * W + cx;
                mean[i] = 0.f;
                v[i] = stepper(P.x[gi], P.xp[gi], P.g[gi]);
            }
        }
    }

    fdct8x8_rows(v, tileA, j);

    // ---- clamp to the quantisation interval (compute.c:323-331); DCT-distance residual ---------
    const int dw[4] = {draw.x, draw.y, draw.z, draw.w};
    float qv[8], qqv[8], rqv[8];
    {
        const float4 *t0 = reinterpret_cast<const float4 *>(&sq[0][j * 8]);
        const float4 *t1 = reinterpret_cast<const float4 *>(&sq[1][j * 8]);
        const float4 *t2 = reinterpret_cast<const float4 *>(&sq[2][j * 8]);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float4 a = t0[k], bq = t1[k], cq = t2[k];
            qv[k * 4] = a.x; qv[k * 4 + 1] = a.y; qv[k * 4 + 2] = a.z; qv[k * 4 + 3] = a.w;
            qqv[k * 4] = bq.x; qqv[k * 4 + 1] = bq.y; qqv[k * 4 + 2] = bq.z; qqv[k * 4 + 3] = bq.w;
            rqv[k * 4] = cq.x; rqv[k * 4 + 1] = cq.y; rqv[k * 4 + 2] = cq.z; rqv[k * 4 + 3] = cq.w;
        }
    }
    float r[8], num[8];
    unsigned rkey = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int di = (i & 1) ? (dw[i >> 1] >> 16) : (int)(short)(dw[i >> 1] & 0xffff);
        const float d = (float)di;
        const float q = qv[i];
        const float lo = fmul(fsub(d, 0.5f), q), hi = fmul(fadd(d, 0.5f), q);
        float t = v[i];
        t = t > hi ? hi : (t < lo ? lo : t);
        v[i] = t;
        num[i] = fsub(t, fmul(d, q));                                        // compute.c:47
        rkey = min(rkey, qdiv_key(num[i]));
        r[i] = qdiv_core(num[i], qqv[i], rqv[i]);                            // compute.c:49; q*q in [1, 2^32] is always a valid divisor
    }
    if (rkey < QDIV_KEY_MIN) {                                               // a residual below 2^-60: IEEE division
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = fdiv(num[i], qqv[i]);
    }

    idct8x8_rows(v, tileA, j);
    if (P.use_prob) {
        idct8x8_rows(r, tileB, j);
        float4 *gprow = reinterpret_cast<float4 *>(P.gp + (size_t)cy * P.cw + bx * 8);
        const float pa = P.p_alpha;                                          // compute.c:62 (the product)
        gprow[0] = make_float4(fmul(pa, r[0]), fmul(pa, r[1]), fmul(pa, r[2]), fmul(pa, r[3]));
        gprow[1] = make_float4(fmul(pa, r[4]), fmul(pa, r[5]), fmul(pa, r[6]), fmul(pa, r[7]));
    }

    // ---- write x_{k+1} (compute.c:387-403) -------------------------------------------------------
    if constexpr (SW > 0) {
        if (SW * SH > 1 || P.resample) {
#pragma unroll
            for (int sy = 0; sy < SH; sy++) {
                float4 *o = reinterpret_cast<float4 *>(P.xp + (size_t)(cy * SH + sy) * W + (size_t)bx * 8 * SW);
#pragma unroll
                for (int k = 0; k < SW * 2; k++) {
                    float e[4];
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        const int col = k * 4 + m, i = col / SW;
                        e[m] = fadd(fsub(z[sy][col], mean[i]), v[i]);
                    }
                    o[k] = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        } else {
            float4 *o = reinterpret_cast<float4 *>(P.xp + (size_t)cy * W + (size_t)bx * 8);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int cx = bx * 8 + i;
            if (P.resample) {
                for (int sy = 0; sy < sh; sy++)
                    for (int sx = 0; sx < sw; sx++) {
                        const size_t gi = (size_t)(cy * sh + sy) * W + cx * sw + sx;
                        const float zz = stepper(P.x[gi], P.xp[gi], P.g[gi]);
                        P.xp[gi] = fadd(fsub(zz, mean[i]), v[i]);
                    }
            } else {
                P.xp[(size_t)cy * W + cx] = v[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// set-up kernels
// ------------------------------------------------------------------------------------------
// conventional decode of one plane: dequantise + IDCT + raster (jpeg.c:83-92, jpeg2png.c:131-139)
__global__ void __launch_bounds__(P_NT) k_decode(const int16_t *data, const float *q /*[64] device*/, float *out, int cw, int ch) {
    __shared__ __align__(16) float tiles[P_NT / 8][TILE_STRIDE];
    const int tid = threadIdx.x, b = tid >> 3, j = tid & 7;
    const int nb = (cw >> 3) * (ch >> 3);
    const int blk = blockIdx.x * (P_NT / 8) + b;
    if (blk >= nb) return;
    const int bw = cw >> 3, by = blk / bw, bx = blk - by * bw;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int d = data[(size_t)blk * 64 + j * 8 + i];
        v[i] = __int2float_rn(d * (int)q[j * 8 + i]);                         // int product, one rounding (jpeg.c:88)
    }
    idct8x8_rows(v, tiles[b], j);
#pragma unroll
    for (int i = 0; i < 8; i++) out[(size_t)(by * 8 + j) * cw + bx * 8 + i] = v[i];
}

// aux_init (compute.c:295-309): nearest-neighbour upsample with edge clamp into x and xp
__global__ void k_init_plane(const float *fdata, float *x, float *xp, int W, int H, int cw, int ch, int sw, int sh) {
    const size_t n = (size_t)W * H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int py = (int)(i / W), px = (int)(i - (size_t)py * W);
        int cx = px / sw, cy = py / sh;
        cx = cx < cw - 1 ? cx : cw - 1;
        cy = cy < ch - 1 ? cy : ch - 1;
        const float v = fdata[(size_t)cy * cw + cx];
        x[i] = v;
        xp[i] = v;
    }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_project(const FrameDev &F, float factor, cudaStream_t s) {
    for (int c = 0; c < F.nc; c++) {
        const PlaneDev &P = F.pl[c];
        const int tw = 8 * P_BW * P.sw, th = 8 * P_BH * P.sh;
        ProjPlane G;
        G.c = c;
        G.gx = (F.W + tw - 1) / tw;
        const int total = G.gx * ((F.H + th - 1) / th);
        if (P.sw == 1 && P.sh == 1) k_project<1, 1><<<total, P_NT, 0, s>>>(F, G, factor);
        else if (P.sw == 2 && P.sh == 2) k_project<2, 2><<<total, P_NT, 0, s>>>(F, G, factor);
        else if (P.sw == 2 && P.sh == 1) k_project<2, 1><<<total, P_NT, 0, s>>>(F, G, factor);
        else if (P.sw == 1 && P.sh == 2) k_project<1, 2><<<total, P_NT, 0, s>>>(F, G, factor);
        else k_project<0, 0><<<total, P_NT, 0, s>>>(F, G, factor);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_decode(const int16_t *data, const float *q_dev, float *out, int cw, int ch, cudaStream_t s) {
    const int nb = (cw / 8) * (ch / 8);
    k_decode<<<(nb + P_NT / 8 - 1) / (P_NT / 8), P_NT, 0, s>>>(data, q_dev, out, cw, ch);
    return cudaGetLastError();
}

cudaError_t launch_init_plane(const float *fdata, float *x, float *xp, int W, int H, int cw, int ch, int sw, int sh,
                              cudaStream_t s) {
    const size_t n = (size_t)W * H;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_init_plane<<<blocks, 256, 0, s>>>(fdata, x, xp, W, H, cw, ch, sw, sh);
    return cudaGetLastError();
}

}  // namespace j2p
