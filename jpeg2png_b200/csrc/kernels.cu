// kernels.cu — sm_100a kernels of the jpeg2png solver hot path.
//
// One solver iteration (reference compute.c:427-453) is two kernels:
//
//   k_gradient  : FISTA extrapolation y = x_k + f (x_k - x_{k-1}) recomputed on the fly
//                 (compute.c:431-440), TV sub-gradient (compute.c:73-113), second-order TGV
//                 sub-gradient (compute.c:128-186) restated as an ordered per-pixel GATHER
//                 (SURVEY.md §8a), plus the DCT-distance term read from `gp`; writes g and the
//                 per-CTA fp64 partial sums of g^2; the last CTA to finish folds the partials
//                 into the three norms of compute.c:200-206.
//   k_project   : recomputes y, takes the normalised step (compute.c:209-216), projects onto
//                 the quantisation box — block mean split, 8x8 DCT, clamp, IDCT, add back
//                 (compute.c:334-404) — writes x_{k+1} over x_{k-1}, and — from the clamped
//                 coefficients it has in registers — already produces the DCT-distance gradient
//                 of the NEXT iteration (compute.c:38-70) into `gp`.  The reference keeps the
//                 clamped coefficients in aux.cos (compute.c:381) and re-reads them next step;
//                 here they never leave the SM.
//
// box()/unbox() (box.c) are addressing only.  No tensor cores: the path is a stencil plus
// block-local 8-point butterflies in emulated-reference arithmetic (numerics.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"

namespace j2p {

// ------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------
// k_gradient (v1: one CTA = 32x16 frame pixels, all channels; three smem-staged passes)
// ------------------------------------------------------------------------------------------
constexpr int G_TW = 32, G_TH = 16, G_NT = 256;
constexpr int G_YW = G_TW + 4, G_YH = G_TH + 4;   // FISTA point, halo 2
constexpr int G_SW = G_TW + 2, G_SH = G_TH + 2;   // per-source terms, halo 1

template <int NC>
__global__ void __launch_bounds__(G_NT) k_gradient(const __grid_constant__ FrameDev F, const float factor) {
    extern __shared__ float smem[];
    float *sy = smem;                                    // [NC][G_YH][G_YW]
    float *st = smem + NC * G_YH * G_YW;                 // [7][NC][G_SH][G_SW]
    constexpr int TS = G_SH * G_SW;                      // one term plane
    float *tvs = st, *tvr = st + NC * TS, *tvb = st + 2 * NC * TS;
    float *t2s = st + 3 * NC * TS, *t2lr = st + 4 * NC * TS, *t2ud = st + 5 * NC * TS, *t2dg = st + 6 * NC * TS;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * G_TW, y0 = blockIdx.y * G_TH;
    const int W = F.W, H = F.H;

    // pass 1: FISTA point on the tile + halo 2 (0 outside the frame; never used there)
    for (int i = tid; i < G_YW * G_YH; i += G_NT) {
        const int ly = i / G_YW, lx = i - ly * G_YW;
        const int px = x0 - 2 + lx, py = y0 - 2 + ly;
        const bool in = px >= 0 && px < W && py >= 0 && py < H;
        const size_t gi = (size_t)py * W + px;
#pragma unroll
        for (int c = 0; c < NC; c++) {
            float v = 0.f;
            if (in) {
                const float a = F.pl[c].x[gi], b = F.pl[c].xp[gi];
                v = fadd(a, fmul(factor, fsub(a, b)));
            }
            sy[c * G_YH * G_YW + i] = v;
        }
    }
    __syncthreads();

    // pass 2: per-source TV / TGV terms on the tile + halo 1
    for (int i = tid; i < TS; i += G_NT) {
        const int ly = i / G_SW, lx = i - ly * G_SW;
        const int px = x0 - 1 + lx, py = y0 - 1 + ly;
        const bool in = px >= 0 && px < W && py >= 0 && py < H;
        float o_tvs[NC], o_tvr[NC], o_tvb[NC], o_t2s[NC], o_lr[NC], o_ud[NC], o_dg[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) o_tvs[c] = o_tvr[c] = o_tvb[c] = o_t2s[c] = o_lr[c] = o_ud[c] = o_dg[c] = 0.f;
        if (in) {
            const bool has_r = px < W - 1, has_d = py < H - 1, has_l = px > 0, has_u = py > 0;
            const int yi = (ly + 1) * G_YW + (lx + 1);          // same pixel inside sy
            float gx[NC], gy[NC], gxx[NC], gyy[NC], sym[NC];
            float n1 = 0.f, n2 = 0.f;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const float *Y = sy + c * G_YH * G_YW;
                const float y00 = Y[yi];
                // forward differences at this pixel (compute.c:79-81)
                gx[c] = has_r ? fsub(Y[yi + 1], y00) : 0.f;
                gy[c] = has_d ? fsub(Y[yi + G_YW], y00) : 0.f;
                n1 = fadd(n1, fsq(gx[c]));
                n1 = fadd(n1, fsq(gy[c]));
                // forward differences at the left and upper neighbours, then backward
                // differences of those (compute.c:136-145)
                float gx_l = 0.f, gy_l = 0.f, gx_u = 0.f, gy_u = 0.f;
                if (has_l) {
                    const float yl = Y[yi - 1];
                    gx_l = fsub(y00, yl);                                   // x-1 < W-1 always
                    gy_l = has_d ? fsub(Y[yi - 1 + G_YW], yl) : 0.f;
                }
                if (has_u) {
                    const float yu = Y[yi - G_YW];
                    gx_u = has_r ? fsub(Y[yi - G_YW + 1], yu) : 0.f;
                    gy_u = fsub(y00, yu);                                   // y-1 < H-1 always
                }
                gxx[c] = has_l ? fsub(gx[c], gx_l) : 0.f;
                const float gyx = has_l ? fsub(gy[c], gy_l) : 0.f;
                const float gxy = has_u ? fsub(gx[c], gx_u) : 0.f;
                gyy[c] = has_u ? fsub(gy[c], gy_u) : 0.f;
                sym[c] = fmul(fadd(gxy, gyx), 0.5f);                        // (gxy+gyx)/2., exact either way
                n2 = fadd(n2, fadd(fadd(fsq(gxx[c]), fmul(2.f, fsq(sym[c]))), fsq(gyy[c])));
            }
            n1 = fsqrt(n1);
            if (n1 != 0.f) {
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    o_tvs[c] = fdiv(fmul(F.a1, -fadd(gx[c], gy[c])), n1);   // compute.c:98
                    o_tvr[c] = fdiv(fmul(F.a1, gx[c]), n1);                 // compute.c:100
                    o_tvb[c] = fdiv(fmul(F.a1, gy[c]), n1);                 // compute.c:103
                }
            }
            if (F.use_tgv) {
                n2 = fsqrt(n2);
                if (n2 != 0.f) {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const float tw = fadd(fadd(fmul(2.f, gxx[c]), fmul(2.f, sym[c])), fmul(2.f, gyy[c]));
                        o_t2s[c] = fmul(F.a2, fdiv(-tw, n2));                            // compute.c:165
                        o_lr[c] = fmul(F.a2, fdiv(fadd(sym[c], gxx[c]), n2));            // compute.c:167,170
                        o_ud[c] = fmul(F.a2, fdiv(fadd(gyy[c], sym[c]), n2));            // compute.c:173,176
                        o_dg[c] = fmul(F.a2, fdiv(-sym[c], n2));                         // compute.c:179,182
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            tvs[c * TS + i] = o_tvs[c];
            tvr[c * TS + i] = o_tvr[c];
            tvb[c * TS + i] = o_tvb[c];
            t2s[c * TS + i] = o_t2s[c];
            t2lr[c * TS + i] = o_lr[c];
            t2ud[c * TS + i] = o_ud[c];
            t2dg[c * TS + i] = o_dg[c];
        }
    }
    __syncthreads();

    // pass 3: ordered gather into g, and the fp64 sums of squares
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0.;
    for (int i = tid; i < G_TW * G_TH; i += G_NT) {
        const int ly = i / G_TW, lx = i - ly * G_TW;
        const int px = x0 + lx, py = y0 + ly;
        if (px < W && py < H) {
            const int si = (ly + 1) * G_SW + (lx + 1);
            const size_t gi = (size_t)py * W + px;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const PlaneDev &P = F.pl[c];
                float og = 0.f;
                if (P.use_prob) {
                    const int cx = px / P.sw, cy = py / P.sh;
                    if (cx < P.cw && cy < P.ch) og = fadd(0.f, P.gp[(size_t)cy * P.cw + cx]);   // compute.c:62
                }
                const int o = c * TS + si;
                og = fadd(og, tvb[o - G_SW]);          // TV, source above
                og = fadd(og, tvr[o - 1]);             // TV, source left
                og = fadd(og, tvs[o]);                 // TV, self
                if (F.use_tgv) {
                    og = fadd(og, t2ud[o - G_SW]);     // above
                    og = fadd(og, t2dg[o - G_SW + 1]); // above-right
                    og = fadd(og, t2lr[o - 1]);        // left
                    og = fadd(og, t2s[o]);             // self
                    og = fadd(og, t2lr[o + 1]);        // right
                    og = fadd(og, t2dg[o + G_SW - 1]); // below-left
                    og = fadd(og, t2ud[o + G_SW]);     // below
                }
                P.g[gi] = og;
                acc[c] = __dadd_rn(acc[c], (double)fsq(og));                                    // compute.c:203
            }
        }
    }

    // CTA reduction (fixed order => run-to-run deterministic), then the last-CTA fold
    __shared__ double red[3][G_NT / 32];
    __shared__ unsigned ticket;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const double s = warp_sum(acc[c]);
        if (lane == 0) red[c][wid] = s;
    }
    __syncthreads();
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x, ncta = gridDim.x * gridDim.y;
    if (tid < NC) {
        double s = 0.;
        for (int k = 0; k < G_NT / 32; k++) s = __dadd_rn(s, red[tid][k]);
        F.partials[(size_t)tid * F.grad_ctas + cta] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) ticket = atomicAdd(F.counter, 1u);
    __syncthreads();
    if (ticket == ncta - 1) {
        __threadfence();
#pragma unroll
        for (int c = 0; c < NC; c++) {
            double s = 0.;
            for (unsigned k = tid; k < ncta; k += G_NT) s = __dadd_rn(s, __ldcg(&F.partials[(size_t)c * F.grad_ctas + k]));
            s = warp_sum(s);
            if (lane == 0) red[c][wid] = s;
        }
        __syncthreads();
        if (tid < NC) {
            double s = 0.;
            for (int k = 0; k < G_NT / 32; k++) s = __dadd_rn(s, red[tid][k]);
            F.norms[tid] = fsqrt(__double2float_rn(s));                                         // compute.c:205
        }
        if (tid == 0) *F.counter = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// 8x8 block transposes among the 8 lanes that own one block (lane j holds row j).
// tile: 8 rows x 9 floats (padded), private to the block.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void transpose8(float (&v)[8], float *tile, int j) {
#pragma unroll
    for (int i = 0; i < 8; i++) tile[j * 9 + i] = v[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[i * 9 + j];
    __syncwarp();
}

// 2-D transforms for a thread that holds row j of the block and ends holding row j.
// Vertical pass first, horizontal second (ooura/dct.c:39-94, :103-158).
__device__ __forceinline__ void fdct8x8_rows(float (&v)[8], float *tile, int j) {
    transpose8(v, tile, j);
    fdct8(v);
    transpose8(v, tile, j);
    fdct8(v);
}
__device__ __forceinline__ void idct8x8_rows(float (&v)[8], float *tile, int j) {
    transpose8(v, tile, j);
    idct8(v);
    transpose8(v, tile, j);
    idct8(v);
}

// ------------------------------------------------------------------------------------------
// k_project (v1: 8 threads per coefficient block, 32 blocks per CTA)
// ------------------------------------------------------------------------------------------
constexpr int P_NT = 256, P_BW = 8, P_BH = 4;   // CTA tile: 8 x 4 coefficient blocks

struct ProjGrid {
    int first[4];   // first linear CTA of plane c; first[nc] = total
    int gx[3];      // CTAs per row of plane c
};

__global__ void __launch_bounds__(P_NT) k_project(const __grid_constant__ FrameDev F, const __grid_constant__ ProjGrid G,
                                                   const float factor) {
    __shared__ float tiles[P_NT / 8][8 * 9];
    const int tid = threadIdx.x;
    int c = 0;
    if ((int)blockIdx.x >= G.first[1]) c = 1;
    if ((int)blockIdx.x >= G.first[2]) c = 2;
    const int rel = blockIdx.x - G.first[c];
    const int ctay = rel / G.gx[c], ctax = rel - ctay * G.gx[c];
    const PlaneDev &P = F.pl[c];
    const int W = F.W, H = F.H;
    const int b = tid >> 3, j = tid & 7;
    const int bx = ctax * P_BW + (b & (P_BW - 1)), by = ctay * P_BH + (b >> 3);
    const bool real = bx < (P.cw >> 3) && by < (P.ch >> 3);
    const float norm = F.norms[c];
    const float step = F.step;
    const int sw = P.sw, sh = P.sh;
    float *tile = tiles[b];

    // the stepped point at one frame pixel (compute.c:436 then :213)
    auto stepped = [&](size_t gi) -> float {
        const float a = P.x[gi], p = P.xp[gi];
        float y = fadd(a, fmul(factor, fsub(a, p)));
        if (norm != 0.f) y = fsub(y, fmul(step, fdiv(P.g[gi], norm)));
        return y;
    };

    const int cy = by * 8 + j;
    if (!real) {
        // pixels of the frame that no coefficient block covers: step only (compute.c:349-350 never visits them)
        for (int i = 0; i < 8; i++)
            for (int sy = 0; sy < sh; sy++)
                for (int sx = 0; sx < sw; sx++) {
                    const int px = (bx * 8 + i) * sw + sx, py = cy * sh + sy;
                    if (px < W && py < H) {
                        const size_t gi = (size_t)py * W + px;
                        P.xp[gi] = stepped(gi);
                    }
                }
        return;   // whole 8-lane groups leave together; the remaining lanes still __syncwarp among themselves
    }

    float v[8], mean[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int cx = bx * 8 + i;
        if (P.resample) {
            float m = 0.f;                                                   // compute.c:351
            for (int sy = 0; sy < sh; sy++)
                for (int sx = 0; sx < sw; sx++) m = fadd(m, stepped((size_t)(cy * sh + sy) * W + cx * sw + sx));
            m = fdiv(m, P.cnt);                                      // compute.c:359
            mean[i] = m;
            v[i] = m;
        } else {
            mean[i] = 0.f;
            v[i] = stepped((size_t)cy * W + cx);
        }
    }

    fdct8x8_rows(v, tile, j);

    // clamp to the quantisation interval (compute.c:323-331) and form the DCT-distance residual
    const int16_t *drow = P.data + ((size_t)(by * (P.cw >> 3) + bx) * 64 + j * 8);
    const int4 draw = *reinterpret_cast<const int4 *>(drow);
    const int dw[4] = {draw.x, draw.y, draw.z, draw.w};
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int di = (i & 1) ? (dw[i >> 1] >> 16) : (int)(short)(dw[i >> 1] & 0xffff);
        const float d = (float)di;
        const float q = F.q[c][j * 8 + i];
        const float lo = fmul(fsub(d, 0.5f), q), hi = fmul(fadd(d, 0.5f), q);
        float t = v[i];
        t = t > hi ? hi : (t < lo ? lo : t);
        v[i] = t;
        r[i] = fdiv(fsub(t, fmul(d, q)), F.qq[c][j * 8 + i]);                // compute.c:47,49
    }

    idct8x8_rows(v, tile, j);
    if (P.use_prob) {
        idct8x8_rows(r, tile, j);
        float *gprow = P.gp + (size_t)cy * P.cw + bx * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) gprow[i] = fmul(P.p_alpha, r[i]);         // compute.c:62 (the product)
    }

    // write x_{k+1} (compute.c:387-403)
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int cx = bx * 8 + i;
        if (P.resample) {
            for (int sy = 0; sy < sh; sy++)
                for (int sx = 0; sx < sw; sx++) {
                    const size_t gi = (size_t)(cy * sh + sy) * W + cx * sw + sx;
                    const float z = stepped(gi);
                    P.xp[gi] = fadd(fsub(z, mean[i]), v[i]);
                }
        } else {
            P.xp[(size_t)cy * W + cx] = v[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// set-up kernels
// ------------------------------------------------------------------------------------------
// conventional decode of one plane: dequantise + IDCT + raster (jpeg.c:83-92, jpeg2png.c:131-139)
__global__ void __launch_bounds__(P_NT) k_decode(const int16_t *data, const float *q /*[64] device*/, float *out, int cw, int ch) {
    __shared__ float tiles[P_NT / 8][8 * 9];
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
static size_t grad_smem(int nc) { return (size_t)(nc * G_YH * G_YW + 7 * nc * G_SH * G_SW) * sizeof(float); }

int grad_cta_count(int W, int H) { return ((W + G_TW - 1) / G_TW) * ((H + G_TH - 1) / G_TH); }

cudaError_t configure_kernels() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_gradient<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem(1));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_gradient<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem(2));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_gradient<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem(3));
    return e;
}

cudaError_t launch_gradient(const FrameDev &F, float factor, cudaStream_t s) {
    dim3 grid((F.W + G_TW - 1) / G_TW, (F.H + G_TH - 1) / G_TH);
    switch (F.nc) {
        case 1: k_gradient<1><<<grid, G_NT, grad_smem(1), s>>>(F, factor); break;
        case 2: k_gradient<2><<<grid, G_NT, grad_smem(2), s>>>(F, factor); break;
        default: k_gradient<3><<<grid, G_NT, grad_smem(3), s>>>(F, factor); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_project(const FrameDev &F, float factor, cudaStream_t s) {
    ProjGrid G;
    int total = 0;
    for (int c = 0; c < 3; c++) {
        G.first[c] = total;
        G.gx[c] = 1;
        if (c < F.nc) {
            const int tw = 8 * P_BW * F.pl[c].sw, th = 8 * P_BH * F.pl[c].sh;
            G.gx[c] = (F.W + tw - 1) / tw;
            total += G.gx[c] * ((F.H + th - 1) / th);
        }
    }
    G.first[3] = total;
    for (int c = F.nc; c < 3; c++) G.first[c] = total;
    k_project<<<total, P_NT, 0, s>>>(F, G, factor);
    return cudaGetLastError();
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
