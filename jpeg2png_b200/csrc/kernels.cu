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
// k_gradient — register-marching, warp-autonomous stencil.
//
// A warp owns a vertical strip of 64 frame columns (60 target columns + a 2-column halo on each
// side), two adjacent columns per lane, and walks down a band of rows.  All three stages of the
// sub-gradient live in registers; horizontal neighbours are exchanged with warp shuffles (eight
// per channel and row), vertical neighbours are the values the lane itself produced one and two
// rows earlier.  No shared memory, no block barrier in the main loop.
//
// Row pipeline at step i (the FISTA point of row i has just been formed):
//   source row s = i-1 : forward differences, joint TV norm and the three TV quotients
//                        (compute.c:73-113); backward differences of the differences, joint TGV
//                        norm and the four TGV quotients (compute.c:128-186)
//   target row s-1     : receives its last two addends (below-left, below) and is stored
//   target row s       : receives its first nine addends, in the one order that reproduces the
//                        reference's scan-order scatter
// ------------------------------------------------------------------------------------------
constexpr int GM_WARPS = 4, GM_NT = GM_WARPS * 32, GM_USE = 60;

// the seven quotients of one source pixel and channel, exact division (rare fallback path)
struct SrcTerms {
    float tvs, tvr, tvb, t2s, lr, ud, dg;
};

template <int NC>
__global__ void __launch_bounds__(GM_NT) k_gradient(const __grid_constant__ FrameDev F, const float factor, const int band_rows) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int W = F.W, H = F.H;
    const int X0 = (blockIdx.x * GM_WARPS + wid) * GM_USE;   // first target column of this warp
    const int yb = blockIdx.y * band_rows;                   // first target row of this CTA
    const int ye = min(yb + band_rows, H);
    const int px0 = X0 - 2 + 2 * lane;                       // even; W is even => the pair is in or out together
    const bool pair_in = px0 >= 0 && px0 < W;
    const bool is_target = pair_in && lane >= 1 && lane <= 30;
    const bool has_l0 = px0 > 0, has_r1 = px0 + 1 < W - 1;   // k=1 always has a left neighbour, k=0 a right one
    const float a1 = F.a1, a2 = F.a2;
    const bool use_tgv = F.use_tgv != 0;

    double acc[NC];
    float yP[NC][2], gxP[NC][2], gyP[NC][2], ogp[NC][2], sv_tvb[NC][2], sv_ud[NC][2], sv_dg[NC][2];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        acc[c] = 0.;
#pragma unroll
        for (int k = 0; k < 2; k++) yP[c][k] = gxP[c][k] = gyP[c][k] = ogp[c][k] = sv_tvb[c][k] = sv_ud[c][k] = sv_dg[c][k] = 0.f;
    }
    // coefficient-grid column of each of the two pixels (DCT-distance term is stored at coefficient resolution)
    int gpx[NC][2];
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int cx = (px0 + k) / F.pl[c].sw;
            gpx[c][k] = (F.pl[c].use_prob && pair_in && cx < F.pl[c].cw) ? cx : -1;
        }

    if (X0 < W) {
        for (int i = yb - 2; i <= ye + 1; i++) {
            // ---- FISTA point of row i (compute.c:436) ------------------------------------------
            float yN[NC][2];
            {
                const bool ld = pair_in && i >= 0 && i < H;
                const size_t gi = (size_t)(ld ? i : 0) * W + (ld ? px0 : 0);
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    float2 a = make_float2(0.f, 0.f), b = a;
                    if (ld) {
                        a = *reinterpret_cast<const float2 *>(F.pl[c].x + gi);
                        b = *reinterpret_cast<const float2 *>(F.pl[c].xp + gi);
                    }
                    yN[c][0] = fadd(a.x, fmul(factor, fsub(a.x, b.x)));
                    yN[c][1] = fadd(a.y, fmul(factor, fsub(a.y, b.y)));
                }
            }
            if (i >= yb - 1) {
                const int s = i - 1;
                const bool src_in = pair_in && s >= 0 && s < H;
                const bool has_d = s < H - 1, has_u = s > 0;

                // ---- source row s: TV (compute.c:79-105) ---------------------------------------
                float gx0[NC][2], gy0[NC][2], tvs0[NC][2], tvr0[NC][2], tvb0[NC][2];
                float n1[2] = {0.f, 0.f};
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const float yr1 = __shfl_down_sync(0xffffffffu, yP[c][0], 1);
                    gx0[c][0] = fsub(yP[c][1], yP[c][0]);                         // px0 < W-1 whenever the pair is in the frame
                    gx0[c][1] = has_r1 ? fsub(yr1, yP[c][1]) : 0.f;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        gy0[c][k] = has_d ? fsub(yN[c][k], yP[c][k]) : 0.f;
                        n1[k] = fadd(n1[k], fsq(gx0[c][k]));
                        n1[k] = fadd(n1[k], fsq(gy0[c][k]));
                    }
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const float n = fsqrt(n1[k]);
                    const bool live = src_in && n != 0.f;                         // compute.c:97
                    const float y = __frcp_rn(n);
                    bool ok = qdiv_divisor_ok(n);
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        tvs0[c][k] = qdiv_fast(fmul(a1, -fadd(gx0[c][k], gy0[c][k])), n, y, ok);   // compute.c:98
                        tvr0[c][k] = qdiv_fast(fmul(a1, gx0[c][k]), n, y, ok);                     // compute.c:100
                        tvb0[c][k] = qdiv_fast(fmul(a1, gy0[c][k]), n, y, ok);                     // compute.c:103
                    }
                    if (live && !ok) {
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            tvs0[c][k] = fdiv(fmul(a1, -fadd(gx0[c][k], gy0[c][k])), n);
                            tvr0[c][k] = fdiv(fmul(a1, gx0[c][k]), n);
                            tvb0[c][k] = fdiv(fmul(a1, gy0[c][k]), n);
                        }
                    }
                    if (!live) {
#pragma unroll
                        for (int c = 0; c < NC; c++) tvs0[c][k] = tvr0[c][k] = tvb0[c][k] = 0.f;
                    }
                }

                // ---- source row s: second-order TGV (compute.c:136-183) ------------------------
                float t2s0[NC][2], lr0[NC][2], ud0[NC][2], dg0[NC][2];
                if (use_tgv && i >= yb) {
                    float gxx[NC][2], gyy[NC][2], sym[NC][2];
                    float n2[2] = {0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const float gxl = __shfl_up_sync(0xffffffffu, gx0[c][1], 1);
                        const float gyl = __shfl_up_sync(0xffffffffu, gy0[c][1], 1);
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const bool has_l = k ? true : has_l0;
                            const float gx_l = k ? gx0[c][0] : gxl, gy_l = k ? gy0[c][0] : gyl;
                            gxx[c][k] = has_l ? fsub(gx0[c][k], gx_l) : 0.f;
                            const float gyx = has_l ? fsub(gy0[c][k], gy_l) : 0.f;
                            const float gxy = has_u ? fsub(gx0[c][k], gxP[c][k]) : 0.f;
                            gyy[c][k] = has_u ? fsub(gy0[c][k], gyP[c][k]) : 0.f;
                            sym[c][k] = fmul(fadd(gxy, gyx), 0.5f);               // (gxy+gyx)/2., exact either way
                            n2[k] = fadd(n2[k], fadd(fadd(fsq(gxx[c][k]), fmul(2.f, fsq(sym[c][k]))), fsq(gyy[c][k])));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const float n = fsqrt(n2[k]);
                        const bool live = src_in && n != 0.f;                     // compute.c:158
                        const float y = __frcp_rn(n);
                        bool ok = qdiv_divisor_ok(n);
                        float num_s[NC], num_lr[NC], num_ud[NC], num_dg[NC];
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            num_s[c] = -fadd(fadd(fmul(2.f, gxx[c][k]), fmul(2.f, sym[c][k])), fmul(2.f, gyy[c][k]));
                            num_lr[c] = fadd(sym[c][k], gxx[c][k]);
                            num_ud[c] = fadd(gyy[c][k], sym[c][k]);
                            num_dg[c] = -sym[c][k];
                            t2s0[c][k] = qdiv_fast(num_s[c], n, y, ok);
                            lr0[c][k] = qdiv_fast(num_lr[c], n, y, ok);
                            ud0[c][k] = qdiv_fast(num_ud[c], n, y, ok);
                            dg0[c][k] = qdiv_fast(num_dg[c], n, y, ok);
                        }
                        if (live && !ok) {
#pragma unroll
                            for (int c = 0; c < NC; c++) {
                                t2s0[c][k] = fdiv(num_s[c], n);
                                lr0[c][k] = fdiv(num_lr[c], n);
                                ud0[c][k] = fdiv(num_ud[c], n);
                                dg0[c][k] = fdiv(num_dg[c], n);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            t2s0[c][k] = live ? fmul(a2, t2s0[c][k]) : 0.f;       // compute.c:165
                            lr0[c][k] = live ? fmul(a2, lr0[c][k]) : 0.f;         // compute.c:167,170
                            ud0[c][k] = live ? fmul(a2, ud0[c][k]) : 0.f;         // compute.c:173,176
                            dg0[c][k] = live ? fmul(a2, dg0[c][k]) : 0.f;         // compute.c:179,182
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NC; c++)
#pragma unroll
                        for (int k = 0; k < 2; k++) t2s0[c][k] = lr0[c][k] = ud0[c][k] = dg0[c][k] = 0.f;
                }

                // ---- target row s-1: last two addends, store, sum of squares -------------------
                if (i >= yb + 2) {
                    const size_t gi = (size_t)(s - 1) * W + (is_target ? px0 : 0);
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        float o0 = ogp[c][0], o1 = ogp[c][1];
                        if (use_tgv) {
                            const float dgl = __shfl_up_sync(0xffffffffu, dg0[c][1], 1);
                            o0 = fadd(fadd(o0, dgl), ud0[c][0]);                  // below-left, below
                            o1 = fadd(fadd(o1, dg0[c][0]), ud0[c][1]);
                        }
                        if (is_target) {
                            *reinterpret_cast<float2 *>(F.pl[c].g + gi) = make_float2(o0, o1);
                            acc[c] = __dadd_rn(acc[c], (double)fsq(o0));          // compute.c:203
                            acc[c] = __dadd_rn(acc[c], (double)fsq(o1));
                        }
                    }
                }

                // ---- target row s: first nine addends ------------------------------------------
                if (s >= yb && s < ye) {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const PlaneDev &P = F.pl[c];
                        float p0 = 0.f, p1 = 0.f;
                        if (P.use_prob) {
                            const int cy = s / P.sh;
                            if (cy < P.ch) {
                                const float *row = P.gp + (size_t)cy * P.cw;
                                if (gpx[c][0] >= 0) p0 = fadd(0.f, row[gpx[c][0]]);   // compute.c:62 onto a zeroed gradient
                                if (gpx[c][1] >= 0) p1 = fadd(0.f, row[gpx[c][1]]);
                            }
                        }
                        const float tvr_l = __shfl_up_sync(0xffffffffu, tvr0[c][1], 1);
                        float o0 = fadd(fadd(fadd(p0, sv_tvb[c][0]), tvr_l), tvs0[c][0]);          // above, left, self
                        float o1 = fadd(fadd(fadd(p1, sv_tvb[c][1]), tvr0[c][0]), tvs0[c][1]);
                        if (use_tgv) {
                            const float dg_r = __shfl_down_sync(0xffffffffu, sv_dg[c][0], 1);
                            const float lr_l = __shfl_up_sync(0xffffffffu, lr0[c][1], 1);
                            const float lr_r = __shfl_down_sync(0xffffffffu, lr0[c][0], 1);
                            // above, above-right, left, self, right
                            o0 = fadd(fadd(fadd(fadd(fadd(o0, sv_ud[c][0]), sv_dg[c][1]), lr_l), t2s0[c][0]), lr0[c][1]);
                            o1 = fadd(fadd(fadd(fadd(fadd(o1, sv_ud[c][1]), dg_r), lr0[c][0]), t2s0[c][1]), lr_r);
                        }
                        ogp[c][0] = o0;
                        ogp[c][1] = o1;
                    }
                }

                // ---- rotate ----------------------------------------------------------------------
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        sv_tvb[c][k] = tvb0[c][k];
                        sv_ud[c][k] = ud0[c][k];
                        sv_dg[c][k] = dg0[c][k];
                        gxP[c][k] = gx0[c][k];
                        gyP[c][k] = gy0[c][k];
                    }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                yP[c][0] = yN[c][0];
                yP[c][1] = yN[c][1];
            }
        }
    }

    // CTA reduction (fixed order => run-to-run deterministic), then the last-CTA fold
    __shared__ double red[3][GM_WARPS];
    __shared__ unsigned ticket;
    const int tid = threadIdx.x;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const double sum = warp_sum(acc[c]);
        if (lane == 0) red[c][wid] = sum;
    }
    __syncthreads();
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x, ncta = gridDim.x * gridDim.y;
    if (tid < NC) {
        double sum = 0.;
        for (int k = 0; k < GM_WARPS; k++) sum = __dadd_rn(sum, red[tid][k]);
        F.partials[(size_t)tid * F.grad_ctas + cta] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) ticket = atomicAdd(F.counter, 1u);
    __syncthreads();
    if (ticket == ncta - 1) {
        __threadfence();
#pragma unroll
        for (int c = 0; c < NC; c++) {
            double sum = 0.;
            for (unsigned k = tid; k < ncta; k += GM_NT) sum = __dadd_rn(sum, __ldcg(&F.partials[(size_t)c * F.grad_ctas + k]));
            sum = warp_sum(sum);
            if (lane == 0) red[c][wid] = sum;
        }
        __syncthreads();
        if (tid < NC) {
            double sum = 0.;
            for (int k = 0; k < GM_WARPS; k++) sum = __dadd_rn(sum, red[tid][k]);
            const float norm = fsqrt(__double2float_rn(sum));                                   // compute.c:205
            F.norms[tid] = norm;
            F.norms[4 + tid] = __frcp_rn(norm);                                                 // shared reciprocal for k_project
        }
        if (tid == 0) *F.counter = 0u;
    }
}

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
    bool stepping, div_ok;
    __device__ __forceinline__ float operator()(float x, float xp, float g) const {
        float y = fadd(x, fmul(factor, fsub(x, xp)));
        if (stepping) {
            bool ok = div_ok;
            float q = qdiv_fast(g, norm, rn, ok);
            if (!ok) q = fdiv(g, norm);
            y = fsub(y, fmul(step, q));
        }
        return y;
    }
};

template <int SW, int SH>
__global__ void __launch_bounds__(P_NT) k_project(const __grid_constant__ FrameDev F, const ProjPlane G, const float factor) {
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
    stepper.div_ok = qdiv_divisor_ok(stepper.norm);
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

    // ---- stepped point of the footprint, block-row means (compute.c:348-370) ------------------
    constexpr int ZW = SW ? SW * 8 : 1, ZH = SW ? SH : 1;
    float z[ZH][ZW];
    float v[8], mean[8];
    if constexpr (SW > 0) {
#pragma unroll
        for (int sy = 0; sy < SH; sy++) {
            const size_t gi = (size_t)(cy * SH + sy) * W + (size_t)bx * 8 * SW;
            const float4 *xr = reinterpret_cast<const float4 *>(P.x + gi);
            const float4 *pr = reinterpret_cast<const float4 *>(P.xp + gi);
            const float4 *gr = reinterpret_cast<const float4 *>(P.g + gi);
#pragma unroll
            for (int k = 0; k < SW * 2; k++) {
                const float4 a = xr[k], p = pr[k], g = gr[k];
                z[sy][k * 4 + 0] = stepper(a.x, p.x, g.x);
                z[sy][k * 4 + 1] = stepper(a.y, p.y, g.y);
                z[sy][k * 4 + 2] = stepper(a.z, p.z, g.z);
                z[sy][k * 4 + 3] = stepper(a.w, p.w, g.w);
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
    const int16_t *drow = P.data + ((size_t)(by * (P.cw >> 3) + bx) * 64 + j * 8);
    const int4 draw = *reinterpret_cast<const int4 *>(drow);
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
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int di = (i & 1) ? (dw[i >> 1] >> 16) : (int)(short)(dw[i >> 1] & 0xffff);
        const float d = (float)di;
        const float q = qv[i];
        const float lo = fmul(fsub(d, 0.5f), q), hi = fmul(fadd(d, 0.5f), q);
        float t = v[i];
        t = t > hi ? hi : (t < lo ? lo : t);
        v[i] = t;
        const float num = fsub(t, fmul(d, q));                               // compute.c:47
        bool ok = true;                                                      // q*q in [1, 2^32]: always a valid divisor
        float rr = qdiv_fast(num, qqv[i], rqv[i], ok);
        if (!ok) rr = fdiv(num, qqv[i]);                                     // compute.c:49
        r[i] = rr;
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
// Band height: one resident wave of CTAs if the frame allows it (long bands amortise the two
// extra source rows each band recomputes), never fewer than 8 rows per band.
static int g_grad_slots = 0;   // CTAs resident on the whole device, set by configure_kernels()

void grad_geometry(int W, int H, int *ctas_x, int *bands, int *band_rows) {
    const int strips = (W + GM_USE - 1) / GM_USE;
    *ctas_x = (strips + GM_WARPS - 1) / GM_WARPS;
    const int slots = g_grad_slots > 0 ? g_grad_slots : 148 * 3;
    int want = slots / *ctas_x;
    if (want < 1) want = 1;
    int rows = (H + want - 1) / want;
    if (rows < 8) rows = 8;
    *band_rows = rows;
    *bands = (H + rows - 1) / rows;
}

int grad_cta_count(int W, int H) {
    int cx, b, r;
    grad_geometry(W, H, &cx, &b, &r);
    return cx * b;
}

cudaError_t configure_kernels() {
    int per_sm = 0, dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gradient<3>, GM_NT, 0);
    if (e != cudaSuccess) return e;
    g_grad_slots = sms * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

cudaError_t launch_gradient(const FrameDev &F, float factor, cudaStream_t s) {
    int cx, bands, rows;
    grad_geometry(F.W, F.H, &cx, &bands, &rows);
    dim3 grid(cx, bands);
    switch (F.nc) {
        case 1: k_gradient<1><<<grid, GM_NT, 0, s>>>(F, factor, rows); break;
        case 2: k_gradient<2><<<grid, GM_NT, 0, s>>>(F, factor, rows); break;
        default: k_gradient<3><<<grid, GM_NT, 0, s>>>(F, factor, rows); break;
    }
    return cudaGetLastError();
}

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
