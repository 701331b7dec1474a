// kernels_project_tile22.cu — step + projection of a 2x2-subsampled plane (4:2:0 chroma) with the
// coalesced, swizzled staging of kernels_project_tile.cu.
//
// Default for 2x2 planes since round 2 (validated bit for bit on the whole GPU suite against
// k_project<2,2>, whose threads fetch their own 64-byte row pieces; 1080p: 56 -> 42 us per
// projection, 8K: 508 -> 362 us; profiles/r02_notes.md).
//
// A coefficient block of a 2x2 plane covers 16 x 16 frame pixels.  Thread j of a block owns
// coefficient row j = frame rows 2j and 2j+1 of that footprint (2 x 16 stepped values in
// registers, compute.c:348-370), exactly like k_project<2,2>.  A CTA owns 16 blocks in a row:
// 256 x 16 frame pixels; its 128 threads copy the three 16 KB arrays with cp.async, consecutive
// lanes on consecutive 16-byte pieces, into shared memory whose 16-byte columns are XOR-swizzled by
// (row >> 1) — the eight threads of a block read rows 2j (+sy), so that is the index that must
// spread them over the banks.  x_{k+1} goes back the same way; gp (coefficient resolution, 8 rows
// of 128 floats per tile) through its own small staging array.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"
#include "pdl.cuh"
#include "project_common.cuh"
#include "strip_sync.cuh"

namespace j2p {

constexpr int P22_NB = 16;                 // coefficient blocks per CTA tile (16 x 1)
constexpr int P22_NT = P22_NB * 8;         // 128 threads
constexpr int P22_C4 = P22_NB * 4;         // float4 columns per frame row of the tile (256 pixels)
constexpr int P22_G4 = P22_NB * 2;         // float4 columns per coefficient row of the tile (128 samples)
constexpr size_t P22_SMEM = (size_t)3 * 16 * P22_C4 * sizeof(float4)      // x_k, x_{k-1}, g
                            + (size_t)8 * P22_G4 * sizeof(float4)          // gp staging
                            + (size_t)P22_NB * TILE_STRIDE * sizeof(float) // transpose tiles
                            + 3 * 64 * sizeof(float) + 4 * sizeof(float);  // tables, norm

__global__ void __launch_bounds__(P22_NT, 3) k_project_tile22(const __grid_constant__ FrameDev F, const int c0, const float factor) {
    extern __shared__ __align__(16) unsigned char smem22[];
    float4 *sx = reinterpret_cast<float4 *>(smem22);                 // [16][P22_C4]  x_k  -> later x_{k+1}
    float4 *sp = sx + 16 * P22_C4;                                   // [16][P22_C4]  x_{k-1}
    float4 *sg = sp + 16 * P22_C4;                                   // [16][P22_C4]  g
    float4 *sgp = sg + 16 * P22_C4;                                  // [8][P22_G4]   gp out
    float *tiles = reinterpret_cast<float *>(sgp + 8 * P22_G4);      // [P22_NB][TILE_STRIDE]
    float *sq = tiles + P22_NB * TILE_STRIDE;                        // [3][64]
    float *snorm = sq + 3 * 64;                                      // [2]

    const int tid = threadIdx.x;
    const int c = c0 + blockIdx.z;                                   // planes of equal geometry share one launch
    const PlaneDev &P = F.pl[c];
    const int W = F.W;
    const int bw = P.cw >> 3;
    const int bx0 = blockIdx.x * P22_NB, by = strip_row_order(F.sync, blockIdx.y, gridDim.y);   // the grid covers real blocks only
    const int nbx = min(P22_NB, bw - bx0);
    const int valid_c4 = nbx * 4, valid_g4 = nbx * 2;
    const size_t row0 = (size_t)(by * 16) * W + (size_t)bx0 * 16;    // first frame pixel of the tile

    // ---- coalesced, swizzled copy-in: 16 rows x 64 pieces per array, 8 pieces per thread ----------
    // x_k and x_{k-1} of these planes are not written by the kernels this launch may overlap (the
    // gradient kernel, the luma projection; pdl.cuh): their tiles are requested
    // BEFORE the wait; the g tile follows it at once (kernels_project_tile.cu).
#pragma unroll
    for (int i = 0; i < 16 * P22_C4 / P22_NT; i++) {
        const int e = tid + P22_NT * i, row = e / P22_C4, c4 = e % P22_C4;
        if (c4 < valid_c4) {
            const size_t gi = row0 + (size_t)row * W + (size_t)c4 * 4;
            const int pc = row * P22_C4 + (c4 ^ ((row >> 1) & 7));
            cp_async16(&sx[pc], P.x + gi);
            cp_async16(&sp[pc], P.xp + gi);
        }
    }
    pdl_wait();                                                      // the gradient and its norm are complete and visible
    pdl_launch_dependents();
#pragma unroll
    for (int i = 0; i < 16 * P22_C4 / P22_NT; i++) {
        const int e = tid + P22_NT * i, row = e / P22_C4, c4 = e % P22_C4;
        if (c4 < valid_c4) cp_async16(&sg[row * P22_C4 + (c4 ^ ((row >> 1) & 7))], P.g + row0 + (size_t)row * W + (size_t)c4 * 4);
    }
    cp_async_commit();
    const int b = tid >> 3, j = tid & 7;
    const bool real = b < nbx;
    int4 draw = make_int4(0, 0, 0, 0);
    if (real) draw = __ldg(reinterpret_cast<const int4 *>(P.data + ((size_t)(by * bw + bx0 + b) * 64 + j * 8)));
    if (tid < 64) {
        sq[tid] = F.q[c][tid];
        sq[64 + tid] = F.qq[c][tid];
        sq[128 + tid] = F.rqq[c][tid];
    }
    if (tid >= 64 && tid < 96) strip_norm(F, c, snorm, tid - 64);    // whole frame: what k_gradient left; strips: fold of every rank's sums
    cp_async_wait<0>();
    __syncthreads();

    Stepper stepper;
    stepper.factor = factor;
    stepper.step = F.step;
    stepper.norm = snorm[0];
    stepper.rn = snorm[1];
    stepper.stepping = stepper.norm != 0.f;                          // compute.c:211
    const bool norm_ok = qdiv_divisor_ok(stepper.norm);
    const bool use_prob = P.use_prob != 0;
    const unsigned gmask = 0xffu << (tid & 24);
    float *tile = tiles + b * TILE_STRIDE;

    if (real) {
        // ---- stepped point of the 2 x 16 footprint (compute.c:436, :213) --------------------------
        float z[2][16], v[8], mean[8];
        {
            unsigned key = 0xffffffffu;
#pragma unroll
            for (int sy = 0; sy < 2; sy++) {
                const int rowbase = (2 * j + sy) * P22_C4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int pc = rowbase + ((4 * b + k) ^ j);      // (row >> 1) & 7 == j
                    const float4 a = sx[pc], p = sp[pc], g = sg[pc];
                    z[sy][k * 4 + 0] = stepper.fast(a.x, p.x, g.x, key);
                    z[sy][k * 4 + 1] = stepper.fast(a.y, p.y, g.y, key);
                    z[sy][k * 4 + 2] = stepper.fast(a.z, p.z, g.z, key);
                    z[sy][k * 4 + 3] = stepper.fast(a.w, p.w, g.w, key);
                }
            }
            if (stepper.stepping && !(norm_ok && key >= QDIV_KEY_MIN)) {   // outside the proven range: IEEE division
#pragma unroll
                for (int sy = 0; sy < 2; sy++) {
                    const int rowbase = (2 * j + sy) * P22_C4;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int pc = rowbase + ((4 * b + k) ^ j);
                        const float4 a = sx[pc], p = sp[pc], g = sg[pc];
                        z[sy][k * 4 + 0] = stepper(a.x, p.x, g.x);
                        z[sy][k * 4 + 1] = stepper(a.y, p.y, g.y);
                        z[sy][k * 4 + 2] = stepper(a.z, p.z, g.z);
                        z[sy][k * 4 + 3] = stepper(a.w, p.w, g.w);
                    }
                }
            }
        }
        // block-row means over the 2 x 2 samples, sy outer, sx inner (compute.c:351-360)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float m = 0.f;
            m = fadd(m, z[0][2 * i]);
            m = fadd(m, z[0][2 * i + 1]);
            m = fadd(m, z[1][2 * i]);
            m = fadd(m, z[1][2 * i + 1]);
            m = fmul(m, 0.25f);                                      // / (float)4: exact, power of two
            mean[i] = m;
            v[i] = m;
        }

        fdct8x8_rows(v, tile, j, gmask);

        // ---- clamp to the quantisation interval (compute.c:323-331); residual (compute.c:47-49) --
        const int dw[4] = {draw.x, draw.y, draw.z, draw.w};
        float r[8], num[8];
        unsigned rkey = 0xffffffffu;
        {
            const float4 *t0 = reinterpret_cast<const float4 *>(&sq[j * 8]);
            const float4 *t1 = reinterpret_cast<const float4 *>(&sq[64 + j * 8]);
            const float4 *t2 = reinterpret_cast<const float4 *>(&sq[128 + j * 8]);
            float qv[8], qqv[8], rqv[8];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float4 a = t0[k], bq = t1[k], cq = t2[k];
                qv[k * 4] = a.x; qv[k * 4 + 1] = a.y; qv[k * 4 + 2] = a.z; qv[k * 4 + 3] = a.w;
                qqv[k * 4] = bq.x; qqv[k * 4 + 1] = bq.y; qqv[k * 4 + 2] = bq.z; qqv[k * 4 + 3] = bq.w;
                rqv[k * 4] = cq.x; rqv[k * 4 + 1] = cq.y; rqv[k * 4 + 2] = cq.z; rqv[k * 4 + 3] = cq.w;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int di = (i & 1) ? (dw[i >> 1] >> 16) : (int)(short)(dw[i >> 1] & 0xffff);
                const float d = (float)di;
                const float q = qv[i];
                const float lo = fmul(fsub(d, 0.5f), q), hi = fmul(fadd(d, 0.5f), q);
                float t = v[i];
                t = t > hi ? hi : (t < lo ? lo : t);
                v[i] = t;
                num[i] = fsub(t, fmul(d, q));
                rkey = min(rkey, qdiv_key(num[i]));
                r[i] = qdiv_core(num[i], qqv[i], rqv[i]);
            }
            if (rkey < QDIV_KEY_MIN) {                               // a residual below 2^-60: IEEE division
#pragma unroll
                for (int i = 0; i < 8; i++) r[i] = fdiv(num[i], qqv[i]);
            }
        }

        idct8x8_rows(v, tile, j, gmask);
        if (use_prob) idct8x8_rows(r, tile, j, gmask);

        // ---- x_{k+1} = (z - mean) + projected mean (compute.c:390-403), into this thread's own cells
#pragma unroll
        for (int sy = 0; sy < 2; sy++) {
            const int rowbase = (2 * j + sy) * P22_C4;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float e[4];
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int col = k * 4 + m, i = col >> 1;
                    e[m] = fadd(fsub(z[sy][col], mean[i]), v[i]);
                }
                sx[rowbase + ((4 * b + k) ^ j)] = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
        if (use_prob) {
            const float pa = P.p_alpha;                              // compute.c:62 (the product)
#pragma unroll
            for (int h = 0; h < 2; h++)
                sgp[j * P22_G4 + ((2 * b + h) ^ j)] =
                    make_float4(fmul(pa, r[h * 4 + 0]), fmul(pa, r[h * 4 + 1]), fmul(pa, r[h * 4 + 2]), fmul(pa, r[h * 4 + 3]));
        }
    }
    __syncthreads();

    // ---- coalesced copy-out: x_{k+1} over x_{k-1} (compute.c:387), gp for the next iteration ------
#pragma unroll
    for (int i = 0; i < 16 * P22_C4 / P22_NT; i++) {
        const int e = tid + P22_NT * i, row = e / P22_C4, c4 = e % P22_C4;
        if (c4 < valid_c4)
            *reinterpret_cast<float4 *>(P.xp + row0 + (size_t)row * W + (size_t)c4 * 4) = sx[row * P22_C4 + (c4 ^ ((row >> 1) & 7))];
    }
    if (use_prob) {
        float *gp0 = P.gp + (size_t)(by * 8) * P.cw + (size_t)bx0 * 8;
#pragma unroll
        for (int i = 0; i < 8 * P22_G4 / P22_NT; i++) {
            const int e = tid + P22_NT * i, row = e / P22_G4, c4 = e % P22_G4;
            if (c4 < valid_g4) *reinterpret_cast<float4 *>(gp0 + (size_t)row * P.cw + (size_t)c4 * 4) = sgp[row * P22_G4 + (c4 ^ row)];
        }
    }

    // ---- strips over peer memory: border rows into the neighbours' halo rows (kernels_project_tile.cu)
    const StripSync &S = F.sync;
    if (S.nranks > 1 && S.fused_halo) {
        const bool top = by == 0 && S.has_up, bottom = by == (int)gridDim.y - 1 && S.has_down;
        if (top || bottom) {
            for (int e = tid; e < 4 * P22_C4; e += P22_NT) {             // 2 rows x 64 pieces, top then bottom
                const int side = e / (2 * P22_C4), r = (e / P22_C4) & 1, c4 = e % P22_C4;
                if (c4 >= valid_c4 || !(side ? bottom : top)) continue;
                const int row = side ? 14 + r : r;
                float *dst = (side ? S.down_dst[c] : S.up_dst[c]) + (size_t)r * W + (size_t)bx0 * 16 + (size_t)c4 * 4;
                *reinterpret_cast<float4 *>(dst) = sx[row * P22_C4 + (c4 ^ ((row >> 1) & 7))];
            }
            if (top) strip_border_done(S, 0);
            if (bottom) strip_border_done(S, 1);
        }
    }
}

// frame pixels of a 2x2 plane beyond its coefficient grid (W > 2 cw or H > 2 ch): step only
__global__ void k_step_uncovered22(const __grid_constant__ FrameDev F, const int c, const float factor) {
    const PlaneDev &P = F.pl[c];
    const int W = F.W, H = F.H, cwf = 2 * P.cw, chf = 2 * P.ch;
    Stepper stepper;
    stepper.factor = factor;
    stepper.step = F.step;
    stepper.norm = F.norms[c];
    stepper.rn = 0.f;
    stepper.stepping = stepper.norm != 0.f;
    const unsigned bottom = (unsigned)(H - chf) * (unsigned)W, right_w = (unsigned)(W - cwf);
    const unsigned n = bottom + (unsigned)chf * right_w;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned px, py;
        if (i < bottom) {
            py = (unsigned)chf + i / (unsigned)W;
            px = i % (unsigned)W;
        } else {
            const unsigned k = i - bottom;
            py = k / right_w;
            px = (unsigned)cwf + k % right_w;
        }
        const size_t gi = (size_t)py * W + px;
        P.xp[gi] = stepper(P.x[gi], P.xp[gi], P.g[gi]);
    }
}

cudaError_t configure_project_tile22() {
    return cudaFuncSetAttribute(k_project_tile22, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P22_SMEM);
}

// F: already restricted to the rows the session owns (launch_project).  Projects planes
// c .. c+count-1, which must all be 2x2 planes with the same coefficient grid.
cudaError_t launch_project_tile22(const FrameDev &F, int c, int count, float factor, cudaStream_t s, int *nlaunch) {
    const PlaneDev &P = F.pl[c];
    const int bw = P.cw >> 3, bh = P.ch >> 3;
    const dim3 grid((bw + P22_NB - 1) / P22_NB, bh, count);
    cudaError_t e = launch_chain(k_project_tile22, grid, dim3(P22_NT), P22_SMEM, s, F, c, factor);
    *nlaunch += 1;
    for (int k = c; k < c + count && e == cudaSuccess; k++) {
        const PlaneDev &Q = F.pl[k];
        if (2 * Q.cw < F.W || 2 * Q.ch < F.H) {
            const size_t n = (size_t)(F.H - 2 * Q.ch) * F.W + (size_t)2 * Q.ch * (F.W - 2 * Q.cw);
            int blocks = (int)((n + 255) / 256);
            if (blocks > 148 * 8) blocks = 148 * 8;
            k_step_uncovered22<<<blocks, 256, 0, s>>>(F, k, factor);
            e = cudaGetLastError();
            *nlaunch += 1;
        }
    }
    return e;
}

}  // namespace j2p
