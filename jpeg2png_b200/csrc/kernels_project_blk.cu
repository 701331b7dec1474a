// kernels_project_blk.cu — step + projection of a full-resolution (1x1) plane, ONE THREAD PER
// 8x8 COEFFICIENT BLOCK.
//
// What bounds this kernel is the XU pipe: every expression of the 8-point transforms that touches
// one of Ooura's double constants is an fp64 product/sum between an f32->f64 and an f64->f32
// conversion, and sm_100 converts 16 values per clock and SM (profiles/r01_microbench_ops.txt).
// Three 2-D transforms per block = 960 conversions = 60 clocks*SM per block, i.e. >= 27 us for the
// 129 600 blocks of a 4K plane.  The organisation with 8 threads per block and shared-memory
// transposes (kernels_project.cu) sits at 45 % of that limit no matter the occupancy, the ILP or
// the load pipelining (profiles/r01_notes.md): its threads keep waiting for LDS results queued
// behind other warps' conversions in the in-order MIO queue.  A thread that holds the whole block
// in registers has no transposes at all and eight independent 1-D transforms in flight per pass;
// profiles/r01_microbench2.txt shows it saturating XU with three warps per scheduler.
//
// Schedule: persistent 128-thread CTAs, no barrier in the loop.  While a thread transforms block
// i from registers, the 896 bytes of block i+1 (8 rows of x_k, x_{k-1}, g and the coefficient
// block) stream into its PRIVATE shared-memory slots with cp.async.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"
#include "project_common.cuh"

namespace j2p {

constexpr int PB_NT = 128;
constexpr int PB_SLOTS = 56;                                   // 16 x, 16 xp, 16 g, 8 coefficient float4s
constexpr size_t PB_DYN_SMEM = (size_t)PB_SLOTS * PB_NT * sizeof(float4);

// vertical pass = 1-D transform of each column, horizontal pass = of each row (ooura/dct.c:39-94, :103-158)
template <bool FWD>
__device__ __forceinline__ void pass_cols(float (&a)[64]) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = a[k * 8 + c];
        if (FWD) fdct8(t); else idct8(t);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k * 8 + c] = t[k];
    }
}
template <bool FWD>
__device__ __forceinline__ void pass_rows(float (&a)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = a[r * 8 + k];
        if (FWD) fdct8(t); else idct8(t);
#pragma unroll
        for (int k = 0; k < 8; k++) a[r * 8 + k] = t[k];
    }
}

__global__ void __launch_bounds__(PB_NT, 2) k_project_blk(const __grid_constant__ FrameDev F, const int c, const float factor) {
    extern __shared__ __align__(16) float4 stage[];              // [PB_SLOTS][PB_NT], thread-private columns
    __shared__ __align__(16) float sq[3][64];                    // q, q*q, RN(1/(q*q))
    __shared__ float snorm[2];
    const int tid = threadIdx.x;
    const PlaneDev &P = F.pl[c];
    const int W = F.W;
    const int bw = P.cw >> 3, nblocks = bw * (P.ch >> 3);
    float4 *slot = stage + tid;

    auto issue = [&](int blk) {
        if (blk < nblocks) {
            const int by = blk / bw, bx = blk - by * bw;
            const size_t base = (size_t)(by * 8) * W + (size_t)bx * 8;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const size_t gi = base + (size_t)r * W;
                cp_async16(slot + (0 + r * 2) * PB_NT, P.x + gi);
                cp_async16(slot + (1 + r * 2) * PB_NT, P.x + gi + 4);
                cp_async16(slot + (16 + r * 2) * PB_NT, P.xp + gi);
                cp_async16(slot + (17 + r * 2) * PB_NT, P.xp + gi + 4);
                cp_async16(slot + (32 + r * 2) * PB_NT, P.g + gi);
                cp_async16(slot + (33 + r * 2) * PB_NT, P.g + gi + 4);
                cp_async16(slot + (48 + r) * PB_NT, P.data + ((size_t)blk * 64 + r * 8));
            }
        }
        cp_async_commit();
    };

    int blk = blockIdx.x * PB_NT + tid;
    const int stride = gridDim.x * PB_NT;
    issue(blk);
    if (tid < 64) {
        sq[0][tid] = F.q[c][tid];
        sq[1][tid] = F.qq[c][tid];
        sq[2][tid] = F.rqq[c][tid];
    } else if (tid == 64) {
        snorm[0] = F.norms[c];
        snorm[1] = F.norms[4 + c];
    }
    __syncthreads();
    Stepper stepper;
    stepper.factor = factor;
    stepper.step = F.step;
    stepper.norm = snorm[0];
    stepper.rn = snorm[1];
    stepper.stepping = stepper.norm != 0.f;                        // compute.c:211
    const bool norm_ok = qdiv_divisor_ok(stepper.norm);
    const float pa = P.p_alpha;
    const bool use_prob = P.use_prob != 0, resample = P.resample != 0;

    for (; blk < nblocks; blk += stride) {
        cp_async_wait<0>();                                        // this thread's copies of block `blk` have landed
        const int by = blk / bw, bx = blk - by * bw;
        const size_t base = (size_t)(by * 8) * W + (size_t)bx * 8;

        // ---- stepped point (compute.c:436, :213) -----------------------------------------------
        float v[64];
        unsigned key = 0xffffffffu;
        unsigned long long negzero = 0ull;                         // samples whose stepped value is -0 (see below)
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float4 a = slot[(0 + r * 2 + h) * PB_NT], p = slot[(16 + r * 2 + h) * PB_NT], g = slot[(32 + r * 2 + h) * PB_NT];
                v[r * 8 + h * 4 + 0] = stepper.fast(a.x, p.x, g.x, key);
                v[r * 8 + h * 4 + 1] = stepper.fast(a.y, p.y, g.y, key);
                v[r * 8 + h * 4 + 2] = stepper.fast(a.z, p.z, g.z, key);
                v[r * 8 + h * 4 + 3] = stepper.fast(a.w, p.w, g.w, key);
            }
        if (stepper.stepping && !(norm_ok && key >= QDIV_KEY_MIN)) {   // outside the proven range: IEEE division
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float4 a = slot[(0 + r * 2 + h) * PB_NT], p = slot[(16 + r * 2 + h) * PB_NT], g = slot[(32 + r * 2 + h) * PB_NT];
                    v[r * 8 + h * 4 + 0] = stepper(a.x, p.x, g.x);
                    v[r * 8 + h * 4 + 1] = stepper(a.y, p.y, g.y);
                    v[r * 8 + h * 4 + 2] = stepper(a.z, p.z, g.z);
                    v[r * 8 + h * 4 + 3] = stepper(a.w, p.w, g.w);
                }
        }
        if (resample) {
            // Sampling 1x1 on a coefficient grid smaller than the frame (1080p luma): the reference
            // still splits z into mean = (0 + z)/1 and d = z - mean (compute.c:351-367) and returns
            // d + v'.  d is +0 unless z is -0 (then -0), and 0 + z turns a -0 sample into +0.
#pragma unroll
            for (int i = 0; i < 64; i++) {
                if (__float_as_uint(v[i]) == 0x80000000u) negzero |= 1ull << i;
                v[i] = fadd(0.f, v[i]);
            }
        }

        pass_cols<true>(v);
        pass_rows<true>(v);

        // ---- clamp to the quantisation interval (compute.c:323-331); residual (compute.c:47-49) --
        float r[64];
        {
            unsigned rkey = 0xffffffffu;
#pragma unroll
            for (int row = 0; row < 8; row++) {
                const float4 dq = slot[(48 + row) * PB_NT];
                const int dw[4] = {__float_as_int(dq.x), __float_as_int(dq.y), __float_as_int(dq.z), __float_as_int(dq.w)};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int j = row * 8 + i;
                    const int di = (i & 1) ? (dw[i >> 1] >> 16) : (int)(short)(dw[i >> 1] & 0xffff);
                    const float d = (float)di;
                    const float q = sq[0][j];
                    const float lo = fmul(fsub(d, 0.5f), q), hi = fmul(fadd(d, 0.5f), q);
                    float t = v[j];
                    t = t > hi ? hi : (t < lo ? lo : t);
                    v[j] = t;
                    const float num = fsub(t, fmul(d, q));
                    rkey = min(rkey, qdiv_key(num));
                    r[j] = qdiv_core(num, sq[1][j], sq[2][j]);
                }
            }
            if (rkey < QDIV_KEY_MIN) {                             // a residual below 2^-60: IEEE division
#pragma unroll
                for (int row = 0; row < 8; row++) {
                    const float4 dq = slot[(48 + row) * PB_NT];
                    const int dw[4] = {__float_as_int(dq.x), __float_as_int(dq.y), __float_as_int(dq.z), __float_as_int(dq.w)};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int j = row * 8 + i;
                        const int di = (i & 1) ? (dw[i >> 1] >> 16) : (int)(short)(dw[i >> 1] & 0xffff);
                        r[j] = fdiv(fsub(v[j], fmul((float)di, sq[0][j])), sq[1][j]);
                    }
                }
            }
        }

        // every read of the staging slots is done: start streaming the next block in
        issue(blk + stride);

        pass_cols<false>(v);
        pass_rows<false>(v);
        if (use_prob) {
            pass_cols<false>(r);
            pass_rows<false>(r);
            float *gp = P.gp + (size_t)(by * 8) * P.cw + (size_t)bx * 8;
#pragma unroll
            for (int row = 0; row < 8; row++) {
                float4 *o = reinterpret_cast<float4 *>(gp + (size_t)row * P.cw);
                o[0] = make_float4(fmul(pa, r[row * 8 + 0]), fmul(pa, r[row * 8 + 1]), fmul(pa, r[row * 8 + 2]), fmul(pa, r[row * 8 + 3]));   // compute.c:62
                o[1] = make_float4(fmul(pa, r[row * 8 + 4]), fmul(pa, r[row * 8 + 5]), fmul(pa, r[row * 8 + 6]), fmul(pa, r[row * 8 + 7]));
            }
        }

        // ---- write x_{k+1} over x_{k-1} (compute.c:387-403) ------------------------------------
        if (resample) {
#pragma unroll
            for (int i = 0; i < 64; i++) v[i] = fadd((negzero >> i) & 1ull ? -0.f : 0.f, v[i]);
        }
#pragma unroll
        for (int row = 0; row < 8; row++) {
            float4 *o = reinterpret_cast<float4 *>(P.xp + base + (size_t)row * W);
            o[0] = make_float4(v[row * 8 + 0], v[row * 8 + 1], v[row * 8 + 2], v[row * 8 + 3]);
            o[1] = make_float4(v[row * 8 + 4], v[row * 8 + 5], v[row * 8 + 6], v[row * 8 + 7]);
        }
    }
    cp_async_wait<0>();
}

// frame pixels of a 1x1 plane that no coefficient block covers (1080p: luma rows 1080..1087): step only
__global__ void k_step_uncovered(const __grid_constant__ FrameDev F, const int c, const float factor) {
    const PlaneDev &P = F.pl[c];
    const int W = F.W, H = F.H;
    Stepper stepper;
    stepper.factor = factor;
    stepper.step = F.step;
    stepper.norm = F.norms[c];
    stepper.rn = 0.f;
    stepper.stepping = stepper.norm != 0.f;
    const size_t n = (size_t)W * H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int py = (int)(i / W), px = (int)(i - (size_t)py * W);
        if (px >= P.cw || py >= P.ch) P.xp[i] = stepper(P.x[i], P.xp[i], P.g[i]);
    }
}

cudaError_t launch_step_uncovered(const FrameDev &F, int c, float factor, cudaStream_t s) {
    const size_t n = (size_t)F.W * F.H;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_step_uncovered<<<blocks, 256, 0, s>>>(F, c, factor);
    return cudaGetLastError();
}

static int g_blk_slots = 148 * 2;

cudaError_t configure_project_blk() {
    cudaError_t e = cudaFuncSetAttribute(k_project_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_DYN_SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 0, dev = 0, sms = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_project_blk, PB_NT, PB_DYN_SMEM);
    if (e != cudaSuccess) return e;
    g_blk_slots = sms * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

// F: already restricted to the rows the session owns (launch_project)
cudaError_t launch_project_blk(const FrameDev &F, int c, float factor, cudaStream_t s) {
    const PlaneDev &P = F.pl[c];
    const int nblocks = (P.cw >> 3) * (P.ch >> 3);
    int ctas = (nblocks + PB_NT - 1) / PB_NT;
    if (ctas > g_blk_slots) ctas = g_blk_slots;
    k_project_blk<<<ctas, PB_NT, PB_DYN_SMEM, s>>>(F, c, factor);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (P.cw < F.W || P.ch < F.H) e = launch_step_uncovered(F, c, factor, s);
    return e;
}

}  // namespace j2p
