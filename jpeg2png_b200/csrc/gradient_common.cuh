// gradient_common.cuh — pieces shared by the two builds of the sub-gradient kernel
// (kernels_gradient_packed.cu: the production kernel on packed fp32; kernels_gradient.cu: the
// scalar kernel, which also sums the objective terms for -c csv logging).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"

namespace j2p {

#ifndef J2P_GM_WARPS
#define J2P_GM_WARPS 4
#endif
constexpr int GM_WARPS = J2P_GM_WARPS, GM_NT = GM_WARPS * 32, GM_USE = 60;
#ifndef J2P_GM_DEPTH
#define J2P_GM_DEPTH 4
#endif
constexpr int GM_DEPTH = J2P_GM_DEPTH;      // packed kernel: rows in flight per warp (cp.async ring in shared memory); a power of two
#ifndef J2P_GRAD_MIN_CTAS
#define J2P_GRAD_MIN_CTAS 2     // resident CTAs per SM the register allocation is bounded for.  Measured on the one-block row step at 4K
                                // (profiles/r02_ab_gradient_geometry.txt): 2 CTAs (194 registers, 8 warps/SM) 117.6 us, 3 CTAs (162 registers,
                                // 12 warps) 123.3 us, 4 CTAs (128 registers, spills) 134.9 us — the kernel is bound by dependency latency, and
                                // registers for the scheduler to overlap chains buy more than resident warps do
#endif

// IEEE fallbacks of the two quotient stages, for rows the guards reject (numerics.cuh).  Out of line
// and fed through local memory on purpose: they run once in millions of rows, and kept inline they
// cost the hot path register copies at every row.
template <int NC>
struct TvSlow {
    float gx[NC][2], gy[NC][2];     // in: forward differences of the source row
    float q[3][NC][2];              // out: self / right / below quotients (compute.c:98-103)
    float n[2];                     // out: the norms (for the objective log)
};
// A value read back from local memory after the fallback call is passed through one ALU
// instruction inside the cold branch.  Without it the first instruction after the join waits on
// the scoreboard the compiler gave those local loads — the same one the row prefetch uses — and
// every row stalls there until its prefetch has landed (15 % of all stall samples, 5 us per 4K
// iteration; profiles/r01_notes.md).  `zero` is 0, but not to the compiler.
__device__ __forceinline__ float settle(float v, unsigned zero) { return __uint_as_float(__float_as_uint(v) ^ zero); }
template <int NC>
__device__ __noinline__ void tv_slow(TvSlow<NC> *io, float a1, bool src_in) {
    for (int k = 0; k < 2; k++) {
        float ssq = 0.f;
        for (int c = 0; c < NC; c++) ssq = fadd(fadd(ssq, fsq(io->gx[c][k])), fsq(io->gy[c][k]));   // compute.c:84-89
        const float n = fsqrt(ssq);
        const bool live = src_in && n != 0.f;                                                       // compute.c:97
        io->n[k] = n;
        for (int c = 0; c < NC; c++) {
            const float gx = io->gx[c][k], gy = io->gy[c][k];
            io->q[0][c][k] = live ? fdiv(fmul(a1, -fadd(gx, gy)), n) : 0.f;
            io->q[1][c][k] = live ? fdiv(fmul(a1, gx), n) : 0.f;
            io->q[2][c][k] = live ? fdiv(fmul(a1, gy), n) : 0.f;
        }
    }
}
template <int NC>
struct TgvSlow {
    float gxx[NC][2], gyy[NC][2], sym[NC][2];   // in: second differences of the source row
    float q[4][NC][2];                          // out: a2 * (self / left-right / up-down / diagonal quotients) (compute.c:165-182)
    float n[2];
};
template <int NC>
__device__ __noinline__ void tgv_slow(TgvSlow<NC> *io, float a2, bool src_in) {
    for (int k = 0; k < 2; k++) {
        float ssq = 0.f;
        for (int c = 0; c < NC; c++)
            ssq = fadd(ssq, fadd(fadd(fsq(io->gxx[c][k]), fmul(2.f, fsq(io->sym[c][k]))), fsq(io->gyy[c][k])));   // compute.c:148-152
        const float n = fsqrt(ssq);
        const bool live = src_in && n != 0.f;                                                       // compute.c:158
        io->n[k] = n;
        for (int c = 0; c < NC; c++) {
            const float gxx = io->gxx[c][k], gyy = io->gyy[c][k], sym = io->sym[c][k];
            const float self = -fadd(fadd(fmul(2.f, gxx), fmul(2.f, sym)), fmul(2.f, gyy));
            io->q[0][c][k] = live ? fmul(a2, fdiv(self, n)) : 0.f;
            io->q[1][c][k] = live ? fmul(a2, fdiv(fadd(sym, gxx), n)) : 0.f;
            io->q[2][c][k] = live ? fmul(a2, fdiv(fadd(gyy, sym), n)) : 0.f;
            io->q[3][c][k] = live ? fmul(a2, fdiv(-sym, n)) : 0.f;
        }
    }
}

}  // namespace j2p
