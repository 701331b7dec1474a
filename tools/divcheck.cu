// tools/divcheck.cu — brute-force check that the shared-reciprocal divisions used by the kernels
// (numerics.cuh: qdiv_fast = the five-operation sequence; qdiv4_core / qdiv2x = the four-operation
// sequence on a two-term reciprocal, scalar and packed) return exactly div.rn.f32 on every operand
// pair inside their guard.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../jpeg2png_b200/csrc/numerics.cuh"

__device__ __forceinline__ uint32_t rng(uint64_t &s) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

// mode 0: random mantissas, exponents of a in [-58,58], b in [-38,38]
// mode 1: divisor mantissa near all-ones / all-zeros (the classical hard cases for reciprocal-
//         based division), numerator mantissa random or extreme
// mode 2: quotients near k/1024 with the numerator nudged by a few ulps (near-tie hunting)
// mode 3: quotients near a MIDPOINT of two adjacent floats (a 25-bit odd significand): a = RN(b*m)
//         nudged by a few ulps — the operands on which a faithful-but-not-correct rounding shows
__global__ void check(int mode, uint64_t seed, unsigned long long *bad, unsigned long long *n_fast, float *ex) {
    uint64_t s = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1);
    for (int w = 0; w < 8; w++) rng(s);
    unsigned long long nb = 0, nf = 0;
    for (int it = 0; it < 2048; it++) {
        uint32_t ma = rng(s) & 0x7fffff, mb = rng(s) & 0x7fffff;
        int ea = 127 + (int)(rng(s) % 117) - 58, eb = 127 + (int)(rng(s) % 77) - 38;
        if (mode == 1) {
            uint32_t r = rng(s);
            mb = (r & 1) ? (0x7fffff - (rng(s) & 0x3f)) : (rng(s) & 0x3f);
            if (r & 2) ma = (r & 4) ? (0x7fffff - (rng(s) & 0xff)) : (rng(s) & 0xff);
        }
        float b = __uint_as_float(((uint32_t)eb << 23) | mb);
        float a = __uint_as_float(((rng(s) & 1u) << 31) | ((uint32_t)ea << 23) | ma);
        if (mode == 2) {
            float k = (float)(1 + (rng(s) % 4096)) * (1.0f / 1024.0f);
            a = b * k;
            a = __uint_as_float(__float_as_uint(a) + (rng(s) % 5) - 2);
        }
        if (mode == 3) {
            const double m = ((double)(0x800000u | (rng(s) & 0x7fffff)) + 0.5) * 1.1920928955078125e-07;   // in [1, 2): halfway between two floats
            a = (float)((double)b * m);
            a = __uint_as_float(__float_as_uint(a) + (rng(s) % 5) - 2);
        }
        if (!j2p::qdiv_divisor_ok(b)) continue;
        const float y = __frcp_rn(b);
        bool ok = true;
        const float q = j2p::qdiv_fast(a, b, y, ok);
        if (!ok) continue;
        nf++;
        const float t = __fdiv_rn(a, b);
        // the four-operation sequence, scalar and packed (the other half carries an unrelated pair)
        const float yl = j2p::rcp_low(b, y);
        const float q4 = j2p::qdiv4_core(a, b, y, yl);
        const j2p::f2 nb2 = j2p::pk(-b, -1.0f), y2 = j2p::pk(y, 1.0f);
        const j2p::f2 q42 = j2p::qdiv2x(j2p::pk(a, 3.0f), nb2, y2, j2p::rcp2_low(nb2, y2));
        const bool z = t == 0.f;                                   // the sign of a zero quotient is not part of the contract
        if ((__float_as_uint(q4) != __float_as_uint(t) && !(z && q4 == 0.f)) || __float_as_uint(j2p::lo(q42)) != __float_as_uint(q4) || j2p::hi(q42) != 3.0f) {
            if (nb == 0) { ex[0] = a; ex[1] = b; ex[2] = q4; ex[3] = t; }
            nb++;
        }
        if (__float_as_uint(q) != __float_as_uint(t) && !(z && q == 0.f)) {
            if (nb == 0) { ex[0] = a; ex[1] = b; ex[2] = q; ex[3] = t; }
            nb++;
        }
    }
    if (nb) atomicAdd(bad, nb);
    atomicAdd(n_fast, nf);
}

int main() {
    unsigned long long *bad, *nf; float *ex;
    cudaMallocManaged(&bad, 8); cudaMallocManaged(&nf, 8); cudaMallocManaged(&ex, 16);
    int rc = 0;
    for (int mode = 0; mode < 4; mode++) {
        *bad = 0; *nf = 0;
        for (int rep = 0; rep < 24; rep++) check<<<148 * 16, 256>>>(mode, 0x1234567ull + rep * 7919 + mode * 104729, bad, nf, ex);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %llu fast-path quotients checked, %llu mismatches (%s)\n", mode, *nf, *bad, cudaGetErrorString(e));
        if (*bad) { printf("  example a=%a b=%a fast=%a exact=%a\n", ex[0], ex[1], ex[2], ex[3]); rc = 1; }
    }
    return rc;
}
