// project_common.cuh — pieces shared by the step+projection kernels.
#pragma once
#include <cuda_runtime.h>

#include "numerics.cuh"

namespace j2p {

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

// The same on two adjacent pixels at once (packed fp32, numerics.cuh): half the issue slots.
//   y - step * q  ==  y + (-step) * q  bit for bit; the sum with the product goes through addm2().
// Branch-free: when the norm is zero (compute.c:211 skips the step) the constants are replaced so that
// the same instruction sequence leaves y untouched, bit for bit.  Then every g is +0 (a sum of squares
// is zero only if every term is, and the gradient kernel never produces -0: its first addend is
// fma(gp, mask, +0)), the quotient sequence on (g = +0, -b = -1, 1/b = 0) yields +0, the product with
// -0.0f is -0, and y + (-0) == y for every y including -0.
struct Stepper2 {
    f2 fac, nstep, nnorm, rn, one;
    __device__ __forceinline__ void init(const Stepper &s, float one_) {
        fac = splat(s.factor); one = splat(one_);
        nstep = splat(s.stepping ? -s.step : -0.0f);
        nnorm = splat(s.stepping ? -s.norm : -1.0f);
        rn = splat(s.stepping ? s.rn : 0.0f);
    }
    __device__ __forceinline__ f2 fast(f2 x, f2 xp, f2 g, unsigned &key) const {
        const f2 y = addm2(mul2(fac, sub2(x, xp)), x, one);
        key = min(key, min(qdiv_key(lo(g)), qdiv_key(hi(g))));
        return addm2(mul2(nstep, qdiv2(g, nnorm, rn)), y, one);
    }
};

// (float)d for |d| < 2^22 without the conversion pipe (XU, a quarter of the fp32 rate and the busiest
// pipe of the projection): 1.5 * 2^23 + d is exact as an integer add on the bit pattern, and
// subtracting 1.5 * 2^23 again is an exact fp32 subtraction.  Quantised coefficients are int16.
__device__ __forceinline__ float small_int_to_float(int d) { return __fsub_rn(__int_as_float(0x4B400000 + d), 12582912.0f); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------
// 8x8 block transposes among the 8 lanes that own one block (lane j holds row j).
// tile: 8 rows x 8 floats; element (r, c) lives at r*8 + (c ^ (((r>>2)&1)<<2)) — the two float4
// halves of rows 4..7 are swapped, which makes both the 128-bit row accesses and the scalar
// column accesses bank-conflict free (tiles of the four blocks of a warp are 72 floats apart).
// ------------------------------------------------------------------------------------------
constexpr int TILE_STRIDE = 72;

__device__ __forceinline__ void rows_to_cols(float (&v)[8], float *tile, int j, unsigned mask = 0xffffffffu) {
    const int h = (j >> 2) & 1;
    float4 *row = reinterpret_cast<float4 *>(tile + j * 8);
    row[h] = make_float4(v[0], v[1], v[2], v[3]);
    row[h ^ 1] = make_float4(v[4], v[5], v[6], v[7]);
    __syncwarp(mask);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))];
    __syncwarp(mask);
}
__device__ __forceinline__ void cols_to_rows(float (&v)[8], float *tile, int j, unsigned mask = 0xffffffffu) {
#pragma unroll
    for (int i = 0; i < 8; i++) tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))] = v[i];
    __syncwarp(mask);
    const int h = (j >> 2) & 1;
    const float4 *row = reinterpret_cast<const float4 *>(tile + j * 8);
    const float4 lo = row[h], hi = row[h ^ 1];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    __syncwarp(mask);
}

// 2-D transforms for a thread that holds row j of the block and ends holding row j.
// Vertical pass first, horizontal second (ooura/dct.c:39-94, :103-158).
__device__ __forceinline__ void fdct8x8_rows(float (&v)[8], float *tile, int j, unsigned mask = 0xffffffffu) {
    rows_to_cols(v, tile, j, mask);
    fdct8(v);
    cols_to_rows(v, tile, j, mask);
    fdct8(v);
}
__device__ __forceinline__ void idct8x8_rows(float (&v)[8], float *tile, int j, unsigned mask = 0xffffffffu) {
    rows_to_cols(v, tile, j, mask);
    idct8(v);
    cols_to_rows(v, tile, j, mask);
    idct8(v);
}
// Two independent inverse transforms in lockstep (separate tiles, shared warp barriers): the
// fp64 conversions of one fill the XU latency of the other.
__device__ __forceinline__ void idct8x8_rows_x2(float (&a)[8], float (&b)[8], float *tile_a, float *tile_b, int j) {
    const int h = (j >> 2) & 1;
    {
        float4 *ra = reinterpret_cast<float4 *>(tile_a + j * 8), *rb = reinterpret_cast<float4 *>(tile_b + j * 8);
        ra[h] = make_float4(a[0], a[1], a[2], a[3]);
        ra[h ^ 1] = make_float4(a[4], a[5], a[6], a[7]);
        rb[h] = make_float4(b[0], b[1], b[2], b[3]);
        rb[h ^ 1] = make_float4(b[4], b[5], b[6], b[7]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int o = i * 8 + (j ^ (((i >> 2) & 1) << 2));
        a[i] = tile_a[o];
        b[i] = tile_b[o];
    }
    __syncwarp();
    idct8(a);
    idct8(b);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int o = i * 8 + (j ^ (((i >> 2) & 1) << 2));
        tile_a[o] = a[i];
        tile_b[o] = b[i];
    }
    __syncwarp();
    {
        const float4 *ra = reinterpret_cast<const float4 *>(tile_a + j * 8), *rb = reinterpret_cast<const float4 *>(tile_b + j * 8);
        const float4 al = ra[h], ah = ra[h ^ 1], bl = rb[h], bh = rb[h ^ 1];
        a[0] = al.x; a[1] = al.y; a[2] = al.z; a[3] = al.w; a[4] = ah.x; a[5] = ah.y; a[6] = ah.z; a[7] = ah.w;
        b[0] = bl.x; b[1] = bl.y; b[2] = bl.z; b[3] = bl.w; b[4] = bh.x; b[5] = bh.y; b[6] = bh.z; b[7] = bh.w;
    }
    __syncwarp();
    idct8(a);
    idct8(b);
}


}  // namespace j2p
