// tools/microbench2.cu — what bounds k_project?  fp64 conversion rates in isolation, mixed with
// shared-memory traffic, and two complete 8x8 transform organisations (8 threads per block with
// shared-memory transposes vs one thread per block entirely in registers).
// Not part of the product; results land in profiles/.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../jpeg2png_b200/csrc/numerics.cuh"

using namespace j2p;

#define ITERS 2048
#define ILP 8

__global__ void __launch_bounds__(256) k_f2d(float *out, float seed) {       // f32->f64 only (result folded with integer ops)
    float v[ILP];
    for (int i = 0; i < ILP; i++) v[i] = seed + threadIdx.x * 1e-3f + i;
    unsigned acc = 0;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const double d = (double)v[i];
            acc ^= (unsigned)__double2hiint(d);
            v[i] = __uint_as_float(__float_as_uint(v[i]) + 1u);
        }
    if (acc == 0x12345u) out[0] = 1.f;
}
__global__ void __launch_bounds__(256) k_d2f(float *out, float seed) {       // f64->f32 only
    double v[ILP];
    for (int i = 0; i < ILP; i++) v[i] = (double)seed + threadIdx.x * 1e-3 + i;
    unsigned acc = 0;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const float f = __double2float_rn(v[i]);
            acc ^= __float_as_uint(f);
            v[i] = __hiloint2double(__double2hiint(v[i]), __double2loint(v[i]) + 0x20000000);
        }
    if (acc == 0x12345u) out[0] = 1.f;
}
__global__ void __launch_bounds__(256) k_pair(float *out, float seed) {      // f2d + dmul + d2f (one dscale)
    float v[ILP];
    for (int i = 0; i < ILP; i++) v[i] = seed + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = dscale(1.0000001, v[i]);
    float s = 0;
    for (int i = 0; i < ILP; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}
// the same with K2's proportion of shared-memory traffic mixed in: per 8 dscales (16 conversions)
// one 8x8 transpose step (2 STS.128 + 8 LDS.32 per thread)
__global__ void __launch_bounds__(256) k_pair_smem(float *out, float seed) {
    __shared__ __align__(16) float tiles[32][72];
    float v[8];
    const int b = threadIdx.x >> 3, j = threadIdx.x & 7;
    for (int i = 0; i < 8; i++) v[i] = seed + threadIdx.x * 1e-3f + i;
    float *tile = tiles[b];
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = dscale(1.0000001, v[i]);
        const int h = (j >> 2) & 1;
        float4 *row = reinterpret_cast<float4 *>(tile + j * 8);
        row[h] = make_float4(v[0], v[1], v[2], v[3]);
        row[h ^ 1] = make_float4(v[4], v[5], v[6], v[7]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))];
        __syncwarp();
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_smem_only(float *out, float seed) {
    __shared__ __align__(16) float tiles[32][72];
    float v[8];
    const int b = threadIdx.x >> 3, j = threadIdx.x & 7;
    for (int i = 0; i < 8; i++) v[i] = seed + threadIdx.x * 1e-3f + i;
    float *tile = tiles[b];
    for (int it = 0; it < ITERS; it++) {
        const int h = (j >> 2) & 1;
        float4 *row = reinterpret_cast<float4 *>(tile + j * 8);
        row[h] = make_float4(v[0], v[1], v[2], v[3]);
        row[h ^ 1] = make_float4(v[4], v[5], v[6], v[7]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))] + 1.0f;
        __syncwarp();
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

// ---- complete transforms: forward + inverse + inverse (what k_project does per block) --------
__device__ __forceinline__ void t_r2c(float (&v)[8], float *tile, int j) {
    const int h = (j >> 2) & 1;
    float4 *row = reinterpret_cast<float4 *>(tile + j * 8);
    row[h] = make_float4(v[0], v[1], v[2], v[3]);
    row[h ^ 1] = make_float4(v[4], v[5], v[6], v[7]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))];
    __syncwarp();
}
__device__ __forceinline__ void t_c2r(float (&v)[8], float *tile, int j) {
#pragma unroll
    for (int i = 0; i < 8; i++) tile[i * 8 + (j ^ (((i >> 2) & 1) << 2))] = v[i];
    __syncwarp();
    const int h = (j >> 2) & 1;
    const float4 *row = reinterpret_cast<const float4 *>(tile + j * 8);
    const float4 lo = row[h], hi = row[h ^ 1];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    __syncwarp();
}
#define XITERS 64
__global__ void __launch_bounds__(256) k_xform_8thr(float *out, float seed) {     // 8 threads / block
    __shared__ __align__(16) float tiles[32][72];
    const int b = threadIdx.x >> 3, j = threadIdx.x & 7;
    float v[8], r[8];
    for (int i = 0; i < 8; i++) { v[i] = seed + threadIdx.x * 1e-3f + i; }
    float *tile = tiles[b];
    for (int it = 0; it < XITERS; it++) {
        t_r2c(v, tile, j); fdct8(v); t_c2r(v, tile, j); fdct8(v);
        for (int i = 0; i < 8; i++) r[i] = v[i] * 0.5f;
        t_r2c(v, tile, j); idct8(v); t_c2r(v, tile, j); idct8(v);
        t_r2c(r, tile, j); idct8(r); t_c2r(r, tile, j); idct8(r);
        for (int i = 0; i < 8; i++) v[i] += r[i];
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}
// one thread per block, whole block in registers, no shared memory
__device__ __forceinline__ void col_pass_f(float (&a)[64]) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = a[k * 8 + c];
        fdct8(t);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k * 8 + c] = t[k];
    }
}
__device__ __forceinline__ void row_pass_f(float (&a)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = a[r * 8 + k];
        fdct8(t);
#pragma unroll
        for (int k = 0; k < 8; k++) a[r * 8 + k] = t[k];
    }
}
__device__ __forceinline__ void col_pass_i(float (&a)[64]) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = a[k * 8 + c];
        idct8(t);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k * 8 + c] = t[k];
    }
}
__device__ __forceinline__ void row_pass_i(float (&a)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = a[r * 8 + k];
        idct8(t);
#pragma unroll
        for (int k = 0; k < 8; k++) a[r * 8 + k] = t[k];
    }
}
__global__ void __launch_bounds__(128) k_xform_1thr(float *out, float seed) {
    float v[64], r[64];
#pragma unroll
    for (int i = 0; i < 64; i++) v[i] = seed + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < XITERS; it++) {
        col_pass_f(v); row_pass_f(v);
#pragma unroll
        for (int i = 0; i < 64; i++) r[i] = v[i] * 0.5f;
        col_pass_i(v); row_pass_i(v);
        col_pass_i(r); row_pass_i(r);
#pragma unroll
        for (int i = 0; i < 64; i++) v[i] += r[i];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

template <typename K>
float time_kernel(K k, int blocks, int threads, float *out) {
    k<<<blocks, threads>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 3; r++) k<<<blocks, threads>>>(out, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    float *out; cudaMalloc(&out, 4);
    const int blocks = sms * 8;
    auto rate = [&](const char *name, float ms, double ops_per_thread, int threads, int nblocks) {
        const double ops = (double)nblocks * threads * ops_per_thread;
        printf("%-34s %8.3f ms  %7.2f op/clk/SM\n", name, ms, ops / (ms * 1e-3) / sms / (ghz * 1e9));
    };
    rate("F2F.F64.F32 only", time_kernel(k_f2d, blocks, 256, out), (double)ITERS * ILP, 256, blocks);
    rate("F2F.F32.F64 only", time_kernel(k_d2f, blocks, 256, out), (double)ITERS * ILP, 256, blocks);
    rate("dscale (f2d+dmul+d2f), per cvt", time_kernel(k_pair, blocks, 256, out), (double)ITERS * ILP * 2, 256, blocks);
    rate("dscale + 8x8 transpose, per cvt", time_kernel(k_pair_smem, blocks, 256, out), (double)ITERS * 8 * 2, 256, blocks);
    rate("8x8 transpose only, per 16 'cvt'", time_kernel(k_smem_only, blocks, 256, out), (double)ITERS * 8 * 2, 256, blocks);
    {
        const float ms = time_kernel(k_xform_8thr, blocks, 256, out);
        const double blocks8 = (double)blocks * 32 * XITERS;       // 8x8 blocks transformed (x3 transforms each)
        printf("%-34s %8.3f ms  %7.2f Gblock/s  (k_project needs 0.39 Gblock per 4K plane-iteration... %6.1f us/plane)\n",
               "fdct+2 idct, 8 thr/block, smem", ms, blocks8 / (ms * 1e-3) / 1e9, 129600.0 / (blocks8 / (ms * 1e-3)) * 1e6);
    }
    {
        const int nb = sms * 24;
        const float ms = time_kernel(k_xform_1thr, nb, 128, out);
        const double blocks8 = (double)nb * 128 * XITERS;
        printf("%-34s %8.3f ms  %7.2f Gblock/s  (%6.1f us/plane)\n", "fdct+2 idct, 1 thr/block, regs", ms,
               blocks8 / (ms * 1e-3) / 1e9, 129600.0 / (blocks8 / (ms * 1e-3)) * 1e6);
    }
    return 0;
}
