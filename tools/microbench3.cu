// tools/microbench3.cu — issue rates of Blackwell's packed fp32 instructions (FADD2 / FMUL2 / FFMA2,
// PTX add/mul/fma.rn.f32x2) next to their scalar forms, alone and mixed with ALU / XU work, and of
// the integer widening of a float to a double.  Round-2 design input for k_gradient / k_project;
// not part of the product.  Numbers land in profiles/.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ILP 8
#define ITERS 2048
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1},%2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0,%1},%2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// MODE: 0 scalar fadd x2 (two chains per slot), 1 FADD2, 2 scalar fmul x2, 3 FMUL2, 4 scalar ffma x2, 5 FFMA2,
//       6 scalar 5-op quotient x2, 7 packed 5-op quotient, 8 FFMA2 + 1 LOP3 per packed op, 9 FFMA2 + 2 LOP3,
//       10 scalar ffma x2 + 2 LOP3, 11 FFMA2 + shfl every 4th, 12 f2d by F2F, 13 f2d by integer ops
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float seed, float c0, float c1) {
    u64 v[ILP];
    unsigned w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        const float s = seed + (float)(threadIdx.x * ILP + i) * 1e-3f;
        v[i] = pk(s, s + 0.5f);
        w[i] = threadIdx.x * 77u + i;
    }
    const u64 k0 = pk(c0, c0), k1 = pk(c1, c1);
    double dacc = 0.;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) v[i] = pk(__fadd_rn(lo(v[i]), c0), __fadd_rn(hi(v[i]), c0));
            if (MODE == 1) v[i] = add2(v[i], k0);
            if (MODE == 2) v[i] = pk(__fmul_rn(lo(v[i]), c1), __fmul_rn(hi(v[i]), c1));
            if (MODE == 3) v[i] = mul2(v[i], k1);
            if (MODE == 4) v[i] = pk(__fmaf_rn(lo(v[i]), c1, c0), __fmaf_rn(hi(v[i]), c1, c0));
            if (MODE == 5) v[i] = fma2(v[i], k1, k0);
            if (MODE == 6) {
                float q[2] = {lo(v[i]), hi(v[i])};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float a = q[h], b = c1, y = c0;
                    const float q0 = __fmul_rn(a, y), r0 = __fmaf_rn(-b, q0, a), q1 = __fmaf_rn(r0, y, q0), r1 = __fmaf_rn(-b, q1, a);
                    q[h] = __fmaf_rn(r1, y, q1);
                }
                v[i] = pk(q[0], q[1]);
            }
            if (MODE == 7) {
                const u64 a = v[i], nb = pk(-c1, -c1), y = k0;
                const u64 q0 = mul2(a, y), r0 = fma2(nb, q0, a), q1 = fma2(r0, y, q0), r1 = fma2(nb, q1, a);
                v[i] = fma2(r1, y, q1);
            }
            if (MODE == 8) { v[i] = fma2(v[i], k1, k0); w[i] = (w[i] ^ 0x5bd1e995u) & (w[i] | 0x1234567u); }
            if (MODE == 9) { v[i] = fma2(v[i], k1, k0); w[i] = (w[i] ^ 0x5bd1e995u) & (w[i] | 0x1234567u); w[i] = (w[i] ^ 0x2545F491u) | (w[i] & 0x7654321u); }
            if (MODE == 10) { v[i] = pk(__fmaf_rn(lo(v[i]), c1, c0), __fmaf_rn(hi(v[i]), c1, c0)); w[i] = (w[i] ^ 0x5bd1e995u) & (w[i] | 0x1234567u); w[i] = (w[i] ^ 0x2545F491u) | (w[i] & 0x7654321u); }
            if (MODE == 11) { v[i] = fma2(v[i], k1, k0); if ((i & 3) == 0) v[i] = pk(__shfl_xor_sync(0xffffffffu, lo(v[i]), 1), hi(v[i])); }
            if (MODE == 12) { const float a = lo(v[i]); dacc = __dadd_rn(dacc, (double)a); v[i] = pk(__fadd_rn(a, c0), hi(v[i])); }
            if (MODE == 13) {
                const float a = lo(v[i]);
                const unsigned u = __float_as_uint(a);
                // normal finite non-zero input: exponent rebias 127 -> 1023, significand shifted by 29
                const unsigned hi32 = ((u >> 3) & 0x0fffffffu) + 0x38000000u | (u & 0x80000000u), lo32 = u << 29;
                dacc = __dadd_rn(dacc, __hiloint2double((int)hi32, (int)lo32));
                v[i] = pk(__fadd_rn(a, c0), hi(v[i]));
            }
        }
    }
    float s = (float)dacc;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += lo(v[i]) + hi(v[i]) + (float)w[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char *name, double flop_per_slot, int sms, double ghz) {
    float *out; cudaMalloc(&out, 4);
    const int blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(out, 1.0f, 1.0009765625f, 0.99951171875f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, 256>>>(out, 1.0f, 1.0009765625f, 0.99951171875f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const double slots = (double)blocks * 256 * ITERS * ILP;                 // per-thread loop slots
    const double warp_slots_per_clk_smsp = slots / 32 / (ms * 1e-3) / (sms * 4) / (ghz * 1e9);
    printf("%-44s %8.3f ms  %7.3f slots/clk/SMSP  %7.1f lane-op/clk/SM (%g per slot)\n", name, ms, warp_slots_per_clk_smsp,
           slots * flop_per_slot / (ms * 1e-3) / sms / (ghz * 1e9), flop_per_slot);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    const int sms = p.multiProcessorCount;
    printf("%s  SMs=%d  max clock %.3f GHz; a 'slot' = one loop slot of one thread (2 fp32 results unless noted)\n", p.name, sms, ghz);
    run<0>("2 x fadd.rn (scalar)", 2, sms, ghz);
    run<1>("add.rn.f32x2 (FADD2)", 2, sms, ghz);
    run<2>("2 x fmul.rn (scalar)", 2, sms, ghz);
    run<3>("mul.rn.f32x2 (FMUL2)", 2, sms, ghz);
    run<4>("2 x fma.rn (scalar)", 2, sms, ghz);
    run<5>("fma.rn.f32x2 (FFMA2)", 2, sms, ghz);
    run<6>("2 x 5-op quotient (scalar)", 2, sms, ghz);
    run<7>("5-op quotient on f32x2", 2, sms, ghz);
    run<8>("FFMA2 + 2 LOP3-class ALU ops", 2, sms, ghz);
    run<9>("FFMA2 + 4 LOP3-class ALU ops", 2, sms, ghz);
    run<10>("2 x ffma scalar + 4 LOP3-class ALU ops", 2, sms, ghz);
    run<11>("FFMA2 + shfl on every 4th", 2, sms, ghz);
    run<12>("f32->f64 by F2F + dadd + fadd", 1, sms, ghz);
    run<13>("f32->f64 by integer ops + dadd + fadd", 1, sms, ghz);
    return 0;
}
