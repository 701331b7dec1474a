// tools/microbench.cu — per-SM issue rates of the operations the bit-exact solver is made of.
// Not part of the product; numbers land in profiles/ and drive kernel design decisions.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ILP 8
#define ITERS 4096

template <typename Op>
__global__ void __launch_bounds__(256) k(float *out, float seed, Op op) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = seed + (float)(threadIdx.x * ILP + i) * 1e-3f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = op(v[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

struct OpFadd { __device__ float operator()(float a) const { return __fadd_rn(a, 1.0009765625f); } };
struct OpFmul { __device__ float operator()(float a) const { return __fmul_rn(a, 1.0000001f); } };
struct OpDiv { float d; __device__ float operator()(float a) const { return __fdiv_rn(a, d); } };
struct OpSqrt { __device__ float operator()(float a) const { return __fsqrt_rn(a) + 1.0f; } };
struct OpCvtRound { __device__ float operator()(float a) const { return __double2float_rn(__dadd_rn((double)a, 1e-3)); } };   // f2d + dadd + d2f
struct OpCvtOnly { __device__ float operator()(float a) const {                       // f2d + d2f pairs without fp64 math
        double d = (double)a; long long b = __double_as_longlong(d) ^ 0x10000000ll; return __double2float_rn(__longlong_as_double(b)); } };
struct OpDadd2 { __device__ float operator()(float a) const {                          // one f2d, 8 dadd, one d2f
        double d = (double)a;
#pragma unroll
        for (int i = 0; i < 8; i++) d = __dadd_rn(d, 1e-3);
        return __double2float_rn(d); } };
struct OpDmul2 { __device__ float operator()(float a) const {
        double d = (double)a;
#pragma unroll
        for (int i = 0; i < 8; i++) d = __dmul_rn(d, 1.0000001);
        return __double2float_rn(d); } };
struct OpShfl { __device__ float operator()(float a) const { return __shfl_xor_sync(0xffffffffu, a, 1) + 1.0f; } };
struct OpMark { float d, r; __device__ float operator()(float a) const {              // Markstein division with shared reciprocal
        float q = __fmul_rn(a, r); float rem = __fmaf_rn(-q, d, a); return __fmaf_rn(rem, r, q); } };
struct OpRot { __device__ float operator()(float a) const {                           // one DCT rotation pair: 2 f2d, 4 dmul, 2 dadd, 2 d2f
        const double du = (double)a, dv = (double)(a + 1.0f);
        float p = __double2float_rn(__dsub_rn(__dmul_rn(0.49, du), __dmul_rn(0.097, dv)));
        float q = __double2float_rn(__dadd_rn(__dmul_rn(0.49, dv), __dmul_rn(0.097, du)));
        return p + q; } };

template <typename Op>
void run(const char *name, Op op, double ops_per_call, int sms, double clk_ghz) {
    float *out; cudaMalloc(&out, 4);
    const int blocks = sms * 8;
    k<<<blocks, 256>>>(out, 1.0f, op);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) k<<<blocks, 256>>>(out, 1.0f, op);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const double calls = (double)blocks * 256 * ITERS * ILP;
    const double per_sm_clk = calls * ops_per_call / (ms * 1e-3) / sms / (clk_ghz * 1e9);
    printf("%-28s %8.3f ms  %8.2f Gcall/s  ~%6.1f op/clk/SM (at %.2f GHz, %g op/call)\n", name, ms, calls / ms * 1e-6, per_sm_clk, clk_ghz, ops_per_call);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("%s  SMs=%d  max clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    run("fadd.rn", OpFadd(), 1, sms, ghz);
    run("fmul.rn", OpFmul(), 1, sms, ghz);
    run("fdiv.rn (IEEE)", OpDiv{1.0000001f}, 1, sms, ghz);
    run("fsqrt.rn (IEEE) + fadd", OpSqrt(), 1, sms, ghz);
    run("markstein div (mul+2fma)", OpMark{1.0000001f, 1.0f / 1.0000001f}, 1, sms, ghz);
    run("f2d + dadd + d2f", OpCvtRound(), 1, sms, ghz);
    run("f2d + lop + d2f", OpCvtOnly(), 1, sms, ghz);
    run("f2d + 8 dadd + d2f", OpDadd2(), 8, sms, ghz);
    run("f2d + 8 dmul + d2f", OpDmul2(), 8, sms, ghz);
    run("dct rotation pair", OpRot(), 1, sms, ghz);
    run("shfl.xor + fadd", OpShfl(), 1, sms, ghz);
    return 0;
}
