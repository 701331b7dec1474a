// tools/rootcheck.cu — exhaustive check of numerics.cuh sqrt_core / rcp_core against
// sqrt.rn.f32 / rcp.rn.f32: every fp32 significand (2^23) at a spread of exponents inside the
// guarded range [2^-80, 2^80].  Also checks the composition used by the gradient kernel:
// n = sqrt_core(s), y = rcp_core(n).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../jpeg2png_b200/csrc/numerics.cuh"

__global__ void check(int exp_biased, unsigned long long *bad_sqrt, unsigned long long *bad_rcp, unsigned long long *bad_comp, float *ex) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;      // 0 .. 2^23-1
    if (m >= (1u << 23)) return;
    const float x = __uint_as_float(((uint32_t)exp_biased << 23) | m);
    const float s1 = j2p::sqrt_core(x), s2 = __fsqrt_rn(x);
    if (__float_as_uint(s1) != __float_as_uint(s2)) { atomicAdd(bad_sqrt, 1ull); ex[0] = x; ex[1] = s1; ex[2] = s2; }
    const float r1 = j2p::rcp_core(x), r2 = __frcp_rn(x);
    if (__float_as_uint(r1) != __float_as_uint(r2)) { atomicAdd(bad_rcp, 1ull); ex[3] = x; ex[4] = r1; ex[5] = r2; }
    const float c1 = j2p::rcp_core(s1), c2 = __frcp_rn(s2);
    if (__float_as_uint(c1) != __float_as_uint(c2)) atomicAdd(bad_comp, 1ull);
}

int main() {
    unsigned long long *bad; float *ex;
    cudaMallocManaged(&bad, 24); cudaMallocManaged(&ex, 32);
    bad[0] = bad[1] = bad[2] = 0;
    int n_exp = 0;
    for (int e = 127 - 80; e <= 127 + 79; e += 1) {        // every exponent of the guarded range
        check<<<(1 << 23) / 256, 256>>>(e, bad, bad + 1, bad + 2, ex);
        n_exp++;
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("%d exponents x 2^23 significands: sqrt mismatches %llu, rcp mismatches %llu, rcp(sqrt) mismatches %llu (%s)\n",
           n_exp, bad[0], bad[1], bad[2], cudaGetErrorString(err));
    if (bad[0]) printf("  sqrt example x=%a core=%a rn=%a\n", ex[0], ex[1], ex[2]);
    if (bad[1]) printf("  rcp  example x=%a core=%a rn=%a\n", ex[3], ex[4], ex[5]);
    return (bad[0] || bad[1] || bad[2]) ? 1 : 0;
}
