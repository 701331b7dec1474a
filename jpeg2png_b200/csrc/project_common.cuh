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

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

}  // namespace j2p
