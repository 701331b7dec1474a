// kernels_strip.cu — the two exchanges of a strip iteration over NVLink peer memory.
//
// Row strips of one frame on several GPUs need, per iteration, the sum of g^2
// of every rank (three doubles each) between k_gradient and k_project, and the two border rows
// of the new iterate from each neighbour before the next k_gradient.  Both are tiny and latency
// bound.  Instead of an NCCL launch each, the ranks store straight into each other's memory
// (cudaIpc mappings, NVLink) and synchronise with sequence-numbered flags:
//
//   k_sums_exchange  one warp: lane p writes this rank's three sums into rank p's mailbox and
//                    releases a flag there; lane r then acquires the flag rank r set here; the
//                    sums are folded in rank order (deterministic, identical on every rank) into
//                    the norms and reciprocals k_project reads (compute.c:200-206).
//   k_halo_exchange  a few CTAs copy this strip's first / last two rows of every plane into the
//                    neighbours' halo rows, the last CTA to finish releases a flag on each
//                    neighbour, then waits for the neighbours' flags.
//
// Ordering argument (why two mailbox slots and one halo flag per side suffice) is in DESIGN.md §7.
// Opt-in in round 1 (J2P_STRIP_P2P=1): written when the round's multi-GPU budget was spent, to be
// validated against the NCCL path (tests/test_gpu_strips.py, J2P_TEST_P2P=1) before it becomes the default.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"

namespace j2p {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p == want; gives up after ~4 s of GPU clock and reports through *err (never hangs a box)
__device__ __forceinline__ bool wait_flag(const unsigned *p, unsigned want, int *err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) != want) {
        if (clock64() - t0 > 8000000000ll) {
            atomicExch(err, 1);
            return false;
        }
        __nanosleep(100);
    }
    return true;
}

__global__ void k_sums_exchange(const __grid_constant__ StripPeers P, const double *my_sums, double *my_mail, unsigned *my_flag,
                                unsigned seq, int nc, float *norms, int *err) {
    const int lane = threadIdx.x, slot = (int)(seq & 1u);
    if (lane < P.nranks) {
        double *dst = P.mail[lane] + ((size_t)slot * P.nranks + P.rank) * 4;
        dst[0] = my_sums[0];
        dst[1] = my_sums[1];
        dst[2] = my_sums[2];
        __threadfence_system();
        st_release_sys(P.mail_flag[lane] + slot * P.nranks + P.rank, seq);
    }
    bool ok = true;
    if (lane < P.nranks) ok = wait_flag(my_flag + slot * P.nranks + lane, seq, err);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok || lane >= nc) return;
    double s = 0.;
    for (int r = 0; r < P.nranks; r++) s = __dadd_rn(s, __ldcv(my_mail + ((size_t)slot * P.nranks + r) * 4 + lane));
    const float norm = fsqrt(__double2float_rn(s));                       // compute.c:205
    norms[lane] = norm;
    norms[4 + lane] = __frcp_rn(norm);
}

__global__ void k_halo_exchange(const __grid_constant__ HaloPeers P, unsigned seq, unsigned *ticket, int *err) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (int c = 0; c < P.nc; c++) {
        if (P.has_up) {
            const float4 *s = reinterpret_cast<const float4 *>(P.up_src[c]);
            float4 *d = reinterpret_cast<float4 *>(P.up_dst[c]);
            for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n4; i += stride) d[i] = s[i];
        }
        if (P.has_down) {
            const float4 *s = reinterpret_cast<const float4 *>(P.down_src[c]);
            float4 *d = reinterpret_cast<float4 *>(P.down_dst[c]);
            for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n4; i += stride) d[i] = s[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        *ticket = 0u;
        __threadfence_system();
        if (P.has_up) st_release_sys(P.up_flag, seq);
        if (P.has_down) st_release_sys(P.down_flag, seq);
        if (P.has_up) wait_flag(P.from_up, seq, err);
        if (P.has_down) wait_flag(P.from_down, seq, err);
    }
}

cudaError_t launch_sums_exchange(const StripPeers &P, const double *my_sums, double *my_mail, unsigned *my_flag, unsigned seq, int nc,
                                 float *norms, int *err, cudaStream_t s) {
    k_sums_exchange<<<1, 32, 0, s>>>(P, my_sums, my_mail, my_flag, seq, nc, norms, err);
    return cudaGetLastError();
}

cudaError_t launch_halo_exchange(const HaloPeers &P, unsigned seq, unsigned *ticket, int *err, cudaStream_t s) {
    k_halo_exchange<<<8, 256, 0, s>>>(P, seq, ticket, err);
    return cudaGetLastError();
}

}  // namespace j2p
