// kernels_strip.cu — the stand-alone halo kernel of the strip protocol over NVLink peer memory.
//
// Row strips of one frame on several GPUs exchange, per iteration, the sum of g^2
// of every rank (three doubles each) between the gradient and the projection, and the two border
// rows of the new iterate with each neighbour before the next gradient.  Both exchanges are part
// of the two solver kernels (strip_sync.cuh): stores into the peers' memory (cudaIpc mappings,
// NVLink) plus sequence-numbered flags; nothing of it is a launch on the critical path.
//
// This file holds what is left outside the solver kernels:
//   k_halo_exchange   copies this strip's first / last two rows of every plane into the
//                     neighbours' halo rows and raises their flags.  Used (a) once per solve for
//                     the halo rows of the initial iterate, waiting for the neighbours' rows to
//                     arrive before it ends, and (b) every iteration, without waiting, for frames
//                     whose projection kernels cannot deliver the rows themselves (a plane that
//                     does not span the frame width, or sampling factors other than 1x1 / 2x2).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "numerics.cuh"
#include "strip_sync.cuh"

namespace j2p {

__global__ void k_halo_exchange(const __grid_constant__ HaloPeers P, unsigned seq, unsigned *ticket, int *err, int wait_for_arrival) {
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
        if (wait_for_arrival) {
            if (P.has_up) wait_seq(P.from_up, seq, err);
            if (P.has_down) wait_seq(P.from_down, seq, err);
        }
    }
}

cudaError_t launch_halo_exchange(const HaloPeers &P, unsigned seq, unsigned *ticket, int *err, int wait_for_arrival, cudaStream_t s) {
    k_halo_exchange<<<8, 256, 0, s>>>(P, seq, ticket, err, wait_for_arrival);
    return cudaGetLastError();
}

}  // namespace j2p
