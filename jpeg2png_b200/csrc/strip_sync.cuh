// strip_sync.cuh — device side of the strip exchanges over NVLink peer memory (StripSync in
// kernels.cuh, protocol and ordering argument in DESIGN.md §7).  Included by the gradient and the
// projection kernels: the exchanges are part of those kernels, there is no launch of their own on
// the critical path.
#pragma once
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
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Spin until the sequence number at *p has reached `want` (sequence numbers only grow; the
// comparison is wrap-safe).  Gives up after ~2 s of GPU clock and reports through *err: a lost peer
// must never hang a box.
__device__ __forceinline__ bool wait_seq(const unsigned *p, unsigned want, int *err) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - want) < 0) {
        if (clock64() - t0 > 4000000000ll) {
            atomicExch(err, 1);
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

// k_gradient, first / last row band of a strip: the halo rows of x_k come from the neighbours'
// projection of the previous iteration.  Called by every thread of the CTA before its first load.
__device__ __forceinline__ void strip_wait_halo(const StripSync &S, bool top_band, bool bottom_band) {
    if (S.nranks <= 1) return;
    if (threadIdx.x == 0) {
        if (top_band && S.has_up) wait_seq(S.from_up, S.halo_seq, S.err);
        if (bottom_band && S.has_down) wait_seq(S.from_down, S.halo_seq, S.err);
    }
    __syncthreads();
}

// k_gradient, last CTA: this rank's sums of g^2 to every rank's mailbox (own included).
// Called by the first `nranks` threads of the CTA after `fin[0..2]` (shared memory) are complete.
__device__ __forceinline__ void strip_post_sums(const StripSync &S, const double *fin, int tid) {
    if (tid >= S.nranks) return;
    const int slot = (int)(S.seq & 1u);
    double *dst = S.mail[tid] + ((size_t)slot * S.nranks + S.rank) * 4;
    dst[0] = fin[0];
    dst[1] = fin[1];
    dst[2] = fin[2];
    // the release store orders this thread's three stores before the flag (same thread, same peer):
    // no separate system-scope fence on the critical path of every iteration
    st_release_sys(S.mail_flag[tid] + slot * S.nranks + S.rank, S.seq);
}

// Projection kernels: norm of g of plane c and its reciprocal (compute.c:200-206) into out[0..1].
// Whole-frame sessions read what k_gradient's last CTA left.  Strip sessions wait for the sums of
// every rank and fold them IN RANK ORDER (deterministic, identical on every rank).  The first CTAs
// to get there publish the result locally (F.norms[c], F.norms[4+c], then norms_seq[c] = seq with
// gpu-scope release), so the thousands of CTAs behind them take one cheap local acquire instead of
// `nranks` system-scope ones.  Called by ONE WARP of the CTA (all 32 lanes); the caller
// synchronises the CTA afterwards.
__device__ __forceinline__ void strip_norm(const FrameDev &F, int c, float *out, int lane) {
    const StripSync &S = F.sync;
    if (S.nranks <= 1) {
        if (lane == 0) {
            out[0] = F.norms[c];
            out[1] = F.norms[4 + c];
        }
        return;
    }
    unsigned *ready = reinterpret_cast<unsigned *>(F.norms + 8) + c;      // norms[8..10] hold the published sequence numbers
    unsigned have = 0;
    if (lane == 0) have = ld_acquire_gpu(ready);
    have = __shfl_sync(0xffffffffu, have, 0);
    if (have == S.seq) {
        if (lane == 0) {
            out[0] = __ldcg(F.norms + c);
            out[1] = __ldcg(F.norms + 4 + c);
        }
        return;
    }
    const int slot = (int)(S.seq & 1u);
    bool ok = true;
    if (lane < S.nranks) ok = wait_seq(S.my_flag + slot * S.nranks + lane, S.seq, S.err);
    ok = __all_sync(0xffffffffu, ok);
    double v = 0.;
    if (ok && lane < S.nranks) v = __ldcv(S.my_mail + ((size_t)slot * S.nranks + lane) * 4 + c);
    double s = 0.;
    for (int r = 0; r < S.nranks; r++) s = __dadd_rn(s, __shfl_sync(0xffffffffu, v, r));   // rank order
    if (lane == 0) {
        const float norm = fsqrt(__double2float_rn(s));                      // compute.c:205
        const float rn = __frcp_rn(norm);
        out[0] = norm;
        out[1] = rn;
        if (ok) {
            F.norms[c] = norm;
            F.norms[4 + c] = rn;
            __threadfence();
            st_release_gpu(ready, S.seq);
        }
    }
}

// Projection kernels of a strip session walk their block rows in the order first, last, then the
// interior: the rows the neighbours wait for are projected, stored across NVLink and flagged while
// the interior is still being worked on, so the neighbours' next k_gradient finds its halo rows in
// place instead of waiting for the tail of this kernel (the two exchanges of an iteration then share
// one rank-to-rank synchronisation point, the sums, instead of two).
__device__ __forceinline__ int strip_row_order(const StripSync &S, int y, int rows) {
    if (S.nranks <= 1 || rows < 3) return y;
    return y == 0 ? 0 : (y == 1 ? rows - 1 : y - 1);
}

// Projection kernels with fused halo delivery: a CTA that has stored its share of the strip's
// first (side 0) / last (side 1) two rows into the neighbour calls this with all its threads after
// those stores.  The last such CTA of the iteration raises the neighbour's flag.
__device__ __forceinline__ void strip_border_done(const StripSync &S, int side) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned n = atomicAdd(S.border_ticket + side, 1u) + 1u;
        if (n == S.border_ctas[side]) {
            S.border_ticket[side] = 0u;
            __threadfence_system();
            st_release_sys(side == 0 ? S.up_flag : S.down_flag, S.halo_seq + 1u);
        }
    }
}

}  // namespace j2p
