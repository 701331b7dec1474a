// kernels.cuh — device-side parameter blocks shared by the kernels and session.cu.
#pragma once
#include <stdint.h>

namespace j2p {

// One colour plane as the kernels see it.  For a strip session (multi-GPU row tiling) the
// frame-sized buffers hold the strip's rows plus the halo rows, the coefficient-sized ones only
// the strip's own coefficient rows.
struct PlaneDev {
    float *x;             // current iterate x_k, H x W raster              (reference aux.fdata)
    float *xp;            // previous iterate x_{k-1}; receives x_{k+1}     (reference aux.fista)
    float *g;             // objective sub-gradient, H x W raster           (reference aux.obj_gradient)
    float *gp;            // DCT-distance gradient for the NEXT step at coefficient resolution,
                          // ch x cw raster: p_alpha * idct((cos - data*q)/q^2)   (compute.c:38-70)
    const int16_t *data;  // quantised coefficients, [blocks][64] natural order
    int cw, ch;           // coefficient grid in samples (ch: rows held by this session)
    int sw, sh;           // upsampling factors
    int resample;         // !(cw == W && ch == H) for the WHOLE frame      (compute.c:338)
    int use_prob;         // pweight != 0                                   (compute.c:244)
    float p_alpha;        // pweight*2*255*sqrtf(2)                         (compute.c:245)
    float cnt;            // (float)(sw*sh), the divisor of the block mean     (compute.c:359)
};

// Strip sessions over peer memory (multi-GPU row tiling): the two exchanges of an
// iteration happen INSIDE the two kernels.  k_gradient's last CTA stores this rank's three sums of
// g^2 into every rank's mailbox (NVLink stores through cudaIpc mappings) and raises a flag there;
// every CTA of the projection waits for all flags of the iteration and folds the sums in rank
// order; the projection CTAs that produce the strip's first / last two rows also store them into
// the neighbours' halo rows, and the last of them raises the neighbours' halo flag, which the
// first / last row band of the neighbours' next k_gradient waits for.  nranks <= 1: no in-kernel
// exchange (whole-frame session, or a strip driven through NCCL / the host).
struct StripSync {
    int nranks, rank;
    unsigned seq;                   // 1-based index of this iteration's sums exchange (same on every rank)
    unsigned halo_seq;              // the halo flags must have reached this before k_gradient reads the halo rows
    int fused_halo;                 // the projection kernels deliver the border rows themselves
    int has_up, has_down;
    unsigned border_ctas[2];        // projection CTAs per iteration that hold the top / bottom border rows
    double *mail[8];                // rank p's mailbox as mapped here: [2 slots][nranks][4] doubles
    unsigned *mail_flag[8];         // rank p's mailbox flags: [2 slots][nranks]
    const double *my_mail;          // this rank's own mailbox / flags (local addresses)
    const unsigned *my_flag;
    float *up_dst[3], *down_dst[3]; // where rows [t0, t0+2) / [t1-2, t1) of x_{k+1} go in the neighbours, per plane
    unsigned *up_flag, *down_flag;  // the neighbours' words "my lower / upper neighbour has delivered"
    const unsigned *from_up, *from_down;   // this rank's own words
    unsigned *border_ticket;        // [2] local counters of finished border CTAs (top, bottom)
    int *err;                       // set when a wait timed out (results invalid; never hangs a box)
};

struct FrameDev {
    int W, H, nc;         // H: rows held locally (strip + halo rows); whole frame: H == Hg
    int Hg;               // height of the whole frame
    int y0g;              // frame row of local row 0
    int t0, t1;           // local rows [t0, t1) this session owns (targets); the rest is halo
    PlaneDev pl[3];
    float *slab;          // the allocation that holds x, xp, g, gp of every plane
    const void *host_maps;// host pointer to the session's TileMaps (tma_maps.h), or null: no TMA path for this session
    int buf_sel;          // 0: pl[c].x is the session's first iterate buffer, 1: the second (which tensor map is x_k)
    unsigned plane_stride;// elements between consecutive planes of one array: pl[c].x == pl[0].x + c * plane_stride, same for xp, g, gp
    float q[3][64];       // quantisation tables as float
    float qq[3][64];      // q*q (fp32 product, compute.c:49)
    float rqq[3][64];     // RN(1/(q*q)), the shared reciprocal of the residual division
    float a1;             // (float)(1./sqrtf(nc))                          (compute.c:90)
    float a2;             // (float)(alpha*1./sqrtf(nc)), alpha = weight/sqrtf(2)   (compute.c:154,258)
    int use_tgv;          // weight != 0                                    (compute.c:257)
    float step;           // radius / sqrtf(1 + iterations)                 (compute.c:425,443)
    float one;            // 1.0f, opaque to the compiler: addm2() in numerics.cuh
    double *partials;     // [5][grad_ctas] per-CTA sums of g^2 (and, when logging, of the TV / TGV norms)
    double *sums;         // [3] this session's sum of g^2 (strip mode: combined across ranks by the driver)
    float *norms;         // [0..2] sqrtf((float)sum g^2) (compute.c:200-206); [4..6] RN(1/norm)
    unsigned *counter;    // CTAs-done ticket for the last-CTA reduction
    int grad_ctas;
    int grad_slots;       // CTAs of k_gradient resident on this device at once (band geometry)
    StripSync sync;
    // objective logging (compute.c:271-272), only when the caller asked for a CSV log
    int log_on;
    int log_slot;         // which of the two prob_dist slots k_project accumulates into
    double *logsums;      // [0]=tv, [1]=tv2, [2+3*slot+c] = sum over plane c of (residual/q)^2
};

// ---- the stand-alone halo kernel (kernels_strip.cu); pointers into OTHER ranks' memory are cudaIpc
// mappings made by session.cu
struct HaloPeers {
    float *up_dst[3];         // where this strip's first two rows go: the upper neighbour's bottom halo rows, per plane
    float *down_dst[3];       // where the last two rows go: the lower neighbour's top halo rows
    const float *up_src[3];   // this strip's first two owned rows
    const float *down_src[3]; // this strip's last two owned rows
    unsigned *up_flag, *down_flag;         // the neighbours' words for "my lower / upper neighbour has delivered"
    const unsigned *from_up, *from_down;   // this rank's own words
    int has_up, has_down, nc;
    unsigned n4;              // float4s per plane and side: 2 rows * W / 4
};

}  // namespace j2p
