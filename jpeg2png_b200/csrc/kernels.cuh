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

struct FrameDev {
    int W, H, nc;         // H: rows held locally (strip + halo rows); whole frame: H == Hg
    int Hg;               // height of the whole frame
    int y0g;              // frame row of local row 0
    int t0, t1;           // local rows [t0, t1) this session owns (targets); the rest is halo
    PlaneDev pl[3];
    float q[3][64];       // quantisation tables as float
    float qq[3][64];      // q*q (fp32 product, compute.c:49)
    float rqq[3][64];     // RN(1/(q*q)), the shared reciprocal of the residual division
    float a1;             // (float)(1./sqrtf(nc))                          (compute.c:90)
    float a2;             // (float)(alpha*1./sqrtf(nc)), alpha = weight/sqrtf(2)   (compute.c:154,258)
    int use_tgv;          // weight != 0                                    (compute.c:257)
    float step;           // radius / sqrtf(1 + iterations)                 (compute.c:425,443)
    double *partials;     // [5][grad_ctas] per-CTA sums of g^2 (and, when logging, of the TV / TGV norms)
    double *sums;         // [3] this session's sum of g^2 (strip mode: combined across ranks by the driver)
    float *norms;         // [0..2] sqrtf((float)sum g^2) (compute.c:200-206); [4..6] RN(1/norm)
    unsigned *counter;    // CTAs-done ticket for the last-CTA reduction
    int grad_ctas;
    // objective logging (compute.c:271-272), only when the caller asked for a CSV log
    int log_on;
    int log_slot;         // which of the two prob_dist slots k_project accumulates into
    double *logsums;      // [0]=tv, [1]=tv2, [2+3*slot+c] = sum over plane c of (residual/q)^2
};

// ---- strip exchanges over peer memory (kernels_strip.cu); pointers into OTHER ranks' memory are
// cudaIpc mappings made by session.cu
struct StripPeers {
    double *mail[8];          // rank p's mailbox base, as mapped here: [2 slots][nranks][4] doubles
    unsigned *mail_flag[8];   // rank p's mailbox flags: [2 slots][nranks]
    int nranks, rank;
};
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
