/* jpeg2png_b200.h — C ABI of the B200-native jpeg2png solver (libjpeg2png_b200.so).
 *
 * Plain C, plain pointers and sizes.  Two layers:
 *
 *   1. The DROP-IN layer: `compute()` with exactly the signature, ownership rules and side
 *      effects of the reference solver entry (reference compute.h:8, compute.c:407-465) and the
 *      data contract `struct coef` (reference jpeg2png.h:7-20).  A reference build links this
 *      library instead of its own compute.o/box.o/ooura/dct.o and nothing else changes
 *      (see INTEGRATION.md).
 *
 *   2. The SESSION layer (`j2p_*`): the same solver with the device residency made explicit,
 *      so that a caller (bench, batch driver, multi-GPU strip driver) can keep coefficient
 *      planes resident in HBM, time the iteration loop without the host<->device copies, and
 *      run several frames on several streams/devices.  `compute()` is implemented on top of it.
 *
 * There is NO CPU fallback anywhere behind this header: if no CUDA device is usable every entry
 * point fails (drop-in layer: die() -> exit(EXIT_FAILURE) like the reference, utils.c:20-28;
 * session layer: negative return code + j2p_last_error()).
 */
#ifndef JPEG2PNG_B200_H
#define JPEG2PNG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------
 * Data contract — layout-identical to the reference's `struct coef` (jpeg2png.h:7-20).
 *   h, w            plane size in samples (multiples of 8; jpeg.c:52-53)
 *   h_samp, w_samp  upsampling factors of this plane w.r.t. the frame (jpeg.c:57-58)
 *   data            quantised DCT coefficients, int16, [blocks][64] natural order, blocks in
 *                   raster order (jpeg.c:68-77); read-only for the solver, owned by the caller
 *   fdata           in : conventional decode, raster h x w, 0-centred (jpeg2png.c:127-139)
 *                   out: solver result, raster H x W of the working frame (compute.c:455-461)
 *   quant_table     natural order, all entries non-zero (jpeg.c:41-46)
 * ------------------------------------------------------------------------------------- */
struct coef {
        unsigned h;
        unsigned w;
        unsigned h_samp;
        unsigned w_samp;
        int16_t *data;
        float *fdata;
        uint16_t quant_table[64];
};

/* Reference logger.h:6-11 and progressbar.h:4-7: the solver writes log->iteration every
 * iteration (compute.c:428), calls logger_log once per iteration (compute.c:271-272) and
 * progressbar_inc once per iteration when pb != NULL (compute.c:449-452). */
struct logger;
struct progressbar;

/* ---------------------------------------------------------------------------------------
 * 1. Drop-in entry.  Replaces reference compute.c:407-465.
 *
 *   - consumes coefs[c].fdata (it must come from aligned_alloc/malloc, utils.h:89-106) and sets
 *     it to a 16-byte-aligned malloc-family buffer of H*W floats owned by the caller — the
 *     consumed buffer itself when the plane already has frame size, a new one (and the old one
 *     freed, compute.c:304-305, :458) otherwise;
 *   - overwrites coefs[c].w/h with the frame size W,H (compute.c:460-461);
 *   - re-entrant: may be called concurrently from several host threads (jpeg2png.c:147, :330);
 *     each call has its own CUDA stream and device buffers.  What the calls share is internally
 *     locked and holds no solver state: a cache of idle device blocks, pinned staging buffers and
 *     the host copy threads;
 *   - errors: message on stderr prefixed "jpeg2png: " and exit(EXIT_FAILURE) (utils.c:11-28).
 *
 * The callbacks are resolved at link time exactly like in the reference: the host program
 * provides logger_log() and progressbar_inc() (reference logger.c:20-27, progressbar.c:52-54).
 * When the library is loaded stand-alone (tests, bench) weak no-op defaults are used.
 * ------------------------------------------------------------------------------------- */
void compute(unsigned nchannel, struct coef *coefs, struct logger *log, struct progressbar *pb,
             float weight, float *pweight, unsigned iterations);

/* ---------------------------------------------------------------------------------------
 * 2. Session layer.
 * ------------------------------------------------------------------------------------- */
typedef struct j2p_session j2p_session;

enum {
        J2P_OK = 0,
        J2P_ERR_ARG = -1,      /* bad argument (sizes not multiples of 8, nchannel > 3, ...) */
        J2P_ERR_CUDA = -2,     /* a CUDA runtime call failed; see j2p_last_error()           */
        J2P_ERR_NODEVICE = -3  /* no usable sm_100 device                                    */
};

/* Thread-local text of the last failure on this host thread ("" if none). */
const char *j2p_last_error(void);

/* Number of CUDA devices visible to this process (0 if none / no driver). */
int j2p_device_count(void);

/* Which device the drop-in compute() uses when called from THIS host thread (the reference's
 * compute() has no device argument, compute.h:8).  -1 (the initial state of every thread) = the
 * process default: environment J2P_DEVICE, else device 0.  The reference calls compute() from
 * OpenMP threads for the three planes (jpeg2png.c:147-152) and for files (jpeg2png.c:330); a host
 * that binds each of those threads to its own device spreads them over the GPUs of the box. */
int j2p_set_thread_device(int device);
int j2p_thread_device(void);

/* Describes one frame to solve: the planes that are optimised TOGETHER (reference joint mode:
 * nchannel = 3, jpeg2png.c:144; separate mode: three sessions with nchannel = 1, :147-152).
 * A horizontal strip of the frame (multi-GPU spatial tiling) is selected at creation time with
 * j2p_session_create_strip(). */
struct j2p_frame_desc {
        unsigned nchannel;       /* 1..3 */
        unsigned plane_w[3];     /* coef->w  */
        unsigned plane_h[3];     /* coef->h  */
        unsigned w_samp[3];      /* coef->w_samp */
        unsigned h_samp[3];      /* coef->h_samp */
        float weight;            /* TGV weight (-w), compute.c:257-260 */
        float pweight[3];        /* DCT-distance weights (-p), compute.c:244-245 */
        unsigned iterations;     /* total iteration count (fixes the step size, compute.c:443) */
};

/* Create a session on `device` (cudaSetDevice ordinal).  Allocates all HBM working buffers. */
int j2p_session_create(j2p_session **out, int device, const struct j2p_frame_desc *desc);
void j2p_session_destroy(j2p_session *s);

/* ---- row strips of one frame across several GPUs (BASELINE config 4) ------------
 * A strip session holds frame rows [row0, row0+rows) of the frame described by `desc` (which
 * always describes the WHOLE frame) plus two halo rows on every side that has a neighbour.
 * row0 and row0+rows must be multiples of 8*h_samp of every plane (row0+rows may also be the
 * frame height).  Uploads then carry only the coefficient rows of the strip.  One iteration is
 *     j2p_session_gradient(s);                       k_gradient on the owned rows
 *     <all-gather the 3 doubles at j2p_session_sums_ptr(s) over the ranks>
 *     j2p_session_project(s, gathered, nranks);      fold in rank order -> norms, step + projection
 *     <exchange halos: j2p_session_halo(s, c, side, &send, &recv, &count)>
 * The driver (jpeg2png_b200/strips.py) does the two communication steps with torch.distributed
 * (NCCL over NVLink on GPUs).  After (re)arming a strip session the driver exchanges the halos of
 * the initial iterate once and calls j2p_session_copy_halo_to_prev(). */
int j2p_session_create_strip(j2p_session **out, int device, const struct j2p_frame_desc *desc,
                             unsigned row0, unsigned rows);
int j2p_session_strip_info(const j2p_session *s, unsigned *local_rows, unsigned *first_owned,
                           unsigned *owned_rows);
int j2p_session_gradient(j2p_session *s);
void *j2p_session_sums_ptr(j2p_session *s);
int j2p_session_project(j2p_session *s, const double *sums_by_rank, unsigned nranks);
/* side 0 = top, 1 = bottom.  send: first/last two owned rows of the current iterate; recv: the
 * halo rows beyond them; count: floats to move (0 if there is no neighbour on that side). */
int j2p_session_halo(j2p_session *s, unsigned channel, int side, void **send, void **recv,
                     size_t *count);
int j2p_session_copy_halo_to_prev(j2p_session *s);

/* Native strip loop.  One process per GPU: rank 0 obtains an id with j2p_comm_unique_id(), hands
 * its 128 bytes to the other ranks by any means (the Python driver broadcasts it with
 * torch.distributed), every rank calls j2p_comm_create().  Ranks are the strips in top-to-bottom
 * order.  j2p_session_iterate_strip() is collective; its first call after (re)arming a session also
 * exchanges the halos of the initial iterate; the host never waits inside the loop.
 *
 * On one node (up to 8 ranks) the first call maps the peers' memory (cudaIpc over NVLink) and from
 * then on an iteration is exactly the two solver kernels: the gradient kernel's last CTA stores
 * this rank's sums of g^2 into every rank's mailbox, the projection kernels wait for all of them,
 * fold them in rank order and store the strip's border rows straight into the neighbours' halo
 * rows, the next gradient kernel's border bands wait for those (sequence-numbered flags,
 * st.release.sys / ld.acquire.sys).  When peer memory cannot be mapped, or with J2P_STRIP_P2P=0, the
 * exchanges are ncclAllGather + ncclSend/ncclRecv queued between the kernels.  NCCL is resolved at
 * run time with dlopen("libnccl.so.2"): no link-time dependency.  Same results either way.
 * Destroy the communicator before the session it was used with. */
typedef struct j2p_comm j2p_comm;
#define J2P_COMM_ID_BYTES 128
int j2p_comm_unique_id(void *out, size_t bytes);
int j2p_comm_create(j2p_comm **out, int device, int nranks, int rank, const void *id, size_t bytes);
void j2p_comm_destroy(j2p_comm *c);
int j2p_session_iterate_strip(j2p_session *s, j2p_comm *c, unsigned n);
/* 0, or an error if an in-kernel wait of the peer-memory protocol timed out on the device (a peer
 * died; results are invalid; every wait gives up after about two seconds of GPU clock instead of
 * hanging the device).  Synchronises the device. */
int j2p_comm_status(j2p_comm *c);
/* 1 if j2p_session_iterate_strip exchanges through peer memory, 0 if through NCCL. */
int j2p_comm_protocol(const j2p_comm *c);

/* Working-frame size W x H = max over planes of (plane_w*w_samp, plane_h*h_samp) (compute.c:410-416). */
unsigned j2p_session_width(const j2p_session *s);
unsigned j2p_session_height(const j2p_session *s);

/* Host -> HBM.  `data`: int16 [blocks][64]; `quant`: uint16[64]; `fdata`: raster plane_h x plane_w
 * conventional decode.  Performs the reference aux_init (compute.c:278-310) on the device:
 * cos = data*quant, nearest-neighbour upsample with edge clamp, fista = fdata.  Returns as soon
 * as the host arrays have been read (into the session's pinned staging ring): the caller may free
 * them; the DMA of the last chunks and the set-up kernels are still queued on the session stream. */
int j2p_session_upload(j2p_session *s, unsigned channel, const int16_t *data,
                       const uint16_t *quant, const float *fdata);

/* Re-arm the iteration state from the already-resident coefficient planes and the resident copy
 * of the conventional decode (no host traffic) — lets a benchmark time the loop repeatedly. */
int j2p_session_reset(j2p_session *s);

/* Run `n` solver iterations starting at iteration index `first` (0-based) on the session stream.
 * Asynchronous.  The FISTA momentum sequence (compute.c:431-432) restarts when first == 0. */
int j2p_session_iterate(j2p_session *s, unsigned first, unsigned n);

/* Measurement aid: runs `n` iterations from a freshly re-armed state with CUDA events recorded
 * on the session stream around each kernel, and returns the mean device time per launch of the
 * gradient kernel and of the step+projection kernel (milliseconds). */
int j2p_session_profile(j2p_session *s, unsigned n, float *ms_gradient, float *ms_project);

/* Block the host until iteration `iter` (0-based, already queued by j2p_session_iterate) has
 * finished on the device.  This is what lets `compute()` advance the reference's progress bar
 * (compute.c:449-452) at the pace of the device without draining the stream. */
int j2p_session_wait_iteration(j2p_session *s, unsigned iter);

/* HBM -> host: the current iterate of `channel`, H x W floats raster. */
int j2p_session_download(j2p_session *s, unsigned channel, float *out);

/* HBM -> host, joint (3-plane) whole-frame sessions: the image as the reference hands it to libpng,
 * computed on the device — luma += 128 (jpeg2png.c:156-159), YCbCr -> RGB in double, clamp,
 * scale by (1 << bits)/256, truncate (png.c:39-47), 8-bit or 16-bit big-endian samples
 * (png.c:51-62) — laid out as PNG scanlines: h rows of 1 + w*3*bits/8 bytes, every row starting
 * with filter type 0.  w x h is the visible image (jpeg->w, jpeg->h), at most the frame size.
 * 3 or 6 bytes per pixel cross PCIe instead of 12. */
int j2p_session_download_scanlines(j2p_session *s, unsigned w, unsigned h, unsigned bits,
                                   unsigned char *out);

/* Objective terms of the most recent iteration, as logged by the reference (compute.c:271-272):
 * out[0]=objective, out[1]=prob_dist, out[2]=tv, out[3]=tv2.  Only tracked when logging was
 * enabled with j2p_session_set_logging(s, 1) before iterating. */
int j2p_session_set_logging(j2p_session *s, int enabled);
int j2p_session_objective(j2p_session *s, double out[4]);

/* Block the host until everything queued on the session stream has finished. */
int j2p_session_sync(j2p_session *s);

/* Make a freshly allocated host buffer resident (huge-page hint + parallel first touch).  The
 * drop-in compute() calls it for the result buffers it hands back (compute.c:458) while the device
 * is still iterating, so the page faults of 100 MB of new memory are off the download path. */
void j2p_host_prefault(void *p, size_t bytes);

/* Raw handles for callers that schedule their own work around the session (bench timing with
 * CUDA events on the launching stream; torch interop).  The stream is a cudaStream_t. */
void *j2p_session_stream(j2p_session *s);
/* Device pointer of the current iterate of `channel` (H x W floats). */
void *j2p_session_plane_ptr(j2p_session *s, unsigned channel);

/* Launch counters: number of kernel launches issued by this session since creation. */
unsigned long long j2p_session_launches(const j2p_session *s);

/* Version string of the library build. */
const char *j2p_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JPEG2PNG_B200_H */
