/* compute.c — the drop-in solver entry, host side, C11.
 *
 * Same symbol, signature, ownership rules and callbacks as the reference's compute()
 * (reference compute.h:8, compute.c:407-465); the body only marshals `struct coef` into a device
 * session (session.cu) and back.  There is no CPU implementation of the solver behind it: when no
 * sm_100 device is usable this dies like every other fatal error of the reference
 * (utils.c:11-28: "jpeg2png: <message>" on stderr, exit(EXIT_FAILURE)).
 */
#define _POSIX_C_SOURCE 199309L
#include <stdarg.h>
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/jpeg2png_b200.h"

/* Layout mirrors of the two callback structs (reference logger.h:6-11, progressbar.h:4-7); the
 * host program owns the real definitions. */
struct logger {
        FILE *f;
        const char *filename;
        unsigned channel;
        unsigned iteration;
};
struct progressbar {
        unsigned current;
        unsigned max;
};

/* Provided by the host program when this library replaces the reference solver objects
 * (reference logger.c:20-27, progressbar.c:52-54, progressbar.c:56-66 + utils.c:11-17).  Weak:
 * a stand-alone load (tests, bench) simply has no callbacks. */
extern void logger_log(struct logger *log, double objective, double prob_dist, double tv, double tv2) __attribute__((weak));
extern void progressbar_inc(struct progressbar *pb) __attribute__((weak));
extern void progressbar_clear(struct progressbar *pb) __attribute__((weak));
extern struct progressbar *main_progressbar __attribute__((weak));

_Noreturn static void die(const char *msg, ...) {
        /* utils.c:11-28 */
        if (&main_progressbar && main_progressbar) {
                if (progressbar_clear) progressbar_clear(main_progressbar);
                main_progressbar = NULL;
        }
        fprintf(stderr, "jpeg2png: ");
        va_list l;
        va_start(l, msg);
        vfprintf(stderr, msg, l);
        va_end(l);
        fprintf(stderr, "\n");
        exit(EXIT_FAILURE);
}

/* J2P_TRACE=1: wall-clock of each phase of the call on stderr (a measurement aid, off by default) */
static double now_ms(void) {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* free the consumed input buffers that could not be recycled and make the result buffers that are
 * still missing (alloc_simd contract, utils.h:89-98), pages already touched */
static void host_buffers(unsigned nchannel, float **consumed, float **outs, size_t out_bytes) {
        for (unsigned c = 0; c < nchannel; c++) {
                free(consumed[c]);
                consumed[c] = NULL;
                if (!outs[c]) {
                        outs[c] = aligned_alloc(16, out_bytes);
                        if (!outs[c]) die("allocation error");
                        j2p_host_prefault(outs[c], out_bytes);
                }
        }
}

/* how many iterations may be queued on the device ahead of the progress bar */
#define J2P_PROGRESS_LAG 8u

void compute(unsigned nchannel, struct coef *coefs, struct logger *log, struct progressbar *pb,
             float weight, float *pweight, unsigned iterations) {
        if (nchannel < 1 || nchannel > 3) die("compute: nchannel must be 1..3");

        struct j2p_frame_desc d;
        memset(&d, 0, sizeof d);
        d.nchannel = nchannel;
        d.weight = weight;
        d.iterations = iterations;
        for (unsigned c = 0; c < nchannel; c++) {
                d.plane_w[c] = coefs[c].w;
                d.plane_h[c] = coefs[c].h;
                d.w_samp[c] = coefs[c].w_samp;
                d.h_samp[c] = coefs[c].h_samp;
                d.pweight[c] = pweight[c];
        }

        /* device choice: this thread's binding (j2p_set_thread_device), else J2P_DEVICE, else 0 */
        const int device = j2p_thread_device();

        const char *trace_env = getenv("J2P_TRACE");
        const int trace = trace_env && *trace_env == '1';
        double t0 = trace ? now_ms() : 0, t1;

        j2p_session *s = NULL;
        if (j2p_session_create(&s, device, &d) != J2P_OK) die("%s", j2p_last_error());
        if (trace) { t1 = now_ms(); fprintf(stderr, "j2p trace: create %.2f ms\n", t1 - t0); t0 = t1; }
        const int want_log = log && log->f != NULL;
        if (want_log && j2p_session_set_logging(s, 1) != J2P_OK) die("%s", j2p_last_error());

        /* Buffer ownership as in the reference: the caller's fdata is consumed (compute.c:304-305)
         * and a malloc-family buffer of frame size comes back (compute.c:458).  Where a plane
         * already has frame size (4:4:4 planes, luma) the consumed buffer IS the one handed back —
         * no free, no allocation, no page faults.  Otherwise the free (an munmap: ~5 ms for a 4K
         * plane on the 128-thread hosts) and the allocation + first touch of the new buffer are
         * done while the device iterates, not on the transfer path. */
        const unsigned w = j2p_session_width(s), h = j2p_session_height(s);
        size_t out_bytes = (size_t)w * h * sizeof(float);
        out_bytes = (out_bytes + 15) & ~(size_t)15;
        float *outs[3] = {NULL, NULL, NULL}, *consumed[3] = {NULL, NULL, NULL};
        for (unsigned c = 0; c < nchannel; c++) {
                if (j2p_session_upload(s, c, coefs[c].data, coefs[c].quant_table, coefs[c].fdata) != J2P_OK)
                        die("%s", j2p_last_error());
                if (coefs[c].fdata && (size_t)coefs[c].w * coefs[c].h == (size_t)w * h) outs[c] = coefs[c].fdata;
                else consumed[c] = coefs[c].fdata;
                coefs[c].fdata = NULL;
        }

        if (trace) { t1 = now_ms(); fprintf(stderr, "j2p trace: upload %.2f ms\n", t1 - t0); t0 = t1; }
        const unsigned housekeeping_at = iterations > 16 ? 15 : (iterations ? iterations - 1 : 0);
        int housekeeping_done = 0;

        unsigned reported = 0;
        for (unsigned i = 0; i < iterations; i++) {
                if (log) log->iteration = i;                            /* compute.c:428 */
                if (j2p_session_iterate(s, i, 1) != J2P_OK) die("%s", j2p_last_error());
                if (i == housekeeping_at) {
                        host_buffers(nchannel, consumed, outs, out_bytes);
                        housekeeping_done = 1;
                }
                if (want_log) {
                        double o[4];
                        if (j2p_session_objective(s, o) != J2P_OK) die("%s", j2p_last_error());
                        if (logger_log) logger_log(log, o[0], o[1], o[2], o[3]);   /* compute.c:271-272 */
                }
                if (pb && progressbar_inc) {                            /* compute.c:449-452 */
                        while (reported + J2P_PROGRESS_LAG <= i) {
                                if (j2p_session_wait_iteration(s, reported) != J2P_OK) die("%s", j2p_last_error());
#pragma omp critical(progressbar)
                                progressbar_inc(pb);
                                reported++;
                        }
                }
        }
        if (pb && progressbar_inc) {
                for (; reported < iterations; reported++) {
                        if (j2p_session_wait_iteration(s, reported) != J2P_OK) die("%s", j2p_last_error());
#pragma omp critical(progressbar)
                        progressbar_inc(pb);
                }
        }

        if (trace) {
                t1 = now_ms(); fprintf(stderr, "j2p trace: queue %u iterations %.2f ms\n", iterations, t1 - t0); t0 = t1;
                j2p_session_sync(s);
                t1 = now_ms(); fprintf(stderr, "j2p trace: device drain %.2f ms\n", t1 - t0); t0 = t1;
        }
        if (!housekeeping_done) host_buffers(nchannel, consumed, outs, out_bytes);   /* iterations == 0 */
        for (unsigned c = 0; c < nchannel; c++) {                       /* compute.c:455-463 */
                if (j2p_session_download(s, c, outs[c]) != J2P_OK) die("%s", j2p_last_error());
                coefs[c].fdata = outs[c];
                coefs[c].w = w;
                coefs[c].h = h;
        }
        if (trace) { t1 = now_ms(); fprintf(stderr, "j2p trace: download %.2f ms\n", t1 - t0); t0 = t1; }
        j2p_session_destroy(s);
        if (trace) { t1 = now_ms(); fprintf(stderr, "j2p trace: destroy %.2f ms\n", t1 - t0); }
}
