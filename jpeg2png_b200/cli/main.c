/* main.c — the jpeg2png command line on top of libjpeg2png_b200.so.
 *
 * Behavioural restatement of reference jpeg2png.c:120-357 (decode_file + main): same flags, same
 * defaults (-w 0.3, -p 0.001, -i 50; jpeg2png.c:22-24), same validation rules and messages for the
 * numeric options (jpeg2png.c:206-257), same output naming and overwrite rules (:273-316), same
 * per-file flow (read -> conventional decode -> solve -> +128 on luma -> PNG; :120-172).
 *
 * What differs, on purpose:
 *   - the JPEG reader and PNG writer are the self-contained ones of this directory (no libjpeg /
 *     libpng on the build box);
 *   - the conventional decode runs on the GPU (j2p_session_upload with fdata == NULL) and the
 *     solve goes through the session layer, so each file needs one upload and one download;
 *   - files are spread round-robin over the visible GPUs (the reference's OpenMP file loop,
 *     jpeg2png.c:330, becomes one host thread per file; -t bounds the number of host threads).
 */
#define _POSIX_C_SOURCE 200809L
#include <stdarg.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_thread_num(void) { return 0; }
#endif

#include "../../include/jpeg2png_b200.h"
#include "jpeg_reader.h"
#include "png_writer.h"

#define J2P_CLI_VERSION "1.01-b200"

static const float default_weight = 0.3f;        /* jpeg2png.c:22 */
static const float default_pweight = 0.001f;     /* jpeg2png.c:23 */
static const unsigned default_iterations = 50;   /* jpeg2png.c:24 */

/* ---- progress bar (reference progressbar.c) and die (utils.c:11-40) -------------------------- */
struct progressbar { unsigned current, max; };
static struct progressbar *main_pb;
static const unsigned pb_width = 70;

static void pb_show(const struct progressbar *pb) {
        const unsigned n = pb->max ? pb_width * pb->current / pb->max : pb_width, pct = pb->max ? 100 * pb->current / pb->max : 100;
        printf("\r[");
        for (unsigned i = 0; i < pb_width; i++) putchar(i < n ? '#' : ' ');
        printf("] %3u%%", pct);
        fflush(stdout);
}
static void pb_clear(void) {
        printf("\r%*s\r", (int)pb_width + 7, "");
        fflush(stdout);
}
static void pb_add(struct progressbar *pb, unsigned n) {
#pragma omp critical(progressbar)
        {
                const unsigned before_n = pb->max ? pb_width * pb->current / pb->max : 0, before_p = pb->max ? 100 * pb->current / pb->max : 0;
                pb->current += n;
                const unsigned after_n = pb->max ? pb_width * pb->current / pb->max : 0, after_p = pb->max ? 100 * pb->current / pb->max : 0;
                if (before_n != after_n || before_p != after_p) pb_show(pb);
        }
}

static _Noreturn void die(const char *msg, ...) {
        if (main_pb) { pb_clear(); main_pb = NULL; }
        fprintf(stderr, "jpeg2png: ");
        va_list l;
        va_start(l, msg);
        vfprintf(stderr, msg, l);
        va_end(l);
        fprintf(stderr, "\n");
        exit(EXIT_FAILURE);
}

/* ---- CSV log (reference logger.c) ------------------------------------------------------------ */
static FILE *csv_log;
static void log_row(const char *filename, unsigned channel, unsigned iteration, const double o[4]) {
        if (!csv_log) return;
#pragma omp critical(write_log)
        if (fprintf(csv_log, "%s,%u,%u,%f,%f,%f,%f\n", filename, channel, iteration, o[0], o[1], o[2], o[3]) < 0) die("could not write to csv log");
}

static void usage(void) {
        printf("usage: jpeg2png [options] picture.jpg ...\n\n"
               "Decodes JPEG files into PNG files with the smoothest picture that still encodes to the same JPEG\n"
               "(total generalised variation regularised decoding), computed on an NVIDIA B200.\n\n"
               "  -o FILE, --output FILE            output file name; give it once per input file or not at all\n"
               "                                    (default: the input name with .jpg/.jpeg replaced by .png)\n"
               "  -f, --force                       overwrite existing output files\n"
               "  -i N[,N,N], --iterations          optimisation steps (default %u); three values need -s\n"
               "  -w W[,W,W], --second-order-weight weight of the second-order smoothness term (default %g); 0 is faster;\n"
               "                                    three values need -s\n"
               "  -p P[,P,P], --probability-weight  weight of the distance to the decoded coefficients (default %g); 0 is faster\n"
               "  -s, --separate-components         optimise Y, Cb and Cr separately (faster, slightly worse)\n"
               "  -1, --16-bits-png                 write 16 bits per sample\n"
               "  -t N, --threads N                 host threads used to drive files / GPUs\n"
               "  -q, --quiet                       no progress bar\n"
               "  -c FILE, --csv-log FILE           write the objective of every step to a CSV file\n"
               "  -h, --help                        this text\n"
               "  -V, --version                     version\n",
               default_iterations, (double)default_weight, (double)default_pweight);
        exit(EXIT_FAILURE);
}

static double now_ms(void);

/* ---- one file (reference decode_file, jpeg2png.c:120-172) ------------------------------------ */
struct job {
        unsigned iterations[3];
        float weights[3], pweights[3];
        unsigned png_bits;
        bool all_together, quiet;
};

/* out: the float planes (separate mode), or — joint mode, scanlines != NULL — nothing: the image comes
 * back as PNG scanlines, converted on the device (j2p_session_download_scanlines) */
static void solve(const struct j2p_jpeg *jpeg, const unsigned *chan, unsigned nchan, int device, const struct job *job, unsigned iterations,
                  float weight, struct progressbar *pb, const char *infile, unsigned log_channel, float **out, unsigned *out_w, unsigned *out_h,
                  uint8_t **scanlines) {
        struct j2p_frame_desc d;
        memset(&d, 0, sizeof d);
        d.nchannel = nchan;
        d.weight = weight;
        d.iterations = iterations;
        for (unsigned k = 0; k < nchan; k++) {
                const struct coef *c = &jpeg->coefs[chan[k]];
                d.plane_w[k] = c->w; d.plane_h[k] = c->h; d.w_samp[k] = c->w_samp; d.h_samp[k] = c->h_samp;
                d.pweight[k] = job->pweights[chan[k]];
        }
        const char *tr = getenv("J2P_TRACE");
        const int trace = tr && *tr == '1';
        double t0 = trace ? now_ms() : 0, t1;
        j2p_session *s = NULL;
        if (j2p_session_create(&s, device, &d) != J2P_OK) die("%s", j2p_last_error());
        if (trace) { t1 = now_ms(); fprintf(stderr, "j2p trace:   %s: session create %.1f ms", infile, t1 - t0); t0 = t1; }
        if (csv_log && j2p_session_set_logging(s, 1) != J2P_OK) die("%s", j2p_last_error());
        for (unsigned k = 0; k < nchan; k++) {
                const struct coef *c = &jpeg->coefs[chan[k]];
                if (j2p_session_upload(s, k, c->data, c->quant_table, NULL) != J2P_OK) die("%s", j2p_last_error());   /* decode on the device */
        }
        if (trace) { t1 = now_ms(); fprintf(stderr, ", upload %.1f ms", t1 - t0); t0 = t1; }
        unsigned reported = 0;
        for (unsigned i = 0; i < iterations; i++) {
                if (j2p_session_iterate(s, i, 1) != J2P_OK) die("%s", j2p_last_error());
                if (csv_log) {
                        double o[4];
                        if (j2p_session_objective(s, o) != J2P_OK) die("%s", j2p_last_error());
                        log_row(infile, log_channel, i, o);
                }
                if (pb) while (reported + 8 <= i) { j2p_session_wait_iteration(s, reported++); pb_add(pb, 1); }
        }
        if (pb) for (; reported < iterations; reported++) { j2p_session_wait_iteration(s, reported); pb_add(pb, 1); }
        *out_w = j2p_session_width(s);
        *out_h = j2p_session_height(s);
        if (trace) { t1 = now_ms(); fprintf(stderr, ", %u iterations queued %.1f ms", iterations, t1 - t0); t0 = t1; }
        if (scanlines) {
                *scanlines = malloc((size_t)jpeg->h * ((size_t)jpeg->w * 3 * (job->png_bits / 8) + 1));
                if (!*scanlines) die("allocation error");
                if (j2p_session_download_scanlines(s, jpeg->w, jpeg->h, job->png_bits, *scanlines) != J2P_OK) die("%s", j2p_last_error());
                if (trace) { t1 = now_ms(); fprintf(stderr, ", drain + scanlines %.1f ms", t1 - t0); t0 = t1; }
                j2p_session_destroy(s);
                if (trace) fprintf(stderr, ", destroy %.1f ms\n", now_ms() - t0);
                return;
        }
        for (unsigned k = 0; k < nchan; k++) {
                out[k] = aligned_alloc(16, (((size_t)*out_w * *out_h * sizeof(float)) + 15) & ~(size_t)15);
                if (!out[k]) die("allocation error");
                if (j2p_session_download(s, k, out[k]) != J2P_OK) die("%s", j2p_last_error());
        }
        j2p_session_destroy(s);
}

/* CPUs this process may really use: the affinity mask capped by the cgroup quota.  The GPU boxes
 * show 128 hardware threads and grant 16 CPUs; an OpenMP team as wide as the machine then spends
 * its quota on context switches (round 1: threads 90..126 in the batch trace). */
static int usable_cpus(void) {
        int n = 1;
#ifdef _OPENMP
        n = omp_get_num_procs();
#endif
        FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r");
        if (f) {
                char quota[32];
                long period = 0;
                if (fscanf(f, "%31s %ld", quota, &period) == 2 && period > 0 && strcmp(quota, "max") != 0) {
                        const long q = (atol(quota) + period - 1) / period;
                        if (q > 0 && q < n) n = (int)q;
                }
                fclose(f);
        }
        return n > 0 ? n : 1;
}

static double now_ms(void) {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void decode_file(const char *infile, const char *outfile, const struct job *job, int device, struct progressbar *pb) {
        const char *tr = getenv("J2P_TRACE");
        const int trace = tr && *tr == '1';
        const double t_begin = trace ? now_ms() : 0;
        FILE *in = fopen(infile, "rb");
        if (!in) { if (main_pb) { pb_clear(); main_pb = NULL; } fprintf(stderr, "jpeg2png: could not open input file `%s`: ", infile); perror(NULL); exit(EXIT_FAILURE); }
        fseek(in, 0, SEEK_END);
        const long len = ftell(in);
        fseek(in, 0, SEEK_SET);
        uint8_t *buf = malloc(len > 0 ? (size_t)len : 1);
        if (!buf || fread(buf, 1, (size_t)len, in) != (size_t)len) die("could not read input file `%s`", infile);
        fclose(in);
        struct j2p_jpeg jpeg;
        char err[256];
        if (j2p_read_jpeg_mem(buf, (size_t)len, &jpeg, err, sizeof err) != 0) die("%s", err);
        free(buf);

        const double t_read = trace ? now_ms() : 0;
        float *planes[3] = {NULL, NULL, NULL};
        uint8_t *scanlines = NULL;
        unsigned pw[3], ph[3];
        if (job->all_together) {                                                 /* jpeg2png.c:142-144 */
                /* the +128 of jpeg2png.c:156-159 and the colour conversion of png.c:39-62 run on the device */
                const unsigned chan[3] = {0, 1, 2};
                unsigned w, h;
                solve(&jpeg, chan, 3, device, job, job->iterations[0], job->weights[0], pb, infile, 3, planes, &w, &h, &scanlines);
        } else {                                                                 /* jpeg2png.c:146-152: each plane computes its own frame size */
                /* three independent solves, concurrently like the reference's OpenMP loop, on up to three
                 * devices (inside the file-parallel loop nested parallelism is off and they run in turn) */
                const int ndev = j2p_device_count();
#pragma omp parallel for schedule(static, 1) num_threads(3)
                for (unsigned i = 0; i < 3; i++)
                        solve(&jpeg, &i, 1, (device + (int)i) % (ndev > 0 ? ndev : 1), job, job->iterations[i], job->weights[i], pb, infile, i, &planes[i], &pw[i], &ph[i], NULL);
                for (size_t i = 0; i < (size_t)pw[0] * ph[0]; i++) planes[0][i] += 128.f;   /* jpeg2png.c:156-159 */
        }
        const double t_solve = trace ? now_ms() : 0;

        FILE *out = fopen(outfile, "wb");
        if (!out) { if (main_pb) { pb_clear(); main_pb = NULL; } fprintf(stderr, "jpeg2png: could not open output file `%s`: ", outfile); perror(NULL); exit(EXIT_FAILURE); }
        const int wrc = scanlines ? j2p_write_png_scanlines(out, jpeg.w, jpeg.h, job->png_bits, scanlines)
                                  : j2p_write_png(out, jpeg.w, jpeg.h, job->png_bits, planes[0], pw[0], planes[1], pw[1], planes[2], pw[2]);
        if (wrc != 0) die("could not write PNG file `%s`", outfile);
        fclose(out);
        free(scanlines);
        for (int i = 0; i < 3; i++) { free(planes[i]); free(jpeg.coefs[i].data); }
        if (trace)
                fprintf(stderr, "j2p trace: %s: read+parse %.1f ms, solve (upload..download) %.1f ms, colour+PNG %.1f ms (thread %d, started at %.1f)\n", infile,
                        t_read - t_begin, t_solve - t_read, now_ms() - t_solve, omp_get_thread_num(), t_begin);
}

/* ---- option parsing -------------------------------------------------------------------------- */
struct opt { char s; const char *l; bool arg; };
static const struct opt opts[] = {{'h', "help", false}, {'V', "version", false}, {'o', "output", true}, {'f', "force", false},
                                  {'c', "csv-log", true}, {'t', "threads", true}, {'q', "quiet", false}, {'s', "separate-components", false},
                                  {'1', "16-bits-png", false}, {'i', "iterations", true}, {'p', "probability-weight", true},
                                  {'w', "second-order-weight", true}};

int main(int argc, char **argv) {
        const char *arg_w = NULL, *arg_p = NULL, *arg_i = NULL, *arg_t = NULL, *arg_c = NULL;
        bool help = false, version = false, force = false, quiet = false, separate = false, png16 = false;
        const char **inputs = malloc(sizeof(char *) * (size_t)argc), **outputs = malloc(sizeof(char *) * (size_t)argc);
        unsigned nin = 0, nout = 0;
        bool only_files = false;
        for (int a = 1; a < argc; a++) {
                const char *s = argv[a];
                if (only_files || s[0] != '-' || s[1] == 0) { inputs[nin++] = s; continue; }
                if (strcmp(s, "--") == 0) { only_files = true; continue; }
                const struct opt *o = NULL;
                const char *val = NULL;
                if (s[1] == '-') {
                        const char *eq = strchr(s + 2, '=');
                        const size_t n = eq ? (size_t)(eq - (s + 2)) : strlen(s + 2);
                        for (size_t k = 0; k < sizeof opts / sizeof *opts; k++)
                                if (strlen(opts[k].l) == n && strncmp(opts[k].l, s + 2, n) == 0) o = &opts[k];
                        if (eq) val = eq + 1;
                } else {
                        for (size_t k = 0; k < sizeof opts / sizeof *opts; k++) if (opts[k].s == s[1]) o = &opts[k];
                        if (s[1] == '?') o = &opts[0];
                        if (o && o->arg && s[2]) val = s + 2;
                        else if (o && !o->arg && s[2]) die("unknown option `%s`", s);
                }
                if (!o) die("unknown option `%s`", s);
                if (o->arg && !val) { if (a + 1 >= argc) die("option `%s` needs an argument", s); val = argv[++a]; }
                switch (o->s) {
                        case 'h': help = true; break;
                        case 'V': version = true; break;
                        case 'o': outputs[nout++] = val; break;
                        case 'f': force = true; break;
                        case 'c': arg_c = val; break;
                        case 't': arg_t = val; break;
                        case 'q': quiet = true; break;
                        case 's': separate = true; break;
                        case '1': png16 = true; break;
                        case 'i': arg_i = val; break;
                        case 'p': arg_p = val; break;
                        case 'w': arg_w = val; break;
                }
        }
        if (version) { printf("jpeg2png version " J2P_CLI_VERSION " licensed GPLv3+\n"); exit(EXIT_FAILURE); }   /* jpeg2png.c:196-199 */
        if (nin < 1 || help) usage();

        struct job job;
        job.all_together = !separate;
        job.quiet = quiet;
        job.png_bits = png16 ? 16 : 8;
        job.weights[0] = default_weight; job.weights[1] = job.weights[2] = 0.f;                     /* jpeg2png.c:206 */
        if (arg_w) {
                const int n = sscanf(arg_w, "%f,%f,%f", &job.weights[0], &job.weights[1], &job.weights[2]);
                if (n == 3) { if (job.all_together) die("different weights are only possible when using separated components"); }
                else if (n != 1) die("invalid weight");
        }
        for (int i = 0; i < 3; i++) job.pweights[i] = default_pweight;
        if (arg_p) {
                const int n = sscanf(arg_p, "%f,%f,%f", &job.pweights[0], &job.pweights[1], &job.pweights[2]);
                if (n == 1) job.pweights[1] = job.pweights[2] = job.pweights[0];
                else if (n != 3) die("invalid probability weight");
        }
        for (int i = 0; i < 3; i++) job.iterations[i] = default_iterations;
        if (arg_i) {
                const int n = sscanf(arg_i, "%u,%u,%u", &job.iterations[0], &job.iterations[1], &job.iterations[2]);
                if (n == 3) { if (job.all_together) die("different iteration counts are only possible when using separated components"); }
                else if (n == 1) job.iterations[1] = job.iterations[2] = job.iterations[0];
                else die("invalid number of iterations");
        }
        unsigned threads = 0;
        if (arg_t) {
                if (sscanf(arg_t, "%u", &threads) != 1 || threads == 0) die("invalid number of threads");
#ifdef _OPENMP
                omp_set_num_threads((int)threads);
#endif
        } else {
#ifdef _OPENMP
                if (!getenv("OMP_NUM_THREADS")) omp_set_num_threads(usable_cpus());
#endif
        }
        if (arg_c) {
                csv_log = fopen(arg_c, "wb");
                if (!csv_log) die("could not open csv log `%s`", arg_c);
                if (fprintf(csv_log, "filename,channel,iteration,objective,prob_dist,tv,tv2\n") < 0) die("could not write to csv log");   /* logger.c:13 */
        }
        if (!(nout == 0 || nout == nin)) die("must give output file names for all input files or none");

        char **outfiles = malloc(sizeof(char *) * nin);
        for (unsigned i = 0; i < nin; i++) {
                if (nout) { outfiles[i] = (char *)outputs[i]; continue; }
                const char *infile = inputs[i];                                                     /* jpeg2png.c:283-313 */
                FILE *in = fopen(infile, "rb");
                if (!in) die("could not open input file `%s`", infile);
                fclose(in);
                const size_t l = strlen(infile);
                size_t e = l;
                if (l >= 5 && memcmp(".jpeg", infile + l - 5, 5) == 0) e = l - 5;
                else if (l >= 4 && memcmp(".jpg", infile + l - 4, 4) == 0) e = l - 4;
                char *outfile = malloc(e + 5);
                memcpy(outfile, infile, e);
                memcpy(outfile + e, ".png", 5);
                if (!force) {
                        FILE *probe = fopen(outfile, "rb");
                        if (probe) die("not overwriting output file `%s`", outfile);
                }
                FILE *out = fopen(outfile, "wb");
                if (!out) die("could not open output file `%s`", outfile);
                fclose(out);
                remove(outfile);
                outfiles[i] = outfile;
        }

        const int ndev = j2p_device_count();
        if (ndev <= 0) die("no CUDA device available (the solver has no CPU fallback)");

        struct progressbar pb;
        if (!quiet) {                                                                               /* jpeg2png.c:319-327 */
                pb.current = 0;
                pb.max = job.all_together ? nin * job.iterations[0] : nin * (job.iterations[0] + job.iterations[1] + job.iterations[2]);
                pb_show(&pb);
                main_pb = &pb;
        }
#pragma omp parallel for schedule(dynamic) if (nin > 1)
        for (unsigned i = 0; i < nin; i++) decode_file(inputs[i], outfiles[i], &job, (int)(i % (unsigned)ndev), quiet ? NULL : &pb);

        if (!quiet) { pb_clear(); main_pb = NULL; }
        if (csv_log) fclose(csv_log);
        /* Every output file is closed.  Leave without the CUDA runtime's exit handlers: unpinning the
         * staging buffers and tearing the context down costs a batch of short solves a noticeable part of
         * its wall time and releases nothing the operating system does not release anyway. */
        fflush(NULL);
        _exit(EXIT_SUCCESS);
}
