/* tests/cb_helper.c — stands in for the reference host program's callbacks (logger.c:20-27,
 * progressbar.c:52-54) so the drop-in compute() can be exercised with a logger and a progress bar
 * from Python.  Loaded RTLD_GLOBAL before libjpeg2png_b200.so, whose weak references then bind here. */
#include <stdio.h>
#include <string.h>

struct logger { FILE *f; const char *filename; unsigned channel; unsigned iteration; };
struct progressbar { unsigned current; unsigned max; };

#define CB_MAX 4096
static double cb_rows[CB_MAX][5];
static unsigned cb_n = 0, cb_pb = 0;
struct progressbar *main_progressbar = 0;

void logger_log(struct logger *log, double objective, double prob_dist, double tv, double tv2) {
        if (log->f && cb_n < CB_MAX) {
                cb_rows[cb_n][0] = (double)log->iteration;
                cb_rows[cb_n][1] = objective; cb_rows[cb_n][2] = prob_dist; cb_rows[cb_n][3] = tv; cb_rows[cb_n][4] = tv2;
                cb_n++;
        }
}
void progressbar_inc(struct progressbar *pb) { pb->current++; cb_pb++; }
void progressbar_clear(struct progressbar *pb) { (void)pb; }

unsigned cb_log_count(void) { return cb_n; }
unsigned cb_progress_count(void) { return cb_pb; }
void cb_get_row(unsigned i, double out[5]) { memcpy(out, cb_rows[i], sizeof cb_rows[i]); }
void cb_reset(void) { cb_n = 0; cb_pb = 0; }
