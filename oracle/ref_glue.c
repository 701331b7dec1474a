/* oracle/ref_glue.c — TEST INFRASTRUCTURE ONLY.
 *
 * The reference solver objects reference one global that normally lives in the CLI
 * translation unit (`struct progressbar *main_progressbar`, reference jpeg2png.c:175, used by
 * utils.c:12).  When only the solver sources are compiled into oracle/_ref/ that symbol has to
 * come from somewhere; this file is that somewhere.  It also exposes the thread count so the
 * bench can report how many host cores the reference arm used.
 */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct progressbar;
struct progressbar *main_progressbar = NULL;

int ref_glue_max_threads(void) {
#ifdef _OPENMP
        return omp_get_max_threads();
#else
        return 1;
#endif
}

void ref_glue_set_threads(int n) {
#ifdef _OPENMP
        if (n > 0) omp_set_num_threads(n);
#else
        (void)n;
#endif
}
