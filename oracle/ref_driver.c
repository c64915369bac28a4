/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into libmdzcuda).
 *
 * Thin ctypes-friendly entry into the UNMODIFIED reference hot path.  Built by
 * oracle/Makefile into oracle/_ref/libmdzref.so together with the reference's
 * own fractal.c, frac_*.c, render_threads.c, timer.c, debug.c (compiled from
 * /root/reference/src where they lie).  It fills an image_info exactly as
 * render_to_file does before it starts the pool (reference render.c:34-41:
 * xmin/xmax/ymax/width + gxmin/gymax/gwidth already set by the caller), then
 * drives the reference's own thread pool through its public rth_* API the way
 * render.c:39-92 does, and returns raw_data.
 *
 * mode: 0 = long double (fractal_calculate_line), 1 = MPFR, 2 = GMP mpf.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "image_info.h"
#include "fractal.h"
#include "render_threads.h"

/* palette globals normally defined in globals.c / palette.c; the hot path
 * itself never touches them but image_info.h pulls the declarations in. */

static void cp_mpfr(mpfr_t dst, const __mpfr_struct* src, mpfr_prec_t prec)
{
    mpfr_init2(dst, prec);
    if (src) mpfr_set(dst, src, GMP_RNDN); else mpfr_set_si(dst, 0, GMP_RNDN);
}

#include <pthread.h>

static void* ref_shutdown(void* p)
{
    image_info* img = p;
    rth_ui_quit((rthdata*)img->rth_ptr);
    free(img->raw_data);
    mpfr_clear(img->xmin); mpfr_clear(img->xmax); mpfr_clear(img->ymax);
    mpfr_clear(img->width);
    mpfr_clear(img->u.julia.c_re); mpfr_clear(img->u.julia.c_im);
    mpf_clear(img->gxmin); mpf_clear(img->gxmax); mpf_clear(img->gymax);
    mpf_clear(img->gwidth);
    free(img);
    return 0;
}

int ref_render(int mode, long prec, int family, int fractal, long depth,
               int width, int height, int aa,
               const __mpfr_struct* xmin, const __mpfr_struct* xmax,
               const __mpfr_struct* ymax, const __mpfr_struct* w,
               const __mpf_struct* gxmin, const __mpf_struct* gymax,
               const __mpf_struct* gwidth,
               const __mpfr_struct* jre, const __mpfr_struct* jim,
               int threads, int* raw_out, double* seconds)
{
    image_info* img = calloc(1, sizeof(image_info));
    if (!img) return 0;
    if (aa < 1) aa = 1;
    img->family = family;
    img->fractal = fractal;
    img->depth = depth;
    img->user_width = width;
    img->user_height = height;
    img->aa_factor = aa;
    img->real_width = width * aa;
    img->real_height = height * aa;
    img->precision = prec;
    img->use_multi_prec = (mode != 0);
    img->use_rounding = (mode == 1);
    img->thread_count = threads;
    img->draw_lines = 64;
    /* the caller hands over values already rounded to the precision they
     * carry (image_info.c:261-267 keeps img->xmin.. at max(P,80)) */
    cp_mpfr(img->xmin, xmin, xmin ? xmin->_mpfr_prec : prec);
    cp_mpfr(img->xmax, xmax, xmax ? xmax->_mpfr_prec : prec);
    cp_mpfr(img->ymax, ymax, ymax ? ymax->_mpfr_prec : prec);
    cp_mpfr(img->width, w,   w ? w->_mpfr_prec : prec);
    cp_mpfr(img->u.julia.c_re, jre, jre ? jre->_mpfr_prec : prec);
    cp_mpfr(img->u.julia.c_im, jim, jim ? jim->_mpfr_prec : prec);
    mpf_init2(img->gxmin, prec);  if (gxmin)  mpf_set(img->gxmin, gxmin);
    mpf_init2(img->gxmax, prec);
    mpf_init2(img->gymax, prec);  if (gymax)  mpf_set(img->gymax, gymax);
    mpf_init2(img->gwidth, prec); if (gwidth) mpf_set(img->gwidth, gwidth);

    size_t npx = (size_t)img->real_width * img->real_height;
    img->raw_data = malloc(npx * sizeof(int));
    memset(img->raw_data, 0xff, npx * sizeof(int));

    rthdata* rth = rth_create();
    img->rth_ptr = rth;
    if (!rth || !rth_init(rth, threads, img->draw_lines, img)) return 0;
    rth_set_next_line_cb(rth, mode == 0 ? fractal_calculate_line
                            : mode == 1 ? fractal_mpfr_calculate_line
                                        : fractal_gmp_calculate_line);
    rth_ui_init(rth);
    rth_ui_start_render(rth);
    /* consumer loop shaped like render.c:49-92, without the colouring -- and without its
     * rth_ui_wait_for_line_done(): that wait has no timeout (render_threads.c:568) and now and then misses the
     * last signal (the reference's BUGS:1-6), which holds a test run for ever.  rth_process_lines_rendered()
     * waits half a millisecond by itself; the pool that renders is untouched. */
    int y = 0, linesdone;
    do {
        usleep(100);
        linesdone = rth_process_lines_rendered(rth);
        if (linesdone) {
            int undrawn = 0;
            int miny = rth->min_line_drawn;
            int maxy = miny + rth->line_draw_count + 1;
            if (maxy >= img->user_height) maxy = img->user_height;
            if (linesdone > 0 && maxy > linesdone) maxy = linesdone;
            char* ld = &rth->lines_drawn[miny];
            for (y = miny; y < maxy; ++y, ++ld) {
                if (*ld == 1) { *ld = 2; if (!undrawn) rth->min_line_drawn = y; }
                else if (*ld == 0) undrawn = 1;
            }
        }
    } while (y < img->user_height);
    double t = rth_ui_get_render_time(rth);
    if (seconds) *seconds = t;
    memcpy(raw_out, img->raw_data, npx * sizeof(int));

    /* rth_ui_quit() signals the watch thread's condition without setting the flag it waits on
     * (render_threads.c:445-462 against :199-203): when the watch thread is not yet back in its wait -- a render
     * of a few milliseconds -- the signal is lost and the join in rth_ui_quit never returns.  So the shutdown
     * runs on a detached thread: in that rare case it leaks two sleeping threads instead of holding the caller. */
    pthread_t th;
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
    if (pthread_create(&th, &at, ref_shutdown, img) != 0) ref_shutdown(img);
    pthread_attr_destroy(&at);
    return 1;
}

/*
 * Selected lines only, at any image size: calls the reference's own line driver
 * (fractal_calculate_line / fractal_mpfr_calculate_line / fractal_gmp_calculate_line,
 * reference src/fractal.c:29/120/260) for each requested real line, from `threads`
 * pthreads, and returns them packed as out[k][real_width].  The full-size parity tests
 * use it where rendering every line of a BASELINE config on host cores would take
 * minutes (MPFR-320 at 1920x1080, 23040x12960 supersamples, ...).  The image buffer is
 * full size but only the requested lines are ever touched.
 */
#include <pthread.h>

typedef struct {
    image_info* img;
    int (*cb)(image_info*, int);
    const int* lines;
    int nlines;
    int next;
    pthread_mutex_t mu;
} line_job;

static void* line_worker(void* arg)
{
    line_job* j = arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int k = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (k >= j->nlines) return 0;
        j->cb(j->img, j->lines[k]);
    }
}

int ref_render_lines(int mode, long prec, int family, int fractal, long depth,
                     int width, int height, int aa,
                     const __mpfr_struct* xmin, const __mpfr_struct* xmax,
                     const __mpfr_struct* ymax, const __mpfr_struct* w,
                     const __mpf_struct* gxmin, const __mpf_struct* gymax,
                     const __mpf_struct* gwidth,
                     const __mpfr_struct* jre, const __mpfr_struct* jim,
                     const int* lines, int nlines, int threads, int* out)
{
    image_info* img = calloc(1, sizeof(image_info));
    if (!img) return 0;
    if (aa < 1) aa = 1;
    img->family = family;
    img->fractal = fractal;
    img->depth = depth;
    img->user_width = width;
    img->user_height = height;
    img->aa_factor = aa;
    img->real_width = width * aa;
    img->real_height = height * aa;
    img->precision = prec;
    img->use_multi_prec = (mode != 0);
    img->use_rounding = (mode == 1);
    img->thread_count = threads;
    img->draw_lines = 64;
    cp_mpfr(img->xmin, xmin, xmin ? xmin->_mpfr_prec : prec);
    cp_mpfr(img->xmax, xmax, xmax ? xmax->_mpfr_prec : prec);
    cp_mpfr(img->ymax, ymax, ymax ? ymax->_mpfr_prec : prec);
    cp_mpfr(img->width, w,   w ? w->_mpfr_prec : prec);
    cp_mpfr(img->u.julia.c_re, jre, jre ? jre->_mpfr_prec : prec);
    cp_mpfr(img->u.julia.c_im, jim, jim ? jim->_mpfr_prec : prec);
    mpf_init2(img->gxmin, prec);  if (gxmin)  mpf_set(img->gxmin, gxmin);
    mpf_init2(img->gxmax, prec);
    mpf_init2(img->gymax, prec);  if (gymax)  mpf_set(img->gymax, gymax);
    mpf_init2(img->gwidth, prec); if (gwidth) mpf_set(img->gwidth, gwidth);

    size_t npx = (size_t)img->real_width * img->real_height;
    img->raw_data = malloc(npx * sizeof(int));          /* pages are touched per line only */
    rthdata* rth = rth_create();
    img->rth_ptr = rth;
    if (!img->raw_data || !rth || !rth_init(rth, 1, img->draw_lines, img)) return 0;
    rth->check_stop_px = 64;                             /* render_threads.c:291-292 */

    line_job job;
    job.img = img;
    job.cb = mode == 0 ? fractal_calculate_line : mode == 1 ? fractal_mpfr_calculate_line
                                                            : fractal_gmp_calculate_line;
    job.lines = lines; job.nlines = nlines; job.next = 0;
    pthread_mutex_init(&job.mu, 0);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], 0, line_worker, &job);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], 0);

    for (int k = 0; k < nlines; ++k)
        memcpy(out + (size_t)k * img->real_width,
               img->raw_data + (size_t)lines[k] * img->real_width, (size_t)img->real_width * sizeof(int));
    free(img->raw_data);
    mpfr_clear(img->xmin); mpfr_clear(img->xmax); mpfr_clear(img->ymax);
    mpfr_clear(img->width);
    mpfr_clear(img->u.julia.c_re); mpfr_clear(img->u.julia.c_im);
    mpf_clear(img->gxmin); mpf_clear(img->gxmax); mpf_clear(img->gymax);
    mpf_clear(img->gwidth);
    free(img);
    return 1;
}
