/* tools/preview_latency.c -- the second caller of the rth_* boundary:
 * MDZ's Julia preview, a 160 x (160/aspect) image at anti-aliasing 2 that is re-started on
 * every mouse-motion event (reference src/main_gui.c:28-29, :786-793), driven through the
 * reference's own API only.  Links against libmdzcuda.so or, unchanged, against the
 * reference's own pool (oracle/_ref/libmdzref.so exports the same rth_* symbols), so the
 * two can be timed side by side:
 *
 *   complete:   set a new Julia constant, rth_ui_start_render, consume lines until the
 *               frame is complete (idle_draw_callback's loop); per-frame latency.
 *   interrupt:  a new constant every 2 ms whether or not the previous frame finished (the
 *               watch thread stops and joins the render in progress first); how many frames
 *               complete, and how long the last one takes to appear.
 *
 * usage: preview_latency [frames] [depth] [precision: 64 = long double mode, else MPFR] [modes: 1 complete, 2 interrupt, 3 both]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include "../include/mdz_rth.h"

/* MDZ installs the line driver for the arithmetic mode (image_info.c:243-248).  The reference's
 * pool needs it; libmdzcuda records it and never calls it.  Weak: absent in libmdzcuda. */
extern int fractal_calculate_line(mdz_image_info*, int) __attribute__((weak));
extern int fractal_mpfr_calculate_line(mdz_image_info*, int) __attribute__((weak));

static double now_ms(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* idle_draw_callback's bookkeeping (main_gui.c:533-599 / render.c:49-92) without the drawing;
 * returns 1 when the frame is complete, 0 if told to give up at `deadline_ms` */
static int consume(rthdata* rth, mdz_image_info* img, double deadline_ms)
{
    int y = 0, linesdone;
    do {
        rth_ui_wait_for_line_done(rth);
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
        if (deadline_ms > 0 && now_ms() >= deadline_ms) return 0;
    } while (y < img->user_height);
    return 1;
}

static int cmp_double(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

int main(int argc, char** argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 200;
    const int depth = argc > 2 ? atoi(argv[2]) : 300;          /* DEFAULT_DEPTH, image_info.c:15 */
    const int prec = argc > 3 ? atoi(argv[3]) : 64;
    const int modes = argc > 4 ? atoi(argv[4]) : 3;
    const int UW = 160, UH = 90, AA = 2;                       /* JPRE_SIZE, JPRE_AAFACTOR at 16:9 */
    mdz_image_info* img = calloc(1, sizeof *img);
    img->family = 1; img->fractal = 0; img->depth = depth;
    img->user_width = UW; img->user_height = UH; img->aa_factor = AA;
    img->real_width = UW * AA; img->real_height = UH * AA;
    img->precision = prec < 80 ? 80 : prec;
    img->use_multi_prec = prec != 64; img->use_rounding = true;
    const int ip = img->precision;
    mpfr_init2(img->xmin, ip); mpfr_init2(img->xmax, ip); mpfr_init2(img->ymax, ip); mpfr_init2(img->width, ip);
    mpfr_set_d(img->xmin, -2.0, MPFR_RNDN); mpfr_set_d(img->xmax, 2.0, MPFR_RNDN);
    mpfr_set_d(img->ymax, 1.125, MPFR_RNDN); mpfr_set_d(img->width, 4.0, MPFR_RNDN);
    mpfr_init2(img->u.julia.c_re, ip); mpfr_init2(img->u.julia.c_im, ip);
    img->raw_data = malloc(sizeof(int) * img->real_width * img->real_height);

    rthdata* rth = rth_create();
    if (!rth || !rth_init(rth, (int)sysconf(_SC_NPROCESSORS_ONLN), 64, img)) { puts("FAIL init"); return 1; }
    img->rth_ptr = rth;
    if (prec == 64 ? fractal_calculate_line != 0 : fractal_mpfr_calculate_line != 0)
        rth_set_next_line_cb(rth, prec == 64 ? fractal_calculate_line : fractal_mpfr_calculate_line);
    if (!rth_ui_init(rth)) { puts("FAIL ui_init"); return 1; }

    double* lat = malloc(sizeof(double) * frames);
    long long checksum = 0;
    if (modes & 1) {
    /* the mouse walks round the main cardioid's neighbourhood */
    for (int f = -3; f < frames; ++f) {                        /* 3 untimed frames: context, module load, pool */
        const double t = 6.283185307179586 * (f + 3) / (frames + 3);
        mpfr_set_d(img->u.julia.c_re, 0.7885 * cos(t), MPFR_RNDN);
        mpfr_set_d(img->u.julia.c_im, 0.7885 * sin(t), MPFR_RNDN);
        const double t0 = now_ms();
        rth_ui_start_render(rth);
        rth_ui_wait_until_started(rth);
        consume(rth, img, 0);
        if (f >= 0) lat[f] = now_ms() - t0;
        for (int i = 0; i < img->real_width * img->real_height; i += 97) checksum += img->raw_data[i];
    }
    qsort(lat, frames, sizeof(double), cmp_double);
    double sum = 0; for (int f = 0; f < frames; ++f) sum += lat[f];
    printf("complete:  %d frames of %dx%d aa %d depth %d %s: mean %.3f ms, median %.3f, p95 %.3f, max %.3f (checksum %lld)\n",
           frames, UW, UH, AA, depth, prec == 64 ? "long double" : "mpfr", sum / frames, lat[frames / 2],
           lat[(int)(frames * 0.95)], lat[frames - 1], checksum);
    }
    if (!(modes & 2)) { rth_ui_quit(rth); return 0; }

    /* interrupt mode: a motion event every 2 ms */
    int completed = 0;
    double last_latency = 0;
    const double T0 = now_ms();
    for (int f = 0; f < frames; ++f) {
        const double t = 6.283185307179586 * f / frames;
        mpfr_set_d(img->u.julia.c_re, 0.7885 * cos(t), MPFR_RNDN);
        mpfr_set_d(img->u.julia.c_im, 0.7885 * sin(t), MPFR_RNDN);
        const double t0 = now_ms();
        rth_ui_start_render(rth);                              /* may arrive while rendering */
        rth_ui_wait_until_started(rth);
        const int last = f == frames - 1;
        const int done = consume(rth, img, last ? 0 : t0 + 2.0);
        completed += done;
        if (last) last_latency = now_ms() - t0;
    }
    printf("interrupt: %d motion events 2 ms apart: %d frames completed in time, total %.1f ms, final frame %.3f ms after its event\n",
           frames, completed, now_ms() - T0, last_latency);
    rth_ui_quit(rth);
    return 0;
}
