/* tests/host_emu/rth_preview.c -- the second caller of the boundary: MDZ's Julia preview (main_gui.c:28-29,
 * 533-599, 786-793).  A 160x90 preview with 2x2 anti-aliasing and line_draw_count 2 (image_info.c:50) is
 * restarted on every "mouse motion": the constant changes, rth_ui_start_render arrives while the previous
 * frame may still be rendering, rth_ui_wait_until_started, then the idle callback's consumer loop until the
 * next motion event.  Every frame that completes before the next restart is appended to the output file as
 * (frame index, raw_data); tests/test_preview_gpu.py compares each with the reference's line driver for
 * that frame's constant.
 * usage: rth_preview PRECISION(0 = long double) FRAMES GAP_US OUTFILE */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "../../include/mdz_rth.h"

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int prec_arg = atoi(argv[1]), frames = atoi(argv[2]), gap = atoi(argv[3]);
    const int UW = 160, UH = 90, AA = 2, W = UW * AA, H = UH * AA;
    const int prec = prec_arg ? prec_arg : 80;
    mdz_image_info* img = calloc(1, sizeof *img);
    img->family = 1; img->fractal = 0; img->depth = 300; img->j_pre = 1;
    img->user_width = UW; img->user_height = UH; img->aa_factor = AA; img->real_width = W; img->real_height = H;
    img->precision = prec; img->use_multi_prec = prec_arg != 0; img->use_rounding = true;
    mpfr_init2(img->xmin, prec); mpfr_init2(img->xmax, prec); mpfr_init2(img->ymax, prec); mpfr_init2(img->width, prec);
    mpfr_set_d(img->xmin, -1.625, MPFR_RNDN); mpfr_set_d(img->xmax, 1.625, MPFR_RNDN);
    mpfr_set_d(img->ymax, 0.9140625, MPFR_RNDN); mpfr_set_d(img->width, 3.25, MPFR_RNDN);
    mpfr_init2(img->u.julia.c_re, prec); mpfr_init2(img->u.julia.c_im, prec);
    img->raw_data = malloc(sizeof(int) * W * H);
    FILE* out = fopen(argv[4], "wb");
    if (!out) return 3;

    rthdata* rth = rth_create();
    if (!rth || !rth_init(rth, 4, 2, img)) { puts("FAIL init"); return 1; }
    img->rth_ptr = rth;
    if (!rth_ui_init(rth)) { puts("FAIL ui_init"); return 1; }

    int completed = 0;
    for (int i = 0; i < frames; ++i) {
        /* motion_event: new constant, start over (main_gui.c:786-793) */
        mpfr_set_d(img->u.julia.c_re, -0.8 + 0.25 * cos(0.1 * i), MPFR_RNDN);
        mpfr_set_d(img->u.julia.c_im, 0.156 + 0.25 * sin(0.13 * i), MPFR_RNDN);
        rth_ui_start_render(rth);
        rth_ui_wait_until_started(rth);
        /* idle_draw_callback until the next event (main_gui.c:533-599) */
        const double t0 = now_us();
        const double wait = (i % 3 == 2) ? gap / 8.0 : gap;        /* every third event follows quickly */
        int done = 0;
        while (now_us() - t0 < wait) {
            const int r = rth_process_lines_rendered(rth);
            if (r != 0) {
                int miny = rth->min_line_drawn, maxy = miny + rth->line_draw_count + 1, undrawn = 0, y;
                if (maxy >= UH) maxy = UH;
                if (r > 0 && maxy > r) maxy = r;
                for (y = miny; y < maxy; ++y) {
                    char* ld = &rth->lines_drawn[y];
                    if (*ld == 1) { *ld = 2; if (!undrawn) rth->min_line_drawn = y; }
                    else if (*ld == 0) undrawn = 1;
                }
                if (r < 0 && y >= UH) { done = 1; break; }
            }
        }
        if (done) {
            fwrite(&i, sizeof i, 1, out);
            fwrite(img->raw_data, sizeof(int), (size_t)W * H, out);
            ++completed;
            /* the rest of the interval passes idle */
            while (now_us() - t0 < wait) { struct timespec nap = { 0, 50 * 1000 }; nanosleep(&nap, 0); }
        }
    }
    rth_ui_stop_render_and_wait(rth);
    rth_ui_quit(rth);
    fclose(out);
    printf("OK %d of %d frames completed\n", completed, frames);
    return 0;
}
