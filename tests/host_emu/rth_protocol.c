/* tests/host_emu/rth_protocol.c -- drives libmdzcuda's rth_* API the way MDZ's
 * two callers do (render.c:39-92 consumer loop; main_gui.c restart-while-
 * rendering), without MDZ: start, stop in the middle, restart over a running
 * render, run to completion, quit.  Prints "OK <checksum>" on success.
 * Built and run by tests/test_dropin_gpu.py on the GPU box. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../../include/mdz_rth.h"

static void consume(rthdata* rth, mdz_image_info* img, int stop_after_lines)
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
        if (stop_after_lines && linesdone != 0 && (linesdone < 0 || linesdone >= stop_after_lines)) return;
    } while (y < img->user_height);
}

int main(int argc, char** argv)
{
    const int W = 320, H = 200, prec = argc > 1 ? atoi(argv[1]) : 128;
    mdz_image_info* img = calloc(1, sizeof *img);
    img->family = 0; img->fractal = 0; img->depth = 20000;
    img->user_width = img->real_width = W; img->user_height = img->real_height = H;
    img->aa_factor = 1; img->precision = prec; img->use_multi_prec = true; img->use_rounding = true;
    mpfr_init2(img->xmin, prec); mpfr_init2(img->xmax, prec); mpfr_init2(img->ymax, prec); mpfr_init2(img->width, prec);
    mpfr_set_d(img->xmin, -2.0, MPFR_RNDN); mpfr_set_d(img->xmax, 1.0, MPFR_RNDN);
    mpfr_set_d(img->ymax, 0.9375, MPFR_RNDN); mpfr_set_d(img->width, 3.0, MPFR_RNDN);
    mpfr_init2(img->u.julia.c_re, prec); mpfr_init2(img->u.julia.c_im, prec);
    img->raw_data = malloc(sizeof(int) * W * H);
    memset(img->raw_data, 0xff, sizeof(int) * W * H);

    rthdata* rth = rth_create();
    if (!rth || !rth_init(rth, 4, 64, img)) { puts("FAIL init"); return 1; }
    img->rth_ptr = rth;
    if (!rth_ui_init(rth)) { puts("FAIL ui_init"); return 1; }

    /* 1: start, let a few lines arrive, stop and wait */
    rth_ui_start_render(rth);
    rth_ui_wait_until_started(rth);
    consume(rth, img, 8);
    rth_ui_stop_render_and_wait(rth);
    if (!rth_render_should_stop(rth)) { puts("FAIL should_stop after stop"); return 1; }

    /* 2: start, then start again while rendering (Julia preview pattern) */
    rth_ui_start_render(rth);
    rth_ui_wait_until_started(rth);
    usleep(2000);
    rth_ui_start_render(rth);
    rth_ui_wait_until_started(rth);
    rth_ui_stop_render_and_wait(rth);       /* callers stop a render before they touch its buffers (image_info_gui.c:397) */
    memset(img->raw_data, 0xff, sizeof(int) * W * H);

    /* 3: and once more from scratch, consumed to completion */
    rth_ui_start_render(rth);
    rth_ui_wait_until_started(rth);
    consume(rth, img, 0);
    double t = rth_ui_get_render_time(rth);
    if (rth_process_lines_rendered(rth) > 0) { puts("FAIL not complete"); return 1; }
    for (int i = 0; i < H; ++i) if (rth->lines_drawn[i] != 2) { printf("FAIL line %d not drawn\n", i); return 1; }
    unsigned long long sum = 0;
    for (int i = 0; i < W * H; ++i) {
        if (img->raw_data[i] < 0) { printf("FAIL pixel %d unset\n", i); return 1; }
        sum = sum * 1000003ull + (unsigned)img->raw_data[i];
    }
    if (!(t > 0.0 && t < 600.0)) { puts("FAIL time"); return 1; }
    rth_ui_quit(rth);
    printf("OK %llu\n", sum);
    return 0;
}
