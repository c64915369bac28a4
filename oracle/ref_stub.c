/*
 * oracle/ref_stub.c -- TEST INFRASTRUCTURE ONLY (never linked into libmdzcuda).
 *
 * The headless reference build (oracle/Makefile, target _ref/mdz) compiles the
 * reference's non-GUI sources unmodified from /root/reference/src.  Those
 * sources call five GUI / libpng entry points (main.c:162,168,215;
 * render.c:96; image_info.c:299).  This file supplies them: the GUI ones are
 * no-ops, and save_png_file writes what a test needs instead of a PNG:
 *   <name>      : binary PPM (P6) of rgb_data  (R,G,B bytes as my_png.c strips them)
 *   <name>.raw  : int32 raw_data dump, header "MDZRAW w h aa depth\n"
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "image_info.h"
#include "palette.h"

int  gui_init(int* argc, char*** argv, image_info* img)
{
    (void)argc; (void)argv; (void)img;
    fprintf(stderr, "headless oracle build: no GUI, use -R\n");
    return 0;
}
void gui_close_display(void) {}
void my_png_reset_last_used_filename(void) {}
void my_png_cleanup(void) {}

void save_png_file(image_info* img, char* filename)
{
    FILE* f = fopen(filename, "wb");
    if (!f) { perror(filename); return; }
    fprintf(f, "P6\n%d %d\n255\n", img->user_width, img->user_height);
    long n = (long)img->user_width * img->user_height;
    for (long i = 0; i < n; ++i) {
        guint32 c = img->rgb_data[i];
        unsigned char px[3] = { RED(c), GREEN(c), BLUE(c) };
        fwrite(px, 1, 3, f);
    }
    fclose(f);

    size_t len = strlen(filename);
    char* rawname = malloc(len + 5);
    memcpy(rawname, filename, len);
    memcpy(rawname + len, ".raw", 5);
    f = fopen(rawname, "wb");
    if (f) {
        fprintf(f, "MDZRAW %d %d %d %ld\n", img->real_width, img->real_height,
                img->aa_factor, (long)img->depth);
        fwrite(img->raw_data, sizeof(int),
               (size_t)img->real_width * img->real_height, f);
        /* the rect the hot path saw, as exact hex (mpfr_out_str base 16) */
        fprintf(f, "\nRECT prec %ld\n", (long)img->precision);
        mpfr_out_str(f, 16, 0, img->xmin, GMP_RNDN);  fputc('\n', f);
        mpfr_out_str(f, 16, 0, img->xmax, GMP_RNDN);  fputc('\n', f);
        mpfr_out_str(f, 16, 0, img->ymax, GMP_RNDN);  fputc('\n', f);
        mpfr_out_str(f, 16, 0, img->width, GMP_RNDN); fputc('\n', f);
        fclose(f);
    }
    free(rawname);
}
