/* oracle/layout_probe.c -- TEST INFRASTRUCTURE ONLY.  Compiled against the
 * reference's own headers (oracle/Makefile, target _ref/layout_probe); prints the
 * offset of every image_info / rthdata field so tests/test_rth_layout.py can
 * compare them with include/mdz_rth.h. */
#include <stdio.h>
#include <stddef.h>
#include "image_info.h"
#include "render_threads.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void)
{
    P(image_info, xmin); P(image_info, xmax); P(image_info, ymax); P(image_info, width);
    P(image_info, gxmin); P(image_info, gxmax); P(image_info, gymax); P(image_info, gwidth);
    P(image_info, old_cx); P(image_info, old_cy); P(image_info, old_size);
    P(image_info, pcoords); P(image_info, depth); P(image_info, thread_count); P(image_info, draw_lines);
    P(image_info, raw_data); P(image_info, rgb_data); P(image_info, j_pre); P(image_info, drawing_area);
    P(image_info, rnd_pal); P(image_info, real_width); P(image_info, real_height);
    P(image_info, user_width); P(image_info, user_height); P(image_info, aspect);
    P(image_info, aa_factor); P(image_info, family); P(image_info, fractal); P(image_info, colour_scale);
    P(image_info, u); P(image_info, palette_ip); P(image_info, zoom_new_win);
    P(image_info, use_multi_prec); P(image_info, use_rounding); P(image_info, precision);
    P(image_info, multi_prec_init_done); P(image_info, rth_ptr); P(image_info, lines_drawn);
    P(image_info, ui_ref_center);
    printf("image_info.sizeof %zu\n", sizeof(image_info));
    P(rthdata, img); P(rthdata, lines_drawn); P(rthdata, min_line_drawn); P(rthdata, line_draw_count);
    P(rthdata, thread_count); P(rthdata, check_stop_px); P(rthdata, data);
    printf("rthdata.sizeof %zu\n", sizeof(rthdata));
    return 0;
}
