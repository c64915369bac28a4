/* tests/host_emu/layout_mine.c -- prints the same table as oracle/layout_probe.c
 * from include/mdz_rth.h (the mirror libmdzcuda is compiled against). */
#include <stdio.h>
#include <stddef.h>
#include "../../include/mdz_rth.h"
#define P(T, N, f) printf(N "." #f " %zu\n", offsetof(T, f))
int main(void)
{
#define I(f) P(mdz_image_info, "image_info", f)
    I(xmin); I(xmax); I(ymax); I(width); I(gxmin); I(gxmax); I(gymax); I(gwidth);
    I(old_cx); I(old_cy); I(old_size); I(pcoords); I(depth); I(thread_count); I(draw_lines);
    I(raw_data); I(rgb_data); I(j_pre); I(drawing_area); I(rnd_pal); I(real_width); I(real_height);
    I(user_width); I(user_height); I(aspect); I(aa_factor); I(family); I(fractal); I(colour_scale);
    I(u); I(palette_ip); I(zoom_new_win); I(use_multi_prec); I(use_rounding); I(precision);
    I(multi_prec_init_done); I(rth_ptr); I(lines_drawn); I(ui_ref_center);
    printf("image_info.sizeof %zu\n", sizeof(mdz_image_info));
#define R(f) P(rthdata, "rthdata", f)
    R(img); R(lines_drawn); R(min_line_drawn); R(line_draw_count); R(thread_count); R(check_stop_px); R(data);
    printf("rthdata.sizeof %zu\n", sizeof(rthdata));
    return 0;
}
