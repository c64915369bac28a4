/*
 * mdz_rth.h -- the render-pool API libmdzcuda exports in place of MDZ's
 * src/render_threads.c, plus the host structs it has to read.
 *
 * Everything here restates an interface of the reference so that MDZ links
 * against libmdzcuda unchanged:
 *   - rthdata and the fourteen rth_* prototypes: reference
 *     src/render_threads.h:8-9 (constants), :20-32 (struct), :41-87 (functions);
 *   - mdz_image_info: the memory layout of `image_info`
 *     (src/image_info.h:63-122) and `julia_info` (:50-55).  MDZ's own header
 *     needs GTK; this mirror needs nothing, and tests/test_rth_layout.py checks
 *     every offset against the reference's header when it is available.
 *
 * Callers read and write rthdata.lines_drawn / min_line_drawn /
 * line_draw_count directly (src/render.c:56-73) and src/fractal.c:33 reads
 * check_stop_px, so the public struct is bit-for-bit the reference's; the
 * private part (rthpridata) is opaque there and redesigned here.
 */
#ifndef MDZ_RTH_H
#define MDZ_RTH_H

#include <stdint.h>
#include <stdbool.h>
#include "mdz_mp_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DEFAULT_THREAD_COUNT 2      /* render_threads.h:8 */
#define MAX_THREAD_COUNT 512        /* render_threads.h:9 */

typedef int64_t mdz_depth_t;        /* types.h:8 */

typedef struct {                    /* image_info.h:50-55 */
    mpfr_t c_re;
    mpfr_t c_im;
} mdz_julia_info;

typedef struct mdz_image_info {     /* image_info.h:63-122, same order, same types */
    mpfr_t  xmin, xmax, ymax, width;
    mpf_t   gxmin, gxmax, gymax, gwidth;
    mpfr_t  old_cx, old_cy, old_size;
    void*   pcoords;                /* coords*                */
    mdz_depth_t depth;
    int     thread_count;
    int     draw_lines;
    int*      raw_data;
    uint32_t* rgb_data;
    int     j_pre;                  /* gboolean               */
    void*   drawing_area;           /* GtkWidget*             */
    void*   rnd_pal;                /* random_palette*        */
    int     real_width;
    int     real_height;
    int     user_width;
    int     user_height;
    double  aspect;
    int     aa_factor;
    int     family;
    int     fractal;
    double  colour_scale;
    union { mdz_julia_info julia; } u;
    bool    palette_ip;
    bool    zoom_new_win;
    bool    use_multi_prec;
    bool    use_rounding;
    mpfr_prec_t precision;
    bool    multi_prec_init_done;
    void*   rth_ptr;
    int     lines_drawn;
    bool    ui_ref_center;
} mdz_image_info;

typedef struct rthpridata rthpridata;

typedef struct RTH_DATA {           /* render_threads.h:20-32 */
    mdz_image_info* img;
    char*   lines_drawn;
    int     min_line_drawn;
    int     line_draw_count;
    int     thread_count;
    int     check_stop_px;
    rthpridata* data;
} rthdata;

/* render_threads.h:41-87; return conventions of render_threads.c:77-183, :485-585 */
rthdata* rth_create(void);
int     rth_init(rthdata* rth, int thread_count, int line_draw_count, mdz_image_info* img);
int     rth_ui_init(rthdata* rth);
void    rth_ui_start_render(rthdata* rth);
void    rth_ui_stop_render(rthdata* rth);
void    rth_ui_stop_render_and_wait(rthdata* rth);
void    rth_ui_quit(rthdata* rth);
void    rth_ui_stop_timer(rthdata* rth);
void    rth_ui_wait_until_started(rthdata* rth);
double  rth_ui_get_render_time(rthdata* rth);
void    rth_set_next_line_cb(rthdata* rth, int (*next_line_cb)(mdz_image_info*, int));
int     rth_process_lines_rendered(rthdata* rth);
int     rth_render_should_stop(rthdata* rth);
int     rth_ui_wait_for_line_done(rthdata* rth);

#ifdef __cplusplus
}
#endif
#endif /* MDZ_RTH_H */
