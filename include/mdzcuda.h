/*
 * mdzcuda.h -- C ABI of libmdzcuda, the B200 replacement for MDZ's per-pixel
 * escape-time path.
 *
 * Two layers are exported from the same shared library:
 *
 *  1. The reference's own render-pool API, unchanged: the fourteen rth_*
 *     functions and the rthdata struct of reference src/render_threads.h:20-87.
 *     Those are declared in include/mdz_rth.h.  Linking MDZ against libmdzcuda
 *     instead of compiling src/render_threads.c is the whole integration
 *     (INTEGRATION.md).
 *
 *  2. The plain entry points below, which the rth_* layer is built on and which
 *     a host without MDZ's image_info (tests, bench.py, another language's FFI)
 *     can bind directly.  Only C scalars, plain pointers and the public
 *     mpfr_t / mpf_t structs cross this boundary.
 *
 * A render is described by an mdzcuda_view: exactly the fields that the
 * reference's three line drivers read from image_info at render time
 * (src/fractal.c:29-117, :120-257, :260-397; field list in
 * src/image_info.h:63-122).
 */
#ifndef MDZCUDA_H
#define MDZCUDA_H

#include <stdint.h>
#include "mdz_mp_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* which reference line driver is being replaced */
/* MODE_LD reproduces x87 extended precision: 64-bit significand, round to nearest even.  The exponent is
 * an int32 here, so an orbit that squares itself below 2^-16382 (Julia, c = 0) keeps being squared where
 * the x87 would go through denormals to zero; such a pixel never escapes in either arithmetic, so no
 * iteration count depends on it (DESIGN.md "exponent range").  Coordinates beyond 2^+-16000 are rejected
 * as zero / infinity by the prologue. */
#define MDZCUDA_MODE_LD    0   /* fractal_calculate_line       src/fractal.c:29  (x87 long double)   */
#define MDZCUDA_MODE_MPFR  1   /* fractal_mpfr_calculate_line  src/fractal.c:120 (MPFR, RNDN)        */
#define MDZCUDA_MODE_GMP   2   /* fractal_gmp_calculate_line   src/fractal.c:260 (GMP mpf, truncate) */

/* src/fractal.h:8-26 */
#define MDZCUDA_FAMILY_MANDEL 0
#define MDZCUDA_FAMILY_JULIA  1
#define MDZCUDA_FRACTAL_MANDELBROT         0
#define MDZCUDA_FRACTAL_BURNING_SHIP       1
#define MDZCUDA_FRACTAL_GENERALIZED_CELTIC 2
#define MDZCUDA_FRACTAL_VARIANT            3

typedef struct mdzcuda_view {
    int   mode;                 /* MDZCUDA_MODE_*  (image_info.c:238-248 picks the callback) */
    long  precision;            /* img->precision, bits (ignored for MODE_LD)                */
    int   family;               /* img->family                                               */
    int   fractal;              /* img->fractal                                              */
    long  depth;                /* img->depth, 1..INT32_MAX                                  */
    int   real_width;           /* img->real_width  = user_width  * aa_factor                */
    int   real_height;          /* img->real_height = user_height * aa_factor                */
    int   aa_factor;            /* img->aa_factor, >= 1                                      */
    /* MPFR rect as filled by coords_get_rect (render.c:34-35); LD mode reads
     * xmin/xmax/ymax (fractal.c:50-53), MPFR mode xmin/ymax/width (fractal.c:160-170) */
    const __mpfr_struct* xmin;
    const __mpfr_struct* xmax;
    const __mpfr_struct* ymax;
    const __mpfr_struct* width;
    /* GMP rect as filled by coords_get_rect_gmp (render.c:36-37; fractal.c:300-310) */
    const __mpf_struct*  gxmin;
    const __mpf_struct*  gymax;
    const __mpf_struct*  gwidth;
    /* img->u.julia.c_re / c_im, read only when family is julia (fractal.c:55-59,197-198,341-342) */
    const __mpfr_struct* julia_re;
    const __mpfr_struct* julia_im;
    /* Optional, GMP mode + julia family only: the constant as mpf.  NULL (what the rth_* layer passes)
     * reproduces the reference literally: julia_re / julia_im through the decimal text that
     * mpfr_snprintf("%.Re") prints (src/my_mpfr_to_str.c:68, src/coords.c:13-18, src/fractal.c:341-342),
     * which keeps ONE significant digit with MPFR >= 4.  A host that converts differently (the
     * "%Re" build of the oracle, the Python harness with fixed_re) fills these instead. */
    const __mpf_struct*  gjulia_re;
    const __mpf_struct*  gjulia_im;
} mdzcuda_view;

typedef struct mdzcuda_plan mdzcuda_plan;

/* Last error text for the calling thread ("" if none). */
const char* mdzcuda_last_error(void);

/* Number of CUDA devices visible; 0 (with an error text) if CUDA is unusable. */
int mdzcuda_device_count(void);

/*
 * 1 when a GPU kernel is instantiated for the view's mode and precision, else 0 with the reason in
 * mdzcuda_last_error().  MDZ's settings admit 80..99999999 bits (src/image_info.c:535); kernels exist
 * for long double, MPFR 33..8192 bits (one thread per pixel to 1024 bits, 8, 16 or 32 lanes per pixel above) and GMP mpf to 8000 bits (one thread to 896, lane groups above).  The rth_* layer renders everything else
 * with the line callback the host installed (src/image_info.c:243-248), i.e. on the CPU, with one line
 * on stderr -- see mdzcuda_fallback_lines.
 */
int mdzcuda_view_supported(const mdzcuda_view* view);

/*
 * Image lines (real lines) that the rth_* layer has rendered through the host's next_line callback
 * instead of on the GPU since the library was loaded: unsupported precision / mode, no CUDA device, or
 * a CUDA failure in mid-render.  0 means every line so far came from the CUDA kernels; the GPU parity
 * tests and bench.py assert exactly that.  The plain entry points below never fall back: they fail.
 */
long mdzcuda_fallback_lines(void);

/*
 * Build a render plan on one device: runs the O(W+H) host prologue (column
 * and row coordinates computed with the same libmpfr / libgmp / long double
 * operations the reference's line drivers use) and uploads the tables.
 * The plan renders bands band_first, band_first+band_stride, ... where a band
 * is aa_factor consecutive real lines (the unit rth_next_line hands out,
 * src/render_threads.c:366-377).  band_first=0, band_stride=1 is the whole image.
 * Returns NULL on failure.
 */
mdzcuda_plan* mdzcuda_plan_create(const mdzcuda_view* view, int device,
                                  int band_first, int band_stride);

/*
 * Band scheduler (several devices on one image; the reference's analogue is rth_next_line,
 * src/render_threads.c:360-393, handing out one band at a time under a mutex).  A *fed* plan covers
 * the whole image (band_first 0, band_stride 1) but renders only the bands the host feeds it, while
 * its persistent kernel is running: mdzcuda_plan_feed appends `count` bands to the plan's queue and,
 * with close != 0, tells the kernel that no more will come (it then ends once its queue is done).
 * mdzcuda_plan_backlog is the number of pixels fed but not yet claimed by a lane, as of the last
 * mdzcuda_plan_poll_bands.  mdzcuda_render drives these for ndev > 1; they are exported for hosts
 * that run their own scheduler.  mdzcuda_plan_stream is the plan's own non-blocking stream.
 */
int       mdzcuda_plan_set_fed(mdzcuda_plan*, int on);
int       mdzcuda_plan_feed(mdzcuda_plan*, const int* bands, int count, int close);
long long mdzcuda_plan_backlog(mdzcuda_plan*);
void*     mdzcuda_plan_stream(mdzcuda_plan*);

/*
 * The sequence in which a plan's pixel queue visits its bands.  RASTER (default; what the rth_* layer uses):
 * top to bottom, the order in which the reference's pool hands lines out (src/render_threads.c:366-369), so
 * that a consumer sees the image grow from the top.  CENTRE_OUT: from the middle of the image outwards -- a
 * plan's bands are cut into up to 16 tiles each and the tiles taken by distance from the image centre (fed
 * plans: whole bands, from the middle band outwards).  A pixel that
 * runs to depth occupies its lane for `depth` dependent iterations -- half a second at 512 bits and depth
 * 100000 -- and a render cannot end before the last one started has finished; deep-zoom views keep their
 * dense part (a minibrot) near the centre, so starting there takes that latency out of the tail: it is what
 * mdzcuda_render does (MDZCUDA_ORDER=raster turns it off) and what SURVEY 8(e) asks of the tile order.
 * raw_data does not depend on it.
 */
#define MDZCUDA_ORDER_RASTER     0
#define MDZCUDA_ORDER_CENTRE_OUT 1
int mdzcuda_plan_set_order(mdzcuda_plan*, int order);

/* Tunables (before launch): iterations between queue refills (0 = default; a
 * negative value -n means n with the speculative iteration body switched off,
 * for A/B measurements), resident blocks per SM (0 = occupancy maximum). */
int mdzcuda_plan_tune(mdzcuda_plan*, int chunk_iters, int blocks_per_sm);

/*
 * Exact periodicity check (off by default for a plan; MDZCUDA_CYCLE_DETECT=1 turns it on
 * for every plan, and the rth_* layer turns it on unless MDZCUDA_CYCLE_DETECT=0).  When the
 * orbit's state (wre, wim) repeats bit for bit, the reference's loop
 * (`for (wz = 1; wz <= depth; ++wz)`, src/frac_mandel.c:11-19 / :34-50) can only run to
 * depth and return 0, so the pixel is finished with 0 at once: identical raw_data, a
 * fraction of the iterations for views with interior pixels.  Long double and MPFR modes;
 * ignored in GMP mode.  Throughput figures (bench.py `value`, `e2e`) are measured with it
 * off, because with it on "iterations performed" is no longer the reference's count.
 */
int mdzcuda_plan_set_cycle_detection(mdzcuda_plan*, int on);

/* Tail compaction ("parking").  The render becomes two launches of the same kernel:
 * the first stops when the pixel queue runs dry and writes the state of every pixel still in
 * flight to device memory; the second resumes them, 32 per warp, spread evenly over the SMs.  It
 * pays when the plan's image is only a few times the persistent grid -- one GPU's share of a
 * strong-scaled render -- where the last generation of long-running pixels would otherwise occupy
 * every warp sparsely.  The state is carried over bit for bit: raw_data is identical with it on or
 * off.  mode: -1 automatic (MDZCUDA_PARK=0/1 overrides), 0 off, 1 on.  Compiled into the long double and MPFR kernels of up to 128 bits (4 limbs); a no-op for wider ones and in GMP mode.
 * Scheduling only: the reference hands out whole lines under a mutex (src/render_threads.c:360-393)
 * and has no counterpart. */
int mdzcuda_plan_set_parking(mdzcuda_plan*, int mode);

/* Kernels this plan has launched so far (escape-time launches count 1, or 3 with parking:
 * phase 0, the ordering pass, phase 1; a recolour counts 1).  For reports (bench.py's
 * gpu_launches); -1 on a null plan. */
int mdzcuda_plan_kernels_launched(mdzcuda_plan*);

/* Enqueue the reset + escape-time kernel on `cuda_stream` (a cudaStream_t; NULL
 * = the legacy default stream).  Asynchronous. */
int mdzcuda_plan_launch(mdzcuda_plan*, void* cuda_stream);

/* Block until the last launch has finished. */
int mdzcuda_plan_wait(mdzcuda_plan*);

/* Ask a running launch to stop at its next poll (rth_ui_stop_render). */
int mdzcuda_plan_cancel(mdzcuda_plan*);

/* Bands completed so far by the current / last launch, and the plan's total. */
int mdzcuda_plan_bands_done(mdzcuda_plan*);
int mdzcuda_plan_bands_total(mdzcuda_plan*);

/* Progressive delivery while a launch is running (what rth_process_lines_rendered
 * is fed from): copy the per-band completion flags (bands_total bytes, 1 = all
 * aa_factor*real_width supersamples of that band are final) and return how many
 * are set; copy `count` local bands starting at `first_local_band` into their
 * place in a full-size host raw_data array. */
int mdzcuda_plan_poll_bands(mdzcuda_plan*, unsigned char* flags_host);
int mdzcuda_plan_fetch_bands(mdzcuda_plan*, int32_t* raw_host, int first_local_band, int count);

/* Copy this plan's lines into a full-size host raw_data array
 * (real_width*real_height int32, img->raw_data layout: line*real_width+ix). */
int mdzcuda_plan_fetch(mdzcuda_plan*, int32_t* raw_host);

/* Device pointer of the plan's iteration buffer ([local_lines][real_width] int32)
 * and its line count, for callers that keep results on the GPU. */
void* mdzcuda_plan_device_raw(mdzcuda_plan*);
int   mdzcuda_plan_local_lines(mdzcuda_plan*);

/*
 * Colour epilogue (optional).  With colour parameters set before a launch, the warp
 * that completes a band of aa_factor lines colours it on the spot -- palette lookup
 * or interpolation, anti-aliasing box average and pal_offset exactly as
 * palette_apply (src/palette.c:406-463), get_pixel_colour (:466-503) and
 * do_anti_aliasing (src/render.c:108-158) do -- into a device rgb buffer
 * ([user_height][user_width] guint32, R | G<<8 | B<<16).  mdzcuda_plan_recolour
 * re-runs only that step from the resident iteration counts (palette cycling,
 * src/main_gui.c:200-219).  Passing NULL switches the epilogue off again.
 */
typedef struct mdzcuda_colour {
    const uint32_t* palette;    /* globals.c:3 `palette`, pal_indexes entries   */
    int    pal_indexes;         /* palette.c:15, 2..256                         */
    int    pal_offset;          /* palette.c:14                                 */
    double colour_scale;        /* img->colour_scale                            */
    int    palette_ip;          /* img->palette_ip                              */
} mdzcuda_colour;
int mdzcuda_plan_set_colour(mdzcuda_plan*, const mdzcuda_colour*);
int mdzcuda_plan_recolour(mdzcuda_plan*, void* cuda_stream);
int mdzcuda_plan_fetch_rgb(mdzcuda_plan*, uint32_t* rgb_host /* user_width*user_height */);

/* Static facts about the kernel chosen for this plan (for reports). */
typedef struct mdzcuda_kernel_info {
    int limbs;              /* 32-bit limbs of significand                */
    int regs_per_thread;
    int local_bytes;        /* spill / local memory per thread            */
    int shared_bytes;       /* dynamic shared memory per block            */
    int block_threads;
    int blocks_per_sm;
    int grid_blocks;
    int sm_count;
    int lanes_per_pixel;    /* 1: one thread per pixel; 8 / 16 / 32: a group of lanes per pixel (MPFR above 1024 bits, GMP above 896) */
} mdzcuda_kernel_info;
int mdzcuda_plan_kernel_info(mdzcuda_plan*, mdzcuda_kernel_info* out);

/* Launch on `cuda_stream` and deliver: finished bands are copied into the full-size
 * host raw_data array while the kernel is still running, so the call returns shortly
 * after the last pixel.  Equivalent to launch + fetch. */
int mdzcuda_plan_run(mdzcuda_plan*, void* cuda_stream, int32_t* raw_host);

/* Returns the plan's device buffers, side stream and event to a per-device pool that
 * the next plan draws from (MDZ re-renders on every mouse move in its Julia preview,
 * src/main_gui.c:786-793; cudaMalloc/cudaFree per render cost milliseconds). */
void mdzcuda_plan_destroy(mdzcuda_plan*);

/* Give the pooled device memory (at most MDZCUDA_POOL_MB, default 4096), streams
 * and events back to the driver. */
void mdzcuda_trim(void);

/*
 * One-call render: host view in, host raw_data out, over ndev devices (devices == NULL means
 * 0..ndev-1).  With one device the image is one plan; with several, every device gets a fed plan and
 * a host-side scheduler hands out chunks of bands on demand -- large chunks first, small ones at the
 * end -- so that a device that is slower (busy with other work, or holding the deep part of the
 * image) simply takes fewer (MDZCUDA_SCHED=static: fixed interleave instead, for A/B runs).  This is
 * what the rth_* layer runs per render.  Returns 1 on success.
 */
int mdzcuda_render(const mdzcuda_view* view, int32_t* raw_host,
                   int ndev, const int* devices);

/*
 * Measured integer-multiply peaks of a device (register-only microbenchmarks,
 * about `ms` milliseconds each), in operations per second:
 *   mdzcuda_imad_peak    32x32->64 multiply-accumulates issued as IMAD.WIDE.U32.X
 *                        carry chains, the unit SURVEY 8(d) counts (W(N) = 2N^2+N);
 *   mdzcuda_imad32_peak  independent 32-bit IMAD (low half), the pipe's nominal
 *                        "64 IMAD/clk/SM" issue rate.
 * On B200 the 64-bit form runs at half the rate of the 32-bit one.
 */
double mdzcuda_imad_peak(int device, int ms);
double mdzcuda_imad32_peak(int device, int ms);

/* Test hook: occupy `blocks` SMs of `device` for `ms` milliseconds with a kernel that leaves no room
 * beside it (a device busy with someone else's work, for the band scheduler's tests).  Asynchronous. */
int mdzcuda_debug_occupy(int device, int blocks, int ms);

#ifdef __cplusplus
}
#endif
#endif /* MDZCUDA_H */
