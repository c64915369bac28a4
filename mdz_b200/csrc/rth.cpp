// rth.cpp -- MDZ's render-pool API (reference src/render_threads.h:41-87) on top
// of the CUDA plans of mdzcuda.cu.  Linking MDZ against libmdzcuda instead of
// compiling src/render_threads.c replaces the pthread worker pool with the GPU;
// render.c, main_gui.c, image_info.c and the rest of MDZ stay as they are.
//
// What is kept from the reference, because its callers depend on it:
//   - rth_create / rth_init allocation and return conventions
//     (render_threads.c:77-183): 1 ok, 0 fail, NULL on OOM;
//   - one persistent watch thread per instance that, on every start signal,
//     stops and joins a render in progress before launching the next one --
//     start may arrive while rendering (render_threads.c:185-257; the Julia
//     preview does this on every mouse move, main_gui.c:786-793);
//   - the lines_rendered (0 -> 1 -> 2) / lines_drawn (0 -> 1) hand-off through
//     a window of line_draw_count+1 lines, its return values and its 0.5 ms
//     timed wait (render_threads.c:485-541, :557-574);
//   - the clock: started just before the work, stopped after it, and re-stopped
//     by every rth_ui_get_render_time call (render_threads.c:301,327,581-585);
//   - thread_count is halved until <= real_height/2 and published
//     (render_threads.c:293-297) although no host worker threads exist here.
// What replaces it: instead of N workers pulling lines under a mutex
// (render_threads.c:342-393) the render thread builds one plan per visible GPU
// (bands of aa_factor lines interleaved across devices), launches the
// persistent kernels, and polls band-completion flags, copying finished bands
// into img->raw_data and publishing them line by line.
//
// Views the GPU path cannot render -- a precision or mode no kernel is instantiated
// for (image_info.c:535 admits 80..99999999 bits), no CUDA device, a CUDA failure --
// are rendered by the line callback the HOST installed (image_info.c:243-248:
// fractal_calculate_line / fractal_mpfr_calculate_line / fractal_gmp_calculate_line),
// from a worker pool that restates render_threads.c:342-393, with one line on stderr;
// mdzcuda_fallback_lines() counts those lines, and every GPU test asserts it is 0.
// The library itself still contains no CPU implementation of the loop.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <errno.h>
#include <vector>
#include <string>

#include "../../include/mdz_rth.h"
#include "../../include/mdzcuda.h"
#include "mdz_run.h"

enum { RT_STOP = 0x0002, RT_QUIT = 0x0004, RT_RENDERING = 0x0008 };     // render_threads.c:9-14

struct rthpridata {
    char initialized;
    short thread_count;
    char start, started;
    int status;
    char* lines_rendered;
    int (*next_line_cb)(mdz_image_info*, int);
    int min_line_rendered;
    int total_lines_rendered;
    bool watch_running, render_running;
    pthread_t start_watch_thread, render_thread;
    pthread_mutex_t start_mutex, started_mutex, status_mutex, lines_rendered_mutex;
    pthread_cond_t start_cond, started_cond, lines_rendered_cond;
    struct timeval tv_start, tv_end;
};

static void* rth_watch(void* ptr);
static void* rth_render_main(void* ptr);

extern "C" rthdata* rth_create(void)
{
    rthdata* rth = (rthdata*)calloc(1, sizeof(rthdata));
    if (!rth) return 0;
    rth->data = (rthpridata*)calloc(1, sizeof(rthpridata));
    if (!rth->data) { free(rth); return 0; }
    return rth;
}

extern "C" int rth_init(rthdata* rth, int thread_count, int line_draw_count, mdz_image_info* img)
{
    if (img) rth->img = img;
    rthpridata* d = rth->data;
    if (thread_count) d->thread_count = (short)thread_count;
    d->start = 0;
    d->status = 0;
    free(d->lines_rendered);
    free(rth->lines_drawn);
    d->lines_rendered = (char*)malloc((size_t)rth->img->user_height);
    rth->lines_drawn = (char*)malloc((size_t)rth->img->user_height);
    if (!d->lines_rendered || !rth->lines_drawn) return 0;
    d->min_line_rendered = 0;
    rth->min_line_drawn = 0;
    rth->line_draw_count = line_draw_count;
    if (!d->initialized) {
        d->initialized = 1;
        pthread_mutex_init(&d->start_mutex, 0);
        pthread_mutex_init(&d->started_mutex, 0);
        pthread_mutex_init(&d->status_mutex, 0);
        pthread_mutex_init(&d->lines_rendered_mutex, 0);
        pthread_cond_init(&d->start_cond, 0);
        pthread_cond_init(&d->started_cond, 0);
        pthread_cond_init(&d->lines_rendered_cond, 0);
    }
    return 1;
}

extern "C" int rth_ui_init(rthdata* rth)
{
    if (pthread_create(&rth->data->start_watch_thread, 0, rth_watch, rth)) {
        fprintf(stderr, "Failed to create start_watch thread\n");
        return 0;
    }
    rth->data->watch_running = true;
    return 1;
}

static void* rth_watch(void* ptr)
{
    rthdata* rth = (rthdata*)ptr;
    rthpridata* d = rth->data;
    for (;;) {
        pthread_mutex_lock(&d->start_mutex);
        while (!d->start) {
            // rth_ui_quit signals start_cond without setting start
            pthread_mutex_lock(&d->status_mutex);
            const int q = d->status & RT_QUIT;
            pthread_mutex_unlock(&d->status_mutex);
            if (q) break;
            pthread_cond_wait(&d->start_cond, &d->start_mutex);
        }
        d->start = 0;
        pthread_mutex_unlock(&d->start_mutex);

        pthread_mutex_lock(&d->started_mutex);
        d->started = 0;
        pthread_mutex_unlock(&d->started_mutex);

        pthread_mutex_lock(&d->status_mutex);
        const int quit = d->status & RT_QUIT;
        int rendering = 0;
        if (d->status & RT_RENDERING) { d->status = RT_STOP | (quit ? RT_QUIT : 0); rendering = 1; }
        pthread_mutex_unlock(&d->status_mutex);

        if (d->render_running) {            // finished or stopping: reap it either way
            pthread_join(d->render_thread, 0);
            d->render_running = false;
        }
        (void)rendering;
        if (quit) return 0;
        if (pthread_create(&d->render_thread, 0, rth_render_main, rth)) return 0;
        d->render_running = true;
    }
}

static int stop_requested(rthpridata* d)
{
    pthread_mutex_lock(&d->status_mutex);
    const int s = d->status & (RT_STOP | RT_QUIT);
    pthread_mutex_unlock(&d->status_mutex);
    return s != 0;
}

static void publish_band(rthdata* rth, int band)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->lines_rendered_mutex);
    d->lines_rendered[band] = 1;
    d->total_lines_rendered += rth->img->aa_factor;
    pthread_cond_signal(&d->lines_rendered_cond);
    pthread_mutex_unlock(&d->lines_rendered_mutex);
}

static std::vector<int> pick_devices()
{
    std::vector<int> devs;
    const int n = mdzcuda_device_count();
    const char* env = getenv("MDZCUDA_DEVICES");            // e.g. "0,2,3"
    if (env && *env) {
        for (const char* p = env; *p;) {
            char* end;
            long v = strtol(p, &end, 10);
            if (end == p) break;
            if (v >= 0 && v < n) devs.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    if (devs.empty()) for (int i = 0; i < n; ++i) devs.push_back(i);
    return devs;
}

// ---- the GPU path: mdz_run_view (mdzcuda.cu) with progressive, in-order publication ----------------
struct RenderCtx {
    rthdata* rth;
    std::vector<unsigned char> ready;       // band fetched into img->raw_data
    int published;
};

static int hook_should_stop(void* u) { return stop_requested(((RenderCtx*)u)->rth->data); }

// Bands finish out of order on the device(s) but are published strictly in line order: the reference's
// consumers (render.c:49-92, main_gui.c:533-599) clip their window with the *count* of finished lines,
// which only describes the image when that count is a prefix of it.
static void hook_bands_ready(void* u, int first, int count, int stride)
{
    RenderCtx* c = (RenderCtx*)u;
    const int total = (int)c->ready.size();
    for (int k = 0; k < count; ++k) { const int b = first + k * stride; if (b >= 0 && b < total) c->ready[b] = 1; }
    while (c->published < total && c->ready[c->published]) publish_band(c->rth, c->published++);
}

// ---- the host path: the caller's own line callback, as the reference's pool runs it ------------------
// rth_render + rth_next_line (src/render_threads.c:342-393): N workers take the next band of aa_factor
// lines under a mutex, call next_line_cb(img, line) for each of its lines (the callback returns 0 when
// it saw a stop request, fractal.c:113), mark the band rendered and signal.  Used only when the GPU
// path cannot render the view -- a precision or mode without a kernel, no CUDA device, a CUDA failure --
// and counted in mdzcuda_fallback_lines().
struct HostPool {
    rthdata* rth;
    std::vector<int> bands;         // bands still to render, in order
    size_t next;
    pthread_mutex_t mu;
};

static void* host_worker(void* ptr)
{
    HostPool* hp = (HostPool*)ptr;
    rthdata* rth = hp->rth;
    rthpridata* d = rth->data;
    const int aa = rth->img->aa_factor < 1 ? 1 : rth->img->aa_factor;
    for (;;) {
        pthread_mutex_lock(&hp->mu);
        const size_t k = hp->next++;
        pthread_mutex_unlock(&hp->mu);
        if (k >= hp->bands.size()) return 0;
        const int band = hp->bands[k];
        for (int i = 0; i < aa; ++i)
            if (!d->next_line_cb(rth->img, band * aa + i)) return 0;
        mdz_count_fallback_lines(aa);
        pthread_mutex_lock(&d->lines_rendered_mutex);
        d->lines_rendered[band] = 1;
        d->total_lines_rendered += aa;
        pthread_cond_signal(&d->lines_rendered_cond);
        pthread_mutex_unlock(&d->lines_rendered_mutex);
        if (stop_requested(d)) return 0;
    }
}

static void run_host_pool(rthdata* rth, int first_band)
{
    HostPool hp;
    hp.rth = rth; hp.next = 0;
    for (int b = first_band; b < rth->img->user_height; ++b) hp.bands.push_back(b);
    pthread_mutex_init(&hp.mu, 0);
    int nt = rth->thread_count > 0 ? rth->thread_count : 1;
    if (nt > MAX_THREAD_COUNT) nt = MAX_THREAD_COUNT;
    std::vector<pthread_t> th((size_t)nt);
    int made = 0;
    for (; made < nt; ++made)
        if (pthread_create(&th[made], 0, host_worker, &hp)) {
            fprintf(stderr, "\nrender thread #%d creation failed\n", made);       // render_threads.c:308-313
            break;
        }
    if (made == 0) host_worker(&hp);
    for (int i = 0; i < made; ++i) pthread_join(th[i], 0);
    pthread_mutex_destroy(&hp.mu);
}

static void* rth_render_main(void* ptr)
{
    rthdata* rth = (rthdata*)ptr;
    rthpridata* d = rth->data;
    mdz_image_info* img = rth->img;

    pthread_mutex_lock(&d->lines_rendered_mutex);
    memset(d->lines_rendered, 0, (size_t)img->user_height);
    d->total_lines_rendered = 0;
    d->min_line_rendered = 0;
    pthread_mutex_unlock(&d->lines_rendered_mutex);

    pthread_mutex_lock(&d->status_mutex);
    d->status = RT_RENDERING;
    pthread_mutex_unlock(&d->status_mutex);

    if (!rth->check_stop_px) rth->check_stop_px = 64;
    rth->thread_count = d->thread_count;
    while (rth->thread_count > img->real_height / 2) rth->thread_count /= 2;

    gettimeofday(&d->tv_start, 0);

    pthread_mutex_lock(&d->started_mutex);
    d->started = 1;
    pthread_cond_signal(&d->started_cond);
    pthread_mutex_unlock(&d->started_mutex);

    // the view, exactly the fields the three line drivers read (fractal.c:29-117, :120-257, :260-397)
    mdzcuda_view v;
    memset(&v, 0, sizeof v);
    v.mode = !img->use_multi_prec ? MDZCUDA_MODE_LD : (img->use_rounding ? MDZCUDA_MODE_MPFR : MDZCUDA_MODE_GMP);
    v.precision = img->precision;
    v.family = img->family;
    v.fractal = img->fractal;
    v.depth = (long)img->depth;
    v.real_width = img->real_width;
    v.real_height = img->real_height;
    v.aa_factor = img->aa_factor < 1 ? 1 : img->aa_factor;
    v.xmin = img->xmin; v.xmax = img->xmax; v.ymax = img->ymax; v.width = img->width;
    v.gxmin = img->gxmin; v.gymax = img->gymax; v.gwidth = img->gwidth;
    v.julia_re = img->u.julia.c_re; v.julia_im = img->u.julia.c_im;

    RenderCtx ctx;
    ctx.rth = rth;
    ctx.ready.assign((size_t)(img->user_height > 0 ? img->user_height : 0), 0);
    ctx.published = 0;

    std::string err;
    int rc = 0;
    static const bool force_host = [] { const char* e = getenv("MDZCUDA_FORCE_HOST"); return e && *e && *e != '0'; }();
    std::vector<int> devs;
    if (force_host) err = "MDZCUDA_FORCE_HOST is set";          // test hook: exercise the host path on a GPU box
    else if (!mdzcuda_view_supported(&v)) err = mdzcuda_last_error();
    else {
        devs = pick_devices();
        if (devs.empty()) err = std::string("no CUDA device: ") + mdzcuda_last_error();
    }
    if (err.empty()) {
        mdz_run_hooks h;
        h.user = &ctx;
        h.should_stop = hook_should_stop;
        h.bands_ready = hook_bands_ready;
        h.min_run = 1;                                          // progressive delivery
        h.order = MDZCUDA_ORDER_RASTER;                         // top to bottom, as the reference's pool hands lines out (render_threads.c:366-369)
        // same raw_data, fewer iterations for interior pixels (include/mdzcuda.h); MDZ's own callers only
        // ever see the result, so it is on unless MDZCUDA_CYCLE_DETECT=0
        { const char* e = getenv("MDZCUDA_CYCLE_DETECT"); h.cycle_detection = !(e && *e == '0'); }
        rc = mdz_run_view(&v, img->raw_data, devs.data(), (int)devs.size(), &h);
        if (rc == 0) err = mdzcuda_last_error();
    }

    if (!err.empty() && !stop_requested(d)) {
        if (d->next_line_cb) {
            // "keep the callback as the CPU fallback when no GPU or an unsupported precision is
            // present ... a CUDA failure should fall back to the CPU callback rather than change these codes"
            fprintf(stderr, "\nlibmdzcuda: %s -- rendering %d band(s) with the host's own line callback on %d thread(s)\n",
                    err.c_str(), img->user_height - ctx.published, rth->thread_count > 0 ? rth->thread_count : 1);
            run_host_pool(rth, ctx.published);
        } else {
            // No callback was ever installed (MDZ always installs one, image_info.c:243-248): nothing can render
            // this view.  Reporting completion would hand the caller a cleared raw_data as if it were an image.
            fprintf(stderr, "\nlibmdzcuda: render failed: %s\nlibmdzcuda: no line callback installed to fall back on; giving up\n", err.c_str());
            exit(EXIT_FAILURE);
        }
    }

    gettimeofday(&d->tv_end, 0);
    pthread_mutex_lock(&d->status_mutex);
    d->status = RT_STOP | (d->status & RT_QUIT);
    pthread_mutex_unlock(&d->status_mutex);
    // wake anyone blocked in rth_ui_wait_for_line_done after a stop
    pthread_mutex_lock(&d->lines_rendered_mutex);
    pthread_cond_broadcast(&d->lines_rendered_cond);
    pthread_mutex_unlock(&d->lines_rendered_mutex);
    return 0;
}

extern "C" void rth_ui_start_render(rthdata* rth)
{
    rthpridata* d = rth->data;
    memset(rth->lines_drawn, 0, (size_t)rth->img->user_height);
    rth->min_line_drawn = 0;
    // The reference clears `started` in the watch thread after it has taken the start
    // signal (render_threads.c:206-208), so a caller that goes straight on to
    // rth_ui_wait_until_started can see the previous render's flag.  Clearing it here
    // closes that window; the new render thread sets it once its state is reset.
    pthread_mutex_lock(&d->started_mutex);
    d->started = 0;
    pthread_mutex_unlock(&d->started_mutex);
    pthread_mutex_lock(&d->start_mutex);
    d->start = 1;
    pthread_cond_signal(&d->start_cond);
    pthread_mutex_unlock(&d->start_mutex);
}

extern "C" void rth_ui_stop_render(rthdata* rth)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->status_mutex);
    d->status = RT_STOP | (d->status & RT_RENDERING);
    pthread_mutex_unlock(&d->status_mutex);
}

extern "C" void rth_ui_stop_render_and_wait(rthdata* rth)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->status_mutex);
    if (d->status & RT_STOP) { pthread_mutex_unlock(&d->status_mutex); return; }
    d->status = RT_STOP | (d->status & RT_RENDERING);
    pthread_mutex_unlock(&d->status_mutex);
    // the render thread notices within one poll; wait for it to finish
    for (;;) {
        pthread_mutex_lock(&d->status_mutex);
        const int r = d->status & RT_RENDERING;
        pthread_mutex_unlock(&d->status_mutex);
        if (!r) break;
        struct timespec nap = { 0, 200 * 1000 };
        nanosleep(&nap, 0);
    }
}

extern "C" void rth_ui_quit(rthdata* rth)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->status_mutex);
    d->status |= RT_QUIT;
    pthread_mutex_unlock(&d->status_mutex);
    pthread_mutex_lock(&d->start_mutex);
    pthread_cond_signal(&d->start_cond);
    pthread_mutex_unlock(&d->start_mutex);
    if (d->watch_running) { pthread_join(d->start_watch_thread, 0); d->watch_running = false; }
    if (d->render_running) { pthread_join(d->render_thread, 0); d->render_running = false; }
    free(d->lines_rendered); d->lines_rendered = 0;
    free(rth->lines_drawn); rth->lines_drawn = 0;
}

extern "C" void rth_ui_wait_until_started(rthdata* rth)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->started_mutex);
    while (!d->started) pthread_cond_wait(&d->started_cond, &d->started_mutex);
    pthread_cond_signal(&d->started_cond);
    pthread_mutex_unlock(&d->started_mutex);
}

extern "C" void rth_set_next_line_cb(rthdata* rth, int (*next_line_cb)(mdz_image_info*, int))
{
    rth->data->next_line_cb = next_line_cb;     // the host's own CPU path: only called when the GPU cannot render the view
}

extern "C" int rth_process_lines_rendered(rthdata* rth)
{
    rthpridata* d = rth->data;
    const int height = rth->img->user_height;
    pthread_mutex_lock(&d->lines_rendered_mutex);
    const int linesdone = d->total_lines_rendered / rth->img->aa_factor;
    pthread_mutex_unlock(&d->lines_rendered_mutex);
    const int miny = d->min_line_rendered;
    if (miny == linesdone) return 0;
    int maxy = miny + rth->line_draw_count + 1;
    if (maxy > height) maxy = height;
    if (maxy > linesdone) maxy = linesdone;
    if (miny < maxy) {
        struct timeval tv;
        struct timespec timeout;
        gettimeofday(&tv, 0);
        long ns = tv.tv_usec * 1000L + 500 * 1000L;         // 0.5 ms, as render_threads.c:510-513
        timeout.tv_sec = tv.tv_sec + ns / 1000000000L;
        timeout.tv_nsec = ns % 1000000000L;
        pthread_mutex_lock(&d->lines_rendered_mutex);
        // (Once every line is rendered no further signal can arrive, so there is nothing to wait for: the
        // reference sleeps its 0.5 ms regardless, which makes the Julia preview's consumer -- a window of
        // line_draw_count + 1 = 3 lines per call, image_info.c:50 -- spend 15 ms on a finished 90-line frame.)
        if (d->total_lines_rendered < rth->img->real_height)
            pthread_cond_timedwait(&d->lines_rendered_cond, &d->lines_rendered_mutex, &timeout);
        int unrendered = 0;
        for (int i = miny; i < maxy; ++i) {
            char* lr = &d->lines_rendered[i];
            if (*lr == 1) {
                *lr = 2;
                rth->lines_drawn[i] = 1;
                if (!unrendered) d->min_line_rendered = i;
            } else if (*lr == 0) unrendered = 1;
        }
        pthread_mutex_unlock(&d->lines_rendered_mutex);
    }
    return (linesdone < height) ? linesdone : -1;
}

extern "C" int rth_render_should_stop(rthdata* rth)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->status_mutex);
    const int ret = !!(d->status & RT_STOP);
    pthread_mutex_unlock(&d->status_mutex);
    return ret;
}

extern "C" int rth_ui_wait_for_line_done(rthdata* rth)
{
    rthpridata* d = rth->data;
    pthread_mutex_lock(&d->lines_rendered_mutex);
    if (d->total_lines_rendered == rth->img->real_height) {
        pthread_mutex_unlock(&d->lines_rendered_mutex);
        return -1;
    }
    // the reference waits without a timeout (render_threads.c:568); a bounded wait
    // returns the same values and cannot miss the final signal
    struct timeval tv;
    struct timespec timeout;
    gettimeofday(&tv, 0);
    long ns = tv.tv_usec * 1000L + 50 * 1000 * 1000L;
    timeout.tv_sec = tv.tv_sec + ns / 1000000000L;
    timeout.tv_nsec = ns % 1000000000L;
    pthread_cond_timedwait(&d->lines_rendered_cond, &d->lines_rendered_mutex, &timeout);
    const int linesdone = d->total_lines_rendered / rth->img->aa_factor;
    pthread_mutex_unlock(&d->lines_rendered_mutex);
    return (linesdone < rth->img->user_height) ? linesdone : -1;
}

extern "C" void rth_ui_stop_timer(rthdata* rth) { gettimeofday(&rth->data->tv_end, 0); }

extern "C" double rth_ui_get_render_time(rthdata* rth)
{
    rthpridata* d = rth->data;
    gettimeofday(&d->tv_end, 0);                 // re-stopped on every read (render_threads.c:581-585)
    const double us = (d->tv_end.tv_sec - d->tv_start.tv_sec) * 1e6 + (d->tv_end.tv_usec - d->tv_start.tv_usec);
    return us / 1e6;
}
