// tests/host_emu/rth_stub_backend.cpp -- TEST BUILD ONLY.
// Stands in for mdzcuda.cu under mdz_b200/csrc/rth.cpp so that the rth_* protocol layer (watch thread, render
// thread, in-order publication, stop / restart / quit) can be built with -fsanitize=thread and driven by
// tests/host_emu/rth_protocol.c on a box without a GPU (SURVEY 4 "protocol tests ... under ThreadSanitizer";
// the reference's own protocol has a known intermittent hang, BUGS:1-6).  The "render" writes line + 1 into
// every pixel, delivers bands slightly out of order, a few at a time, and honours the stop hook.
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <vector>
#include "../../mdz_b200/csrc/mdz_run.h"

extern "C" int mdzcuda_device_count(void) { return 1; }
extern "C" const char* mdzcuda_last_error(void) { return ""; }
extern "C" int mdzcuda_view_supported(const mdzcuda_view*) { return 1; }
static long g_fallback = 0;
void mdz_count_fallback_lines(long n) { __atomic_add_fetch(&g_fallback, n, __ATOMIC_RELAXED); }
extern "C" long mdzcuda_fallback_lines(void) { return __atomic_load_n(&g_fallback, __ATOMIC_RELAXED); }

int mdz_run_view(const mdzcuda_view* v, int32_t* raw, const int*, int, const mdz_run_hooks* h)
{
    const int aa = v->aa_factor, W = v->real_width, bands = v->real_height / aa;
    const struct timespec nap = { 0, 150 * 1000 };
    for (int b0 = 0; b0 < bands; b0 += 4) {
        if (h && h->should_stop && h->should_stop(h->user)) return 2;
        const int n = bands - b0 < 4 ? bands - b0 : 4;
        // a group of four, last band first: completion out of order, as on the device
        for (int k = n - 1; k >= 0; --k) {
            const int b = b0 + k;
            for (int l = b * aa; l < (b + 1) * aa; ++l)
                for (int x = 0; x < W; ++x) raw[(size_t)l * W + x] = l + 1;
            if (h && h->bands_ready) h->bands_ready(h->user, b, 1, 1);
        }
        nanosleep(&nap, 0);
    }
    return 1;
}
