// mdz_run.h -- private to libmdzcuda: the one render driver that mdzcuda_render (mdzcuda.cu) and the
// rth_* layer (rth.cpp) share.  Not part of the C ABI.
#pragma once
#include <stdint.h>
#include "../../include/mdzcuda.h"

struct mdz_run_hooks {
    void* user;
    int  (*should_stop)(void* user);                            // polled between deliveries; non-zero cancels the launch
    void (*bands_ready)(void* user, int first_band, int count, int stride); // raw_host now holds bands first, first+stride, ... (any order)
    int  min_run;               // deliver runs of at least this many finished bands while the kernels run (1: at once)
    int  order;                 // MDZCUDA_ORDER_*: the sequence in which bands are started
    int  cycle_detection;       // 0 / 1: mdzcuda_plan_set_cycle_detection for every plan; -1: the plans' default
};

// Render `view` into raw_host over the given devices: one static plan when ndev == 1; otherwise one fed plan
// per device and the host-side band scheduler (band_grants.h), or -- MDZCUDA_SCHED=static -- interleaved
// static plans.  Returns 1 on success, 0 on failure (mdzcuda_last_error), 2 when should_stop ended it early.
int mdz_run_view(const mdzcuda_view* view, int32_t* raw_host, const int* devices, int ndev, const mdz_run_hooks* hooks);

// lines rendered by the host's own line callback instead of the GPU (rth.cpp fallback) -> mdzcuda_fallback_lines()
void mdz_count_fallback_lines(long n);
