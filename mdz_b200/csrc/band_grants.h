// band_grants.h -- how the host-side scheduler cuts an image into chunks of bands for the devices
// that share it (mdzcuda.cu: mdz_run_view).  Plain C++, no CUDA: tests/host_emu/sched_test.cpp
// drives it with simulated devices.
//
// The reference's pool hands out one band (aa_factor lines) at a time under a mutex, to whichever
// worker asks next (src/render_threads.c:360-393).  Here a "worker" is a GPU whose persistent kernel
// eats bands by the dozen, and a grant costs two small copies over PCIe, so the unit is a chunk:
// a quarter of the device's share of what is left (guided self-scheduling: large chunks first, small
// ones at the end), never less than `min_chunk`.  The share is even until the devices have shown
// their pace, then proportional to the bands each has consumed so far -- a device that is slow (busy
// with someone else's work) gets small chunks as well as fewer of them, or one expensive chunk in its
// queue would decide when the render ends.  A device is given its next chunk when the part of its
// queue that no lane has claimed yet falls below its low-water mark.
#pragma once
#include <vector>

namespace mdz {

struct BandGrants {
    int total;          // bands of the image
    int next;           // first band not handed out yet
    int ndev;
    int min_chunk;
    std::vector<double> consumed;       // bands' worth of pixels each device has started so far
    BandGrants(int total_, int ndev_, int min_chunk_)
        : total(total_), next(0), ndev(ndev_ > 0 ? ndev_ : 1), min_chunk(min_chunk_ > 0 ? min_chunk_ : 1),
          consumed((size_t)(ndev_ > 0 ? ndev_ : 1), 0.0) {}
    int remaining() const { return total - next; }
    void progress(int dev, double bands_started) { if (dev >= 0 && dev < ndev) consumed[(size_t)dev] = bands_started; }
    // the next chunk for `dev`: [*first, *first + n); n = 0 when everything has been handed out
    int take(int dev, int* first)
    {
        const int rem = total - next;
        if (rem <= 0) return 0;
        double share = 1.0 / ndev, sum = 0.0;
        for (int i = 0; i < ndev; ++i) sum += consumed[(size_t)i];
        if (sum >= 2.0 * min_chunk * ndev && dev >= 0 && dev < ndev) share = consumed[(size_t)dev] / sum;
        int n = (int)(rem * share / 4.0);
        if (n < min_chunk) n = min_chunk;
        if (n > rem) n = rem;
        *first = next;
        next += n;
        return n;
    }
};

}  // namespace mdz
