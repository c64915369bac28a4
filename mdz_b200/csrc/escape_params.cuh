// escape_params.cuh -- kernel parameter block shared by the host API (mdzcuda.cu) and the
// kernel translation units.
#pragma once
#include <stdint.h>
#include "mpfr_sf.cuh"
#include "escape_step.cuh"
#include "colour.cuh"

namespace mdz {

// Column / row tables, limb-major so that lanes on neighbouring pixels coalesce.
//   m[k * count + i]  limb k (0 = least significant) of entry i
//   e[i]              exponent (E_ZERO for zero)
//   s[i]              1 = negative
struct CoordTable {
    const uint32_t* m;
    const int32_t*  e;
    const uint32_t* s;
    int count;
};

// Tail compaction is compiled into the kernels of at most this many limbs (escape_kernel.cuh Parkable)
// and mdzcuda.cu parks only plans of such kernels.
constexpr int kParkMaxLimbs = 4;

struct EscapeParams {
    CoordTable xs;          // real_width entries: x[ix]      (fractal.c:183-186)
    CoordTable ys;          // one entry per local line: y[line] (fractal.c:167-170)
    CoordTable jc;          // 2 entries: julia c_re, c_im    (fractal.c:197-198)
    RoundCfg rc;
    int32_t* raw;           // [local_lines][width] iteration counts
    unsigned int* queue;    // next pixel index
    unsigned int* band_count;   // finished supersamples per band of aa lines
    volatile unsigned int* bands_done;   // number of completed bands
    unsigned int* band_flag;    // [band] = launch generation once the band is complete (host polls a mirror)
    unsigned int gen;           // this launch's generation: flags left by an earlier launch or plan never match
    const volatile unsigned int* cancel; // == gen: stop (written by the host from a side stream, rth_ui_stop_render)
    int width;              // real width
    int lines;              // local line count (multiple of aa)
    int aa;
    int depth;
    int family;
    int fractal;
    int chunk;              // iterations between refills
    int spec;               // 1: try the speculative branch-free iteration first
    int cycle;              // 1: exact periodicity check (escape_kernel.cuh); MPFR / long double kernels only
    uint32_t* cycle_scratch;    // [2N+3][grid threads] saved states, when cycle != 0
    ColourParams colour;    // fused epilogue: colour a band as soon as it completes (enabled = 0: raw only)
    // Tail compaction (escape_kernel.cuh "Parking"): phase 0 stops when the pixel queue runs dry and
    // writes the state of the pixels still in flight to park_buf; phase 1 -- a second launch of the
    // same kernel -- resumes them, 32 per warp.  park_cap = 0: off (one launch does everything).
    int phase;                      // 0: pixels come from the queue; 1: from the parked list
    unsigned int park_cap;          // entries the list holds (= threads of the grid)
    unsigned int* park_count;       // [0] entries parked, [1] phase 1's queue counter
    uint32_t* park_buf;             // [ParkWords][park_cap]
    const unsigned int* park_perm;  // phase 1 reads the list in this order: by iteration count (park_sort_kernel)
    unsigned int* park_smslot;      // [4096] phase 1: arrivals per SM, dense SM numbers   } zeroed by
    unsigned int* park_claimed;     // [park_cap / 32] phase 1: group of 32 handed out      } park_sort_kernel
    int park_sms;                   // SMs of the device
    // Fed plans (several devices on one image, mdzcuda.cu "band scheduler"): the queue runs over slots of
    // one band each that the host fills while the kernel runs.  nullptr: static plan, queue position = pixel.
    const unsigned int* order;              // [slot] -> band * tiles_per_band + tile (a tile: aa lines x tile_w columns)
    int tile_w;                             // columns per tile (width / tiles_per_band); whole bands: width
    int tiles_per_band;
    const volatile unsigned int* feed;      // [0] slots filled so far, [1] == gen: no more will come
    int prec_bits;          // MPFR precision in bits (the warp-per-pixel kernels derive their rounding position from it)
    Ld64Masks ld_masks;     // ld64_masks(fractal), filled in by the host so that the hot loop reads them as constants
};

constexpr int kBlock = 128;

}  // namespace mdz
