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
    Ld64Masks ld_masks;     // ld64_masks(fractal), filled in by the host so that the hot loop reads them as constants
};

constexpr int kBlock = 128;

}  // namespace mdz
