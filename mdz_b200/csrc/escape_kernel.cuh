// escape_kernel.cuh -- the per-pixel escape-time loop, MPFR-faithful mode.
//
// Restates, for one pixel per thread, the loop bodies of the reference's
// frac_mandel_mpfr (src/frac_mandel.c:25-52), frac_burning_ship_mpfr
// (src/frac_burning_ship.c:27-55), frac_generalized_celtic_mpfr
// (src/frac_generalized_celtic.c:27-55) and frac_variant_mpfr
// (src/frac_variant.c:26-55), and -- at p = 64 -- their long double twins
// (frac_mandel.c:5-21 etc.), with the per-pixel set-up of
// fractal_mpfr_calculate_line (src/fractal.c:183-203) fed from host-built
// column / row tables.
//
// Scheduling: persistent warps pull pixels from a device-side atomic queue
// (the reference hands out lines under a mutex, src/render_threads.c:360-393;
// escape times inside a line are wildly uneven, so the unit here is a pixel).
// Every `chunk` iterations a warp refills its finished lanes with one
// warp-aggregated atomicAdd.
#pragma once
#include "mpfr_sf.cuh"

namespace mdz {

enum { FRACTAL_MANDELBROT = 0, FRACTAL_BURNING_SHIP = 1, FRACTAL_GENERALIZED_CELTIC = 2, FRACTAL_VARIANT = 3 };
enum { FAMILY_MANDEL = 0, FAMILY_JULIA = 1 };

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
    unsigned char* band_flag;   // 1 when band complete (device copy, host polls a mirror)
    const volatile int* cancel; // device stop flag, set by the host from a side stream (rth_ui_stop_render)
    int width;              // real width
    int lines;              // local line count (multiple of aa)
    int aa;
    int depth;
    int family;
    int fractal;
    int chunk;              // iterations between refills
};

constexpr int kBlock = 128;

template <int N>
__device__ __forceinline__ void load_entry(const CoordTable& t, int i, Num<N>& v)
{
#pragma unroll
    for (int k = 0; k < N; ++k) v.m[k] = __ldg(&t.m[(size_t)k * t.count + i]);
    v.e = __ldg(&t.e[i]);
    v.s = __ldg(&t.s[i]);
}

template <int N>
__global__ void __launch_bounds__(kBlock)
escape_mpfr_kernel(const EscapeParams p)
{
    // per-thread copy of c (2N limbs), limb-major: conflict-free
    extern __shared__ uint32_t csm[];
    uint32_t* cre_m = csm + threadIdx.x;
    uint32_t* cim_m = csm + N * kBlock + threadIdx.x;

    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)p.width * (unsigned)p.lines;

    Num<N> wre, wim, wre2, wim2;
    int32_t cre_e = E_ZERO, cim_e = E_ZERO;
    uint32_t cre_s = 0, cim_s = 0;
    set_zero(wre); set_zero(wim); set_zero(wre2); set_zero(wim2);

    bool active = false;
    bool exhausted = false;         // warp-uniform
    int iter = 0;
    unsigned pix = 0;

    const bool abs_im = p.fractal == FRACTAL_BURNING_SHIP;
    const int  abs_re = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 1
                      : p.fractal == FRACTAL_VARIANT ? 2 : 0;

    for (;;) {
        // ---- cooperative cancel (reference polls every 64 px, fractal.c:113) --
        {
            int stop = 0;
            if (lane == 0) stop = *p.cancel;
            if (__shfl_sync(0xffffffffu, stop, 0)) break;
        }
        // ---- refill finished lanes from the queue ------------------------
        if (!exhausted) {
            const unsigned need = __ballot_sync(0xffffffffu, !active);
            if (need) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(p.queue, (unsigned)__popc(need));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + (unsigned)__popc(need) >= total) exhausted = true;
                if (!active) {
                    const unsigned idx = base + (unsigned)__popc(need & ((1u << lane) - 1u));
                    if (idx < total) {
                        pix = idx;
                        const int line = (int)(idx / (unsigned)p.width);
                        const int ix = (int)(idx - (unsigned)line * (unsigned)p.width);
                        Num<N> x, y;
                        load_entry<N>(p.xs, ix, x);
                        load_entry<N>(p.ys, line, y);
                        wre = x; wim = y;
                        fsqr<N>(x, wre2, p.rc);
                        fsqr<N>(y, wim2, p.rc);
                        if (p.family == FAMILY_JULIA) {
                            load_entry<N>(p.jc, 0, x);
                            load_entry<N>(p.jc, 1, y);
                        }
#pragma unroll
                        for (int k = 0; k < N; ++k) { cre_m[k * kBlock] = x.m[k]; cim_m[k * kBlock] = y.m[k]; }
                        cre_e = x.e; cre_s = x.s; cim_e = y.e; cim_s = y.s;
                        iter = 0;
                        active = true;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;

        // ---- iterate ------------------------------------------------------
        for (int k = 0; k < p.chunk; ++k) {
            if (active) {
                ++iter;
                Num<N> t, c;
                // wim = 2*wre*wim + c_im       (|.| on the product for burning ship)
                fmul<N>(wre, wim, t, p.rc);
                if (t.m[N - 1] != 0) t.e += 1;
                if (abs_im) t.s = 0;
#pragma unroll
                for (int q = 0; q < N; ++q) c.m[q] = cim_m[q * kBlock];
                c.e = cim_e; c.s = cim_s;
                fadd<N, MODE_GENERIC>(t, c, wim, p.rc);
                // wre = wre2 - wim2 + c_re     (|.| on the difference for celtic / odd steps of the hybrid)
                fadd<N, MODE_SUB_POS>(wre2, wim2, t, p.rc);
                if (abs_re == 1 || (abs_re == 2 && (iter & 1))) t.s = 0;
#pragma unroll
                for (int q = 0; q < N; ++q) c.m[q] = cre_m[q * kBlock];
                c.e = cre_e; c.s = cre_s;
                fadd<N, MODE_GENERIC>(t, c, wre, p.rc);
                fsqr<N>(wim, wim2, p.rc);
                fsqr<N>(wre, wre2, p.rc);
                // escape: RN(wim2 + wre2) > 4.  Both < 2 cannot exceed 4 even
                // after rounding; either >= 8 certainly does.
                const int32_t emax = wim2.e > wre2.e ? wim2.e : wre2.e;
                bool esc = emax >= 4;
                if (!esc && emax >= 2) {
                    fadd<N, MODE_ADD_POS>(wim2, wre2, t, p.rc);
                    esc = greater_than_4<N>(t);
                }
                if (esc || iter >= p.depth) {
                    p.raw[pix] = esc ? iter : 0;
                    __threadfence();            // raw visible before the band counter moves
                    active = false;
                    const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                    const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                    if (done == (unsigned)p.width * (unsigned)p.aa) {
                        __threadfence();
                        p.band_flag[band] = 1;
                        atomicAdd((unsigned int*)p.bands_done, 1u);
                    }
                }
            }
            if (!__any_sync(0xffffffffu, active)) break;
        }
    }
}

}  // namespace mdz
