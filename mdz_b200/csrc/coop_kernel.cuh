// coop_kernel.cuh -- the escape-time loop with ONE WARP PER PIXEL, for MPFR precisions above 1024 bits
// (33 to 256 limbs: more than a thread can keep in registers).  Arithmetic: coop_ops.cuh.  Same pixel
// queue, band completion, cancel word and fed-plan protocol as the one-thread-per-pixel kernels
// (escape_kernel.cuh); a warp claims one pixel at a time.  Restates the reference's frac_*_mpfr loops
// (src/frac_mandel.c:25-52, src/frac_burning_ship.c:27-55, src/frac_generalized_celtic.c:27-55,
// src/frac_variant.c:26-55) and the per-pixel set-up of fractal_mpfr_calculate_line (src/fractal.c:183-203).
#pragma once
#include "escape_kernel.cuh"
#include "coop_ops.cuh"

namespace mdz {

// shared memory per block: one shifter strip per warp
template <int K> struct CoopSmemWords { static constexpr int value = CoopScratchWords<K>::value * (kBlock / 32); };
template <int K> struct CoopMinBlocks { static constexpr int value = K <= 2 ? 4 : K <= 4 ? 3 : 2; };

// table entry i: limb-major [32 K][count]; lane l takes limbs l*K .. l*K+K-1
template <int K>
__device__ __forceinline__ void load_coop_entry(const CoordTable& t, int i, CNum<K>& v)
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int j = 0; j < K; ++j) v.m[j] = __ldg(&t.m[(size_t)(lane * K + j) * t.count + i]);
    v.e = __ldg(&t.e[i]);
    v.s = __ldg(&t.s[i]);
}

template <int K>
__global__ void __launch_bounds__(kBlock, CoopMinBlocks<K>::value)
escape_coop_kernel(const EscapeParams p)
{
    extern __shared__ uint32_t csm[];
    uint32_t* scr = csm + (threadIdx.x >> 5) * CoopScratchWords<K>::value;
    coop_scratch_init<K>(scr);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)p.width * (unsigned)p.lines;
    CoopCfg cfg;
    cfg.prec = p.prec_bits;
    cfg.R = 32 * 32 * K - p.prec_bits;
    const bool abs_im = p.fractal == FRACTAL_BURNING_SHIP;
    const int  abs_re = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 1
                      : p.fractal == FRACTAL_VARIANT ? 2 : 0;

    CPixel<K> st;
    cset_zero(st.wre); cset_zero(st.wim); cset_zero(st.wre2); cset_zero(st.wim2); cset_zero(st.cre); cset_zero(st.cim);
    st.iter = 0;
    bool active = false;            // warp-uniform: the warp holds a pixel
    bool exhausted = false;
    unsigned pix = 0;               // warp-uniform (lane 0's claim, broadcast)
    int finished_band = -1;

    for (;;) {
        {
            int stop = 0;
            if (lane == 0) stop = *p.cancel == p.gen;
            if (__shfl_sync(0xffffffffu, stop, 0)) break;
        }
        if (!active) {
            bool start = false;
            if (!exhausted || (p.feed && (pix & kReserved) != 0u)) {
                // lane 0 claims for the warp; the others pose as busy
                start = claim_pixels(p, lane, total, lane != 0u, pix, exhausted);
                start = __shfl_sync(0xffffffffu, start ? 1 : 0, 0) != 0;
                pix = __shfl_sync(0xffffffffu, pix, 0);
            }
            if (start) {
                int line, ix;
                pixel_of_claim(p, pix, pix, line, ix);
                CNum<K> x, y;
                load_coop_entry<K>(p.xs, ix, x);
                load_coop_entry<K>(p.ys, line, y);
                if (p.family == FAMILY_JULIA) {
                    CNum<K> cx, cy;
                    load_coop_entry<K>(p.jc, 0, cx);
                    load_coop_entry<K>(p.jc, 1, cy);
                    cpixel_init<K>(st, x, y, cx, cy, cfg);
                } else cpixel_init<K>(st, x, y, x, y, cfg);
                active = true;
            }
            if (!active) { if (queue_idle_wait(p, pix)) continue; break; }
        }
        for (int k = 0; k < p.chunk; ++k) {
            const bool esc = cpixel_step<K>(st, cfg, scr, abs_im, abs_re);
            if (esc || st.iter >= p.depth) {
                if (lane == 0) {
                    p.raw[pix] = esc ? st.iter : 0;
                    __threadfence();
                    const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                    const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                    if (done == (unsigned)p.width * (unsigned)p.aa) finished_band = (int)band;
                }
                active = false;
            }
            if (publish_bands(p, finished_band, lane)) finished_band = -1;
            if (!active) break;
        }
    }
}

}  // namespace mdz
