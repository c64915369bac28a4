// coop_kernel.cuh -- the escape-time loop with A GROUP OF LANES PER PIXEL, for MPFR precisions above 1024 bits
// (33 to 256 limbs: more than a thread can keep in registers).  T lanes hold a pixel's values, K limbs each
// (coop_ops.cuh): 8 x 6, 8 x 8 and 16 x 8 -- four, four and two pixels per warp -- for up to 1536, 2048 and 4096 bits, 32 x 6 and 32 x 8 for
// 6144 and 8192.  Same pixel queue, band completion, cancel word and fed-plan protocol as the one-thread-per-pixel
// kernels (escape_kernel.cuh); a group claims one pixel at a time.  Restates the reference's frac_*_mpfr loops
// (src/frac_mandel.c:25-52, src/frac_burning_ship.c:27-55, src/frac_generalized_celtic.c:27-55,
// src/frac_variant.c:26-55) and the per-pixel set-up of fractal_mpfr_calculate_line (src/fractal.c:183-203); with
// GMP = true the frac_*_gmp loops (src/frac_mandel.c:55-82 ...) and fractal_gmp_calculate_line (src/fractal.c:310-342)
// on coop_mpf.cuh's arithmetic, for mpf precisions above 512 bits.
#pragma once
#include "escape_kernel.cuh"
#include "coop_ops.cuh"
#include "coop_mpf.cuh"

namespace mdz {

// shared memory per block: one shifter strip per group
template <int K, int T> struct CoopSmemWords { static constexpr int value = CoopScratchWords<K, T>::value * (kBlock / T); };
template <int K, int T> struct CoopMinBlocks { static constexpr int value = K <= 4 ? 4 : 3; };

// table entry i: limb-major [T K][count]; lane l of the group takes limbs l*K .. l*K+K-1
template <int K, int T>
__device__ __forceinline__ void load_coop_entry(const CoordTable& t, int i, CNum<K, T>& v)
{
    const unsigned gl = threadIdx.x & (unsigned)(T - 1);
#pragma unroll
    for (int j = 0; j < K; ++j) v.m[j] = __ldg(&t.m[(size_t)(gl * K + j) * t.count + i]);
    v.e = __ldg(&t.e[i]);
    v.s = __ldg(&t.s[i]);
}

template <int K, int T, bool GMP>
__global__ void __launch_bounds__(kBlock, CoopMinBlocks<K, T>::value)
escape_coop_kernel(const EscapeParams p)
{
    extern __shared__ uint32_t csm[];
    uint32_t* scr = csm + (threadIdx.x / T) * CoopScratchWords<K, T>::value;
    coop_scratch_init<K, T>(scr);

    const unsigned lane = threadIdx.x & 31u;
    const bool leader = (threadIdx.x & (unsigned)(T - 1)) == 0u;
    const unsigned total = (unsigned)p.width * (unsigned)p.lines;
    CoopCfg cfg;                     // MPFR: prec_bits is the precision
    cfg.prec = p.prec_bits;
    cfg.R = 32 * T * K - p.prec_bits;
    const CoopGCfg gcfg = make_coop_gcfg<K, T>(GMP ? p.prec_bits : 1);      // GMP: prec_bits carries NL = P + 1 limbs
    const bool abs_im = p.fractal == FRACTAL_BURNING_SHIP;
    const int  abs_re = p.fractal == FRACTAL_GENERALIZED_CELTIC ? 1
                      : p.fractal == FRACTAL_VARIANT ? 2 : 0;

    CPixel<K, T> st;
    cset_zero(st.wre); cset_zero(st.wim); cset_zero(st.wre2); cset_zero(st.wim2); cset_zero(st.cre); cset_zero(st.cim);
    st.iter = 0;
    bool active = false;            // uniform over the group: it holds a pixel
    bool exhausted = false;         // uniform over the warp
    unsigned pix = 0;               // uniform over the group (its leader's claim, broadcast)
    int finished_band = -1;

    for (;;) {
        __syncwarp();
        {
            int stop = 0;
            if (lane == 0) stop = *p.cancel == p.gen;
            if (__shfl_sync(0xffffffffu, stop, 0)) break;
        }
        if (__any_sync(0xffffffffu, !active)) {
            bool start = false;
            if (!exhausted || (p.feed && __any_sync(0xffffffffu, !active && (pix & kReserved) != 0u))) {
                // the leaders of the idle groups claim (one warp-aggregated atomicAdd); everybody else poses as busy
                start = claim_pixels(p, lane, total, !leader || active, pix, exhausted);
                start = bcast<T>(start ? 1u : 0u, 0) != 0u;
                pix = bcast<T>(pix, 0);
            }
            if (start) {
                int line, ix;
                pixel_of_claim(p, pix, pix, line, ix);
                CNum<K, T> x, y;
                load_coop_entry<K, T>(p.xs, ix, x);
                load_coop_entry<K, T>(p.ys, line, y);
                if (GMP) { cg_adopt<K, T>(x); cg_adopt<K, T>(y); }
                if (p.family == FAMILY_JULIA) {
                    CNum<K, T> cx, cy;
                    load_coop_entry<K, T>(p.jc, 0, cx);
                    load_coop_entry<K, T>(p.jc, 1, cy);
                    if (GMP) { cg_adopt<K, T>(cx); cg_adopt<K, T>(cy); }
                    cpixel_load<K, T>(st, x, y, cx, cy);
                } else cpixel_load<K, T>(st, x, y, x, y);
                active = true;
            }
            __syncwarp();
            if (__any_sync(0xffffffffu, start)) {                                   // products are the whole warp's
                if (GMP) cgpixel_squares<K, T>(st, gcfg); else cpixel_squares<K, T>(st, cfg);
            }
            if (!__any_sync(0xffffffffu, active)) { if (queue_idle_wait(p, pix)) continue; break; }
        }
        for (int k = 0; k < p.chunk; ++k) {
            {
                const bool esc = GMP ? cgpixel_step<K, T>(st, gcfg, scr, abs_im, abs_re, active)
                                     : cpixel_step<K, T>(st, cfg, scr, abs_im, abs_re, active);
                if (active && (esc || st.iter >= p.depth)) {
                    if (leader) {
                        p.raw[pix] = esc ? st.iter : 0;
                        __threadfence();
                        const unsigned band = (pix / (unsigned)p.width) / (unsigned)p.aa;
                        const unsigned done = atomicAdd(&p.band_count[band], 1u) + 1u;
                        if (done == (unsigned)p.width * (unsigned)p.aa) finished_band = (int)band;
                    }
                    active = false;
                }
            }
            __syncwarp();
            if (publish_bands(p, finished_band, lane)) finished_band = -1;
            if (__any_sync(0xffffffffu, !active)) {
                // a group is free: refill before going on -- unless the queue has nothing left to give, in which case
                // the groups still at work keep the chunk (no poll of the cancel word per iteration in the tail)
                if (!__any_sync(0xffffffffu, active)) break;
                if (!exhausted || (p.feed && __any_sync(0xffffffffu, !active && (pix & kReserved) != 0u))) break;
            }
        }
    }
}

}  // namespace mdz
