// escape_step.cuh -- one pixel's state and one iteration of the escape-time
// recurrence, MPFR-faithful.  Shared by the CUDA kernel (escape_kernel.cuh) and,
// compiled for the host with MDZ_HOST_EMU, by the CPU-side tests.
//
// pixel_step restates the loop body of the reference's frac_mandel_mpfr
// (src/frac_mandel.c:34-50) and its three variants: frac_burning_ship_mpfr
// takes |.| of the rounded product (src/frac_burning_ship.c:38-41),
// frac_generalized_celtic_mpfr of the rounded difference
// (src/frac_generalized_celtic.c:42-44), frac_variant_mpfr of the difference on
// odd iterations only (src/frac_variant.c:42-43).
#pragma once
#include "mpfr_sf.cuh"

namespace mdz {

enum { FRACTAL_MANDELBROT = 0, FRACTAL_BURNING_SHIP = 1, FRACTAL_GENERALIZED_CELTIC = 2, FRACTAL_VARIANT = 3 };
enum { FAMILY_MANDEL = 0, FAMILY_JULIA = 1 };

// how the fractal type enters the long double mode's fast iteration (ld64_step.cuh: ld64_masks)
struct Ld64Masks { uint32_t im_keep, re_and, re_xor; };

template <int N>
struct PixelState {
    Num<N> wre, wim, wre2, wim2;
    int32_t cre_e, cim_e;       // c's limbs live in a per-thread shared-memory column
    uint32_t cre_s, cim_s;
    int iter;
};

// set-up of one pixel (reference src/fractal.c:188-203): z0 = (x, y), squares
// rounded once, c = pixel (Mandelbrot family) or the Julia constant.
template <int N>
MDZ_HD void pixel_init(PixelState<N>& st, const Num<N>& x, const Num<N>& y,
                       const Num<N>& cx, const Num<N>& cy, const RoundCfg& rc,
                       uint32_t* cre_m, uint32_t* cim_m)
{
    st.wre = x; st.wim = y;
    fsqr<N>(x, st.wre2, rc);
    fsqr<N>(y, st.wim2, rc);
    MDZ_UNROLL
    for (int k = 0; k < N; ++k) { cre_m[k * kScratchStride] = cx.m[k]; cim_m[k * kScratchStride] = cy.m[k]; }
    st.cre_e = cx.e; st.cre_s = cx.s; st.cim_e = cy.e; st.cim_s = cy.s;
    st.iter = 0;
}

// RN(wim2 + wre2) > 4, the reference's bail-out test (mpfr_add + mpfr_greater_p,
// src/frac_mandel.c:46-48).  Both squares below 2 cannot exceed 4 even after rounding,
// either at 8 or more certainly does; in between the top limbs decide (escape_precheck)
// unless the sum is within 2^-27 of 4, and only then is it formed.
template <int N>
MDZ_HD bool escaped(const Num<N>& wim2, const Num<N>& wre2, const RoundCfg& rc, uint32_t* scr)
{
    const int32_t emax = wim2.e > wre2.e ? wim2.e : wre2.e;
    bool esc = emax >= 4;
    if (!esc && emax >= 2) {
        const int pre = escape_precheck<N>(wim2, wre2);
        esc = pre > 0;
        if (pre == 0) {
            MDZ_COUNT(CNT_ESC_ADD);
            Num<N> t;
            fadd<N, MODE_ADD_POS>(wim2, wre2, t, rc, scr);
            esc = greater_than_4<N>(t);
        }
    }
    return esc;
}

// one iteration; returns true when RN(wim2 + wre2) > 4 (the pixel escaped at st.iter)
template <int N>
MDZ_HD bool pixel_step(PixelState<N>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                       uint32_t* scr, const RoundCfg& rc, bool abs_im, int abs_re)
{
    ++st.iter;
    MDZ_COUNT(CNT_ITER);
    Num<N> t, c;
    // An orbit on the real axis stays there.  With wim == 0 and c_im == 0 the product wre*wim, its
    // double and the sum with c_im are exact zeros, wim2 stays 0, and wre2 - 0 is wre2 with nothing to
    // round (it is >= 0, so the |.| of the celtic variants changes nothing either): the nine calls of
    // src/frac_mandel.c:36-48 produce wre = RN(wre2 + c_re), wre2 = RN(wre^2) and test wre2 alone.
    // These are the pixels of the row y = 0, on which every product has a zero operand -- outside the
    // fast iterations' domain, so each of them would otherwise pay the full general step, per lane,
    // for up to `depth` iterations, and set the pace of its warp.
    if (is_zero(st.wim) && cim_m[(N - 1) * kScratchStride] == 0u) {
        MDZ_UNROLL
        for (int q = 0; q < N; ++q) c.m[q] = cre_m[q * kScratchStride];
        c.e = st.cre_e; c.s = st.cre_s;
        t = st.wre2; t.s = 0;
        fadd<N, MODE_GENERIC>(t, c, st.wre, rc, scr);
        fsqr<N>(st.wre, st.wre2, rc);
        return escaped<N>(st.wim2, st.wre2, rc, scr);
    }
    // wim = 2*wre*wim + c_im       (|.| on the product for burning ship)
    fmul<N>(st.wre, st.wim, t, rc);
    if (t.m[N - 1] != 0) t.e += 1;
    if (abs_im) t.s = 0;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c.m[q] = cim_m[q * kScratchStride];
    c.e = st.cim_e; c.s = st.cim_s;
    fadd<N, MODE_GENERIC>(t, c, st.wim, rc, scr);
    // wre = wre2 - wim2 + c_re     (|.| on the difference for celtic / odd steps of the hybrid)
    fadd<N, MODE_SUB_POS>(st.wre2, st.wim2, t, rc, scr);
    if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) t.s = 0;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c.m[q] = cre_m[q * kScratchStride];
    c.e = st.cre_e; c.s = st.cre_s;
    fadd<N, MODE_GENERIC>(t, c, st.wre, rc, scr);
    fsqr<N>(st.wim, st.wim2, rc);
    fsqr<N>(st.wre, st.wre2, rc);
    return escaped<N>(st.wim2, st.wre2, rc, scr);
}


// Speculative iteration: the same arithmetic built from the branch-free
// *_spec operations (mpfr_sf.cuh), i.e. one basic block in which ptxas overlaps
// the IMAD.WIDE chains of the products with the shift/add chains of the sums.
// Any condition the branch-free code does not cover raises `rare`; the caller
// then restores the previous state and runs pixel_step (below in
// pixel_step_auto).  The escape sum is not part of the speculation: when
// max(e) is 2 or 3 it is computed afterwards with the general fadd.
template <int N>
MDZ_HD bool pixel_step_spec(PixelState<N>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                            uint32_t* scr, const RoundCfg& rc, bool abs_im, int abs_re, uint32_t& rare)
{
    ++st.iter;
    Num<N> t, u, c, c2;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c.m[q] = cim_m[q * kScratchStride];
    c.e = st.cim_e; c.s = st.cim_s;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c2.m[q] = cre_m[q * kScratchStride];
    c2.e = st.cre_e; c2.s = st.cre_s;
    fmul_spec<N>(st.wre, st.wim, t, rc, rare);
    fadd_spec<N, MODE_SUB_POS>(st.wre2, st.wim2, u, rc, rare);
    if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) u.s = 0;
    fadd_spec<N, MODE_GENERIC>(u, c2, st.wre, rc, rare);
    if (t.m[N - 1] != 0) t.e += 1;
    if (abs_im) t.s = 0;
    fadd_spec<N, MODE_GENERIC>(t, c, st.wim, rc, rare);
    fsqr_spec<N>(st.wre, st.wre2, rc, rare);
    fsqr_spec<N>(st.wim, st.wim2, rc, rare);
    return rare == 0 ? escaped<N>(st.wim2, st.wre2, rc, scr) : false;
}

// The same step with fadd_spec_wide (gaps up to 126 bits): level 2 of pixel_step_auto.
template <int N>
MDZ_HD bool pixel_step_spec_wide(PixelState<N>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                            uint32_t* scr, const RoundCfg& rc, bool abs_im, int abs_re, uint32_t& rare)
{
    ++st.iter;
    Num<N> t, u, c, c2;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c.m[q] = cim_m[q * kScratchStride];
    c.e = st.cim_e; c.s = st.cim_s;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c2.m[q] = cre_m[q * kScratchStride];
    c2.e = st.cre_e; c2.s = st.cre_s;
    fmul_spec<N>(st.wre, st.wim, t, rc, rare);
    fadd_spec_wide<N, MODE_SUB_POS>(st.wre2, st.wim2, u, rc, rare);
    if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) u.s = 0;
    fadd_spec_wide<N, MODE_GENERIC>(u, c2, st.wre, rc, rare);
    if (t.m[N - 1] != 0) t.e += 1;
    if (abs_im) t.s = 0;
    fadd_spec_wide<N, MODE_GENERIC>(t, c, st.wim, rc, rare);
    fsqr_spec<N>(st.wre, st.wre2, rc, rare);
    fsqr_spec<N>(st.wim, st.wim2, rc, rare);
    return rare == 0 ? escaped<N>(st.wim2, st.wre2, rc, scr) : false;
}


// Hybrid iteration (11 limbs and up): the speculative operations with their fall-backs *inside* the step.
// pixel_step_auto redoes a whole iteration with the general step when any speculative operation declines;
// from 11 limbs up those two steps are ~20 KB of unrolled code each, the SM's instruction cache holds 32 KB,
// and a warp that alternates between them -- or merely runs next to one that does -- waits for instructions
// (ncu: 7 stall cycles per issue; B200, 512 bits, next to a minibrot, where every orbit returns to ~0 once per
// period and two iterations in 707 cancel 200 bits or add across a 400-bit gap: 6.8 G it/s with the adaptive
// levels, 13.1 with the general step alone, against 14.5 on views where nothing ever declines).  Here only the
// part that declined is redone, in place: a product by the exact out-of-line product (probability ~N 2^-29), the
// three additions by the general fadd (~10 KB of code).  Nothing is overwritten before it is known to be good,
// so there is no checkpoint, no second copy of the products, no per-warp level to adapt -- and one lane's rare
// addition costs its warp three general additions instead of an iteration and two cache refills.
// rare_seen counts the iterations in which the speculative additions had to be redone: a warp whose orbits make
// them decline most of the time (gallery/deep_embedded_julia.mdz: wre^2 - wim^2 has a 31..40-bit exponent gap in
// every second iteration) is better off with the general step from the start; escape_kernel.cuh "adapt" decides.
template <int N>
MDZ_HD bool pixel_step_hybrid(PixelState<N>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                              uint32_t* scr, const RoundCfg& rc, bool abs_im, int abs_re, uint32_t& rare_seen)
{
    // an orbit on the real axis stays there: the general step's short form (pixel_step)
    if (is_zero(st.wim) && cim_m[(N - 1) * kScratchStride] == 0u)
        return pixel_step<N>(st, cre_m, cim_m, scr, rc, abs_im, abs_re);
    ++st.iter;
    MDZ_COUNT(CNT_ITER);
    Num<N> t, u, c;
    const bool drop_re = abs_re == 1 || (abs_re == 2 && (st.iter & 1));
    // the product first, settled before anything else: wre and wim are dead after it, which is what keeps the
    // 16-limb kernel's additions in registers (with the check further down ptxas spilled 31 registers per iteration)
    uint32_t rm = 0, ra = 0;
    fmul_spec<N>(st.wre, st.wim, t, rc, rm);
    if (rm) { MDZ_COUNT(CNT_MUL_BAIL); t = fmul_general<N>(st.wre, st.wim, rc); }
    if (t.m[N - 1] != 0) t.e += 1;
    if (abs_im) t.s = 0;
    // the additions write straight into wre / wim: what the fall-back needs (wre2, wim2, t, c) is still intact
    // (forming this difference before the product, to overlap with it, measured 2.7 % slower at 16 limbs: more spills)
    fadd_spec<N, MODE_SUB_POS>(st.wre2, st.wim2, u, rc, ra);
    if (drop_re) u.s = 0;
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c.m[q] = cre_m[q * kScratchStride];
    c.e = st.cre_e; c.s = st.cre_s;
    fadd_spec<N, MODE_GENERIC>(u, c, st.wre, rc, ra);
    MDZ_UNROLL
    for (int q = 0; q < N; ++q) c.m[q] = cim_m[q * kScratchStride];
    c.e = st.cim_e; c.s = st.cim_s;
    fadd_spec<N, MODE_GENERIC>(t, c, st.wim, rc, ra);
    if (ra) {
        rare_seen += 1u;
        MDZ_COUNT(CNT_SPEC_FALLBACK);
        fadd<N, MODE_SUB_POS>(st.wre2, st.wim2, u, rc, scr);
        if (drop_re) u.s = 0;
        MDZ_UNROLL
        for (int q = 0; q < N; ++q) c.m[q] = cre_m[q * kScratchStride];
        c.e = st.cre_e; c.s = st.cre_s;
        fadd<N, MODE_GENERIC>(u, c, st.wre, rc, scr);
        MDZ_UNROLL
        for (int q = 0; q < N; ++q) c.m[q] = cim_m[q * kScratchStride];
        c.e = st.cim_e; c.s = st.cim_s;
        fadd<N, MODE_GENERIC>(t, c, st.wim, rc, scr);
    }
    uint32_t r1 = 0, r2 = 0;
    fsqr_spec<N>(st.wre, st.wre2, rc, r1);
    fsqr_spec<N>(st.wim, st.wim2, rc, r2);
    if (r1) { MDZ_COUNT(CNT_MUL_BAIL); Num<N> f = fmul_general<N>(st.wre, st.wre, rc); f.s = 0; st.wre2 = f; }
    if (r2) { MDZ_COUNT(CNT_MUL_BAIL); Num<N> f = fmul_general<N>(st.wim, st.wim, rc); f.s = 0; st.wim2 = f; }
    return escaped<N>(st.wim2, st.wre2, rc, scr);
}


// Checkpoint of the loop-carried state for the speculative step.  For small limb
// counts it simply stays in registers; for large ones (4N extra registers would cost a
// resident block) it goes to a per-thread shared-memory column of 4N+5 words.
template <int N> struct CkptWords { static constexpr int value = 4 * N + 5; };

template <int N>
MDZ_HD void ckpt_save(const PixelState<N>& st, uint32_t* ck)
{
    MDZ_UNROLL
    for (int k = 0; k < N; ++k) {
        ck[(0 * N + k) * kScratchStride] = st.wre.m[k];
        ck[(1 * N + k) * kScratchStride] = st.wim.m[k];
        ck[(2 * N + k) * kScratchStride] = st.wre2.m[k];
        ck[(3 * N + k) * kScratchStride] = st.wim2.m[k];
    }
    ck[(4 * N + 0) * kScratchStride] = (uint32_t)st.wre.e;
    ck[(4 * N + 1) * kScratchStride] = (uint32_t)st.wim.e;
    ck[(4 * N + 2) * kScratchStride] = (uint32_t)st.wre2.e;
    ck[(4 * N + 3) * kScratchStride] = (uint32_t)st.wim2.e;
    ck[(4 * N + 4) * kScratchStride] = st.wre.s | (st.wim.s << 1);      // squares are non-negative
}

template <int N>
MDZ_HD void ckpt_load(PixelState<N>& st, const uint32_t* ck)
{
    MDZ_UNROLL
    for (int k = 0; k < N; ++k) {
        st.wre.m[k]  = ck[(0 * N + k) * kScratchStride];
        st.wim.m[k]  = ck[(1 * N + k) * kScratchStride];
        st.wre2.m[k] = ck[(2 * N + k) * kScratchStride];
        st.wim2.m[k] = ck[(3 * N + k) * kScratchStride];
    }
    st.wre.e  = (int32_t)ck[(4 * N + 0) * kScratchStride];
    st.wim.e  = (int32_t)ck[(4 * N + 1) * kScratchStride];
    st.wre2.e = (int32_t)ck[(4 * N + 2) * kScratchStride];
    st.wim2.e = (int32_t)ck[(4 * N + 3) * kScratchStride];
    const uint32_t sg = ck[(4 * N + 4) * kScratchStride];
    st.wre.s = sg & 1u; st.wim.s = (sg >> 1) & 1u; st.wre2.s = 0; st.wim2.s = 0;
}

// Speculate when the warp's recent history says it pays (use_spec is
// warp-uniform and maintained by the kernel), fall back per lane otherwise.
// SMEM_CKPT: the previous state is parked in shared memory instead of registers.
// (Measured alternative, rejected: running the general step out of line from the
// checkpoint keeps it out of the kernel's register allocation but costs a state
// round trip per call -- 12% slower at 512 bits on ordinary views, 25% on views
// that need the general step every iteration.)
template <int N, bool SMEM_CKPT>
MDZ_HD bool pixel_step_auto(PixelState<N>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                            uint32_t* scr, uint32_t* ckpt, const RoundCfg& rc, bool abs_im, int abs_re,
                            int level, uint32_t& rare_seen)
{
    // level: 0 general step only, 1 speculative, 2 speculative with the wide-gap additions.  (The multi-limb
    // kernels no longer choose level 2 -- escape_kernel.cuh "adapt" -- but it stays compiled in: without it the
    // register allocator spills more in the 3-, 4- and 10-limb kernels, 320 bits 19.5 -> 18.1 G it/s on the B200.)
    if (level != 0) {
        PixelState<N> keep;
        if (SMEM_CKPT) ckpt_save<N>(st, ckpt); else keep = st;
        uint32_t rare = 0;
        const bool esc = level == 2 ? pixel_step_spec_wide<N>(st, cre_m, cim_m, scr, rc, abs_im, abs_re, rare)
                                    : pixel_step_spec<N>(st, cre_m, cim_m, scr, rc, abs_im, abs_re, rare);
        if (rare == 0) return esc;
        MDZ_COUNT(CNT_SPEC_FALLBACK);
        rare_seen += 1;
        if (SMEM_CKPT) { ckpt_load<N>(st, ckpt); st.iter -= 1; } else st = keep;
    }
    return pixel_step<N>(st, cre_m, cim_m, scr, rc, abs_im, abs_re);
}

}  // namespace mdz
