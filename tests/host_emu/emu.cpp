// tests/host_emu/emu.cpp -- TEST BUILD ONLY.
// Compiles the device arithmetic headers (mdz_b200/csrc/limb_ops.cuh,
// mpfr_sf.cuh) for the host with MDZ_HOST_EMU, where every PTX carry-chain
// primitive is replaced by a C++ emulation with an explicit carry flag.  This
// lets the limb algorithms be differential-tested against libmpfr on a box
// without a GPU.  It is not a product path: libmdzcuda never links this file
// and has no CPU fallback.
#define MDZ_HOST_EMU 1
#include "../../mdz_b200/csrc/escape_step.cuh"
#include "../../mdz_b200/csrc/ld64_step.cuh"
#include "../../mdz_b200/csrc/mp_convert.h"

using namespace mdz;

template <int N>
static int binop(int op, long prec,
                 const uint64_t* al, int as, long ae,
                 const uint64_t* bl, int bs, long be,
                 uint64_t* rl, int* rs, long* re)
{
    RoundCfg rc = make_round_cfg(N, (int)prec);
    Num<N> a, b, r;
    uint32_t scratch[ScratchWords<N>::value] = {0};
    if (as == 0) set_zero(a); else { sig64_to_sig32(al, prec, a.m, N); a.e = (int32_t)ae; a.s = as < 0; }
    if (bs == 0) set_zero(b); else { sig64_to_sig32(bl, prec, b.m, N); b.e = (int32_t)be; b.s = bs < 0; }
    switch (op) {
    case 0: fmul<N>(a, b, r, rc); break;
    case 1: fsqr<N>(a, r, rc); break;
    case 2: fadd<N, MODE_GENERIC>(a, b, r, rc, scratch); break;
    case 3: b.s ^= 1u; fadd<N, MODE_GENERIC>(a, b, r, rc, scratch); break;
    case 4: fadd<N, MODE_SUB_POS>(a, b, r, rc, scratch); break;
    case 5: fadd<N, MODE_ADD_POS>(a, b, r, rc, scratch); break;
    case 6: *rs = greater_than_4<N>(a) ? 1 : 0; return 1;
    case 11: { uint32_t scr2[ScratchWords<N>::value] = {0}; *rs = escaped<N>(a, b, rc, scr2) ? 1 : 0; return 1; }   // RN(a + b) > 4, a, b >= 0
    case 7: case 8: case 9: case 10: {      // fadd_spec_wide: add, sub, a-b with a,b >= 0, a+b with a,b >= 0; returns 2 when it declined
        uint32_t rare = 0;
        if (op == 8) b.s ^= 1u;
        if (op == 7 || op == 8) fadd_spec_wide<N, MODE_GENERIC>(a, b, r, rc, rare);
        else if (op == 9) fadd_spec_wide<N, MODE_SUB_POS>(a, b, r, rc, rare);
        else fadd_spec_wide<N, MODE_ADD_POS>(a, b, r, rc, rare);
        if (rare) return 2;
        break;
    }
    default: return 0;
    }
    if (is_zero(r)) { *rs = 0; *re = 0; for (int i = 0; i < limbs64_for_prec(prec); ++i) rl[i] = 0; return 1; }
    sig32_to_sig64(r.m, N, prec, rl);
    *rs = r.s ? -1 : 1;
    *re = r.e;
    return 1;
}

#define CASE(n) case n: return binop<n>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
extern "C" int emu_binop(int op, long prec,
                         const uint64_t* al, int as, long ae,
                         const uint64_t* bl, int bs, long be,
                         uint64_t* rl, int* rs, long* re)
{
    switch (limbs32_for_prec(prec)) {
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
    CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21)
    CASE(22) CASE(23) CASE(24) CASE(25) CASE(26) CASE(27) CASE(28) CASE(29) CASE(30) CASE(31) CASE(32)
    default: return 0;
    }
}


// ---- p = 64 fast operations (ld64_step.cuh): result + "rare" flag --------------------
// op 0: mul (sign = product of signs), 2: add, 3: sub.  Returns 1 when the operation
// declared itself outside its covered domain (the kernel then uses the general code).
extern "C" int emu_ld64_op(int op, uint64_t am, int as, long ae, uint64_t bm, int bs, long be,
                           uint64_t* rm, int* rs, long* re)
{
    Num<2> a, b, r;
    if (as == 0) set_zero(a); else { a.m[0] = (uint32_t)am; a.m[1] = (uint32_t)(am >> 32); a.e = (int32_t)ae; a.s = as < 0; }
    if (bs == 0) set_zero(b); else { b.m[0] = (uint32_t)bm; b.m[1] = (uint32_t)(bm >> 32); b.e = (int32_t)be; b.s = bs < 0; }
    bool rare = false;
    if (op == 0) { mul64_spec(a, b, r, rare); r.s = a.s ^ b.s; }
    else if (op == 8) { mul64_spec<true>(a, b, r, rare); r.s = a.s ^ b.s; }       // level 2 product
    else if (op == 6 || op == 7) {          // a - b for a, b >= 0 through the dedicated difference: plain / level 2
        Ld64Flags f; ld64_flags_init(f, false);
        a.s = 0; b.s = 0;
        if (op == 6) add64_core<false, true>(a, b, r, f); else add64_core<true, true>(a, b, r, f);
        rare = ld64_flags_rare(f);
    }
    else if (op >= 4) { if (op == 5) b.s ^= 1u; add64_spec<true>(a, b, r, rare); }     // 4 / 5: add / sub, level-2 variant
    else { if (op == 3) b.s ^= 1u; add64_spec(a, b, r, rare); }
    *rm = ((uint64_t)r.m[1] << 32) | r.m[0]; *rs = r.s ? -1 : 1; *re = r.e;
    return rare ? 1 : 0;
}

// ---- whole-pixel emulation: the kernel's pixel_init / pixel_step on the host ----
template <int N>
static long pixel(long prec, int fractal, long depth, int spec,
                  const uint64_t* const* l, const int* sg, const long* ex)
{
    RoundCfg rc = make_round_cfg(N, (int)prec);
    Num<N> v[4];        // x, y, cx, cy
    for (int k = 0; k < 4; ++k) {
        if (sg[k] == 0) set_zero(v[k]);
        else { sig64_to_sig32(l[k], prec, v[k].m, N); v[k].e = (int32_t)ex[k]; v[k].s = sg[k] < 0; }
    }
    uint32_t cre[N], cim[N], scr[ScratchWords<N>::value] = {0};
    PixelState<N> st;
    pixel_init<N>(st, v[0], v[1], v[2], v[3], rc, cre, cim);
    const bool abs_im = fractal == FRACTAL_BURNING_SHIP;
    const int abs_re = fractal == FRACTAL_GENERALIZED_CELTIC ? 1 : fractal == FRACTAL_VARIANT ? 2 : 0;
    uint32_t rare_seen = 0;
    uint32_t ck[CkptWords<N>::value];
    if (spec == 5) {            // the hybrid iteration (fall-backs inside the step), the kernels' choice from 11 limbs up
        while (st.iter < depth)
            if (pixel_step_hybrid<N>(st, cre, cim, scr, rc, abs_im, abs_re, rare_seen)) return st.iter;
        return 0;
    }
    while (st.iter < depth)
        if ((spec == 2 || spec == 4) ? pixel_step_auto<N, true>(st, cre, cim, scr, ck, rc, abs_im, abs_re, spec == 4 ? 2 : 1, rare_seen)
                                     : pixel_step_auto<N, false>(st, cre, cim, scr, ck, rc, abs_im, abs_re, spec == 3 ? 2 : spec, rare_seen)) return st.iter;
    return 0;
}

#define PCASE(n) case n: return pixel<n>(prec, fractal, depth, spec, l, sg, ex);
extern "C" long emu_pixel(long prec, int fractal, long depth, int spec,
                          const uint64_t* xl, int xs, long xe, const uint64_t* yl, int ys, long ye,
                          const uint64_t* cxl, int cxs, long cxe, const uint64_t* cyl, int cys, long cye)
{
    const uint64_t* l[4] = {xl, yl, cxl, cyl};
    const int sg[4] = {xs, ys, cxs, cys};
    const long ex[4] = {xe, ye, cxe, cye};
    switch (limbs32_for_prec(prec)) {
    PCASE(2) PCASE(3) PCASE(4) PCASE(5) PCASE(6) PCASE(7) PCASE(8) PCASE(9) PCASE(10)
    PCASE(11) PCASE(12) PCASE(13) PCASE(14) PCASE(15) PCASE(16) PCASE(17) PCASE(18) PCASE(19) PCASE(20) PCASE(21)
    PCASE(22) PCASE(23) PCASE(24) PCASE(25) PCASE(26) PCASE(27) PCASE(28) PCASE(29) PCASE(30) PCASE(31) PCASE(32)
    default: return -1;
    }
}

extern "C" void emu_counts(unsigned long long* out, int reset)
{
    for (int i = 0; i < CNT_N; ++i) { out[i] = g_counts[i]; if (reset) g_counts[i] = 0; }
}

// ---- GMP mpf mode ---------------------------------------------------------------
#include "../../mdz_b200/csrc/mpf_sf.cuh"

template <int NL>
static int gmp_op(int op, const uint64_t* al, long ae, int as, const uint64_t* bl, long be, int bs,
                  uint64_t* rl, long* re, int* rs)
{
    Mpf<NL> a, b, r;
    for (int i = 0; i < NL; ++i) { a.l[i] = al[i]; b.l[i] = bl[i]; }
    a.e = (int32_t)ae; a.s = as < 0; b.e = (int32_t)be; b.s = bs < 0;
    if (as == 0) gset_zero(a);
    if (bs == 0) gset_zero(b);
    switch (op) {
    case 0: gmul<NL>(a, b, r); break;
    case 1: gmul2<NL>(a, r); break;
    case 2: gadd<NL>(a, b, r, false); break;
    case 3: gadd<NL>(a, b, r, true); break;
    case 4: *rs = ggt4<NL>(a) ? 1 : 0; return 1;
    default: return 0;
    }
    for (int i = 0; i < NL; ++i) rl[i] = r.l[i];
    *re = r.e; *rs = gz(r) ? 0 : (r.s ? -1 : 1);
    return 1;
}

#define GCASE(n) case n: return gmp_op<n>(op, al, ae, as, bl, be, bs, rl, re, rs);
extern "C" int emu_gmp_op(int op, int nl, const uint64_t* al, long ae, int as,
                          const uint64_t* bl, long be, int bs, uint64_t* rl, long* re, int* rs)
{
    switch (nl) {
    GCASE(3) GCASE(4) GCASE(5) GCASE(6) GCASE(7) GCASE(8) GCASE(9) GCASE(10) GCASE(11) GCASE(12)
    GCASE(13) GCASE(14) GCASE(15) GCASE(16) GCASE(17) GCASE(18)
    default: return 0;
    }
}

template <int NL>
static long gmp_pixel(int fractal, long depth, const uint64_t* const* l, const long* ex, const int* sg)
{
    Mpf<NL> v[4];
    for (int k = 0; k < 4; ++k) {
        for (int i = 0; i < NL; ++i) v[k].l[i] = l[k][i];
        v[k].e = (int32_t)ex[k]; v[k].s = sg[k] < 0;
        if (sg[k] == 0) gset_zero(v[k]);
    }
    GmpPixel<NL> st;
    gmp_pixel_init<NL>(st, v[0], v[1], v[2], v[3]);
    const bool abs_im = fractal == FRACTAL_BURNING_SHIP;
    const int abs_re = fractal == FRACTAL_GENERALIZED_CELTIC ? 1 : fractal == FRACTAL_VARIANT ? 2 : 0;
    while (st.iter < depth)
        if (gmp_pixel_step<NL>(st, abs_im, abs_re)) return st.iter;
    return 0;
}

#define GPCASE(n) case n: return gmp_pixel<n>(fractal, depth, l, ex, sg);
extern "C" long emu_gmp_pixel(int nl, int fractal, long depth,
                              const uint64_t* xl, long xe, int xs, const uint64_t* yl, long ye, int ys,
                              const uint64_t* cxl, long cxe, int cxs, const uint64_t* cyl, long cye, int cys)
{
    const uint64_t* l[4] = {xl, yl, cxl, cyl};
    const long ex[4] = {xe, ye, cxe, cye};
    const int sg[4] = {xs, ys, cxs, cys};
    switch (nl) {
    GPCASE(3) GPCASE(4) GPCASE(5) GPCASE(6) GPCASE(7) GPCASE(8) GPCASE(9) GPCASE(10) GPCASE(11) GPCASE(12)
    GPCASE(13) GPCASE(14) GPCASE(15) GPCASE(16) GPCASE(17) GPCASE(18)
    default: return -1;
    }
}

// ---- GMP mpf mode, fast register implementation (mpf_fast.cuh) -------------------------
#include "../../mdz_b200/csrc/mpf_fast.cuh"

template <int NL>
static int gmpf_op(int op, const uint64_t* al, long ae, int as, const uint64_t* bl, long be, int bs,
                   uint64_t* rl, long* re, int* rs)
{
    constexpr int NW = 2 * NL;
    Mpf<NL> a, b, z;
    for (int i = 0; i < NL; ++i) { a.l[i] = al[i]; b.l[i] = bl[i]; }
    a.e = (int32_t)ae; a.s = as < 0; b.e = (int32_t)be; b.s = bs < 0;
    if (as == 0) gset_zero(a);
    if (bs == 0) gset_zero(b);
    GF<NW> x, y, r;
    gf_from_slow<NW>(a, x); gf_from_slow<NW>(b, y);
    uint32_t scratch[GScratchWords<NW>::value] = {0};
    switch (op) {
    case 0: gf_mul<NW, false>(x, y, r); break;
    case 1: gf_mul2<NW>(x, r); break;
    case 2: gf_addsub<NW>(x, y, r, false, scratch); break;
    case 3: gf_addsub<NW>(x, y, r, true, scratch); break;
    case 4: *rs = gf_gt4<NW>(x) ? 1 : 0; return 1;
    case 5: gf_mul<NW, true>(x, x, r); break;
    default: return 0;
    }
    gf_to_slow<NW>(r, z);
    for (int i = 0; i < NL; ++i) rl[i] = z.l[i];
    *re = z.e; *rs = gz(z) ? 0 : (z.s ? -1 : 1);
    return 1;
}

#define GFCASE(n) case n: return gmpf_op<n>(op, al, ae, as, bl, be, bs, rl, re, rs);
extern "C" int emu_gmpf_op(int op, int nl, const uint64_t* al, long ae, int as,
                           const uint64_t* bl, long be, int bs, uint64_t* rl, long* re, int* rs)
{
    switch (nl) {
    GFCASE(4) GFCASE(5) GFCASE(6) GFCASE(7) GFCASE(8) GFCASE(9) GFCASE(10) GFCASE(11) GFCASE(12)
    GFCASE(13) GFCASE(14) GFCASE(15) GFCASE(16) GFCASE(17) GFCASE(18)
    default: return 0;
    }
}

template <int NL>
static long gmpf_pixel(int fractal, long depth, const uint64_t* const* l, const long* ex, const int* sg)
{
    constexpr int NW = 2 * NL;
    GF<NW> v[4];
    for (int k = 0; k < 4; ++k) {
        Mpf<NL> t;
        for (int i = 0; i < NL; ++i) t.l[i] = l[k][i];
        t.e = (int32_t)ex[k]; t.s = sg[k] < 0;
        if (sg[k] == 0) gset_zero(t);
        gf_from_slow<NW>(t, v[k]);
    }
    uint32_t cre[NW], cim[NW], scratch[GScratchWords<NW>::value] = {0};
    GFPixel<NW> st;
    gf_pixel_init<NW>(st, v[0], v[1], v[2], v[3], cre, cim);
    const bool abs_im = fractal == FRACTAL_BURNING_SHIP;
    const int abs_re = fractal == FRACTAL_GENERALIZED_CELTIC ? 1 : fractal == FRACTAL_VARIANT ? 2 : 0;
    while (st.iter < depth)
        if (gf_pixel_step<NW>(st, cre, cim, scratch, abs_im, abs_re)) return st.iter;
    return 0;
}

#define GFPCASE(n) case n: return gmpf_pixel<n>(fractal, depth, l, ex, sg);
extern "C" long emu_gmpf_pixel(int nl, int fractal, long depth,
                               const uint64_t* xl, long xe, int xs, const uint64_t* yl, long ye, int ys,
                               const uint64_t* cxl, long cxe, int cxs, const uint64_t* cyl, long cye, int cys)
{
    const uint64_t* l[4] = {xl, yl, cxl, cyl};
    const long ex[4] = {xe, ye, cxe, cye};
    const int sg[4] = {xs, ys, cxs, cys};
    switch (nl) {
    GFPCASE(4) GFPCASE(5) GFPCASE(6) GFPCASE(7) GFPCASE(8) GFPCASE(9) GFPCASE(10) GFPCASE(11) GFPCASE(12)
    GFPCASE(13) GFPCASE(14) GFPCASE(15) GFPCASE(16) GFPCASE(17) GFPCASE(18)
    default: return -1;
    }
}
