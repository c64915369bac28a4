// mpf_fast.cuh -- GMP-mpf-faithful arithmetic, register-resident 32-bit limbs.
//
// Same contract as mpf_sf.cuh (the reference's mpf_mul / mpf_mul_ui(.,2) /
// mpf_add / mpf_sub / mpf_cmp as GMP 6.3.0 behaves), built
// from the machinery of the MPFR path: IMAD.WIDE carry-chain high products with
// an exactness check and full-product fallback, carry-chain adds, and a
// shared-memory column for the limb-granular alignment (mpf exponents count
// 64-bit limbs, so there are no bit shifts at all except the x2).
//
// Layout: NW = 2*(P+1) 32-bit words, least significant first, top aligned at
// 64-bit granularity: (m[NW-1]:m[NW-2]) != 0 unless the value is zero.
// Anything off the common path (GMP's one-limb-gap "close" subtraction, a
// product whose truncation cannot be decided) converts to mpf_sf.cuh's
// reference implementation in an out-of-line function.
#pragma once
#include "limb_ops.cuh"
#include "mpf_sf.cuh"
#include "mpfr_sf.cuh"      // kScratchStride, MDZ_NOINLINE_DEV

namespace mdz {

template <int NW>
struct GF {
    uint32_t m[NW];
    int32_t  e;         // exponent in 64-bit limbs
    uint32_t s;
};

template <int NW> MDZ_HD bool gfz(const GF<NW>& a) { return (a.m[NW - 1] | a.m[NW - 2]) == 0; }
template <int NW> MDZ_HD void gf_zero(GF<NW>& a)
{
    MDZ_UNROLL
    for (int i = 0; i < NW; ++i) a.m[i] = 0;
    a.e = 0; a.s = 0;
}

template <int NW> MDZ_HD void gf_to_slow(const GF<NW>& a, Mpf<NW / 2>& r)
{
    MDZ_UNROLL
    for (int i = 0; i < NW / 2; ++i) r.l[i] = ((uint64_t)a.m[2 * i + 1] << 32) | a.m[2 * i];
    r.e = a.e; r.s = a.s;
}
template <int NW> MDZ_HD void gf_from_slow(const Mpf<NW / 2>& a, GF<NW>& r)
{
    MDZ_UNROLL
    for (int i = 0; i < NW / 2; ++i) { r.m[2 * i] = (uint32_t)a.l[i]; r.m[2 * i + 1] = (uint32_t)(a.l[i] >> 32); }
    r.e = a.e; r.s = a.s;
}

// out-of-line exact versions (mpf_sf.cuh), by value
template <int NW>
MDZ_NOINLINE_DEV GF<NW> gf_mul_slow(GF<NW> a, GF<NW> b)
{
    Mpf<NW / 2> x, y, z; GF<NW> r;
    gf_to_slow<NW>(a, x); gf_to_slow<NW>(b, y);
    gmul<NW / 2>(x, y, z);
    gf_from_slow<NW>(z, r);
    return r;
}
template <int NW>
MDZ_NOINLINE_DEV GF<NW> gf_addsub_slow(GF<NW> a, GF<NW> b, int subtract)
{
    Mpf<NW / 2> x, y, z; GF<NW> r;
    gf_to_slow<NW>(a, x); gf_to_slow<NW>(b, y);
    gadd<NW / 2>(x, y, z, subtract != 0);
    gf_from_slow<NW>(z, r);
    return r;
}

// ---------------------------------------------------------------------------
// High product with X extra low columns: t[k] = limb position (M-2-X)+k of the sum
// over i+j >= M-2-X of a[i]*b[j], k = 0 .. M+1+X (M-limb operands).  Same even/odd
// accumulator scheme as mul_hi (limb_ops.cuh).
// ---------------------------------------------------------------------------
template <int M, int X, bool SQUARE>
MDZ_HD void mul_hi_x(const uint32_t* a, const uint32_t* b, uint32_t (&t)[M + 2 + X])
{
    constexpr int BASE = M - 2 - X;         // lowest column kept (>= 0 required)
    constexpr int TN = M + 2 + X;
    uint32_t e[TN + 2], o[TN + 2];
    MDZ_UNROLL
    for (int i = 0; i < TN + 2; ++i) { e[i] = 0; o[i] = 0; }
    MDZ_UNROLL
    for (int i = 0; i < (SQUARE ? M - 1 : M); ++i) {
        const int jlo = SQUARE ? i + 1 : 0;
        const int jmin = (BASE - i) > jlo ? (BASE - i) : jlo;
        MDZ_UNROLL
        for (int c = 0; c < 2; ++c) {
            if (jmin + c >= M) continue;
            int last = 0;
            MDZ_UNROLL
            for (int j = jmin + c; j < M; j += 2) {
                const int q = i + j - BASE;
                if ((q & 1) == 0) {
                    if (j == jmin + c) mad_wide_cc(e[q], e[q + 1], a[j], b[i]);
                    else               madc_wide_cc(e[q], e[q + 1], a[j], b[i]);
                } else {
                    if (j == jmin + c) mad_wide_cc(o[q - 1], o[q], a[j], b[i]);
                    else               madc_wide_cc(o[q - 1], o[q], a[j], b[i]);
                }
                last = q;
            }
            const int cp = last + 2;
            if (cp < TN) {
                if ((cp & 1) == 0) e[cp] = addc(e[cp], 0u);
                else               o[cp - 1] = addc(o[cp - 1], 0u);
            }
        }
    }
    uint32_t x[TN];
    x[0] = e[0];
    x[1] = add_cc(e[1], o[0]);
    MDZ_UNROLL
    for (int i = 2; i < TN - 1; ++i) x[i] = addc_cc(e[i], o[i - 1]);
    x[TN - 1] = addc(e[TN - 1], o[TN - 2]);
    if (!SQUARE) {
        MDZ_UNROLL
        for (int i = 0; i < TN; ++i) t[i] = x[i];
        return;
    }
    // 2*cross + diagonal
    MDZ_UNROLL
    for (int i = TN - 1; i >= 1; --i) t[i] = fsl(x[i - 1], x[i], 1);
    t[0] = x[0] << 1;
    constexpr int i0 = (BASE + 1) / 2;      // smallest i with 2i >= BASE
    {
        constexpr int q0 = 2 * i0 - BASE;
        mad_wide_cc(t[q0], t[q0 + 1], a[i0], a[i0]);
    }
    MDZ_UNROLL
    for (int i = i0 + 1; i < M; ++i) {
        const int q = 2 * i - BASE;
        madc_wide_cc(t[q], t[q + 1], a[i], a[i]);
    }
}

// mpf_mul (also the squarings): top P limbs of each operand, exact top P+1 limbs of the
// product after dropping a zero top limb.
template <int NW, bool SQUARE>
MDZ_HD void gf_mul(const GF<NW>& u, const GF<NW>& v, GF<NW>& r)
{
    constexpr int M = NW - 2;               // 32-bit words of an operand's top P limbs
    constexpr int X = 4;                    // columns from M-6 up: position M-5 is a guard word below M-4
    uint32_t t[M + 2 + X];
    mul_hi_x<M, X, SQUARE>(u.m + 2, v.m + 2, t);
    // the omitted columns (i+j < M-6) sum to less than M units of position M-5 (t[1]; a
    // product's high word lands one column up), twice that for a squaring: the words
    // above t[1] are exact unless t[1] is within that distance of wrapping
    const bool unsure = t[1] >= 0xffffffffu - (uint32_t)(2 * M + 4);
    const bool adj = (t[M + 1 + X] | t[M + X]) == 0;                // top 64-bit limb zero: drop it
    if (unsure || gfz(u) || gfz(v)) { r = gf_mul_slow<NW>(u, SQUARE ? u : v); return; }
    // result words = positions [M-2-2adj .. 2M-1-2adj]  ->  t[X - 2adj + i], i = 0..NW-1
    MDZ_UNROLL
    for (int i = 0; i < NW; ++i) r.m[i] = adj ? t[X - 2 + i] : t[X + i];
    r.e = u.e + v.e - (adj ? 1 : 0);
    r.s = SQUARE ? 0u : (u.s ^ v.s);
}

// mpf_mul_ui(r, u, 2)
template <int NW>
MDZ_HD void gf_mul2(const GF<NW>& u, GF<NW>& r)
{
    constexpr int M = NW - 2;
    uint32_t sh[M];
    sh[0] = (u.m[2] << 1) | (u.m[1] >> 31);                        // carry in from the limb below
    MDZ_UNROLL
    for (int i = 1; i < M; ++i) sh[i] = fsl(u.m[1 + i], u.m[2 + i], 1);
    const bool c = (u.m[NW - 1] >> 31) != 0;
    const bool z = gfz(u);
    // carry: P+1 limbs [sh, 1]; else P limbs top aligned
    r.m[0] = c ? sh[0] : 0u;
    r.m[1] = c ? sh[1] : 0u;
    MDZ_UNROLL
    for (int i = 2; i < M; ++i) r.m[i] = c ? sh[i] : sh[i - 2];
    r.m[M] = c ? 1u : sh[M - 2];
    r.m[M + 1] = c ? 0u : sh[M - 1];
    r.e = z ? 0 : u.e + (c ? 1 : 0);
    r.s = z ? 0u : u.s;
}

// scratch column: [0,NW) zeros, [NW,2NW) data, [2NW,3NW) zeros  (kept that way)
template <int NW> struct GScratchWords { static constexpr int value = 3 * NW; };

// x >>= 32*q words (q > 0): through the scratch column
template <int NW>
MDZ_HD void gf_shift_down(uint32_t (&x)[NW], uint32_t q, uint32_t* scratch)
{
    if (q > (uint32_t)NW) q = NW;
    uint32_t* d = scratch + NW * kScratchStride;
    MDZ_UNROLL
    for (int k = 0; k < NW; ++k) d[k * kScratchStride] = x[k];
    const uint32_t* src = d + q * kScratchStride;
    MDZ_UNROLL
    for (int k = 0; k < NW; ++k) x[k] = src[k * kScratchStride];
}

// x <<= 32*q words (q > 0)
template <int NW>
MDZ_HD void gf_shift_up(uint32_t (&x)[NW], uint32_t q, uint32_t* scratch)
{
    if (q > (uint32_t)NW) q = NW;
    uint32_t* d = scratch + NW * kScratchStride;
    MDZ_UNROLL
    for (int k = 0; k < NW; ++k) d[k * kScratchStride] = x[k];
    const uint32_t* src = d - q * kScratchStride;
    MDZ_UNROLL
    for (int k = 0; k < NW; ++k) x[k] = src[k * kScratchStride];
}

// mpf_add / mpf_sub (a + b, or a - b when subtract)
template <int NW>
MDZ_HD void gf_addsub(const GF<NW>& a, const GF<NW>& b, GF<NW>& r, bool subtract, uint32_t* scratch)
{
    constexpr int M = NW - 2;
    constexpr int P = M / 2;
    const uint32_t sb = b.s ^ (subtract ? 1u : 0u);
    const bool az = gfz(a), bz = gfz(b);
    if (az || bz) {
        // a zero operand returns the other one (mpf_set keeps all P+1 limbs)
        MDZ_UNROLL
        for (int i = 0; i < NW; ++i) r.m[i] = az ? b.m[i] : a.m[i];
        r.e = az ? b.e : a.e;
        r.s = az ? (bz ? 0u : sb) : a.s;
        if (az && bz) r.e = 0;
        return;
    }
    const int32_t d = a.e - b.e;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    uint32_t xa[NW], xb[NW];
    MDZ_UNROLL
    for (int i = 0; i < NW; ++i) { xa[i] = a.m[i]; xb[i] = b.m[i]; }
    if (a.s == sb) {
        // ---- same sign: mpf_add, works to P limbs ----
        if (ad >= (uint32_t)P) {                  // the smaller one vanishes: result = top P limbs of the larger
            const bool a_big = d >= 0;
            r.m[0] = 0; r.m[1] = 0;
            MDZ_UNROLL
            for (int i = 2; i < NW; ++i) r.m[i] = a_big ? a.m[i] : b.m[i];
            r.e = a_big ? a.e : b.e; r.s = a.s;
            return;
        }
        if (ad != 0) { if (d > 0) gf_shift_down<NW>(xb, 2u * ad, scratch); else gf_shift_down<NW>(xa, 2u * ad, scratch); }
        // window = words [2, NW) of the larger-exponent frame
        uint32_t t[M];
        t[0] = add_cc(xa[2], xb[2]);
        MDZ_UNROLL
        for (int i = 1; i < M; ++i) t[i] = addc_cc(xa[2 + i], xb[2 + i]);
        const bool c = addc(0u, 0u) != 0;
        r.m[0] = c ? t[0] : 0u;
        r.m[1] = c ? t[1] : 0u;
        MDZ_UNROLL
        for (int i = 2; i < M; ++i) r.m[i] = c ? t[i] : t[i - 2];
        r.m[M] = c ? 1u : t[M - 2];
        r.m[M + 1] = c ? 0u : t[M - 1];
        r.e = (d >= 0 ? a.e : b.e) + (c ? 1 : 0);
        r.s = a.s;
        return;
    }
    // ---- opposite signs: mpf_sub, works to Q = P+1 limbs ----
    {
        // GMP's close path for a one-limb gap (u = 1:0:.., v = ff..ff:..) slides the window
        const GF<NW>& u = d >= 0 ? a : b;
        const GF<NW>& v = d >= 0 ? b : a;
        if (ad == 1 && u.m[NW - 1] == 0 && u.m[NW - 2] == 1 && (v.m[NW - 1] & v.m[NW - 2]) == 0xffffffffu &&
            (u.m[NW - 3] | u.m[NW - 4]) == 0) {
            r = gf_addsub_slow<NW>(a, b, subtract ? 1 : 0);
            return;
        }
    }
    if (ad >= (uint32_t)(P + 1)) {                // d >= Q: result is the larger one, untouched
        const bool a_big = d >= 0;
        MDZ_UNROLL
        for (int i = 0; i < NW; ++i) r.m[i] = a_big ? a.m[i] : b.m[i];
        r.e = a_big ? a.e : b.e; r.s = a_big ? a.s : sb;
        return;
    }
    if (ad != 0) { if (d > 0) gf_shift_down<NW>(xb, 2u * ad, scratch); else gf_shift_down<NW>(xa, 2u * ad, scratch); }
    // x = xa - xb over all NW words; a borrow (only possible when d == 0) means |b| > |a|
    uint32_t x[NW];
    x[0] = sub_cc(xa[0], xb[0]);
    MDZ_UNROLL
    for (int i = 1; i < NW; ++i) x[i] = subc_cc(xa[i], xb[i]);
    const bool neg = subc(0u, 0u) != 0;
    if (neg) {                                    // two's complement
        x[0] = sub_cc(0u, x[0]);
        MDZ_UNROLL
        for (int i = 1; i < NW; ++i) x[i] = subc_cc(0u, x[i]);
    }
    // the exponent frame is the larger exponent; the sign follows the larger magnitude
    const bool a_frame = d >= 0;
    const uint32_t sign = neg ? sb : a.s;
    // strip leading zero 64-bit limbs
    int k = 0;
    if ((x[NW - 1] | x[NW - 2]) == 0) {
        uint32_t any = 0;
        MDZ_UNROLL
        for (int i = 0; i < NW; ++i) any |= x[i];
        if (any == 0) { gf_zero(r); return; }
        k = 1;
#if !defined(MDZ_HOST_EMU)
#pragma unroll 1
#endif
        for (int j = NW / 2 - 2; j >= 0; --j) {   // count further zero limbs
            uint32_t w = 0;
            MDZ_UNROLL
            for (int i = 0; i < NW / 2; ++i) if (i == j) w = x[2 * i] | x[2 * i + 1];
            if (w != 0) break;
            ++k;
        }
        gf_shift_up<NW>(x, 2u * (uint32_t)k, scratch);
    }
    MDZ_UNROLL
    for (int i = 0; i < NW; ++i) r.m[i] = x[i];
    r.e = (a_frame ? a.e : b.e) - k;
    r.s = sign;
}

// mpf_cmp(a, 4) > 0
template <int NW>
MDZ_HD bool gf_gt4(const GF<NW>& a)
{
    if (gfz(a) || a.s) return false;
    if (a.e != 1) return a.e > 1;
    if (a.m[NW - 1] != 0 || a.m[NW - 2] != 4) return a.m[NW - 1] != 0 || a.m[NW - 2] > 4;
    uint32_t low = 0;
    MDZ_UNROLL
    for (int i = 0; i < NW - 2; ++i) low |= a.m[i];
    return low != 0;
}

// ---- one pixel (frac_*_gmp) ----------------------------------------------------------
template <int NW>
struct GFPixel {
    GF<NW> wre, wim, wre2, wim2;
    int32_t cre_e, cim_e;       // c's words live in shared memory
    uint32_t cre_s, cim_s;
    int iter;
};

template <int NW>
MDZ_HD void gf_pixel_init(GFPixel<NW>& st, const GF<NW>& x, const GF<NW>& y, const GF<NW>& cx, const GF<NW>& cy,
                          uint32_t* cre_m, uint32_t* cim_m)
{
    st.wre = x; st.wim = y;
    gf_mul<NW, true>(x, x, st.wre2);
    gf_mul<NW, true>(y, y, st.wim2);
    MDZ_UNROLL
    for (int k = 0; k < NW; ++k) { cre_m[k * kScratchStride] = cx.m[k]; cim_m[k * kScratchStride] = cy.m[k]; }
    st.cre_e = cx.e; st.cre_s = cx.s; st.cim_e = cy.e; st.cim_s = cy.s;
    st.iter = 0;
}

template <int NW>
MDZ_HD bool gf_pixel_step(GFPixel<NW>& st, const uint32_t* cre_m, const uint32_t* cim_m, uint32_t* scratch,
                          bool abs_im, int abs_re)
{
    ++st.iter;
    GF<NW> t1, t2, c;
    gf_mul<NW, false>(st.wre, st.wim, t1);
    if (abs_im) t1.s = 0;
    gf_mul2<NW>(t1, t2);
    MDZ_UNROLL
    for (int q = 0; q < NW; ++q) c.m[q] = cim_m[q * kScratchStride];
    c.e = st.cim_e; c.s = st.cim_s;
    gf_addsub<NW>(t2, c, st.wim, false, scratch);
    gf_addsub<NW>(st.wre2, st.wim2, t1, true, scratch);
    if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) t1.s = 0;
    MDZ_UNROLL
    for (int q = 0; q < NW; ++q) c.m[q] = cre_m[q * kScratchStride];
    c.e = st.cre_e; c.s = st.cre_s;
    gf_addsub<NW>(t1, c, st.wre, false, scratch);
    gf_mul<NW, true>(st.wim, st.wim, st.wim2);
    gf_mul<NW, true>(st.wre, st.wre, st.wre2);
    // mpf_cmp(wim2 + wre2, 4) > 0: both below 1 cannot reach 4, either at 2^64 or more surely does
    const int32_t emax = (gfz(st.wim2) ? -1000000 : st.wim2.e) > (gfz(st.wre2) ? -1000000 : st.wre2.e)
                       ? (gfz(st.wim2) ? -1000000 : st.wim2.e) : (gfz(st.wre2) ? -1000000 : st.wre2.e);
    if (emax >= 2) return true;
    if (emax <= 0) return false;
    // emax == 1: a square with limb exponent 1 is its top limb plus a fraction below 1, so the
    // integer parts bound the sum: ia + ib <= 2 -> below 4; >= 5 -> above 4 also after mpf_add's
    // truncation.  Only 3 and 4 need the sum itself.
    {
        const bool ai = !gfz(st.wim2) && st.wim2.e == 1, bi = !gfz(st.wre2) && st.wre2.e == 1;
        if ((ai && st.wim2.m[NW - 1] != 0) || (bi && st.wre2.m[NW - 1] != 0)) return true;     // >= 2^32
        const uint64_t ip = (uint64_t)(ai ? st.wim2.m[NW - 2] : 0u) + (bi ? st.wre2.m[NW - 2] : 0u);
        if (ip <= 2) return false;
        if (ip >= 5) return true;
    }
    gf_addsub<NW>(st.wim2, st.wre2, t1, false, scratch);
    return gf_gt4<NW>(t1);
}

}  // namespace mdz
