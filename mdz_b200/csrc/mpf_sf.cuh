// mpf_sf.cuh -- GMP-mpf-faithful arithmetic on fixed arrays of 64-bit limbs.
//
// The reference's GMP mode (src/fractal.c:260-397, frac_*_gmp in src/frac_*.c)
// calls mpf_mul, mpf_mul_ui(.,2), mpf_add, mpf_sub, mpf_abs, mpf_cmp.  GMP's
// mpf numbers are limb-granular floating point: value = +- (limbs) * 2^(64*(exp-size)),
// top limb non-zero but not bit-normalised, each operation truncating by its
// own rule.  Those rules, as observed on GMP 6.3.0 (libgmp.so.10) and written
// out in SURVEY Appendix E, are restated here.  GMP's source is not available
// on this box and nothing below derives from it.
//
// Representation: NL = P+1 limbs, top aligned (l[NL-1] != 0 unless the value is
// zero), zero padded below -- value-equivalent to GMP's variable-size storage
// because every truncation is measured from the top limb (Appendix E,
// "Consequences for the kernel").  P = floor((max(53,p)+127)/64) is mpf_init2's
// precision in limbs.  Exponent in limbs (int32), sign bit, zero = all limbs 0.
//
// This first version favours being obviously equal to the spec over speed:
// alignment and normalisation use small index loops (local memory when the
// shift is not a compile-time constant).
#pragma once
#include <stdint.h>
#include "limb_ops.cuh"

namespace mdz {

MDZ_HD void mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo)
{
#if defined(MDZ_HOST_EMU)
    unsigned __int128 p = (unsigned __int128)a * b;
    hi = (uint64_t)(p >> 64); lo = (uint64_t)p;
#else
    hi = __umul64hi(a, b); lo = a * b;
#endif
}

template <int NL>
struct Mpf {
    uint64_t l[NL];     // l[NL-1] most significant
    int32_t  e;         // exponent in limbs: value = 0.l * 2^(64*e)
    uint32_t s;         // 1 = negative
};

template <int NL> MDZ_HD bool gz(const Mpf<NL>& a) { return a.l[NL - 1] == 0; }

template <int NL> MDZ_HD void gset_zero(Mpf<NL>& a)
{
    for (int i = 0; i < NL; ++i) a.l[i] = 0;
    a.e = 0; a.s = 0;
}

// mpf_mul: operands cut to their top P limbs, exact product, a zero top limb
// dropped (exponent - 1), top P+1 limbs kept.
template <int NL>
MDZ_HD void gmul(const Mpf<NL>& u, const Mpf<NL>& v, Mpf<NL>& r)
{
    constexpr int P = NL - 1;
    if (gz(u) || gz(v)) { gset_zero(r); return; }
    uint64_t t[2 * P];
    for (int i = 0; i < 2 * P; ++i) t[i] = 0;
    for (int i = 0; i < P; ++i) {
        uint64_t carry = 0;
        const uint64_t ui = u.l[i + 1];
        for (int j = 0; j < P; ++j) {
            uint64_t hi, lo;
            mul64(ui, v.l[j + 1], hi, lo);
            uint64_t s = t[i + j] + lo;
            hi += (s < lo);
            uint64_t s2 = s + carry;
            hi += (s2 < s);
            t[i + j] = s2;
            carry = hi;
        }
        t[i + P] = carry;
    }
    const int adj = (t[2 * P - 1] == 0) ? 1 : 0;
    // keep limbs [2P-1-adj-P .. 2P-1-adj]
    for (int i = 0; i < NL; ++i) {
        const int k = i + (2 * P - 1 - adj - P);
        r.l[i] = (k >= 0) ? t[k] : 0;
    }
    r.e = u.e + v.e - adj;
    r.s = u.s ^ v.s;
}

// mpf_mul_ui(r, u, 2): the limb below the top P only contributes its carry.
template <int NL>
MDZ_HD void gmul2(const Mpf<NL>& u, Mpf<NL>& r)
{
    if (gz(u)) { gset_zero(r); return; }
    // T = 2*hi + (lo >> 63), hi = top P limbs, as P+1 limbs
    uint64_t t[NL];
    uint64_t cin = u.l[0] >> 63;
    for (int i = 1; i < NL; ++i) {
        t[i - 1] = (u.l[i] << 1) | cin;
        cin = u.l[i] >> 63;
    }
    t[NL - 1] = cin;
    if (t[NL - 1] == 0) {
        // drop the zero top limb: P limbs, top aligned in NL
        for (int i = NL - 1; i >= 1; --i) r.l[i] = t[i - 1];
        r.l[0] = 0;
        r.e = u.e;
    } else {
        for (int i = 0; i < NL; ++i) r.l[i] = t[i];
        r.e = u.e + 1;
    }
    r.s = u.s;
}

// magnitude add, same sign (mpf_add): works to P limbs
template <int NL>
MDZ_HD void gadd_mag(const Mpf<NL>& a, const Mpf<NL>& b, Mpf<NL>& r, uint32_t sign)
{
    constexpr int P = NL - 1;
    const bool swap = a.e < b.e;                    // not on ties
    const Mpf<NL>& u = swap ? b : a;
    const Mpf<NL>& v = swap ? a : b;
    const int d = u.e - v.e;
    if (d >= P) {                                   // v vanishes: result is u cut to P limbs
        for (int i = 1; i < NL; ++i) r.l[i] = u.l[i];
        r.l[0] = 0; r.e = u.e; r.s = sign;
        return;
    }
    // window: the top P limbs below u's top.  index w = 0..P-1 <-> array index w+1
    uint64_t t[NL];
    uint64_t carry = 0;
    for (int w = 0; w < P; ++w) {
        const uint64_t x = u.l[w + 1];
        const int vi = w + 1 + d;                   // v's limb at this position
        const uint64_t y = (vi < NL) ? v.l[vi] : 0;
        uint64_t s = x + y;
        uint64_t c1 = (s < x);
        uint64_t s2 = s + carry;
        c1 += (s2 < s);
        t[w] = s2;
        carry = c1;
    }
    if (carry) {
        for (int w = 0; w < P; ++w) r.l[w] = t[w];
        r.l[P] = 1;
        r.e = u.e + 1;
    } else {
        for (int w = 0; w < P; ++w) r.l[w + 1] = t[w];
        r.l[0] = 0;
        r.e = u.e;
    }
    r.s = sign;
}

// magnitude subtract, same sign (mpf_sub): works to Q = P+1 limbs, exact
// difference over that window, leading zero limbs stripped.  `neg` is the sign
// the result has when |a| > |b|.
template <int NL>
MDZ_HD void gsub_mag(const Mpf<NL>& a, const Mpf<NL>& b, Mpf<NL>& r, uint32_t neg)
{
    const bool swap = a.e < b.e;
    const Mpf<NL>& u = swap ? b : a;
    const Mpf<NL>& v = swap ? a : b;
    if (swap) neg ^= 1u;
    const int d = u.e - v.e;
    if (d >= NL) { r = u; r.s = neg; return; }
    if (d == 1 && u.l[NL - 1] == 1 && v.l[NL - 1] == ~0ull && u.l[NL - 2] == 0) {
        // GMP's "close" path for a gap of one limb (Appendix E): u = 1:0:..., v = ff..ff:...
        // The leading limbs cancel pairwise and the working window slides down with
        // them, so v's lowest limb takes part unless nothing below the top cancelled.
        constexpr int Q = NL;
        uint64_t U[NL], V[NL];
        int nu = NL, nv = NL, e = u.e;
        for (int i = 0; i < NL; ++i) { U[i] = u.l[i]; V[i] = v.l[i]; }
        --nu; --e;                                           // drop u's top limb (the 1)
        while (nu > 0 && nv > 0 && U[nu - 1] == 0 && V[nv - 1] == ~0ull) { --nu; --nv; --e; }
        int ulo = 0, vlo = 0;                                // first limb kept (low end)
        if (nu == 0) { while (nv > 0 && V[nv - 1] == ~0ull) { --nv; --e; } }
        else if (nu > Q - 1) ulo = nu - (Q - 1);
        if (nv > Q - 1) vlo = nv - (Q - 1);
        const int su = nu - ulo, sv = nv - vlo;              // sizes after truncation
        const int n = su > sv ? su : sv;
        uint64_t t[NL + 1];
        for (int i = 0; i <= NL; ++i) t[i] = 0;
        int tn;
        if (sv == 0) {
            for (int i = 0; i < su; ++i) t[i] = U[ulo + i];
            t[su] = 1; tn = su + 1; e += 1;
        } else {
            // t = (u - v) mod 2^(64n), both top aligned in n limbs
            uint64_t borrow = 0;
            for (int w = 0; w < n; ++w) {
                const int iu = w - (n - su), iv = w - (n - sv);
                const uint64_t x = iu >= 0 ? U[ulo + iu] : 0;
                const uint64_t z = iv >= 0 ? V[vlo + iv] : 0;
                const uint64_t s1 = x - z;
                const uint64_t b1 = x < z;
                const uint64_t s2 = s1 - borrow;
                const uint64_t b2 = s1 < borrow;
                t[w] = s2;
                borrow = b1 | b2;
            }
            if (!borrow) { t[n] = 1; tn = n + 1; e += 1; }
            else { tn = n; while (tn > 0 && t[tn - 1] == 0) { --tn; --e; } }
        }
        if (tn == 0) { gset_zero(r); return; }
        // top-align tn limbs into NL (tn <= NL)
        for (int w = NL - 1; w >= 0; --w) { const int k = w - (NL - tn); r.l[w] = k >= 0 ? t[k] : 0; }
        r.e = e;
        r.s = neg;
        return;
    }
    // v aligned into u's window (its limbs below the window are dropped)
    uint64_t y[NL];
    for (int w = 0; w < NL; ++w) { const int vi = w + d; y[w] = (vi < NL) ? v.l[vi] : 0; }
    // order by magnitude inside the window (a flip is only possible when d == 0)
    bool flip = false;
    if (d == 0) {
        for (int w = NL - 1; w >= 0; --w) {
            if (u.l[w] != y[w]) { flip = u.l[w] < y[w]; break; }
        }
    }
    uint64_t t[NL];
    uint64_t borrow = 0;
    for (int w = 0; w < NL; ++w) {
        const uint64_t x = flip ? y[w] : u.l[w];
        const uint64_t z = flip ? u.l[w] : y[w];
        const uint64_t s = x - z;
        const uint64_t b1 = (x < z);
        const uint64_t s2 = s - borrow;
        const uint64_t b2 = (s < borrow);
        t[w] = s2;
        borrow = b1 | b2;
    }
    if (flip) neg ^= 1u;
    // strip leading zero limbs
    int k = 0;
    while (k < NL && t[NL - 1 - k] == 0) ++k;
    if (k == NL) { gset_zero(r); return; }
    for (int w = NL - 1; w >= 0; --w) r.l[w] = (w - k >= 0) ? t[w - k] : 0;
    r.e = u.e - k;
    r.s = neg;
}

// mpf_add / mpf_sub with signs
template <int NL>
MDZ_HD void gadd(const Mpf<NL>& a, const Mpf<NL>& b, Mpf<NL>& r, bool subtract)
{
    const uint32_t sb = b.s ^ (subtract ? 1u : 0u);
    if (gz(a)) { r = b; r.s = gz(b) ? 0u : sb; return; }
    if (gz(b)) { r = a; return; }
    if (a.s == sb) gadd_mag<NL>(a, b, r, a.s);
    else           gsub_mag<NL>(a, b, r, a.s);
}

// mpf_cmp(a, 4) > 0
template <int NL>
MDZ_HD bool ggt4(const Mpf<NL>& a)
{
    if (gz(a) || a.s) return false;
    if (a.e != 1) return a.e > 1;
    if (a.l[NL - 1] != 4) return a.l[NL - 1] > 4;
    for (int i = 0; i < NL - 1; ++i) if (a.l[i]) return true;
    return false;
}

// ---- one pixel: frac_mandel_gmp (src/frac_mandel.c:55-82) and its variants ----
// (src/frac_burning_ship.c:58-86, src/frac_generalized_celtic.c:58-86, src/frac_variant.c:58-87)
template <int NL>
struct GmpPixel {
    Mpf<NL> wre, wim, wre2, wim2, cre, cim;
    int iter;
};

template <int NL>
MDZ_HD void gmp_pixel_init(GmpPixel<NL>& st, const Mpf<NL>& x, const Mpf<NL>& y,
                           const Mpf<NL>& cx, const Mpf<NL>& cy)
{
    st.wre = x; st.wim = y;
    gmul<NL>(x, x, st.wre2);        // src/fractal.c:328, :331-333
    gmul<NL>(y, y, st.wim2);        // :310
    st.cre = cx; st.cim = cy;
    st.iter = 0;
}

template <int NL>
MDZ_HD bool gmp_pixel_step(GmpPixel<NL>& st, bool abs_im, int abs_re)
{
    ++st.iter;
    Mpf<NL> t1, t2;
    gmul<NL>(st.wre, st.wim, t1);
    if (abs_im) t1.s = 0;
    gmul2<NL>(t1, t2);
    gadd<NL>(t2, st.cim, st.wim, false);
    gadd<NL>(st.wre2, st.wim2, t1, true);
    if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) t1.s = 0;
    gadd<NL>(t1, st.cre, st.wre, false);
    gmul<NL>(st.wim, st.wim, st.wim2);
    gmul<NL>(st.wre, st.wre, st.wre2);
    gadd<NL>(st.wim2, st.wre2, t1, false);
    return ggt4<NL>(t1);
}

}  // namespace mdz
