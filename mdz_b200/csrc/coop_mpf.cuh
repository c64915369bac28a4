// coop_mpf.cuh -- GMP-mpf-faithful arithmetic with a GROUP OF LANES PER VALUE, for mpf precisions above 512 bits
// (more than a thread can hold).  Same contract as mpf_sf.cuh -- mpf_mul, mpf_mul_ui(., 2), mpf_add, mpf_sub,
// mpf_cmp(., 4) as GMP 6.3.0 behaves, SURVEY Appendix E, the calls of the reference's frac_*_gmp loops
// (src/frac_mandel.c:55-82 and its three siblings) -- on the lane-blocked layout and the cross-lane machinery of
// coop_ops.cuh: T lanes x K 32-bit words, least significant first, K even so that a 64-bit limb never straddles
// two lanes.
//
// A value's NL = P + 1 limbs are top aligned in the N = T K words (N / 2 >= NL + 1: one spare limb, so that the
// high half of a product still holds the limb an adjusted product needs); everything below word `lowcut` =
// N - 2 NL is zero.  The exponent counts 64-bit limbs; zero is e == E_ZERO with all words zero (mpf_sf.cuh looks at
// the top limb instead: the two agree because every operation tests for zero first).  mpf arithmetic is
// limb-granular: nothing is rounded and the only bit shift is the doubling, so alignment is a word shift through
// the shared-memory strip, normalisation strips whole zero limbs, and a product is coop_mul_full followed by a cut.
// GMP's one-limb-gap "close" subtraction (u = 1:0:..., v = ff..ff:...) is run by one lane over the strip: it is a
// chain of data-dependent scans, and the orbit meets it about never.
//
// As in coop_ops.cuh products are executed by every lane of the warp at once (no group-dependent branch), additions
// by each group on its own.
#pragma once
#include "coop_ops.cuh"

namespace mdz {

struct CoopGCfg {
    int lowcut;         // index of the first word of limb l[0]: N - 2 NL
    int nl;             // NL = P + 1 limbs
};
template <int K, int T> MDZ_HD CoopGCfg make_coop_gcfg(int nl) { CoopGCfg c; c.nl = nl; c.lowcut = T * K - 2 * nl; return c; }

// zero the words whose index in the value is below `from`
template <int K, int T>
MDZ_HD void cg_cut(LW (&x)[K], int from)
{
    const LW i0 = lane_in<T>() * LW((uint32_t)K);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) x[j] = x[j] & ~m_lts(i0 + LW((uint32_t)j), LW((uint32_t)from));
}

// one limb (two words) towards the top; the vacated bottom limb is zero
template <int K, int T, bool W>
MDZ_HD void cg_up1(const LW (&x)[K], LW (&y)[K])
{
    const LW p0 = shfl_up1<T, W>(x[K - 2]), p1 = shfl_up1<T, W>(x[K - 1]);
    MDZ_UNROLL
    for (int j = K - 1; j >= 2; --j) y[j] = x[j - 2];
    y[0] = p0; y[1] = p1;
}
// one limb towards the bottom (its lowest limb leaves); the vacated top limb becomes 1
template <int K, int T>
MDZ_HD void cg_down1_top1(LW (&x)[K])
{
    const LW n0 = shfl_dn1<T>(x[0]), n1 = shfl_dn1<T>(x[1]);
    MDZ_UNROLL
    for (int j = 0; j + 2 < K; ++j) x[j] = x[j + 2];
    x[K - 2] = n0 | sel(m_eq(lane_in<T>(), LW((uint32_t)(T - 1))), LW(1u), LW(0u));
    x[K - 1] = n1;
}

template <int K, int T> MDZ_HD void cg_zero(CNum<K, T>& a) { cset_zero(a); }

// mpf_mul: operands cut to their top P limbs, exact product, a zero top limb dropped (exponent - 1), top P + 1
// limbs kept.  Every lane of the warp at once.
template <int K, int T>
MDZ_HD void cg_mul(const CNum<K, T>& u, const CNum<K, T>& v, CNum<K, T>& r, const CoopGCfg& cfg)
{
    static_assert(K >= 4 && (K & 1) == 0, "whole limbs per lane, and the top two limbs in the top lane");
    const bool zero = cis_zero(u) || cis_zero(v);
    if (T == 32 && zero) { cg_zero(r); return; }
    if (T != 32) warp_converge();
    LW a[K], b[K];
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) { a[j] = u.m[j]; b[j] = v.m[j]; }
    cg_cut<K, T>(a, cfg.lowcut + 2);
    cg_cut<K, T>(b, cfg.lowcut + 2);
    LW h[K], up[K];
    uint32_t lowtop, low_sticky;
    coop_mul_full<K, T>(a, b, h, lowtop, low_sticky);
    const uint32_t top = bcast<T, true>(h[K - 1] | h[K - 2], T - 1);
    const uint32_t adj = top == 0u ? 1u : 0u;
    cg_up1<K, T, true>(h, up);
    const LW am = LW(adj ? 0xffffffffu : 0u);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r.m[j] = sel(am, up[j], h[j]);
    cg_cut<K, T>(r.m, cfg.lowcut);
    r.e = u.e + v.e - (int32_t)adj;
    r.s = u.s ^ v.s;
    if (zero) cg_zero(r);
}

// mpf_mul_ui(r, u, 2): the limb below the top P only contributes its carry
template <int K, int T>
MDZ_HD void cg_mul2(const CNum<K, T>& u, CNum<K, T>& r, const CoopGCfg& cfg)
{
    if (cis_zero(u)) { cg_zero(r); return; }
    const LW below = shfl_up1<T>(u.m[K - 1]);
    const uint32_t cout = bcast<T>(u.m[K - 1], T - 1) >> 31;
    LW y[K];
    MDZ_UNROLL
    for (int j = K - 1; j >= 1; --j) y[j] = fsl(u.m[j - 1], u.m[j], LW(1u));
    y[0] = fsl(below, u.m[0], LW(1u));
    cg_cut<K, T>(y, cfg.lowcut + 2);
    if (cout) cg_down1_top1<K, T>(y);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r.m[j] = y[j];
    r.e = u.e + (int32_t)cout;
    r.s = u.s;
}

// magnitude add, same sign (mpf_add): works to P limbs
template <int K, int T>
MDZ_HD void cg_add_mag(const CNum<K, T>& a, const CNum<K, T>& b, CNum<K, T>& r, uint32_t sign, const CoopGCfg& cfg, uint32_t* scr)
{
    const bool swap = a.e < b.e;                    // not on ties
    const CNum<K, T>& u = swap ? b : a;
    const CNum<K, T>& v = swap ? a : b;
    const int32_t d = u.e - v.e;
    LW x[K];
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) x[j] = u.m[j];
    cg_cut<K, T>(x, cfg.lowcut + 2);
    if (d >= cfg.nl - 1) {                          // v vanishes: u cut to P limbs
        MDZ_UNROLL
        for (int j = 0; j < K; ++j) r.m[j] = x[j];
        r.e = u.e; r.s = sign;
        return;
    }
    LW y[K], t[K];
    uint32_t g, st;
    coop_shr<K, T>(v.m, 64u * (uint32_t)d, y, g, st, scr);
    cg_cut<K, T>(y, cfg.lowcut + 2);
    const uint32_t cout = coop_add_n<K, T>(x, y, t, 0u);
    if (cout) cg_down1_top1<K, T>(t);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r.m[j] = t[j];
    r.e = u.e + (int32_t)cout;
    r.s = sign;
}

// ---- GMP's "close" subtraction, by one lane over the strip --------------------------------------------------
// U, V: the operands' NL limbs, least significant first, as pairs of words; R: NL limbs of result, top aligned.
// Returns the result's limb count (0: zero) and its exponent.  (mpf_sf.cuh gsub_mag, first branch.)
MDZ_HD uint64_t cg_limb(const uint32_t* a, int i) { return ((uint64_t)a[2 * i + 1] << 32) | a[2 * i]; }
MDZ_HD void cg_set_limb(uint32_t* a, int i, uint64_t w) { a[2 * i] = (uint32_t)w; a[2 * i + 1] = (uint32_t)(w >> 32); }
MDZ_HD int cg_sub_close_serial(const uint32_t* U, const uint32_t* V, uint32_t* R, int NL, int32_t& e_io)
{
    const int Q = NL;
    int nu = NL, nv = NL;
    int32_t e = e_io;
    --nu; --e;                                              // u's top limb (the 1)
    while (nu > 0 && nv > 0 && cg_limb(U, nu - 1) == 0 && cg_limb(V, nv - 1) == ~0ull) { --nu; --nv; --e; }
    int ulo = 0, vlo = 0;
    if (nu == 0) { while (nv > 0 && cg_limb(V, nv - 1) == ~0ull) { --nv; --e; } }
    else if (nu > Q - 1) ulo = nu - (Q - 1);
    if (nv > Q - 1) vlo = nv - (Q - 1);
    const int su = nu - ulo, sv = nv - vlo;
    const int n = su > sv ? su : sv;
    // t = R used as n + 1 limbs, least significant first; top aligned into NL afterwards (in place, from the top)
    int tn;
    if (sv == 0) {
        for (int i = 0; i < su; ++i) cg_set_limb(R, i, cg_limb(U, ulo + i));
        cg_set_limb(R, su, 1); tn = su + 1; e += 1;
    } else {
        uint64_t borrow = 0;
        for (int w = 0; w < n; ++w) {
            const int iu = w - (n - su), iv = w - (n - sv);
            const uint64_t x = iu >= 0 ? cg_limb(U, ulo + iu) : 0;
            const uint64_t z = iv >= 0 ? cg_limb(V, vlo + iv) : 0;
            const uint64_t s1 = x - z;
            const uint64_t b1 = x < z;
            const uint64_t s2 = s1 - borrow;
            const uint64_t b2 = s1 < borrow;
            cg_set_limb(R, w, s2);
            borrow = b1 | b2;
        }
        if (!borrow) { cg_set_limb(R, n, 1); tn = n + 1; e += 1; }
        else { tn = n; while (tn > 0 && cg_limb(R, tn - 1) == 0) { --tn; --e; } }
    }
    if (tn > 0 && tn < NL) {
        for (int w = NL - 1; w >= 0; --w) { const int k = w - (NL - tn); cg_set_limb(R, w, k >= 0 ? cg_limb(R, k) : 0); }
    }
    e_io = e;
    return tn;
}

#if defined(MDZ_HOST_EMU)
template <int T> inline bool serial_lane() { return true; }
static long g_cg_close_calls = 0;       // the differential test checks that its operands reach this path
#define MDZ_CG_COUNT_CLOSE() (++g_cg_close_calls)
#else
#define MDZ_CG_COUNT_CLOSE() ((void)0)
template <int T> MDZ_HD bool serial_lane() { return (threadIdx.x & (unsigned)(T - 1)) == 0u; }
#endif

template <int K, int T>
MDZ_HD void cg_sub_close(const CNum<K, T>& u, const CNum<K, T>& v, CNum<K, T>& r, uint32_t neg, const CoopGCfg& cfg, uint32_t* scr)
{
    constexpr int N = T * K;
    MDZ_CG_COUNT_CLOSE();
    // strip: U at [0, N), V at [N, 2N), R at [2N, 3N + 2), the verdict in the two words after R's NL limbs
    const LW i0 = lane_in<T>() * LW((uint32_t)K);
    warp_sync<T>();
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) {
        sm_store(scr, i0 + LW((uint32_t)j), u.m[j]);
        sm_store(scr, i0 + LW((uint32_t)(N + j)), v.m[j]);
    }
    warp_sync<T>();
    if (serial_lane<T>()) {
        int32_t e = u.e;
        uint32_t* R = scr + 2 * N;
        const int tn = cg_sub_close_serial(scr + cfg.lowcut, scr + N + cfg.lowcut, R, cfg.nl, e);
        R[2 * cfg.nl] = (uint32_t)tn;
        R[2 * cfg.nl + 1] = (uint32_t)e;
    }
    warp_sync<T>();
    const uint32_t tn = bcast<T>(sm_load(scr, LW((uint32_t)(2 * N + 2 * cfg.nl))), 0);
    const int32_t e = (int32_t)bcast<T>(sm_load(scr, LW((uint32_t)(2 * N + 2 * cfg.nl + 1))), 0);
    // R's NL limbs belong at words lowcut .. N of the value
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) {
        const LW at = i0 + LW((uint32_t)j) - LW((uint32_t)cfg.lowcut);              // index into R; negative below lowcut
        const LW in = ~m_lts(at, LW(0u));
        r.m[j] = sm_load(scr, sel(in, at + LW((uint32_t)(2 * N)), LW(0u))) & in;
    }
    warp_sync<T>();
    coop_scratch_init<K, T>(scr);                   // the strip's zero margins are what the shifters rely on
    if (tn == 0u) { cg_zero(r); return; }
    r.e = e;
    r.s = neg;
}

// magnitude subtract, same sign (mpf_sub): works to NL = P + 1 limbs, exact difference over that window, leading
// zero limbs stripped.  `neg` is the sign the result has when |a| > |b|.
template <int K, int T>
MDZ_HD void cg_sub_mag(const CNum<K, T>& a, const CNum<K, T>& b, CNum<K, T>& r, uint32_t neg, const CoopGCfg& cfg, uint32_t* scr)
{
    constexpr int N = T * K;
    const bool swap = a.e < b.e;
    const CNum<K, T>& u = swap ? b : a;
    const CNum<K, T>& v = swap ? a : b;
    if (swap) neg ^= 1u;
    const int32_t d = u.e - v.e;
    if (d >= cfg.nl) { r = u; r.s = neg; return; }
    if (d == 1) {
        // u = 1:0:..., v = ff..ff:... ?   (the top two limbs are the top lane's last four words)
        const LW is1 = m_eq(u.m[K - 1], LW(0u)) & m_eq(u.m[K - 2], LW(1u)) & m_eq(u.m[K - 3] | u.m[K - 4], LW(0u));
        const LW isf = m_eq(v.m[K - 1] & v.m[K - 2], LW(0xffffffffu));
        if (bcast<T>(is1 & isf, T - 1) != 0u) { cg_sub_close<K, T>(u, v, r, neg, cfg, scr); return; }
    }
    LW y[K], t[K];
    uint32_t g, st;
    coop_shr<K, T>(v.m, 64u * (uint32_t)d, y, g, st, scr);
    cg_cut<K, T>(y, cfg.lowcut);
    bool flip = false;
    if (d == 0) flip = coop_cmp<K, T>(u.m, y) < 0;
    if (flip) { (void)coop_sub_n<K, T>(y, u.m, t, 0u); neg ^= 1u; }
    else (void)coop_sub_n<K, T>(u.m, y, t, 0u);
    const int z = coop_clz<K, T>(t);
    if (z == 32 * N) { cg_zero(r); return; }
    const int k = z >> 6;                           // whole zero limbs at the top
    if (k > 0) { uint32_t g0 = 0u; coop_shl<K, T>(t, g0, 64u * (uint32_t)k, scr); }
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r.m[j] = t[j];
    r.e = u.e - k;
    r.s = neg;
}

// mpf_add / mpf_sub with signs
template <int K, int T>
MDZ_HD void cg_add(const CNum<K, T>& a, const CNum<K, T>& b, CNum<K, T>& r, bool subtract, const CoopGCfg& cfg, uint32_t* scr)
{
    const uint32_t sb = b.s ^ (subtract ? 1u : 0u);
    if (cis_zero(a)) { r = b; r.s = cis_zero(b) ? 0u : sb; return; }
    if (cis_zero(b)) { r = a; return; }
    if (a.s == sb) cg_add_mag<K, T>(a, b, r, a.s, cfg, scr);
    else           cg_sub_mag<K, T>(a, b, r, a.s, cfg, scr);
}

// mpf_cmp(a, 4) > 0
template <int K, int T>
MDZ_HD bool cg_gt4(const CNum<K, T>& a)
{
    if (cis_zero(a) || a.s) return false;
    if (a.e != 1) return a.e > 1;
    const uint32_t hi = bcast<T>(a.m[K - 1], T - 1), lo = bcast<T>(a.m[K - 2], T - 1);
    if (hi != 0u || lo != 4u) return hi != 0u || lo > 4u;
    const LW top = m_eq(lane_in<T>(), LW((uint32_t)(T - 1)));
    LW rest = LW(0u);
    MDZ_UNROLL
    for (int j = 0; j < K - 2; ++j) rest = rest | a.m[j];
    rest = rest | ((a.m[K - 2] | a.m[K - 1]) & ~top);
    return ballot_nz<T>(rest) != 0u;
}

// a table entry as it comes from the host (exponent 0 for zero): zero is e == E_ZERO here
template <int K, int T>
MDZ_HD void cg_adopt(CNum<K, T>& a)
{
    if (bcast<T>(a.m[K - 1] | a.m[K - 2], T - 1) == 0u) cg_zero(a);
}

// ---- one pixel: frac_mandel_gmp (src/frac_mandel.c:55-82) and its variants (src/frac_burning_ship.c:58-86,
// src/frac_generalized_celtic.c:58-86, src/frac_variant.c:58-87); set-up src/fractal.c:310, :328-333 --------------
template <int K, int T>
MDZ_HD void cgpixel_squares(CPixel<K, T>& st, const CoopGCfg& cfg)
{
    cg_mul<K, T>(st.wre, st.wre, st.wre2, cfg);
    cg_mul<K, T>(st.wim, st.wim, st.wim2, cfg);
}

template <int K, int T>
MDZ_HD void cgpixel_init(CPixel<K, T>& st, const CNum<K, T>& x, const CNum<K, T>& y, const CNum<K, T>& cx, const CNum<K, T>& cy, const CoopGCfg& cfg)
{
    cpixel_load<K, T>(st, x, y, cx, cy);
    cgpixel_squares<K, T>(st, cfg);
}

// One iteration; `active` as in cpixel_step: the products are the whole warp's, the rest the group's own.
template <int K, int T>
MDZ_HD bool cgpixel_step(CPixel<K, T>& st, const CoopGCfg& cfg, uint32_t* scr, bool abs_im, int abs_re, bool active = true)
{
    CNum<K, T> t1, t2;
    cg_mul<K, T>(st.wre, st.wim, t1, cfg);
    if (active) {
        ++st.iter;
        if (abs_im) t1.s = 0u;
        cg_mul2<K, T>(t1, t2, cfg);
        cg_add<K, T>(t2, st.cim, st.wim, false, cfg, scr);
        cg_add<K, T>(st.wre2, st.wim2, t1, true, cfg, scr);
        if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) t1.s = 0u;
        cg_add<K, T>(t1, st.cre, st.wre, false, cfg, scr);
    }
    cg_mul<K, T>(st.wim, st.wim, st.wim2, cfg);
    cg_mul<K, T>(st.wre, st.wre, st.wre2, cfg);
    if (!active) return false;
    cg_add<K, T>(st.wim2, st.wre2, t1, false, cfg, scr);
    return cg_gt4<K, T>(t1);
}

}  // namespace mdz
