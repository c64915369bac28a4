// mpfr_sf.cuh -- MPFR-faithful soft floating point on N x 32-bit limbs.
//
// Every operation is "exact result, rounded once to p bits, nearest, ties to
// even" -- the contract of mpfr_mul / mpfr_add / mpfr_sub with MPFR_RNDN that
// the reference's hot loop relies on (reference src/frac_mandel.c:36-48, and
// the same nine calls in frac_burning_ship.c / frac_generalized_celtic.c /
// frac_variant.c).  At p = 64 the same rule reproduces x87 `long double`
// (reference src/fractal.c:38-48, src/frac_mandel.c:5-21; SURVEY finding 1).
//
// Representation (registers):  value = (-1)^s * 0.m * 2^e,  m = N limbs,
// most significant limb m[N-1] with bit 31 set (normalised), the low
// R = 32N - p bits of m[0] zero.  Zero is m[N-1] == 0 (all limbs 0) with
// e = E_ZERO so that alignment treats it as infinitely small.  Exponents are
// int32; products whose exponent falls below E_MIN flush to zero (MPFR's range
// is 2^62; only a Julia orbit with c == 0 squares that far down and it never
// escapes either way -- DESIGN.md "exponent range").
#pragma once
#include "limb_ops.cuh"

namespace mdz {

constexpr int32_t E_ZERO = -(1 << 29);
constexpr int32_t E_MIN  = -(1 << 28);

// Rounding position inside limb 0 / the guard limb, uniform for a render.
struct RoundCfg {
    uint32_t half_x0;   // round bit mask inside m[0]      (R > 0)
    uint32_t half_g;    // round bit mask inside the guard (R == 0)
    uint32_t below_x0;  // bits of m[0] strictly below the round bit
    uint32_t below_g;   // bits of the guard strictly below the round bit
    uint32_t ulp;       // 1 << R
    uint32_t keep;      // ~(ulp - 1)
};

inline RoundCfg make_round_cfg(int nlimbs, int prec_bits)
{
    const int R = 32 * nlimbs - prec_bits;          // 0..31
    RoundCfg c;
    c.ulp  = 1u << R;
    c.keep = ~(c.ulp - 1u);
    if (R == 0) { c.half_x0 = 0; c.half_g = 0x80000000u; c.below_x0 = 0; c.below_g = 0x7fffffffu; }
    else { c.half_x0 = 1u << (R - 1); c.half_g = 0; c.below_x0 = c.half_x0 - 1u; c.below_g = 0xffffffffu; }
    return c;
}

template <int N>
struct Num {
    uint32_t m[N];
    int32_t  e;
    uint32_t s;     // 1 = negative
};

template <int N> MDZ_HD bool is_zero(const Num<N>& a) { return a.m[N - 1] == 0; }

template <int N> MDZ_HD void set_zero(Num<N>& a)
{
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) a.m[i] = 0;
    a.e = E_ZERO; a.s = 0;
}

// Round x[1..N] (guard limb x[0], sticky flag below it) to p bits, RN-even.
// Returns 1 if the increment carried out of the top limb (mantissa is then
// 0x80000000:0...; the caller bumps the exponent).
template <int N>
MDZ_HD uint32_t round_rn(uint32_t (&x)[N + 1], uint32_t sticky, const RoundCfg& rc)
{
    const uint32_t x0 = x[1], g = x[0];
    const uint32_t rb  = (x0 & rc.half_x0) | (g & rc.half_g);
    const uint32_t st  = sticky | (x0 & rc.below_x0) | (g & rc.below_g);
    const uint32_t lsb = x0 & rc.ulp;
    const uint32_t inc = (rb != 0 && (st | lsb) != 0) ? rc.ulp : 0u;
    x[1] = add_cc(x0 & rc.keep, inc);
    MDZ_UNROLL
    for (int i = 2; i <= N; ++i) x[i] = addc_cc(x[i], 0u);
    const uint32_t cout = addc(0u, 0u);
    x[N] |= cout << 31;      // all-ones + ulp -> 1000...0
    return cout;
}

// r = RN(a * b) from a full product already in prod[2N]; e = ea + eb.
template <int N>
MDZ_HD void finish_product(const uint32_t (&prod)[2 * N], int32_t e, uint32_t s,
                           Num<N>& r, const RoundCfg& rc)
{
    // the product of two normalised significands has its top bit at 64N-1 or 64N-2
    const uint32_t sh = (prod[2 * N - 1] >> 31) ^ 1u;       // 1 -> shift left one bit
    uint32_t x[N + 1];
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(prod[N - 2 + i], prod[N - 1 + i], sh);
    uint32_t lowtop = (N >= 2) ? prod[N - 2] : 0u;
    x[0] = fsl(lowtop, prod[N - 1], sh);
    uint32_t sticky = (N >= 2) ? (lowtop << sh) : 0u;       // bits of prod[N-2] left below the guard
    MDZ_UNROLL
    for (int i = 0; i + 2 < N; ++i) sticky |= prod[i];
    const uint32_t cout = round_rn<N>(x, sticky, rc);
    e = e - (int32_t)sh + (int32_t)cout;
    const bool nz = x[N] != 0;                              // zero operand -> zero product
    const bool ok = nz && e >= E_MIN;
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = ok ? x[i + 1] : 0u;
    r.e = ok ? e : E_ZERO;
    r.s = ok ? s : 0u;
}

template <int N>
MDZ_HD void fmul(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc)
{
    uint32_t prod[2 * N];
    mul_full<N>(a.m, b.m, prod);
    finish_product<N>(prod, a.e + b.e, a.s ^ b.s, r, rc);
}

template <int N>
MDZ_HD void fsqr(const Num<N>& a, Num<N>& r, const RoundCfg& rc)
{
    uint32_t prod[2 * N];
    sqr_full<N>(a.m, prod);
    finish_product<N>(prod, a.e + a.e, 0u, r, rc);
}

// ---------------------------------------------------------------------------
// r = RN(a + b)            MODE_GENERIC  (signs as given)
// r = RN(a - b), a,b >= 0  MODE_SUB_POS
// r = RN(a + b), a,b >= 0  MODE_ADD_POS
//
// Both operands get a guard limb below and are pre-shifted right by one bit so
// that a same-sign sum cannot carry out; the one with the smaller exponent is
// shifted further by the exponent gap.  Gaps below 31 bits lose nothing (the
// guard limb catches them) and are pure funnel shifts; larger gaps take a
// divergent limb-shift loop that folds what drops off into a sticky flag.
// For an effective subtraction the smaller operand is complemented (chosen
// from exponents and top limbs; exact ties on both take a rare full compare),
// with the sticky bit acting as the borrow, so the difference is never
// negative.  Normalisation is one left funnel shift by clz (0..1 after an
// addition, 1.. after a subtraction; >= 32 only on massive cancellation, which
// is exact and handled by a limb loop).
// ---------------------------------------------------------------------------
enum { MODE_GENERIC = 0, MODE_SUB_POS = 1, MODE_ADD_POS = 2 };

template <int N>
MDZ_HD void limb_shift_right(uint32_t (&x)[N + 1], int q, uint32_t& sticky)
{
    if (q > N + 1) q = N + 1;
    for (int k = 0; k < q; ++k) {
        sticky |= x[0];
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = x[i + 1];
        x[N] = 0;
    }
}

template <int N, int MODE>
MDZ_HD void fadd(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc)
{
    const uint32_t sb = (MODE == MODE_SUB_POS) ? 1u : (MODE == MODE_ADD_POS ? 0u : b.s);
    const uint32_t sa = (MODE == MODE_GENERIC) ? a.s : 0u;
    const bool sub = (MODE == MODE_SUB_POS) ? true : (MODE == MODE_ADD_POS ? false : (sa != sb));

    int32_t d = a.e - b.e;
    uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    if (ad > 32u * (N + 3)) ad = 32u * (N + 3);

    uint32_t xa[N + 1], xb[N + 1];
    xa[0] = 0; xb[0] = 0;
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) { xa[i + 1] = a.m[i]; xb[i + 1] = b.m[i]; }

    uint32_t sticky = 0;
    uint32_t sh = ad + 1u;                 // includes the one-bit headroom pre-shift
    if (sh >= 32u) {                       // rare: exponent gap of 31 bits or more
        const int q = (int)(sh >> 5);
        if (d > 0) limb_shift_right<N>(xb, q, sticky);
        else       limb_shift_right<N>(xa, q, sticky);
        sh &= 31u;
    }
    const uint32_t sha = (d < 0) ? sh : 1u;
    const uint32_t shb = (d > 0) ? sh : 1u;
    // bits that leave the guard limb (only possible after a limb shift)
    sticky |= fsr(0u, xa[0], sha) | fsr(0u, xb[0], shb);
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) { xa[i] = fsr(xa[i], xa[i + 1], sha); xb[i] = fsr(xb[i], xb[i + 1], shb); }
    xa[N] >>= sha; xb[N] >>= shb;
    sticky = (sticky != 0) ? 1u : 0u;

    int32_t e = (d < 0 ? b.e : a.e) + 1;
    uint32_t s = sa;
    uint32_t x[N + 1];

    if (!sub) {
        x[0] = add_cc(xa[0], xb[0]);
        MDZ_UNROLL
        for (int i = 1; i < N; ++i) x[i] = addc_cc(xa[i], xb[i]);
        x[N] = addc(xa[N], xb[N]);
    } else {
        // which magnitude is larger?
        bool a_big;
        if (d != 0) a_big = d > 0;
        else if (a.m[N - 1] != b.m[N - 1]) a_big = a.m[N - 1] > b.m[N - 1];
        else {                              // rare: same exponent, same top limb
            a_big = true;
            bool decided = false;
            MDZ_UNROLL
            for (int i = N - 2; i >= 0; --i)
                if (!decided && a.m[i] != b.m[i]) { a_big = a.m[i] > b.m[i]; decided = true; }
        }
        const uint32_t ma = a_big ? 0u : 0xffffffffu;
        const uint32_t mb = ~ma;
        s = a_big ? sa : sb;
        // big - small - sticky  ==  big + ~small + (1 - sticky)
        const uint32_t cin = sticky ^ 1u;
        (void)add_cc(cin, 0xffffffffu);                    // CC = cin
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = addc_cc(xa[i] ^ ma, xb[i] ^ mb);
        x[N] = addc(xa[N] ^ ma, xb[N] ^ mb);
    }

    // normalise
    if (x[N] == 0) {                        // rare: >= 31 bits cancelled (exact, no sticky)
        uint32_t any = 0;
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) any |= x[i];
        if (any == 0) { set_zero(r); return; }
        while (x[N] == 0) {
            MDZ_UNROLL
            for (int i = N; i >= 1; --i) x[i] = x[i - 1];
            x[0] = 0;
            e -= 32;
        }
    }
    const uint32_t lz = (uint32_t)clz32(x[N]);
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(x[i - 1], x[i], lz);
    x[0] <<= lz;
    e -= (int32_t)lz;

    const uint32_t cout = round_rn<N>(x, sticky, rc);
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = x[i + 1];
    r.e = e + (int32_t)cout;
    r.s = s;
}

// a > 4 ?   (4 = 0.1b * 2^3)
template <int N>
MDZ_HD bool greater_than_4(const Num<N>& a)
{
    if (a.m[N - 1] == 0 || a.s) return false;
    if (a.e != 3) return a.e > 3;
    uint32_t low = a.m[N - 1] ^ 0x80000000u;
    MDZ_UNROLL
    for (int i = 0; i < N - 1; ++i) low |= a.m[i];
    return low != 0;
}

}  // namespace mdz
