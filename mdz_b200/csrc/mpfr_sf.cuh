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

// Round x[1..N] (guard limb x[0], sticky flag below it) to p bits, RN-even,
// propagating the increment through every limb (general version, used by the
// out-of-line slow paths).  Returns 1 if the increment carried out of the top
// limb (mantissa is then 0x80000000:0...; the caller bumps the exponent).
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

// Fast-path rounding: only limb 1 is touched.  Returns true when the increment
// carries out of limb 1 (probability ~2^-32 per operation on generic data);
// the caller then redoes the operation on its general out-of-line path.
template <int N>
MDZ_HD bool round_rn_fast(uint32_t (&x)[N + 1], uint32_t sticky, const RoundCfg& rc)
{
    const uint32_t x0 = x[1], g = x[0];
    const uint32_t rb  = (x0 & rc.half_x0) | (g & rc.half_g);
    const uint32_t st  = sticky | (x0 & rc.below_x0) | (g & rc.below_g);
    const uint32_t lsb = x0 & rc.ulp;
    const uint32_t inc = (rb != 0 && (st | lsb) != 0) ? rc.ulp : 0u;
    const uint32_t lo = (x0 & rc.keep) + inc;
    x[1] = lo;
    return lo < inc;
}

template <int N> struct Wide { uint32_t w[2 * N]; };
template <int N> struct Ext  { uint32_t w[N + 1]; };

// r = RN(a * b) from a full product in prod[2N]; e = ea + eb.  General version.
template <int N>
MDZ_HD void finish_product(const uint32_t (&prod)[2 * N], int32_t e, uint32_t s,
                           Num<N>& r, const RoundCfg& rc, bool fast, bool& bail)
{
    // the product of two normalised significands has its top bit at 64N-1 or 64N-2
    const uint32_t sh = (prod[2 * N - 1] >> 31) ^ 1u;       // 1 -> shift left one bit
    uint32_t x[N + 1];
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(prod[N - 2 + i], prod[N - 1 + i], sh);
    const uint32_t lowtop = prod[N - 2];
    x[0] = fsl(lowtop, prod[N - 1], sh);
    uint32_t sticky = lowtop << sh;                          // bits of prod[N-2] left below the guard
    MDZ_UNROLL
    for (int i = 0; i + 2 < N; ++i) sticky |= prod[i];
    uint32_t cout = 0;
    if (fast) bail = round_rn_fast<N>(x, sticky, rc);
    else      cout = round_rn<N>(x, sticky, rc);
    e = e - (int32_t)sh + (int32_t)cout;
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = x[i + 1];
    // a zero operand gives an all-zero product; an exponent below E_MIN is kept
    // as "tiny non-zero" (mantissa intact, e = E_ZERO): it is shifted out of every
    // sum as a pure sticky bit, which is what a 2^-2^28 value does anyway
    r.e = (x[N] != 0 && e >= E_MIN) ? e : E_ZERO;
    r.s = s;
}

#if defined(MDZ_HOST_EMU)
#define MDZ_NOINLINE_DEV inline
#else
#define MDZ_NOINLINE_DEV __device__ __noinline__
#endif

// out-of-line general versions (by value, so the callers' registers never have
// their address taken)
template <int N>
MDZ_NOINLINE_DEV Num<N> fmul_general(Num<N> a, Num<N> b, RoundCfg rc)
{
    Num<N> r; bool dummy = false;
    uint32_t prod[2 * N];
    mul_full<N>(a.m, b.m, prod);
    finish_product<N>(prod, a.e + b.e, a.s ^ b.s, r, rc, false, dummy);
    return r;
}

// Fast path: round from the high part of the product (mul_hi / sqr_hi).  The
// true product lies in [T, T + N) units of the guard limb's last bit, so the
// round-to-nearest result is decided unless the bits below the rounding
// position are within 2N+4 of wrapping or of zero (a possible tie); then, and
// when the increment would carry out of the lowest limb, the caller redoes the
// operation with the full product.  Probability ~N*2^-29 per multiply.
template <int N>
MDZ_HD bool finish_high(const uint32_t (&t)[N + 2], int32_t e, uint32_t s,
                        Num<N>& r, const RoundCfg& rc)
{
    const uint32_t sh = (t[N + 1] >> 31) ^ 1u;
    uint32_t x[N + 1];
    MDZ_UNROLL
    for (int i = N; i >= 0; --i) x[i] = fsl(t[i], t[i + 1], sh);
    const uint32_t x0 = x[1], g = x[0];
    const uint32_t rb = (x0 & rc.half_x0) | (g & rc.half_g);
    const uint32_t gb = g & rc.below_g;
    constexpr uint32_t c = 2u * N + 4u;
    bool bail = ((gb + c) & rc.below_g) <= 2u * c;
    const uint32_t inc = rb != 0 ? rc.ulp : 0u;             // inexact for sure: sticky = 1
    const uint32_t lo = (x0 & rc.keep) + inc;
    bail = bail || (lo < inc);
    r.m[0] = lo;
    MDZ_UNROLL
    for (int i = 1; i < N; ++i) r.m[i] = x[i + 1];
    e -= (int32_t)sh;
    r.e = e >= E_MIN ? e : E_ZERO;
    r.s = s;
    return bail;
}

template <int N>
MDZ_HD void fmul(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc)
{
    uint32_t t[N + 2];
    mul_hi<N>(a.m, b.m, t);
    if (finish_high<N>(t, a.e + b.e, a.s ^ b.s, r, rc)) { MDZ_COUNT(CNT_MUL_BAIL); r = fmul_general<N>(a, b, rc); }
}

template <int N>
MDZ_HD void fsqr(const Num<N>& a, Num<N>& r, const RoundCfg& rc)
{
    uint32_t t[N + 2];
    sqr_hi<N>(a.m, t);
    if (finish_high<N>(t, a.e + a.e, 0u, r, rc)) { MDZ_COUNT(CNT_MUL_BAIL); Num<N> f = fmul_general<N>(a, a, rc); f.s = 0; r = f; }
}

// ---------------------------------------------------------------------------
// r = RN(a + b)            MODE_GENERIC  (signs as given)
// r = RN(a - b), a,b >= 0  MODE_SUB_POS
// r = RN(a + b), a,b >= 0  MODE_ADD_POS
//
// Both operands get a guard limb below and are pre-shifted right by one bit so
// that a same-sign sum cannot carry out; the one with the smaller exponent is
// shifted further by the exponent gap.  For an effective subtraction the
// smaller operand is complemented, so the difference is never negative.
// Normalisation is one left funnel shift by clz (0..1 after an addition, >= 1
// after a subtraction).
//
// Everything a warp meets with more than negligible probability is inline and
// shares one instruction stream (a per-lane event that happens 3% of the time
// happens in most warps): the bit part of the alignment is a funnel shift for
// every lane; a gap of 31 bits or more adds a whole-limb shift through a
// per-thread shared-memory column (sticky from what drops off); 31..62
// cancelled bits add an in-place limb move; equal exponents and top limbs add
// one borrow chain to order the operands.  Only deeper cancellation / an exact
// zero and a rounding increment that carries out of the lowest limb leave the
// instruction stream (by-value, out-of-line, ~2^-32 per operation).
// ---------------------------------------------------------------------------
enum { MODE_GENERIC = 0, MODE_SUB_POS = 1, MODE_ADD_POS = 2 };

// exact sum/difference x (guard limb x.w[0], no sticky) with top limb zero:
// normalise by whole limbs, then bits, then round.  Out of line, by value.
template <int N>
MDZ_NOINLINE_DEV Num<N> fadd_finish_general(Ext<N> xe, int32_t e, uint32_t s, RoundCfg rc)
{
    Num<N> r;
    uint32_t x[N + 1];
    MDZ_UNROLL
    for (int i = 0; i <= N; ++i) x[i] = xe.w[i];
    uint32_t any = 0;
    MDZ_UNROLL
    for (int i = 0; i <= N; ++i) any |= x[i];
    if (any == 0) { set_zero(r); return r; }
    while (x[N] == 0) {
        MDZ_UNROLL
        for (int i = N; i >= 1; --i) x[i] = x[i - 1];
        x[0] = 0;
        e -= 32;
    }
    const uint32_t lz = (uint32_t)clz32(x[N]);
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(x[i - 1], x[i], lz);
    x[0] <<= lz;
    e -= (int32_t)lz;
    const uint32_t cout = round_rn<N>(x, 0u, rc);
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = x[i + 1];
    r.e = e + (int32_t)cout;
    r.s = s;
    return r;
}

// the rounding increment wrapped limb 0: carry on through the upper limbs
template <int N>
MDZ_NOINLINE_DEV Num<N> round_carry_general(Num<N> r)
{
    r.m[1] = add_cc(r.m[1], 1u);
    MDZ_UNROLL
    for (int i = 2; i < N; ++i) r.m[i] = addc_cc(r.m[i], 0u);
    const uint32_t cout = addc(0u, 0u);
    r.m[N - 1] |= cout << 31;
    r.e += (int32_t)cout;
    return r;
}

// Per-thread scratch column in shared memory used as a limb shifter: words
// [0, N) are written on demand, words [N, 2N+2) stay zero for the whole kernel.
// Limb k of a thread lives at scratch[k * kScratchStride] (conflict-free).
#if defined(MDZ_HOST_EMU)
constexpr int kScratchStride = 1;
#else
constexpr int kScratchStride = 128;     // == kBlock (escape_kernel.cuh)
#endif
template <int N> struct ScratchWords { static constexpr int value = 2 * N + 2; };

// x (N+1 limbs, x[0] = guard) >>= 32*q for 1 <= q <= N+1, through the scratch
// column; the limbs that fall off the bottom are OR-ed into sticky.
template <int N>
MDZ_HD void limb_shift_down(uint32_t (&x)[N + 1], uint32_t q, uint32_t* scratch, uint32_t& sticky)
{
    MDZ_UNROLL
    for (int k = 0; k <= N; ++k) scratch[k * kScratchStride] = x[k];
    const uint32_t* src = scratch + q * kScratchStride;
    MDZ_UNROLL
    for (int i = 0; i <= N; ++i) x[i] = src[i * kScratchStride];
    uint32_t st = 0;
#if !defined(MDZ_HOST_EMU)
#pragma unroll 1
#endif
    for (uint32_t k = 0; k < q; ++k) st |= scratch[k * kScratchStride];
    sticky |= st;
}

template <int N, int MODE>
MDZ_HD void fadd(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc, uint32_t* scratch)
{
    const uint32_t sb = (MODE == MODE_SUB_POS) ? 1u : (MODE == MODE_ADD_POS ? 0u : b.s);
    const uint32_t sa = (MODE == MODE_GENERIC) ? a.s : 0u;
    const bool sub = (MODE == MODE_SUB_POS) ? true : (MODE == MODE_ADD_POS ? false : (sa != sb));

    const int32_t d = a.e - b.e;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    if (ad >= 32u * N + 2u) {
        // the smaller operand (or a zero) lies wholly below the rounding position:
        // RN(big +- tiny) == big  (gap >= p + 2)
        MDZ_COUNT(CNT_ADD_COPY);
        const bool a_is_big = d > 0;
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) r.m[i] = a_is_big ? a.m[i] : b.m[i];
        r.e = a_is_big ? a.e : b.e;
        r.s = a_is_big ? sa : sb;
        return;
    }
    // which magnitude is larger (only matters under subtraction)
    bool a_big = d > 0 || (d == 0 && a.m[N - 1] > b.m[N - 1]);
    if (MODE != MODE_ADD_POS) {
        if (sub && d == 0 && a.m[N - 1] == b.m[N - 1]) {
            // same exponent, same top limb: compare the rest (borrow of a - b)
            MDZ_COUNT(CNT_ADD_TIE);
            (void)sub_cc(a.m[0], b.m[0]);
            MDZ_UNROLL
            for (int i = 1; i < N - 1; ++i) (void)subc_cc(a.m[i], b.m[i]);
            a_big = subc(0u, 0u) == 0u;
        }
    }
    // one bit of headroom for both, the exponent gap for the smaller one:
    // bit part by funnel shift (the guard limb catches what leaves limb 0) ...
    const uint32_t sh = ad + 1u, q = sh >> 5, rr = sh & 31u;
    const uint32_t sha = d < 0 ? rr : 1u;
    const uint32_t shb = d > 0 ? rr : 1u;
    uint32_t xa[N + 1], xb[N + 1];
    xa[0] = fsr(0u, a.m[0], sha); xb[0] = fsr(0u, b.m[0], shb);
    MDZ_UNROLL
    for (int i = 1; i < N; ++i) { xa[i] = fsr(a.m[i - 1], a.m[i], sha); xb[i] = fsr(b.m[i - 1], b.m[i], shb); }
    xa[N] = a.m[N - 1] >> sha; xb[N] = b.m[N - 1] >> shb;
    // ... whole limbs (gap of 31 bits or more) through the shared-memory column
    uint32_t sticky = 0;
    if (q != 0) {
        MDZ_COUNT(CNT_ADD_MEDIUM);
        if (d > 0) limb_shift_down<N>(xb, q, scratch, sticky);
        else       limb_shift_down<N>(xa, q, scratch, sticky);
        sticky = sticky != 0 ? 1u : 0u;
    } else {
        MDZ_COUNT(CNT_ADD_FAST);
    }

    int32_t e = (d < 0 ? b.e : a.e) + 1;
    uint32_t s = sa;
    uint32_t x[N + 1];
    if (MODE == MODE_ADD_POS) {
        x[0] = add_cc(xa[0], xb[0]);
        MDZ_UNROLL
        for (int i = 1; i < N; ++i) x[i] = addc_cc(xa[i], xb[i]);
        x[N] = addc(xa[N], xb[N]);
    } else {
        // masks: complement the smaller operand under subtraction
        const uint32_t ma = (sub && !a_big) ? 0xffffffffu : 0u;
        const uint32_t mb = (sub && a_big) ? 0xffffffffu : 0u;
        if (sub && !a_big) s = sb;
        // big - small - sticky == big + ~small + (1 - sticky)
        (void)add_cc(sub ? (sticky ^ 1u) : 0u, 0xffffffffu);
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = addc_cc(xa[i] ^ ma, xb[i] ^ mb);
        x[N] = addc(xa[N] ^ ma, xb[N] ^ mb);
        if (x[N] == 0) {
            // 31..62 bits cancelled (exact): move up one limb in place
            MDZ_COUNT(CNT_ADD_CANCEL);
            MDZ_UNROLL
            for (int i = N; i >= 1; --i) x[i] = x[i - 1];
            x[0] = 0;
            e -= 32;
            if (x[N] == 0) {                // rare: deeper cancellation or an exact zero
                Ext<N> xe;
                MDZ_UNROLL
                for (int i = 0; i <= N; ++i) xe.w[i] = x[i];
                r = fadd_finish_general<N>(xe, e, s, rc);
                return;
            }
        }
    }
    const uint32_t lz = (uint32_t)clz32(x[N]);
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(x[i - 1], x[i], lz);
    x[0] <<= lz;
    e -= (int32_t)lz;
    const bool carry = round_rn_fast<N>(x, sticky, rc);
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = x[i + 1];
    r.e = e;
    r.s = s;
    if (carry) { MDZ_COUNT(CNT_ROUND_CARRY); r = round_carry_general<N>(r); }   // rare: increment leaves limb 0
}


// ---------------------------------------------------------------------------
// Speculative, branch-free versions: no rare-case handling at all, only a flag.
// A whole iteration built from these is one basic block, so ptxas can overlap
// the IMAD.WIDE chains of a product with the shift/add chains of an independent
// sum.  When `rare` comes back non-zero the caller discards the results.
// ---------------------------------------------------------------------------
template <int N, int MODE>
MDZ_HD void fadd_spec(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc, uint32_t& rare)
{
    const uint32_t sb = (MODE == MODE_SUB_POS) ? 1u : (MODE == MODE_ADD_POS ? 0u : b.s);
    const uint32_t sa = (MODE == MODE_GENERIC) ? a.s : 0u;
    const bool sub = (MODE == MODE_SUB_POS) ? true : (MODE == MODE_ADD_POS ? false : (sa != sb));
    const int32_t d = a.e - b.e;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    rare |= (ad >= 31u) ? 1u : 0u;
    if (ad >= 31u) MDZ_COUNT(CNT_SPEC_GAP);
    const uint32_t sha = 1u + (d < 0 ? ad : 0u);
    const uint32_t shb = 1u + (d > 0 ? ad : 0u);
    uint32_t xa[N + 1], xb[N + 1];
    xa[0] = fsr(0u, a.m[0], sha); xb[0] = fsr(0u, b.m[0], shb);
    MDZ_UNROLL
    for (int i = 1; i < N; ++i) { xa[i] = fsr(a.m[i - 1], a.m[i], sha); xb[i] = fsr(b.m[i - 1], b.m[i], shb); }
    xa[N] = a.m[N - 1] >> sha; xb[N] = b.m[N - 1] >> shb;
    uint32_t x[N + 1];
    uint32_t s = sa;
    if (MODE == MODE_ADD_POS) {
        x[0] = add_cc(xa[0], xb[0]);
        MDZ_UNROLL
        for (int i = 1; i < N; ++i) x[i] = addc_cc(xa[i], xb[i]);
        x[N] = addc(xa[N], xb[N]);
    } else {
        const bool a_big = d > 0 || (d == 0 && a.m[N - 1] > b.m[N - 1]);
        rare |= (sub && d == 0 && a.m[N - 1] == b.m[N - 1]) ? 1u : 0u;
        if (sub && d == 0 && a.m[N - 1] == b.m[N - 1]) MDZ_COUNT(CNT_SPEC_TIE);
        const uint32_t ma = (sub && !a_big) ? 0xffffffffu : 0u;
        const uint32_t mb = (sub && a_big) ? 0xffffffffu : 0u;
        s = (sub && !a_big) ? sb : sa;
        (void)add_cc(sub ? 1u : 0u, 0xffffffffu);
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = addc_cc(xa[i] ^ ma, xb[i] ^ mb);
        x[N] = addc(xa[N] ^ ma, xb[N] ^ mb);
        rare |= (x[N] == 0) ? 1u : 0u;
        if (x[N] == 0) MDZ_COUNT(CNT_SPEC_CANCEL);
    }
    const uint32_t lz = (uint32_t)clz32(x[N]);
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(x[i - 1], x[i], lz);
    x[0] <<= lz;
    rare |= round_rn_fast<N>(x, 0u, rc) ? 1u : 0u;
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = x[i + 1];
    r.e = (d < 0 ? b.e : a.e) + 1 - (int32_t)lz;
    r.s = s;
}

// Wide-gap variant: exponent gaps up to 126 bits (three whole limbs plus 30 bits), still
// branch-free.  The operands are first ordered by exponent (selects), so only the smaller
// one goes through the shifter -- two conditional limb moves (by one and by two limbs) and a
// funnel shift -- and only it is complemented for a subtraction; what falls off the guard
// limb is the sticky bit, which enters the difference as a borrow exactly as in fadd.  About
// 3N+1 instructions more per addition than fadd_spec, so the kernel switches to it per warp,
// when orbits keep meeting gaps of 31 bits and more (views hugging an axis: in
// gallery/deep_embedded_julia.mdz wre^2 - wim^2 has a 31..40-bit gap every other iteration).
template <int N, int MODE>
MDZ_HD void fadd_spec_wide(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc, uint32_t& rare)
{
    const uint32_t sb = (MODE == MODE_SUB_POS) ? 1u : (MODE == MODE_ADD_POS ? 0u : b.s);
    const uint32_t sa = (MODE == MODE_GENERIC) ? a.s : 0u;
    const bool sub = (MODE == MODE_SUB_POS) ? true : (MODE == MODE_ADD_POS ? false : (sa != sb));
    const int32_t d = a.e - b.e;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    rare |= (ad > 126u) ? 1u : 0u;
    if (ad > 126u) MDZ_COUNT(CNT_SPEC_GAP);
    // order by magnitude: exponent, then the whole significand (one borrow chain), so that a
    // difference is never negative -- also when the top limbs are equal
    bool a_big = d >= 0;
    if (MODE != MODE_ADD_POS) {
        (void)sub_cc(a.m[0], b.m[0]);
        MDZ_UNROLL
        for (int i = 1; i < N; ++i) (void)subc_cc(a.m[i], b.m[i]);
        const bool a_ge_b = subc(0u, 0u) == 0u;
        a_big = d > 0 || (d == 0 && a_ge_b);
    }
    uint32_t big[N], y[N + 1];
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) { big[i] = a_big ? a.m[i] : b.m[i]; y[i + 1] = a_big ? b.m[i] : a.m[i]; }
    y[0] = 0u;                                       // guard limb
    const int32_t e_big = a_big ? a.e : b.e;
    const uint32_t s_big = a_big ? sa : sb;
    // small >>= 32*q + rr, with sh = ad + 1 (one bit of headroom, as for the big one)
    const uint32_t sh = ad + 1u, q = sh >> 5, rr = sh & 31u;
    uint32_t sticky = 0;
    {
        const bool q1 = (q & 1u) != 0;               // by one limb: drops the (empty) guard
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) y[i] = q1 ? y[i + 1] : y[i];
        y[N] = q1 ? 0u : y[N];
        const bool q2 = (q & 2u) != 0;               // by two limbs
        sticky = q2 ? (y[0] | y[1]) : 0u;
        MDZ_UNROLL
        for (int i = 0; i + 2 <= N; ++i) y[i] = q2 ? y[i + 2] : y[i];
        if (N >= 1) y[N - 1] = q2 ? 0u : y[N - 1];
        y[N] = q2 ? 0u : y[N];
    }
    sticky |= fsr(0u, y[0], rr);                     // the bits of y[0] that the funnel shift drops
    uint32_t xs[N + 1], xb[N + 1];
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) xs[i] = fsr(y[i], y[i + 1], rr);
    xs[N] = y[N] >> rr;
    xb[0] = big[0] << 31;
    MDZ_UNROLL
    for (int i = 1; i < N; ++i) xb[i] = fsr(big[i - 1], big[i], 1u);
    xb[N] = big[N - 1] >> 1;
    sticky = sticky != 0 ? 1u : 0u;
    uint32_t x[N + 1];
    if (MODE == MODE_ADD_POS) {
        x[0] = add_cc(xb[0], xs[0]);
        MDZ_UNROLL
        for (int i = 1; i < N; ++i) x[i] = addc_cc(xb[i], xs[i]);
        x[N] = addc(xb[N], xs[N]);
    } else {
        const uint32_t ms = sub ? 0xffffffffu : 0u;
        // big - small - sticky == big + ~small + (1 - sticky)
        (void)add_cc(sub ? (sticky ^ 1u) : 0u, 0xffffffffu);
        MDZ_UNROLL
        for (int i = 0; i < N; ++i) x[i] = addc_cc(xb[i], xs[i] ^ ms);
        x[N] = addc(xb[N], xs[N] ^ ms);
    }
    int32_t e_out = e_big + 1;
    if (MODE != MODE_ADD_POS) {
        // 31..62 cancelled bits (exact: the gap is at most one bit, nothing was dropped):
        // move up one limb, by selects.  Deeper cancellation or an exact zero stays rare.
        const bool z = x[N] == 0;
        MDZ_UNROLL
        for (int i = N; i >= 1; --i) x[i] = z ? x[i - 1] : x[i];
        x[0] = z ? 0u : x[0];
        e_out -= z ? 32 : 0;
        rare |= (x[N] == 0) ? 1u : 0u;
        if (x[N] == 0) MDZ_COUNT(CNT_SPEC_CANCEL);
    }
    const uint32_t lz = (uint32_t)clz32(x[N]);
    MDZ_UNROLL
    for (int i = N; i >= 1; --i) x[i] = fsl(x[i - 1], x[i], lz);
    x[0] <<= lz;
    rare |= round_rn_fast<N>(x, sticky, rc) ? 1u : 0u;
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) r.m[i] = x[i + 1];
    r.e = e_out - (int32_t)lz;
    r.s = s_big;
}

template <int N>
MDZ_HD void fmul_spec(const Num<N>& a, const Num<N>& b, Num<N>& r, const RoundCfg& rc, uint32_t& rare)
{
    uint32_t t[N + 2];
    mul_hi<N>(a.m, b.m, t);
    rare |= finish_high<N>(t, a.e + b.e, a.s ^ b.s, r, rc) ? 1u : 0u;
}

template <int N>
MDZ_HD void fsqr_spec(const Num<N>& a, Num<N>& r, const RoundCfg& rc, uint32_t& rare)
{
    uint32_t t[N + 2];
    sqr_hi<N>(a.m, t);
    rare |= finish_high<N>(t, a.e + a.e, 0u, r, rc) ? 1u : 0u;
}

// Escape test RN(a + b) > 4 for a, b >= 0 with exponents below 4, decided from the top limbs
// where that is safe: in units of 2^-28 a value is less than floor(m_top * 2^(e-4)) + 1, so
//   floors + 2 <= 2^30   ->  a + b < 4: no escape whatever the rounding;
//   floors     >  2^30   ->  a + b >= 4 + 2^-28, which rounds above 4 at any precision >= 33;
// returns -1 / +1 for those, 0 when the sum is within 2^-27 of 4 and has to be formed.
// Orbits that hover just below |z|^2 = 4 (gallery/deep_embedded_julia.mdz spends three quarters
// of its iterations with a square in [2, 4)) otherwise pay a fourth addition per iteration.
template <int N>
MDZ_HD int escape_precheck(const Num<N>& a, const Num<N>& b)
{
    const uint32_t sa = (uint32_t)(4 - a.e), sb = (uint32_t)(4 - b.e);       // >= 1 here
    const uint32_t fa = sa < 32u ? a.m[N - 1] >> sa : 0u;
    const uint32_t fb = sb < 32u ? b.m[N - 1] >> sb : 0u;
    const uint32_t f = fa + fb;                                             // < 2^32: each is below 2^31
    if (f <= (1u << 30) - 2u) return -1;
    if (f > (1u << 30)) return 1;
    return 0;
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
