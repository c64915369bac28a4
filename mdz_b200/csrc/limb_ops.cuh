// limb_ops.cuh -- 32-bit limb primitives for the multi-limb significand core.
//
// On the device every primitive is one PTX instruction that takes part in a
// carry chain (add.cc / addc.cc / mad.lo.cc / madc.hi.cc ...).  ptxas for
// sm_100a fuses each `mad(c).lo.cc ; madc.hi.cc` pair into a single
// IMAD.WIDE.U32(.X) with predicate carry-in/out, so a product row is N/2
// back-to-back IMAD.WIDE.U32.X on the FMA pipe (checked with cuobjdump -sass).
// All carry-chain asm is `volatile` so NVVM keeps the chains in program order;
// the condition-code register is virtual in PTX and ptxas tracks it.
//
// With MDZ_HOST_EMU defined (tests/host_emu only, never the product) the same
// names are plain C++ with an emulated carry flag, so the limb algorithms can
// be differential-tested against libmpfr / libgmp on a box with no GPU.
#pragma once
#include <stdint.h>

#if defined(MDZ_HOST_EMU)
#define MDZ_HD inline
#define MDZ_UNROLL
#else
#define MDZ_HD __device__ __forceinline__
#define MDZ_UNROLL _Pragma("unroll")
#endif

namespace mdz {

// path counters: compiled in for the host emulation only (tests/host_emu)
#if defined(MDZ_HOST_EMU)
enum { CNT_ADD_FAST, CNT_ADD_MEDIUM, CNT_ADD_COPY, CNT_ADD_TIE, CNT_ADD_CANCEL, CNT_ROUND_CARRY,
       CNT_MUL_BAIL, CNT_ESC_ADD, CNT_ITER, CNT_SPEC_FALLBACK, CNT_SPEC_GAP, CNT_SPEC_TIE, CNT_SPEC_CANCEL, CNT_SPEC_ROUND, CNT_SPEC_MUL, CNT_N };
static thread_local unsigned long long g_counts[CNT_N];
#define MDZ_COUNT(id) (++::mdz::g_counts[::mdz::id])
#else
#define MDZ_COUNT(id) ((void)0)
#endif

#if defined(MDZ_HOST_EMU)
static thread_local uint32_t g_cc = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b)
{ uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b)
{ uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b)
{ uint64_t t = (uint64_t)a + b + g_cc; return (uint32_t)t; }
inline uint32_t sub_cc(uint32_t a, uint32_t b)
{ uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b)
{ uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b)
{ uint64_t t = (uint64_t)a - b - g_cc; return (uint32_t)t; }
// d(lo,hi) += a*b as one 64-bit multiply-accumulate, starting a chain
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
    uint64_t p = (uint64_t)a * b;
    uint64_t t = (uint64_t)lo + (uint32_t)p; lo = (uint32_t)t;
    uint64_t u = (uint64_t)hi + (uint32_t)(p >> 32) + (t >> 32); hi = (uint32_t)u;
    g_cc = (uint32_t)(u >> 32);
}
// ... continuing a chain (carry in and out)
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
    uint64_t p = (uint64_t)a * b;
    uint64_t t = (uint64_t)lo + (uint32_t)p + g_cc; lo = (uint32_t)t;
    uint64_t u = (uint64_t)hi + (uint32_t)(p >> 32) + (t >> 32); hi = (uint32_t)u;
    g_cc = (uint32_t)(u >> 32);
}
inline uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s)   // (hi:lo) >> (s&31), low word
{ s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
inline uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s)   // (hi:lo) << (s&31), high word
{ s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#else
MDZ_HD uint32_t add_cc(uint32_t a, uint32_t b)
{ uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MDZ_HD uint32_t addc_cc(uint32_t a, uint32_t b)
{ uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MDZ_HD uint32_t addc(uint32_t a, uint32_t b)
{ uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MDZ_HD uint32_t sub_cc(uint32_t a, uint32_t b)
{ uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MDZ_HD uint32_t subc_cc(uint32_t a, uint32_t b)
{ uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MDZ_HD uint32_t subc(uint32_t a, uint32_t b)
{ uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MDZ_HD void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
MDZ_HD void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
MDZ_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
MDZ_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
MDZ_HD int clz32(uint32_t x) { return __clz((int)x); }
#endif

// ---------------------------------------------------------------------------
// Full 2N-limb product r = a * b (schoolbook, N*N IMAD.WIDE).
//
// Two accumulator arrays keep every 64-bit accumulator register pair aligned:
// e[x] holds limb position x for pairs starting at even positions, o[x] holds
// limb position x+1 for pairs starting at odd positions.  Row i (multiplier
// limb b[i]) runs two independent carry chains, one over the even-indexed
// limbs of a and one over the odd-indexed ones; a[j]*b[i] lands at position
// i+j, whose parity picks the array.  The carry that leaves a chain is added
// into the next slot of the same array, which no pair has touched yet (it can
// only hold earlier chain carries), so the single addc cannot overflow.
// ---------------------------------------------------------------------------
// T: the limb type -- uint32_t, or (host emulation of the warp-cooperative code, coop_ops.cuh) a vector of 32 lanes
template <int N, class T = uint32_t>
MDZ_HD void mul_full(const T (&a)[N], const T (&b)[N], T (&r)[2 * N])
{
    T e[2 * N + 2], o[2 * N + 2];
    MDZ_UNROLL
    for (int i = 0; i < 2 * N + 2; ++i) { e[i] = T(0u); o[i] = T(0u); }
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) {
        MDZ_UNROLL
        for (int c = 0; c < 2; ++c) {
            if (c >= N) continue;
            int last = 0;
            MDZ_UNROLL
            for (int j = c; j < N; j += 2) {
                const int pos = i + j;
                if ((pos & 1) == 0) {
                    if (j == c) mad_wide_cc(e[pos], e[pos + 1], a[j], b[i]);
                    else        madc_wide_cc(e[pos], e[pos + 1], a[j], b[i]);
                } else {
                    if (j == c) mad_wide_cc(o[pos - 1], o[pos], a[j], b[i]);
                    else        madc_wide_cc(o[pos - 1], o[pos], a[j], b[i]);
                }
                last = pos;
            }
            const int cp = last + 2;            // where the chain's carry lands
            if (cp < 2 * N) {
                if ((cp & 1) == 0) e[cp] = addc(e[cp], T(0u));
                else               o[cp - 1] = addc(o[cp - 1], T(0u));
            }
        }
    }
    r[0] = e[0];
    r[1] = add_cc(e[1], o[0]);
    MDZ_UNROLL
    for (int i = 2; i < 2 * N - 1; ++i) r[i] = addc_cc(e[i], o[i - 1]);
    r[2 * N - 1] = addc(e[2 * N - 1], o[2 * N - 2]);
}

// ---------------------------------------------------------------------------
// Full 2N-limb square r = a * a with N(N+1)/2 IMAD.WIDE: cross products
// a[i]*a[j] (i<j) accumulated with the same even/odd scheme, doubled by a
// one-bit funnel shift, then the diagonal a[i]^2 added in one carry chain.
// ---------------------------------------------------------------------------
template <int N>
MDZ_HD void sqr_full(const uint32_t (&a)[N], uint32_t (&r)[2 * N])
{
    uint32_t e[2 * N + 2], o[2 * N + 2];
    MDZ_UNROLL
    for (int i = 0; i < 2 * N + 2; ++i) { e[i] = 0; o[i] = 0; }
    MDZ_UNROLL
    for (int i = 0; i < N - 1; ++i) {
        MDZ_UNROLL
        for (int c = 1; c <= 2; ++c) {
            if (i + c >= N) continue;
            int last = 0;
            MDZ_UNROLL
            for (int j = i + c; j < N; j += 2) {
                const int pos = i + j;
                if ((pos & 1) == 0) {
                    if (j == i + c) mad_wide_cc(e[pos], e[pos + 1], a[j], a[i]);
                    else            madc_wide_cc(e[pos], e[pos + 1], a[j], a[i]);
                } else {
                    if (j == i + c) mad_wide_cc(o[pos - 1], o[pos], a[j], a[i]);
                    else            madc_wide_cc(o[pos - 1], o[pos], a[j], a[i]);
                }
                last = pos;
            }
            const int cp = last + 2;
            if (cp < 2 * N) {
                if ((cp & 1) == 0) e[cp] = addc(e[cp], 0u);
                else               o[cp - 1] = addc(o[cp - 1], 0u);
            }
        }
    }
    // cross = e + (o << 32)
    uint32_t x[2 * N];
    x[0] = e[0];
    x[1] = add_cc(e[1], o[0]);
    MDZ_UNROLL
    for (int i = 2; i < 2 * N - 1; ++i) x[i] = addc_cc(e[i], o[i - 1]);
    x[2 * N - 1] = addc(e[2 * N - 1], o[2 * N - 2]);
    // r = 2*cross
    MDZ_UNROLL
    for (int i = 2 * N - 1; i >= 1; --i) r[i] = fsl(x[i - 1], x[i], 1);
    r[0] = x[0] << 1;
    // r += sum a[i]^2 << 64i : one chain of N IMAD.WIDE
    mad_wide_cc(r[0], r[1], a[0], a[0]);
    MDZ_UNROLL
    for (int i = 1; i < N; ++i) madc_wide_cc(r[2 * i], r[2 * i + 1], a[i], a[i]);
}


// ---------------------------------------------------------------------------
// High part of the product: t[0..N+1] = sum over i+j >= N-2 of
// a[i]*b[j] * 2^(32*(i+j-(N-2))), i.e. limb positions N-2 .. 2N-1 of the full
// product without the carries from the columns below.  N(N+1)/2 + 2N - 1
// IMAD.WIDE instead of N^2.  The omitted columns sum to less than
// (N-1) * 2^(32*(N-1)): fewer than N units of t[1]'s least significant bit.
// MPFR itself uses such a "mulhigh" and falls back to the full product when the
// rounding cannot be decided; so does the caller here (mpfr_sf.cuh).
// Same even/odd accumulator scheme as mul_full, indexed by q = i+j-(N-2).
// ---------------------------------------------------------------------------
template <int N>
MDZ_HD void mul_hi(const uint32_t (&a)[N], const uint32_t (&b)[N], uint32_t (&t)[N + 2])
{
    uint32_t e[N + 4], o[N + 4];
    MDZ_UNROLL
    for (int i = 0; i < N + 4; ++i) { e[i] = 0; o[i] = 0; }
    MDZ_UNROLL
    for (int i = 0; i < N; ++i) {
        const int jmin = (N - 2 - i) > 0 ? (N - 2 - i) : 0;
        MDZ_UNROLL
        for (int c = 0; c < 2; ++c) {
            if (jmin + c >= N) continue;
            int last = 0;
            MDZ_UNROLL
            for (int j = jmin + c; j < N; j += 2) {
                const int q = i + j - (N - 2);
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
            if (cp < N + 2) {
                if ((cp & 1) == 0) e[cp] = addc(e[cp], 0u);
                else               o[cp - 1] = addc(o[cp - 1], 0u);
            }
        }
    }
    t[0] = e[0];
    t[1] = add_cc(e[1], o[0]);
    MDZ_UNROLL
    for (int i = 2; i < N + 1; ++i) t[i] = addc_cc(e[i], o[i - 1]);
    t[N + 1] = addc(e[N + 1], o[N]);
}

// High part of the square, same contract as mul_hi.
template <int N>
MDZ_HD void sqr_hi(const uint32_t (&a)[N], uint32_t (&t)[N + 2])
{
    uint32_t e[N + 4], o[N + 4];
    MDZ_UNROLL
    for (int i = 0; i < N + 4; ++i) { e[i] = 0; o[i] = 0; }
    MDZ_UNROLL
    for (int i = 0; i < N - 1; ++i) {
        const int jmin = (N - 2 - i) > (i + 1) ? (N - 2 - i) : (i + 1);
        MDZ_UNROLL
        for (int c = 0; c < 2; ++c) {
            if (jmin + c >= N) continue;
            int last = 0;
            MDZ_UNROLL
            for (int j = jmin + c; j < N; j += 2) {
                const int q = i + j - (N - 2);
                if ((q & 1) == 0) {
                    if (j == jmin + c) mad_wide_cc(e[q], e[q + 1], a[j], a[i]);
                    else               madc_wide_cc(e[q], e[q + 1], a[j], a[i]);
                } else {
                    if (j == jmin + c) mad_wide_cc(o[q - 1], o[q], a[j], a[i]);
                    else               madc_wide_cc(o[q - 1], o[q], a[j], a[i]);
                }
                last = q;
            }
            const int cp = last + 2;
            if (cp < N + 2) {
                if ((cp & 1) == 0) e[cp] = addc(e[cp], 0u);
                else               o[cp - 1] = addc(o[cp - 1], 0u);
            }
        }
    }
    uint32_t x[N + 2];
    x[0] = e[0];
    x[1] = add_cc(e[1], o[0]);
    MDZ_UNROLL
    for (int i = 2; i < N + 1; ++i) x[i] = addc_cc(e[i], o[i - 1]);
    x[N + 1] = addc(e[N + 1], o[N]);
    MDZ_UNROLL
    for (int i = N + 1; i >= 1; --i) t[i] = fsl(x[i - 1], x[i], 1);
    t[0] = x[0] << 1;
    // diagonal a[i]^2 at position 2i >= N-2: q = 2i-(N-2), consecutive pairs
    constexpr int i0 = (N - 1) / 2;          // smallest i with 2i >= N-2
    {
        constexpr int q0 = 2 * i0 - (N - 2);
        mad_wide_cc(t[q0], t[q0 + 1], a[i0], a[i0]);
    }
    MDZ_UNROLL
    for (int i = i0 + 1; i < N; ++i) {
        const int q = 2 * i - (N - 2);
        madc_wide_cc(t[q], t[q + 1], a[i], a[i]);
    }
}

}  // namespace mdz
