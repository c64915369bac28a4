// coop_ops.cuh -- MPFR-faithful soft floating point with a GROUP OF LANES PER VALUE: the significand's
// N = T K limbs are split over T lanes (T = 32: a warp per value; T = 16, 8: two, four values per warp), K
// consecutive limbs each (lane 0 of the group the least significant), sign and exponent are
// group-uniform scalars.  For precisions beyond what one thread can hold in registers (above 1024
// bits: T x K = 8 x 6, 8 x 8, 16 x 8, 32 x 6, 32 x 8 -> 1536, 2048, 4096, 6144, 8192 bits).
//
// Same contract as mpfr_sf.cuh -- "exact result, rounded once to p bits, nearest, ties to even",
// i.e. mpfr_mul / mpfr_add / mpfr_sub with MPFR_RNDN as the reference's loops call them
// (src/frac_mandel.c:36-48 and its three siblings) -- with these mechanics:
//   * a product is T steps: in step t every lane multiplies its K limbs by the K limbs of lane t
//     (broadcast by shuffle; a K x K schoolbook block on IMAD.WIDE carry chains, limb_ops.cuh
//     mul_full), adds the 2K-limb block into a sliding window, hands the window's finished low
//     half to the lane below (shuffle) and takes the one from the lane above.  After the last step
//     each lane holds its K limbs of the product's high half; the low half has left through lane 0,
//     which kept its top limb (guard) and the OR of the rest (sticky): rounding is exact, there is
//     no "high product + fallback" here;
//   * carries and borrows across lanes are resolved in one step from two ballots (which blocks
//     generate a carry, which ones would pass one on) by a 32-bit addition: the carry chain of that
//     addition IS the carry chain of the lanes;
//   * alignment and normalisation shifts go through a per-warp strip of shared memory (store the
//     limbs, read them back at an offset, funnel-shift): O(K) work per lane, against O(32 K^2) for a
//     product;
//   * every decision is uniform over the group (one pixel per group), so no case needs a fallback path; with two
//     groups in a warp each cross-lane operation names only its own group's lanes, and the halves may diverge.
//
// The code is written against a small SIMT vocabulary (LW = one 32-bit word per lane, shuffles,
// ballots, the shared-memory strip).  On the device LW is uint32_t and the vocabulary is the
// intrinsics; with MDZ_HOST_EMU (tests/host_emu only, never the product) LW is a vector of 32
// lanes executed in lock step, so the same source is differential-tested against libmpfr on a box
// without a GPU.
#pragma once
#include "mpfr_sf.cuh"

namespace mdz {

#if defined(MDZ_HOST_EMU)
// ---- host emulation of a warp: 32 lanes in lock step ------------------------------------------
struct V32 {
    uint32_t v[32];
    V32() { for (int i = 0; i < 32; ++i) v[i] = 0u; }
    V32(uint32_t c) { for (int i = 0; i < 32; ++i) v[i] = c; }
};
#define MDZ_V32_BINOP(op) \
    inline V32 operator op(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] op b.v[i]; return r; }
MDZ_V32_BINOP(+) MDZ_V32_BINOP(-) MDZ_V32_BINOP(&) MDZ_V32_BINOP(|) MDZ_V32_BINOP(^) MDZ_V32_BINOP(*)
#undef MDZ_V32_BINOP
inline V32 operator~(const V32& a) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = ~a.v[i]; return r; }
inline V32 operator<<(const V32& a, const V32& n) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = n.v[i] >= 32u ? 0u : a.v[i] << n.v[i]; return r; }
inline V32 operator>>(const V32& a, const V32& n) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = n.v[i] >= 32u ? 0u : a.v[i] >> n.v[i]; return r; }
inline V32& operator|=(V32& a, const V32& b) { a = a | b; return a; }
inline V32& operator&=(V32& a, const V32& b) { a = a & b; return a; }
typedef V32 LW;

static thread_local V32 g_ccv;
inline V32 add_cc(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) { uint64_t t = (uint64_t)a.v[i] + b.v[i]; g_ccv.v[i] = (uint32_t)(t >> 32); r.v[i] = (uint32_t)t; } return r; }
inline V32 addc_cc(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) { uint64_t t = (uint64_t)a.v[i] + b.v[i] + g_ccv.v[i]; g_ccv.v[i] = (uint32_t)(t >> 32); r.v[i] = (uint32_t)t; } return r; }
inline V32 addc(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] + b.v[i] + g_ccv.v[i]; return r; }
inline V32 sub_cc(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) { uint64_t t = (uint64_t)a.v[i] - b.v[i]; g_ccv.v[i] = (uint32_t)(t >> 63); r.v[i] = (uint32_t)t; } return r; }
inline V32 subc_cc(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) { uint64_t t = (uint64_t)a.v[i] - b.v[i] - g_ccv.v[i]; g_ccv.v[i] = (uint32_t)(t >> 63); r.v[i] = (uint32_t)t; } return r; }
inline V32 subc(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] - b.v[i] - g_ccv.v[i]; return r; }
inline void mad_wide_cc(V32& lo, V32& hi, const V32& a, const V32& b)
{
    for (int i = 0; i < 32; ++i) {
        uint64_t p = (uint64_t)a.v[i] * b.v[i];
        uint64_t t = (uint64_t)lo.v[i] + (uint32_t)p; lo.v[i] = (uint32_t)t;
        uint64_t u = (uint64_t)hi.v[i] + (uint32_t)(p >> 32) + (t >> 32); hi.v[i] = (uint32_t)u;
        g_ccv.v[i] = (uint32_t)(u >> 32);
    }
}
inline void madc_wide_cc(V32& lo, V32& hi, const V32& a, const V32& b)
{
    for (int i = 0; i < 32; ++i) {
        uint64_t p = (uint64_t)a.v[i] * b.v[i];
        uint64_t t = (uint64_t)lo.v[i] + (uint32_t)p + g_ccv.v[i]; lo.v[i] = (uint32_t)t;
        uint64_t u = (uint64_t)hi.v[i] + (uint32_t)(p >> 32) + (t >> 32); hi.v[i] = (uint32_t)u;
        g_ccv.v[i] = (uint32_t)(u >> 32);
    }
}
inline V32 fsr(const V32& lo, const V32& hi, const V32& s) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = fsr(lo.v[i], hi.v[i], s.v[i]); return r; }
inline V32 fsl(const V32& lo, const V32& hi, const V32& s) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = fsl(lo.v[i], hi.v[i], s.v[i]); return r; }
inline V32 lw_clz(const V32& x) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = (uint32_t)clz32(x.v[i]); return r; }

// T lanes per value: the emulated warp holds 32 / T identical groups; "uniform" results are group 0's
template <int T> inline V32 lane_in() { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = (uint32_t)(i % T); return r; }
template <int T, bool W = false> inline uint32_t bcast(const V32& x, int src) { return x.v[src & (T - 1)]; }
template <int T, bool W = false> inline V32 shfl_dn1(const V32& x) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = (i % T) == T - 1 ? 0u : x.v[i + 1]; return r; }
template <int T, bool W = false> inline V32 shfl_up1(const V32& x) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = (i % T) == 0 ? 0u : x.v[i - 1]; return r; }
template <int T, bool W = false> inline uint32_t ballot_nz(const V32& x) { uint32_t m = 0; for (int i = 0; i < T; ++i) m |= (x.v[i] != 0u ? 1u : 0u) << i; return m; }
inline V32 m_nz(const V32& a) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] != 0u ? 0xffffffffu : 0u; return r; }
inline V32 m_eq(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] == b.v[i] ? 0xffffffffu : 0u; return r; }
inline V32 m_lt(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = a.v[i] < b.v[i] ? 0xffffffffu : 0u; return r; }
inline V32 m_lts(const V32& a, const V32& b) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = (int32_t)a.v[i] < (int32_t)b.v[i] ? 0xffffffffu : 0u; return r; }
inline void sm_store(uint32_t* base, const V32& idx, const V32& val) { for (int i = 0; i < 32; ++i) base[idx.v[i]] = val.v[i]; }
inline V32 sm_load(const uint32_t* base, const V32& idx) { V32 r; for (int i = 0; i < 32; ++i) r.v[i] = base[idx.v[i]]; return r; }
template <int T> inline void warp_sync() {}
inline void warp_converge() {}
#define MDZ_COOP_LOOP
#else
// ---- the device: one 32-bit word per lane ---------------------------------------------------
typedef uint32_t LW;
MDZ_HD LW lw_clz(LW x) { return (LW)__clz((int)x); }
// T lanes per value (32, or 16: two values per warp).  A cross-lane operation names either the lanes of its own
// group (W = false: the two halves of a warp may be in different branches when they get here -- independent thread
// scheduling -- at the price of the convergence check the compiler puts in front of a shuffle whose mask is not a
// constant: MATCH.ANY + REDUX + a vote) or the whole warp (W = true: every lane of the warp executes this very
// instruction; the width argument keeps the data inside the group).  The product and its rounding, nine tenths of
// an iteration, are written without group-dependent branches and run with W = true.
template <int T> MDZ_HD uint32_t group_shift() { return threadIdx.x & 31u & ~(unsigned)(T - 1); }
template <int T, bool W = false> MDZ_HD uint32_t group_mask() { return (T == 32 || W) ? 0xffffffffu : (((1u << (T & 31)) - 1u) << group_shift<T>()); }
template <int T> MDZ_HD LW lane_in() { return threadIdx.x & (unsigned)(T - 1); }
template <int T, bool W = false> MDZ_HD uint32_t bcast(LW x, int src) { return __shfl_sync(group_mask<T, W>(), x, src, T); }
template <int T, bool W = false> MDZ_HD LW shfl_dn1(LW x) { const LW r = __shfl_down_sync(group_mask<T, W>(), x, 1, T); return lane_in<T>() == (unsigned)(T - 1) ? 0u : r; }
template <int T, bool W = false> MDZ_HD LW shfl_up1(LW x) { const LW r = __shfl_up_sync(group_mask<T, W>(), x, 1, T); return lane_in<T>() == 0u ? 0u : r; }
template <int T, bool W = false> MDZ_HD uint32_t ballot_nz(LW x)
{
    const uint32_t b = __ballot_sync(group_mask<T, W>(), x != 0u);
    return T == 32 ? b : (b >> group_shift<T>()) & ((1u << (T & 31)) - 1u);
}
MDZ_HD void warp_converge() { __syncwarp(); }
MDZ_HD LW m_nz(LW a) { return a != 0u ? 0xffffffffu : 0u; }
MDZ_HD LW m_eq(LW a, LW b) { return a == b ? 0xffffffffu : 0u; }
MDZ_HD LW m_lt(LW a, LW b) { return a < b ? 0xffffffffu : 0u; }
MDZ_HD LW m_lts(LW a, LW b) { return (int32_t)a < (int32_t)b ? 0xffffffffu : 0u; }
MDZ_HD void sm_store(uint32_t* base, LW idx, LW val) { base[idx] = val; }
MDZ_HD LW sm_load(const uint32_t* base, LW idx) { return base[idx]; }
template <int T> MDZ_HD void warp_sync() { __syncwarp(group_mask<T>()); }
#define MDZ_COOP_LOOP _Pragma("unroll 1")
#endif

MDZ_HD LW sel(const LW& mask, const LW& a, const LW& b) { return (a & mask) | (b & ~mask); }

// low `n` bits set, for a per-lane signed count: n <= 0 -> 0, n >= 32 -> all ones
MDZ_HD LW low_bits(const LW& n)
{
    const LW neg = m_lts(n, LW(1u));                       // n <= 0
    const LW big = ~m_lts(n, LW(32u));                     // n >= 32
    const LW sh = n & LW(31u);
    const LW part = (LW(1u) << sh) - LW(1u);
    return sel(neg, LW(0u), sel(big, LW(0xffffffffu), part));
}
// the single bit at position n of a 32-bit word, 0 when n is outside 0..31
MDZ_HD LW one_bit(const LW& n)
{
    const LW in = ~m_lts(n, LW(0u)) & m_lts(n, LW(32u));
    return sel(in, LW(1u) << (n & LW(31u)), LW(0u));
}

template <int K, int T>
struct CNum {
    LW m[K];        // lane l holds limbs l*K .. l*K+K-1; top bit of lane 31's last limb set (normalised)
    int32_t e;      // warp-uniform; E_ZERO for zero (all limbs 0)
    uint32_t s;     // warp-uniform; 1 = negative
};

// Rounding position: the significand is N = 32 K limbs, of which the low R = 32 N - p bits stay zero.
struct CoopCfg {
    int prec;
    int R;
};
template <int K, int T> inline CoopCfg make_coop_cfg(int prec) { CoopCfg c; c.prec = prec; c.R = 32 * T * K - prec; return c; }

// per-warp strip of shared memory for the shifts: [0, N+1) zero, [N+1] guard limb, [N+2, 2N+2) the limbs,
// [2N+2, 3N+4) zero.  The zero margins are written once (coop_scratch_init) and never again.
template <int K, int T> struct CoopScratchWords { static constexpr int value = 3 * T * K + 4; };
template <int K, int T> MDZ_HD void coop_scratch_init(uint32_t* scr)
{
    const LW lane = lane_in<T>();
    for (int i = 0; i < (CoopScratchWords<K, T>::value + T - 1) / T; ++i) {
        const LW idx = lane + LW((uint32_t)(T * i));
        const LW ok = m_lt(idx, LW((uint32_t)CoopScratchWords<K, T>::value));
        sm_store(scr, sel(ok, idx, LW(0u)), LW(0u));
    }
    warp_sync<T>();
}

template <int K, int T> MDZ_HD void cset_zero(CNum<K, T>& a)
{
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) a.m[j] = LW(0u);
    a.e = E_ZERO; a.s = 0u;
}
template <int K, int T> MDZ_HD bool cis_zero(const CNum<K, T>& a) { return a.e == E_ZERO; }

// ---- carries across lanes ------------------------------------------------------------------------
// gen: 1 in the lanes whose block produced a carry (borrow); prop: mask of the lanes whose block would
// pass an incoming one on (all ones after an addition, all zeros after a subtraction; never both).
// Returns the carry INTO each lane (0 / 1) and the one that leaves lane 31.  The adder identity
// c = (A + B) ^ A ^ B with A = gen | prop, B = gen makes one 32-bit addition do the whole ripple.
template <int T, bool W = false>
MDZ_HD LW resolve_carries(const LW& gen, const LW& prop, uint32_t& cout)
{
    const uint32_t G = ballot_nz<T, W>(gen);
    const uint32_t P = ballot_nz<T, W>(prop) & ~G;
    const uint64_t A = (uint64_t)(G | P), B = (uint64_t)G;
    const uint64_t C = (A + B) ^ A ^ B;
    cout = (uint32_t)(C >> T) & 1u;
    return (LW((uint32_t)C) >> lane_in<T>()) & LW(1u);
}

// r = x + y + cin0 (cin0: warp-uniform 0 / 1 into limb 0); returns the carry out of the top limb
template <int K, int T, bool W = false>
MDZ_HD uint32_t coop_add_n(const LW (&x)[K], const LW (&y)[K], LW (&r)[K], uint32_t cin0)
{
    const LW c0 = sel(m_eq(lane_in<T>(), LW(0u)), LW(cin0), LW(0u));
    (void)add_cc(c0, LW(0xffffffffu));
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r[j] = addc_cc(x[j], y[j]);
    const LW co = addc(LW(0u), LW(0u));
    LW all = r[0];
    MDZ_UNROLL
    for (int j = 1; j < K; ++j) all = all & r[j];
    uint32_t cout;
    const LW cin = resolve_carries<T, W>(co, m_eq(all, LW(0xffffffffu)), cout);
    r[0] = add_cc(r[0], cin);
    MDZ_UNROLL
    for (int j = 1; j < K; ++j) r[j] = addc_cc(r[j], LW(0u));
    return cout;
}

// r = x - y - bin0; returns the borrow out of the top limb (0 when x >= y + bin0)
template <int K, int T>
MDZ_HD uint32_t coop_sub_n(const LW (&x)[K], const LW (&y)[K], LW (&r)[K], uint32_t bin0)
{
    const LW b0 = sel(m_eq(lane_in<T>(), LW(0u)), LW(bin0), LW(0u));
    (void)sub_cc(LW(0u), b0);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r[j] = subc_cc(x[j], y[j]);
    const LW bo = subc(LW(0u), LW(0u)) & LW(1u);
    LW any = r[0];
    MDZ_UNROLL
    for (int j = 1; j < K; ++j) any = any | r[j];
    uint32_t bout;
    const LW bin = resolve_carries<T>(bo, ~m_nz(any), bout);
    r[0] = sub_cc(r[0], bin);
    MDZ_UNROLL
    for (int j = 1; j < K; ++j) r[j] = subc_cc(r[j], LW(0u));
    return bout;
}

// -1 / 0 / +1: x < y, x == y, x > y as N-limb integers
template <int K, int T>
MDZ_HD int coop_cmp(const LW (&x)[K], const LW (&y)[K])
{
    LW gt = LW(0u), lt = LW(0u);
    MDZ_UNROLL
    for (int j = K - 1; j >= 0; --j) {
        const LW open = ~(gt | lt);
        gt = gt | (open & m_lt(y[j], x[j]));
        lt = lt | (open & m_lt(x[j], y[j]));
    }
    const uint32_t G = ballot_nz<T>(gt), L = ballot_nz<T>(lt);
    if (G == L) return 0;
    return G > L ? 1 : -1;          // the highest lane that differs has its bit in exactly one of them
}

// leading zero bits of the N-limb number (32 N when it is zero)
template <int K, int T>
MDZ_HD int coop_clz(const LW (&x)[K])
{
    LW any = x[0];
    MDZ_UNROLL
    for (int j = 1; j < K; ++j) any = any | x[j];
    const uint32_t M = ballot_nz<T>(any);
    if (M == 0u) return 32 * T * K;
    const int top = 31 - clz32(M);           // M has T bits at most
    LW lz = LW(0u), found = LW(0u);
    MDZ_UNROLL
    for (int j = K - 1; j >= 0; --j) {
        lz = lz + sel(found, LW(0u), lw_clz(x[j]));
        found = found | m_nz(x[j]);
    }
    return (T - 1 - top) * 32 * K + (int)bcast<T>(lz, top);
}

// ---- shifts through the shared-memory strip -----------------------------------------------------------
// y = x >> s with a guard limb g (the 32 bits below limb 0) and sticky (any bit below that); s >= 0
template <int K, int T>
MDZ_HD void coop_shr(const LW (&x)[K], uint32_t s, LW (&y)[K], uint32_t& g, uint32_t& sticky, uint32_t* scr)
{
    constexpr int N = T * K, B = N + 2;
    const uint32_t q = s >> 5, r = s & 31u;
    const LW i0 = lane_in<T>() * LW((uint32_t)K);
    warp_sync<T>();
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) sm_store(scr, i0 + LW((uint32_t)(B + j)), x[j]);
    sm_store(scr, LW((uint32_t)(B - 1)), LW(0u));
    warp_sync<T>();
    LW st = LW(0u);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) {
        const LW at = i0 + LW((uint32_t)(B + j) + q);
        y[j] = fsr(sm_load(scr, at), sm_load(scr, at + LW(1u)), LW(r));
        // what this lane's own limb i = i0 + j loses below the guard: limbs i < q-1 whole, limb q-1 its low r bits
        const LW i = i0 + LW((uint32_t)j);
        const LW whole = m_lts(i + LW(1u), LW(q));                           // i + 1 < q
        const LW edge = m_eq(i + LW(1u), LW(q));
        st = st | (x[j] & (whole | (edge & LW(r ? ((1u << r) - 1u) : 0u))));
    }
    // the guard: bits [s - 32, s) of x, every lane reads the same two words
    const LW lo = sm_load(scr, LW((uint32_t)(B - 1) + q)), hi = sm_load(scr, LW((uint32_t)B + q));
    g = bcast<T>(fsr(lo, hi, LW(r)), 0);
    sticky = ballot_nz<T>(st) != 0u ? 1u : 0u;
}

// (x : g) <<= z, 0 <= z <= 32 N + 32: the guard limb's bits move up into the limbs
template <int K, int T>
MDZ_HD void coop_shl(LW (&x)[K], uint32_t& g, uint32_t z, uint32_t* scr)
{
    constexpr int N = T * K, B = N + 2;
    const uint32_t q = z >> 5, r = z & 31u;
    const LW i0 = lane_in<T>() * LW((uint32_t)K);
    warp_sync<T>();
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) sm_store(scr, i0 + LW((uint32_t)(B + j)), x[j]);
    sm_store(scr, LW((uint32_t)(B - 1)), LW(g));
    warp_sync<T>();
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) {
        const LW at = i0 + LW((uint32_t)(B + j) - q);
        x[j] = fsl(sm_load(scr, at - LW(1u)), sm_load(scr, at), LW(r));
    }
    g = q == 0u ? (r ? g << r : g) : 0u;
    warp_sync<T>();
    sm_store(scr, LW((uint32_t)(B - 1)), LW(0u));         // the guard slot goes back to zero for the next reader
}

// ---- rounding: (x : g : sticky), x normalised, to p bits, nearest, ties to even --------------------------
// returns 1 when the increment carried out of the top (x is then 1000...0 and the caller bumps the exponent).
// W: called by every lane of the warp at once (the groups of a warp must not part ways inside): the increment is
// then added whether it is one or zero.
template <int K, int T, bool W = false>
MDZ_HD uint32_t coop_round(LW (&x)[K], uint32_t g, uint32_t sticky, const CoopCfg& cfg)
{
    const int R = cfg.R;
    const LW i0 = lane_in<T>() * LW((uint32_t)K);
    LW stv = LW(0u), rbv = LW(0u), lsbv = LW(0u);
    LW ulp[K];
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) {
        const LW pos0 = (i0 + LW((uint32_t)j)) * LW(32u);                    // absolute position of this limb's bit 0
        const LW rel_r = LW((uint32_t)(R - 1)) - pos0;                        // round bit relative to this limb (may be negative)
        const LW rel_u = LW((uint32_t)R) - pos0;                              // unit in the last place
        stv = stv | (x[j] & low_bits(rel_r));
        rbv = rbv | (x[j] & one_bit(rel_r));
        ulp[j] = one_bit(rel_u);
        lsbv = lsbv | (x[j] & ulp[j]);
        x[j] = x[j] & ~low_bits(rel_u);
    }
    uint32_t rb, st;
    if (R == 0) { rb = g >> 31; st = sticky | ((g & 0x7fffffffu) != 0u ? 1u : 0u); }
    else { rb = ballot_nz<T, W>(rbv) != 0u ? 1u : 0u; st = sticky | (g != 0u ? 1u : 0u) | (ballot_nz<T, W>(stv) != 0u ? 1u : 0u); }
    const uint32_t lsb = ballot_nz<T, W>(lsbv) != 0u ? 1u : 0u;
    const bool up = rb && (st | lsb);
    if (T == 32 || !W) { if (!up) return 0u; }
    else {
        const LW keep = LW(up ? 0xffffffffu : 0u);
        MDZ_UNROLL
        for (int j = 0; j < K; ++j) ulp[j] = ulp[j] & keep;
    }
    LW t[K];
    const uint32_t cout = coop_add_n<K, T, W>(x, ulp, t, 0u);
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) x[j] = t[j];
    if (cout) x[K - 1] = x[K - 1] | sel(m_eq(lane_in<T>(), LW((uint32_t)(T - 1))), LW(0x80000000u), LW(0u));
    return cout;
}

// ---- product ---------------------------------------------------------------------------------------------
// hi = the high N limbs of a * b (lane-blocked); lowtop = limb N-1 of the product; low_sticky = any lower bit.
// Called by every lane of the warp at once.
template <int K, int T>
MDZ_HD void coop_mul_full(const LW (&a)[K], const LW (&b)[K], LW (&hi)[K], uint32_t& lowtop, uint32_t& low_sticky)
{
    LW W[2 * K + 1];
    MDZ_UNROLL
    for (int i = 0; i < 2 * K + 1; ++i) W[i] = LW(0u);
    LW lowor = LW(0u), lastl = LW(0u);
    MDZ_COOP_LOOP
    for (int t = 0; t < T; ++t) {
        LW bt[K], P[2 * K];
        MDZ_UNROLL
        for (int j = 0; j < K; ++j) bt[j] = LW(bcast<T, true>(b[j], t));
        mul_full<K, LW>(a, bt, P);
        W[0] = add_cc(W[0], P[0]);
        MDZ_UNROLL
        for (int i = 1; i < 2 * K; ++i) W[i] = addc_cc(W[i], P[i]);
        W[2 * K] = addc(W[2 * K], LW(0u));
        // the low half of the window is final for this lane: down it goes (lane 0's is product block t)
        LW L[K];
        MDZ_UNROLL
        for (int j = 0; j < K; ++j) L[j] = W[j];
        lowor = lowor | lastl;
        MDZ_UNROLL
        for (int j = 0; j + 1 < K; ++j) lowor = lowor | L[j];
        lastl = L[K - 1];
        MDZ_UNROLL
        for (int i = 0; i <= K; ++i) W[i] = W[i + K];
        MDZ_UNROLL
        for (int i = K + 1; i <= 2 * K; ++i) W[i] = LW(0u);
        W[0] = add_cc(W[0], shfl_dn1<T, true>(L[0]));
        MDZ_UNROLL
        for (int j = 1; j < K; ++j) W[j] = addc_cc(W[j], shfl_dn1<T, true>(L[j]));
        W[K] = addc(W[K], LW(0u));
    }
    // what is left above a lane's block belongs to the lane above
    LW add[K], lo[K];
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) { add[j] = LW(0u); lo[j] = W[j]; }
    add[0] = shfl_up1<T, true>(W[K]);
    (void)coop_add_n<K, T, true>(lo, add, hi, 0u);
    lowtop = bcast<T, true>(lastl, 0);
    low_sticky = bcast<T, true>(lowor, 0) != 0u ? 1u : 0u;
}

// r = RN(a * b).  Called by every lane of the warp at once (an idle group multiplies zeros): no branch here depends
// on a group's values.
template <int K, int T>
MDZ_HD void cmul(const CNum<K, T>& a, const CNum<K, T>& b, CNum<K, T>& r, const CoopCfg& cfg)
{
    const bool zero = cis_zero(a) || cis_zero(b);
    if (T == 32 && zero) { cset_zero(r); r.s = a.s ^ b.s; return; }
    if (T != 32) warp_converge();
    uint32_t g, st;
    coop_mul_full<K, T>(a.m, b.m, r.m, g, st);
    int32_t e = a.e + b.e;
    // the product of two normalised significands has its top bit at 64N-1 or 64N-2: shift by one or by nothing
    const uint32_t top = bcast<T, true>(r.m[K - 1], T - 1);
    const uint32_t sh = (top >> 31) ^ 1u;
    {
        const LW below = shfl_up1<T, true>(r.m[K - 1]);                // top limb of the lane below; 0 for lane 0 ...
        const LW in = sel(m_eq(lane_in<T>(), LW(0u)), LW(g), below);   // ... which takes the guard's top bit
        MDZ_UNROLL
        for (int j = K - 1; j >= 1; --j) r.m[j] = fsl(r.m[j - 1], r.m[j], LW(sh));
        r.m[0] = fsl(in, r.m[0], LW(sh));
        g <<= sh;
        e -= (int32_t)sh;
    }
    e += (int32_t)coop_round<K, T, true>(r.m, g, st, cfg);
    r.e = (e >= E_MIN && !zero) ? e : E_ZERO;   // as finish_product: below E_MIN the value acts as an exact-zero sticky
    if (r.e == E_ZERO) { MDZ_UNROLL for (int j = 0; j < K; ++j) r.m[j] = LW(0u); }
    r.s = a.s ^ b.s;
}

// r = RN(a + b) (MODE_GENERIC), RN(a - b) for a, b >= 0 (MODE_SUB_POS), RN(a + b) for a, b >= 0 (MODE_ADD_POS)
template <int K, int T, int MODE>
MDZ_HD void cadd(const CNum<K, T>& a, const CNum<K, T>& b, CNum<K, T>& r, const CoopCfg& cfg, uint32_t* scr)
{
    constexpr int N = T * K;
    const uint32_t sb = (MODE == MODE_SUB_POS) ? 1u : (MODE == MODE_ADD_POS ? 0u : b.s);
    const uint32_t sa = (MODE == MODE_GENERIC) ? a.s : 0u;
    const bool sub = sa != sb;
    if (cis_zero(b)) { r = a; r.s = cis_zero(a) ? 0u : sa; return; }
    if (cis_zero(a)) { r = b; r.s = sb; return; }
    const int32_t d = a.e - b.e;
    bool a_big = d > 0;
    if (d == 0) {
        const int c = sub ? coop_cmp<K, T>(a.m, b.m) : 1;
        if (c == 0) { cset_zero(r); return; }               // exact cancellation: +0
        a_big = c > 0;
    }
    const CNum<K, T>& big = a_big ? a : b;
    const CNum<K, T>& small = a_big ? b : a;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    const uint32_t s_out = a_big ? sa : sb;
    if (ad >= (uint32_t)(32 * N + 2)) {                   // the smaller one lies wholly below the rounding position (gap >= p + 2)
        r = big; r.s = s_out; return;
    }
    LW y[K], x[K];
    uint32_t g, st;
    coop_shr<K, T>(small.m, ad, y, g, st, scr);
    int32_t e = big.e;
    if (!sub) {
        const uint32_t cout = coop_add_n<K, T>(big.m, y, x, 0u);
        if (cout) {
            // (1 : x) >>= 1; the bit that leaves limb 0 goes to the top of the guard
            const LW above = sel(m_eq(lane_in<T>(), LW((uint32_t)(T - 1))), LW(1u), shfl_dn1<T>(x[0]));
            const uint32_t out = bcast<T>(x[0], 0) & 1u;
            MDZ_UNROLL
            for (int j = 0; j + 1 < K; ++j) x[j] = fsr(x[j], x[j + 1], LW(1u));
            x[K - 1] = fsr(x[K - 1], above, LW(1u));
            st |= g & 1u;
            g = (g >> 1) | (out << 31);
            e += 1;
        }
    } else {
        // big - (y : g : sticky): the guard limb and the sticky bit borrow from the limbs
        const uint32_t bin = (g | st) != 0u ? 1u : 0u;
        g = 0u - g - st;
        (void)coop_sub_n<K, T>(big.m, y, x, bin);
        int z = coop_clz<K, T>(x);
        if (z == 32 * N) {
            if (g == 0u) { cset_zero(r); return; }          // cannot happen for unequal operands; kept for safety
            z += clz32(g);
        }
        if (z > 0) { coop_shl<K, T>(x, g, (uint32_t)z, scr); e -= z; }
    }
    MDZ_UNROLL
    for (int j = 0; j < K; ++j) r.m[j] = x[j];
    e += (int32_t)coop_round<K, T>(r.m, g, st, cfg);
    r.e = e;
    r.s = s_out;
}

// a > 4 ?   (4 = 0.1b * 2^3)
template <int K, int T>
MDZ_HD bool cgreater_than_4(const CNum<K, T>& a)
{
    if (cis_zero(a) || a.s) return false;
    if (a.e != 3) return a.e > 3;
    // anything set besides the leading bit?
    LW low = a.m[K - 1] & sel(m_eq(lane_in<T>(), LW((uint32_t)(T - 1))), LW(0x7fffffffu), LW(0xffffffffu));
    MDZ_UNROLL
    for (int j = 0; j + 1 < K; ++j) low = low | a.m[j];
    return ballot_nz<T>(low) != 0u;
}

// ---- one pixel (reference src/frac_mandel.c:34-50 and its three variants; escape_step.cuh pixel_step) ------
template <int K, int T>
struct CPixel {
    CNum<K, T> wre, wim, wre2, wim2, cre, cim;
    int iter;
};

// the two squares the loop starts from; every lane of the warp at once (for a group in mid-pixel this recomputes
// what it already holds)
template <int K, int T>
MDZ_HD void cpixel_squares(CPixel<K, T>& st, const CoopCfg& cfg)
{
    cmul<K, T>(st.wre, st.wre, st.wre2, cfg); st.wre2.s = 0u;
    cmul<K, T>(st.wim, st.wim, st.wim2, cfg); st.wim2.s = 0u;
}

template <int K, int T>
MDZ_HD void cpixel_load(CPixel<K, T>& st, const CNum<K, T>& x, const CNum<K, T>& y, const CNum<K, T>& cx, const CNum<K, T>& cy)
{
    st.wre = x; st.wim = y;
    st.cre = cx; st.cim = cy;
    st.iter = 0;
}

template <int K, int T>
MDZ_HD void cpixel_init(CPixel<K, T>& st, const CNum<K, T>& x, const CNum<K, T>& y, const CNum<K, T>& cx, const CNum<K, T>& cy, const CoopCfg& cfg)
{
    cpixel_load<K, T>(st, x, y, cx, cy);
    cpixel_squares<K, T>(st, cfg);
}

template <int K, int T>
MDZ_HD bool cescaped(const CNum<K, T>& wim2, const CNum<K, T>& wre2, const CoopCfg& cfg, uint32_t* scr)
{
    const int32_t emax = wim2.e > wre2.e ? wim2.e : wre2.e;
    if (emax >= 4) return true;
    if (emax < 2) return false;         // both squares below 2: the sum cannot exceed 4 even after rounding
    CNum<K, T> t;
    cadd<K, T, MODE_ADD_POS>(wim2, wre2, t, cfg, scr);
    return cgreater_than_4<K, T>(t);
}

// One iteration.  With two groups in a warp every lane comes here as long as either group has a pixel: the three
// products are the whole warp's, the additions (whose paths depend on the values) each group's own; `active` is
// uniform over the group.
template <int K, int T>
MDZ_HD bool cpixel_step(CPixel<K, T>& st, const CoopCfg& cfg, uint32_t* scr, bool abs_im, int abs_re, bool active = true)
{
    CNum<K, T> t;
    // wim = 2*wre*wim + c_im       (|.| on the product for burning ship)
    cmul<K, T>(st.wre, st.wim, t, cfg);
    if (active) {
        ++st.iter;
        if (!cis_zero(t)) t.e += 1;
        if (abs_im) t.s = 0u;
        cadd<K, T, MODE_GENERIC>(t, st.cim, st.wim, cfg, scr);
        // wre = wre2 - wim2 + c_re     (|.| on the difference for celtic / odd steps of the hybrid)
        cadd<K, T, MODE_SUB_POS>(st.wre2, st.wim2, t, cfg, scr);
        if (abs_re == 1 || (abs_re == 2 && (st.iter & 1))) t.s = 0u;
        cadd<K, T, MODE_GENERIC>(t, st.cre, st.wre, cfg, scr);
    }
    cmul<K, T>(st.wim, st.wim, st.wim2, cfg); st.wim2.s = 0u;
    cmul<K, T>(st.wre, st.wre, st.wre2, cfg); st.wre2.s = 0u;
    return active && cescaped<K, T>(st.wim2, st.wre2, cfg, scr);
}

}  // namespace mdz
