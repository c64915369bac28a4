// ld64_step.cuh -- the escape-time iteration at p = 64 on a 64-bit significand:
// the hardware ("long double") mode of the reference, src/frac_mandel.c:5-21,
// src/frac_burning_ship.c:5-23, src/frac_generalized_celtic.c:5-23,
// src/frac_variant.c:5-22, whose x87 arithmetic (64-bit significand, round to
// nearest even) is the MPFR rule at precision 64 (SURVEY finding 1).
//
// This is the speculative step of escape_step.cuh specialised for two limbs.  The
// generic limb code spends most of its instructions on cases a 64-bit significand
// does not have (guard limbs, limb shifters, rounding positions inside a limb); here
// a product is four IMAD.WIDE, a sum is formed exactly in a 128-bit frame (so rounding
// needs no sticky bookkeeping), and round-to-nearest-even is one four-instruction
// carry chain.  The whole iteration is one basic block.  Everything outside the
// covered domain raises `rare`, and the caller (pixel_step_auto) redoes the iteration
// with the general step, so results are those of the general code by construction:
//   covered:  exponent gap of an addition <= 62, fewer than 31 cancelled bits,
//             no zero / underflowed operand, no rounding carry out of the top bit.
#pragma once
#include "escape_step.cuh"

namespace mdz {

#if defined(MDZ_HOST_EMU)
inline uint64_t shr64c(uint64_t x, uint32_t n) { return n >= 64 ? 0 : x >> n; }
inline uint64_t shl64c(uint64_t x, uint32_t n) { return n >= 64 ? 0 : x << n; }
// {m1:m0} = {h1:h0} + (({l1:l0} | ({h1:h0} & 1)) > 2^63): round to nearest, ties to even
inline void round_rne64(uint32_t& m0, uint32_t& m1, uint32_t h0, uint32_t h1, uint32_t l0, uint32_t l1)
{
    const uint64_t h = ((uint64_t)h1 << 32) | h0, l = (((uint64_t)l1 << 32) | l0) | (h & 1u);
    const uint64_t m = h + (l > 0x8000000000000000ull ? 1u : 0u);
    m0 = (uint32_t)m; m1 = (uint32_t)(m >> 32);
}
// the same, with the carry out of the increment (the significand was all ones: it is now 0)
inline void round_rne64c(uint32_t& m0, uint32_t& m1, uint32_t& c, uint32_t h0, uint32_t h1, uint32_t l0, uint32_t l1)
{
    const uint64_t h = ((uint64_t)h1 << 32) | h0, l = (((uint64_t)l1 << 32) | l0) | (h & 1u);
    const uint64_t m = h + (l > 0x8000000000000000ull ? 1u : 0u);
    m0 = (uint32_t)m; m1 = (uint32_t)(m >> 32); c = m < h ? 1u : 0u;
}
// x = a + (b ^ mask) + (mask & 1) over 128 bits (a0 == 0)
inline void addsub128(uint32_t (&x)[4], uint32_t a1, uint32_t a2, uint32_t a3, const uint32_t (&b)[4], uint32_t mask)
{
    uint64_t c = mask & 1u, t;
    t = (uint64_t)0  + (b[0] ^ mask) + c; x[0] = (uint32_t)t; c = t >> 32;
    t = (uint64_t)a1 + (b[1] ^ mask) + c; x[1] = (uint32_t)t; c = t >> 32;
    t = (uint64_t)a2 + (b[2] ^ mask) + c; x[2] = (uint32_t)t; c = t >> 32;
    t = (uint64_t)a3 + (b[3] ^ mask) + c; x[3] = (uint32_t)t;
}
// x = a - b over 128 bits (a0 == 0)
inline void sub128(uint32_t (&x)[4], uint32_t a1, uint32_t a2, uint32_t a3, const uint32_t (&b)[4])
{
    uint64_t br = 0, t;
    t = (uint64_t)0  - b[0] - br; x[0] = (uint32_t)t; br = (t >> 32) & 1u;
    t = (uint64_t)a1 - b[1] - br; x[1] = (uint32_t)t; br = (t >> 32) & 1u;
    t = (uint64_t)a2 - b[2] - br; x[2] = (uint32_t)t; br = (t >> 32) & 1u;
    t = (uint64_t)a3 - b[3] - br; x[3] = (uint32_t)t;
}
#else
// PTX shifts clamp the count at the register width, so a count of 64 gives 0
MDZ_HD uint64_t shr64c(uint64_t x, uint32_t n) { uint64_t r; asm("shr.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
MDZ_HD uint64_t shl64c(uint64_t x, uint32_t n) { uint64_t r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
// round up iff (l | (h & 1)) > 2^63  <=>  (l | (h & 1)) + (2^63 - 1) carries out of 64 bits
MDZ_HD void round_rne64(uint32_t& m0, uint32_t& m1, uint32_t h0, uint32_t h1, uint32_t l0, uint32_t l1)
{
    const uint32_t w0 = l0 | (h0 & 1u);
    uint32_t junk;
    asm("{\n\t"
        "add.cc.u32  %2, %3, 0xffffffff;\n\t"
        "addc.cc.u32 %2, %4, 0x7fffffff;\n\t"
        "addc.cc.u32 %0, %5, 0;\n\t"
        "addc.u32    %1, %6, 0;\n\t"
        "}" : "=&r"(m0), "=&r"(m1), "=&r"(junk) : "r"(w0), "r"(l1), "r"(h0), "r"(h1));
}
MDZ_HD void round_rne64c(uint32_t& m0, uint32_t& m1, uint32_t& c, uint32_t h0, uint32_t h1, uint32_t l0, uint32_t l1)
{
    const uint32_t w0 = l0 | (h0 & 1u);
    uint32_t junk;
    asm("{\n\t"
        "add.cc.u32  %3, %4, 0xffffffff;\n\t"
        "addc.cc.u32 %3, %5, 0x7fffffff;\n\t"
        "addc.cc.u32 %0, %6, 0;\n\t"
        "addc.cc.u32 %1, %7, 0;\n\t"
        "addc.u32    %2, 0, 0;\n\t"
        "}" : "=&r"(m0), "=&r"(m1), "=&r"(c), "=&r"(junk) : "r"(w0), "r"(l1), "r"(h0), "r"(h1));
}
MDZ_HD void addsub128(uint32_t (&x)[4], uint32_t a1, uint32_t a2, uint32_t a3, const uint32_t (&b)[4], uint32_t mask)
{
    uint32_t junk;
    asm("{\n\t"
        "add.cc.u32  %4, %5, 0xffffffff;\n\t"       // carry in = 1 when subtracting
        "addc.cc.u32 %0, %6, 0;\n\t"
        "addc.cc.u32 %1, %7, %10;\n\t"
        "addc.cc.u32 %2, %8, %11;\n\t"
        "addc.u32    %3, %9, %12;\n\t"
        "}" : "=&r"(x[0]), "=&r"(x[1]), "=&r"(x[2]), "=&r"(x[3]), "=&r"(junk)
            : "r"(mask & 1u), "r"(b[0] ^ mask), "r"(b[1] ^ mask), "r"(b[2] ^ mask), "r"(b[3] ^ mask),
              "r"(a1), "r"(a2), "r"(a3));
}
// a difference known to be one at compile time: one borrow chain, nothing to complement
MDZ_HD void sub128(uint32_t (&x)[4], uint32_t a1, uint32_t a2, uint32_t a3, const uint32_t (&b)[4])
{
    asm("{\n\t"
        "sub.cc.u32  %0, 0, %4;\n\t"
        "subc.cc.u32 %1, %8, %5;\n\t"
        "subc.cc.u32 %2, %9, %6;\n\t"
        "subc.u32    %3, %10, %7;\n\t"
        "}" : "=&r"(x[0]), "=&r"(x[1]), "=&r"(x[2]), "=&r"(x[3])
            : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(a1), "r"(a2), "r"(a3));
}
#endif

// 1 when bit 31 of x is clear, as a register value: written out because the compiler turns the
// plain expression, applied to the top half of a 64-bit product, into a 64-bit compare + select
// and then needs two more instructions to subtract the resulting predicate from the exponent
#if defined(MDZ_HOST_EMU)
inline uint32_t top_clear(uint32_t x) { return (x >> 31) ^ 1u; }
#else
MDZ_HD uint32_t top_clear(uint32_t x)
{
    uint32_t r;
    asm("{\n\t.reg .u32 t;\n\tshr.u32 t, %1, 31;\n\txor.b32 %0, t, 1;\n\t}" : "=r"(r) : "r"(x));
    return r;
}
#endif

// The fast operations report what they cannot do through three accumulators instead of a
// predicate per operation (thirteen compares per iteration otherwise):
//   topand  AND of the top words of every result: its bit 31 is clear as soon as one result lost
//           its leading bit (zero factor, 31 or more cancelled bits, rounding carried out)
//   negor   OR of the top words of the differences' 128-bit frames: bit 31 set means one came out
//           negative (operands with equal exponents and top words, ordered by those alone)
//   rare    the remaining tests (exponent gap beyond the frame, product exponent below E_MIN)
struct Ld64Flags {
    uint32_t topand, negor;
    bool rare;
};
MDZ_HD void ld64_flags_init(Ld64Flags& f, bool rare) { f.topand = 0xffffffffu; f.negor = 0u; f.rare = rare; }
MDZ_HD bool ld64_flags_rare(const Ld64Flags& f) { return f.rare || (int32_t)(~f.topand | f.negor) < 0; }

// r = RN(a * b): the sign is left to the caller.  A product just below a power of two
// (top bit at 126, all ones after the one-bit shift) can round up past 2^64; the top-bit
// test catches that and a zero operand.
// WIDE (level 2): an exactly zero factor gives an exact zero, and a product that rounds up to the next
// power of two is renormalised instead of declined -- a converged orbit can sit on either for ever
// (c = -1: wre is 0 every other iteration; c = 1/4 + i/8: wre * wim rounds up to 1/8 at the fixed point).
template <bool WIDE = false>
MDZ_HD void mul64_core(const Num<2>& a, const Num<2>& b, Num<2>& r, Ld64Flags& f)
{
    // (Measured alternative, rejected: limb_ops.cuh mul_full<2> -- aligned accumulator pairs, no
    // zero-extension moves -- trades six IMAD.MOV on the idle FMA pipe for one more add on the ALU pipe,
    // which is the one that binds: 13.3 against 12.65 ms per generation.)
    const uint64_t p00 = (uint64_t)a.m[0] * b.m[0];
    const uint64_t t   = (uint64_t)a.m[1] * b.m[0] + (uint32_t)(p00 >> 32);
    const uint64_t u   = (uint64_t)a.m[0] * b.m[1] + (uint32_t)t;
    const uint64_t hi  = (uint64_t)a.m[1] * b.m[1] + (uint32_t)(t >> 32) + (uint32_t)(u >> 32);
    const uint32_t p0 = (uint32_t)p00, p1 = (uint32_t)u, p2 = (uint32_t)hi, p3 = (uint32_t)(hi >> 32);
    const uint32_t sh = top_clear(p3);               // top bit at 127 or 126
    const uint32_t x3 = fsl(p2, p3, sh), x2 = fsl(p1, p2, sh), x1 = fsl(p0, p1, sh), x0 = p0 << sh;
    if (WIDE) {
        uint32_t c;
        round_rne64c(r.m[0], r.m[1], c, x2, x3, x0, x1);
        r.m[1] |= c << 31;
        const bool z = a.m[1] == 0u || b.m[1] == 0u;
        const int32_t e = a.e + b.e - (int32_t)sh + (int32_t)c;
        f.rare = f.rare || (!z && e < E_MIN);
        r.e = z ? E_ZERO : e;
        f.topand &= r.m[1] | (z ? 0x80000000u : 0u);
    } else {
        round_rne64(r.m[0], r.m[1], x2, x3, x0, x1);
        r.e = a.e + b.e - (int32_t)sh;
        f.topand &= r.m[1];
        f.rare = f.rare || r.e < E_MIN;
    }
}
template <bool WIDE = false>
MDZ_HD void mul64_spec(const Num<2>& a, const Num<2>& b, Num<2>& r, bool& rare)
{
    Ld64Flags f; ld64_flags_init(f, rare);
    mul64_core<WIDE>(a, b, r, f);
    rare = ld64_flags_rare(f);
}

// r = RN(a + b) with the signs as given.  The operands are ordered by (exponent, top word) alone --
// one 64-bit compare.  When those are equal an addition does not care about the order; a
// difference then cancels 32 bits or more, which is outside this function's domain anyway: it
// comes out with a zero top word (caught by topand) or negative (caught by negor).
//
// WIDE is the kernel's level 2 in long double mode, for warps that hold pixels on which the plain
// version declines at every iteration (escape_kernel.cuh "adapt"), at ~10 instructions more per
// addition:
//  * a second operand below a quarter of the first one's last place -- exponent gap >= 66, or an
//    exact zero, e.g. c_re on the column x = 0 or wim^2 next to the real axis -- only has to be there,
//    and on which side: it enters as one sticky bit at the bottom of the frame (the clamped shifts
//    have already made it 0).  A - B with A a power of two is the one case a quarter ulp can decide;
//    the difference then normalises to all ones, the increment carries out, and the top-bit test
//    declines.  Gaps of 64 and 65 stay with the general code;
//  * 32 to 62 cancelled bits are normalised by a word move first (orbits that converge to a fixed
//    point on a diagonal, |wre| = |wim|, cancel wre^2 - wim^2 almost completely for ever).
// On BASELINE configs[1] ~1 700 interior pixels of the first kind and 13 of the second used to run
// the general step 10 000 times each, alone in their warps: the end of every render waited for them.
// SUBPOS: r = RN(a - b) for a, b >= 0 (wre^2 - wim^2), b's sign as given is ignored.
template <bool WIDE = false, bool SUBPOS = false>
MDZ_HD void add64_core(const Num<2>& a, const Num<2>& b, Num<2>& r, Ld64Flags& f)
{
    const int32_t d = a.e - b.e;
    const int64_t ka = (int64_t)(((uint64_t)(uint32_t)a.e << 32) | a.m[1]);
    const int64_t kb = (int64_t)(((uint64_t)(uint32_t)b.e << 32) | b.m[1]);
    // |A| >= |B| afterwards, up to the low words; level 2 normalises deep cancellation itself, so
    // there the low words have their say as well
    const bool swap = WIDE ? (d < 0 || (d == 0 && (((uint64_t)a.m[1] << 32) | a.m[0]) < (((uint64_t)b.m[1] << 32) | b.m[0])))
                           : ka < kb;
    const uint32_t A0 = swap ? b.m[0] : a.m[0], A1 = swap ? b.m[1] : a.m[1];
    const uint32_t B0 = swap ? a.m[0] : b.m[0], B1 = swap ? a.m[1] : b.m[1];
    const uint64_t Bm = ((uint64_t)B1 << 32) | B0;
    const int32_t Ae = swap ? b.e : a.e;
    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
    if (WIDE) f.rare = f.rare || (ad - 64u) < 2u; else f.rare = f.rare || ad > 62u;
    // 128-bit frame with one bit of headroom: A >> 1, B >> (ad + 1); nothing of B
    // leaves the frame while ad <= 63, so the sum is exact
    const uint64_t BH = shr64c(Bm, ad + 1u), BL = shl64c(Bm, 63u - ad);
    const uint32_t sticky = (WIDE && ad > 65u) ? 1u : 0u;
    const uint32_t bw[4] = { (uint32_t)BL | sticky, (uint32_t)(BL >> 32), (uint32_t)BH, (uint32_t)(BH >> 32) };
    uint32_t x[4];
    if (SUBPOS) {
        sub128(x, A0 << 31, fsr(A0, A1, 1), A1 >> 1, bw);
        f.negor |= x[3];
    } else {
        const uint32_t mask = (a.s != b.s) ? 0xffffffffu : 0u;
        addsub128(x, A0 << 31, fsr(A0, A1, 1), A1 >> 1, bw, mask);
        f.negor |= x[3] & mask;                                 // bit 127 is the sums' carry; a difference sets it only when negative
    }
    // normalise.  31 or more cancelled bits (top word zero, about one addition in a million
    // on orbit data) are left to the general code: lz is then 32, the funnel shifts move
    // nothing, and the zero top word fails the top-bit test
    int32_t e = Ae + 1;
    if (WIDE) {
        const bool z = x[3] == 0u;                              // 32 bits or more cancelled: one word up
        x[3] = z ? x[2] : x[3]; x[2] = z ? x[1] : x[2]; x[1] = z ? x[0] : x[1]; x[0] = z ? 0u : x[0];
        e -= z ? 32 : 0;
    }
    const uint32_t lz = (uint32_t)clz32(x[3]);
    const uint32_t h1 = fsl(x[2], x[3], lz), h0 = fsl(x[1], x[2], lz), l1 = fsl(x[0], x[1], lz), l0 = x[0] << lz;
    if (WIDE) {
        // level 2 also keeps an exactly zero result and a rounding increment that carries out
        uint32_t c;
        round_rne64c(r.m[0], r.m[1], c, h0, h1, l0, l1);
        r.m[1] |= c << 31;
        const bool z = (x[0] | x[1] | x[2] | x[3]) == 0u;
        f.topand &= r.m[1] | (z ? 0x80000000u : 0u);
        r.e = z ? E_ZERO : e - (int32_t)lz + (int32_t)c;
    } else {
    round_rne64(r.m[0], r.m[1], h0, h1, l0, l1);
    // top bit clear: >= 31 bits cancelled / exact zero, or the increment carried out
    f.topand &= r.m[1];
    r.e = e - (int32_t)lz;
    }
    r.s = SUBPOS ? (swap ? 1u : 0u) : (swap ? b.s : a.s);
}
template <bool WIDE = false>
MDZ_HD void add64_spec(const Num<2>& a, const Num<2>& b, Num<2>& r, bool& rare)
{
    Ld64Flags f; ld64_flags_init(f, rare);
    add64_core<WIDE>(a, b, r, f);
    rare = ld64_flags_rare(f);
}

// One iteration from `in` to `out` (distinct objects: the kernel's hot loop ping-pongs
// between two register sets instead of copying the state back), c passed by value so that it
// can live in registers.  `rare` comes back true when the step declined; `out` is then garbage.
// The fractal type enters as three words (EscapeParams::ld_masks, read straight from the constant
// bank: computed inside the kernel the compiler re-derives them from the type in every iteration, ten
// instructions): the product's sign is kept (im_keep = 1) or dropped (0: burning ship); the
// difference keeps its sign when ((iteration & re_and) ^ re_xor) is 1 -- always (0, 1), never (0, 0:
// generalized celtic), on even iterations (1, 1: the hybrid).
MDZ_HD Ld64Masks ld64_masks(bool abs_im, int abs_re)
{
    Ld64Masks m;
    m.im_keep = abs_im ? 0u : 1u;
    m.re_and = abs_re == 2 ? 1u : 0u;
    m.re_xor = abs_re == 1 ? 0u : 1u;
    return m;
}

template <bool WIDE = false>
MDZ_HD bool ld64_step(const PixelState<2>& in, PixelState<2>& out, const Num<2>& cre, const Num<2>& cim,
                      uint32_t* scr, const RoundCfg& rc, const Ld64Masks& mk, bool& rare)
{
    out.iter = in.iter + 1;
    out.cre_e = in.cre_e; out.cim_e = in.cim_e; out.cre_s = in.cre_s; out.cim_s = in.cim_s;
    Ld64Flags f; ld64_flags_init(f, rare);
    Num<2> t, u;
    // wim = 2*wre*wim + c_im
    mul64_core<WIDE>(in.wre, in.wim, t, f);
    t.e += 1;
    t.s = (in.wre.s ^ in.wim.s) & mk.im_keep;
    // wre = wre2 - wim2 + c_re
    add64_core<WIDE, true>(in.wre2, in.wim2, u, f);
    u.s &= ((uint32_t)out.iter & mk.re_and) ^ mk.re_xor;
    add64_core<WIDE>(t, cim, out.wim, f);
    add64_core<WIDE>(u, cre, out.wre, f);
    mul64_core<WIDE>(out.wim, out.wim, out.wim2, f);
    mul64_core<WIDE>(out.wre, out.wre, out.wre2, f);
    rare = ld64_flags_rare(f);
    out.wim2.s = 0; out.wre2.s = 0;
    const int32_t emax = out.wim2.e > out.wre2.e ? out.wim2.e : out.wre2.e;
    bool esc = emax >= 4;
    if (!rare && !esc && emax >= 2) {
        // RN(wim2 + wre2) > 4 can only be in doubt when the larger square is in [2, 8), and
        // then the top words settle it unless the sum is within 2^-27 of 4
        const int pre = escape_precheck<2>(out.wim2, out.wre2);
        esc = pre > 0;
        if (pre == 0) {
            MDZ_COUNT(CNT_ESC_ADD);
            Num<2> sum; bool r2 = false;
            add64_spec(out.wim2, out.wre2, sum, r2);
            if (r2) fadd<2, MODE_ADD_POS>(out.wim2, out.wre2, sum, rc, scr);
            esc = greater_than_4<2>(sum);
        }
    }
    return esc;
}

template <>
MDZ_HD bool pixel_step_spec<2>(PixelState<2>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                               uint32_t* scr, const RoundCfg& rc, bool abs_im, int abs_re, uint32_t& rare_out)
{
    bool rare = rc.ulp != 1u;                        // precision below 64 bits: general code only
    Num<2> cim, cre;
    cim.m[0] = cim_m[0]; cim.m[1] = cim_m[kScratchStride]; cim.e = st.cim_e; cim.s = st.cim_s;
    cre.m[0] = cre_m[0]; cre.m[1] = cre_m[kScratchStride]; cre.e = st.cre_e; cre.s = st.cre_s;
    PixelState<2> out;
    const bool esc = ld64_step(st, out, cre, cim, scr, rc, ld64_masks(abs_im, abs_re), rare);
    st = out;
    rare_out |= rare ? 1u : 0u;
    return esc;
}

// level 2 at two limbs: the additions also take a far smaller or zero operand and deep cancellation
template <>
MDZ_HD bool pixel_step_spec_wide<2>(PixelState<2>& st, const uint32_t* cre_m, const uint32_t* cim_m,
                                    uint32_t* scr, const RoundCfg& rc, bool abs_im, int abs_re, uint32_t& rare_out)
{
    bool rare = rc.ulp != 1u;
    Num<2> cim, cre;
    cim.m[0] = cim_m[0]; cim.m[1] = cim_m[kScratchStride]; cim.e = st.cim_e; cim.s = st.cim_s;
    cre.m[0] = cre_m[0]; cre.m[1] = cre_m[kScratchStride]; cre.e = st.cre_e; cre.s = st.cre_s;
    PixelState<2> out;
    const bool esc = ld64_step<true>(st, out, cre, cim, scr, rc, ld64_masks(abs_im, abs_re), rare);
    st = out;
    rare_out |= rare ? 1u : 0u;
    return esc;
}

}  // namespace mdz
