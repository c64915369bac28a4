// mp_convert.h -- host-side packing between the MPFR in-memory significand
// (ceil(p/64) 64-bit limbs, least significant first, top-aligned; MPFR manual
// "Internals", layout probed in SURVEY Appendix B) and the device's N x 32-bit
// top-aligned limbs.  N = ceil(p/32); when ceil(p/64)*2 > N the lowest 32-bit
// word of the MPFR limbs lies wholly below bit p and is zero, so it is dropped.
#pragma once
#include <stdint.h>
#include <string.h>

namespace mdz {

inline int limbs32_for_prec(long prec) { return (int)((prec + 31) / 32); }
inline int limbs64_for_prec(long prec) { return (int)((prec + 63) / 64); }

// limbs64: nl = ceil(prec/64) words.  out: n32 words, least significant first.
inline void sig64_to_sig32(const uint64_t* limbs64, long prec, uint32_t* out, int n32)
{
    const int nl = limbs64_for_prec(prec);
    const int off = 2 * nl - n32;           // 0 or 1 words dropped at the bottom
    for (int i = 0; i < n32; ++i) {
        const int w = i + off;
        const uint64_t l = limbs64[w >> 1];
        out[i] = (w & 1) ? (uint32_t)(l >> 32) : (uint32_t)l;
    }
}

inline void sig32_to_sig64(const uint32_t* in, int n32, long prec, uint64_t* limbs64)
{
    const int nl = limbs64_for_prec(prec);
    const int off = 2 * nl - n32;
    for (int i = 0; i < nl; ++i) limbs64[i] = 0;
    for (int i = 0; i < n32; ++i) {
        const int w = i + off;
        limbs64[w >> 1] |= (uint64_t)in[i] << ((w & 1) ? 32 : 0);
    }
}

}  // namespace mdz
