// tests/host_emu/coop_emu.cpp -- TEST BUILD ONLY.
// Compiles the warp-cooperative arithmetic (mdz_b200/csrc/coop_ops.cuh) for the host with
// MDZ_HOST_EMU: a warp is a vector of 32 lanes executed in lock step, shuffles and ballots are array
// operations, the shared-memory strip is a plain array.  tests/test_coop_vs_mpfr.py compares every
// operation with libmpfr, and whole pixels with the reference's frac_*_mpfr functions.
#define MDZ_HOST_EMU 1
#include <vector>
#include "../../mdz_b200/csrc/coop_ops.cuh"
#include "../../mdz_b200/csrc/mp_convert.h"

using namespace mdz;

// top-aligned: the n = ceil(p/32) significant limbs at the top of the N = 32 K, zeros below
template <int K, int T>
static void load(CNum<K, T>& a, const uint64_t* l, int sgn, long e, long prec)
{
    constexpr int N = T * K;
    if (sgn == 0) { cset_zero(a); return; }
    const int n = limbs32_for_prec(prec);
    std::vector<uint32_t> tmp(n), full(N, 0u);
    sig64_to_sig32(l, prec, tmp.data(), n);
    for (int i = 0; i < n; ++i) full[N - n + i] = tmp[i];
    for (int lane = 0; lane < 32; ++lane)                   // 32 / T identical groups
        for (int j = 0; j < K; ++j) a.m[j].v[lane] = full[(lane % T) * K + j];
    a.e = (int32_t)e; a.s = sgn < 0;
}

template <int K, int T>
static void store(const CNum<K, T>& r, uint64_t* rl, int* rs, long* re, long prec)
{
    constexpr int N = T * K;
    const int n = limbs32_for_prec(prec);
    if (cis_zero(r)) { *rs = 0; *re = 0; for (int i = 0; i < limbs64_for_prec(prec); ++i) rl[i] = 0; return; }
    std::vector<uint32_t> full(N), tmp(n);
    for (int lane = 0; lane < T; ++lane)
        for (int j = 0; j < K; ++j) full[lane * K + j] = r.m[j].v[lane];
    for (int lane = T; lane < 32; ++lane)                   // the other groups computed the same thing
        for (int j = 0; j < K; ++j) if (r.m[j].v[lane] != r.m[j].v[lane % T]) { *rs = 98; return; }
    for (int i = 0; i < N - n; ++i) if (full[i] != 0u) { *rs = 99; return; }        // bits below the precision must be zero
    for (int i = 0; i < n; ++i) tmp[i] = full[N - n + i];
    sig32_to_sig64(tmp.data(), n, prec, rl);
    *rs = r.s ? -1 : 1;
    *re = r.e;
}

template <int K, int T>
static int binop(int op, long prec, const uint64_t* al, int as, long ae, const uint64_t* bl, int bs, long be,
                 uint64_t* rl, int* rs, long* re)
{
    const CoopCfg cfg = make_coop_cfg<K, T>((int)prec);
    std::vector<uint32_t> scr(CoopScratchWords<K, T>::value, 0xdeadbeefu);
    coop_scratch_init<K, T>(scr.data());
    CNum<K, T> a, b, r;
    load<K, T>(a, al, as, ae, prec); load<K, T>(b, bl, bs, be, prec);
    switch (op) {
    case 0: cmul<K, T>(a, b, r, cfg); break;
    case 1: cmul<K, T>(a, a, r, cfg); r.s = 0; break;
    case 2: cadd<K, T, MODE_GENERIC>(a, b, r, cfg, scr.data()); break;
    case 3: b.s ^= 1u; cadd<K, T, MODE_GENERIC>(a, b, r, cfg, scr.data()); break;
    case 4: cadd<K, T, MODE_SUB_POS>(a, b, r, cfg, scr.data()); break;
    case 5: cadd<K, T, MODE_ADD_POS>(a, b, r, cfg, scr.data()); break;
    case 6: *rs = cgreater_than_4<K, T>(a) ? 1 : 0; return 1;
    case 11: *rs = cescaped<K, T>(a, b, cfg, scr.data()) ? 1 : 0; return 1;
    default: return 0;
    }
    store<K, T>(r, rl, rs, re, prec);
    return 1;
}

extern "C" int coop_binop(int K, int T, int op, long prec, const uint64_t* al, int as, long ae, const uint64_t* bl, int bs, long be,
                          uint64_t* rl, int* rs, long* re)
{
    switch (K * 100 + T) {
    case 232: return binop<2, 32>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 432: return binop<4, 32>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 632: return binop<6, 32>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 832: return binop<8, 32>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 416: return binop<4, 16>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 816: return binop<8, 16>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 216: return binop<2, 16>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 808: return binop<8, 8>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 608: return binop<6, 8>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    default: return 0;
    }
}

template <int K, int T>
static long pixel(long prec, int fractal, long depth, const uint64_t* const* l, const int* sg, const long* ex)
{
    const CoopCfg cfg = make_coop_cfg<K, T>((int)prec);
    std::vector<uint32_t> scr(CoopScratchWords<K, T>::value, 0u);
    coop_scratch_init<K, T>(scr.data());
    CNum<K, T> v[4];
    for (int k = 0; k < 4; ++k) load<K, T>(v[k], l[k], sg[k], ex[k], prec);
    CPixel<K, T> st;
    cpixel_init<K, T>(st, v[0], v[1], v[2], v[3], cfg);
    const bool abs_im = fractal == 1;
    const int abs_re = fractal == 2 ? 1 : fractal == 3 ? 2 : 0;
    while (st.iter < depth)
        if (cpixel_step<K, T>(st, cfg, scr.data(), abs_im, abs_re)) return st.iter;
    return 0;
}

extern "C" long coop_pixel(int K, int T, long prec, int fractal, long depth,
                           const uint64_t* xl, int xs, long xe, const uint64_t* yl, int ys, long ye,
                           const uint64_t* cxl, int cxs, long cxe, const uint64_t* cyl, int cys, long cye)
{
    const uint64_t* l[4] = {xl, yl, cxl, cyl};
    const int sg[4] = {xs, ys, cxs, cys};
    const long ex[4] = {xe, ye, cxe, cye};
    switch (K * 100 + T) {
    case 232: return pixel<2, 32>(prec, fractal, depth, l, sg, ex);
    case 432: return pixel<4, 32>(prec, fractal, depth, l, sg, ex);
    case 632: return pixel<6, 32>(prec, fractal, depth, l, sg, ex);
    case 832: return pixel<8, 32>(prec, fractal, depth, l, sg, ex);
    case 416: return pixel<4, 16>(prec, fractal, depth, l, sg, ex);
    case 816: return pixel<8, 16>(prec, fractal, depth, l, sg, ex);
    case 216: return pixel<2, 16>(prec, fractal, depth, l, sg, ex);
    case 808: return pixel<8, 8>(prec, fractal, depth, l, sg, ex);
    case 608: return pixel<6, 8>(prec, fractal, depth, l, sg, ex);
    default: return -1;
    }
}
