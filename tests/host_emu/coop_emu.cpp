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
template <int K>
static void load(CNum<K>& a, const uint64_t* l, int sgn, long e, long prec)
{
    constexpr int N = 32 * K;
    if (sgn == 0) { cset_zero(a); return; }
    const int n = limbs32_for_prec(prec);
    std::vector<uint32_t> tmp(n), full(N, 0u);
    sig64_to_sig32(l, prec, tmp.data(), n);
    for (int i = 0; i < n; ++i) full[N - n + i] = tmp[i];
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < K; ++j) a.m[j].v[lane] = full[lane * K + j];
    a.e = (int32_t)e; a.s = sgn < 0;
}

template <int K>
static void store(const CNum<K>& r, uint64_t* rl, int* rs, long* re, long prec)
{
    constexpr int N = 32 * K;
    const int n = limbs32_for_prec(prec);
    if (cis_zero(r)) { *rs = 0; *re = 0; for (int i = 0; i < limbs64_for_prec(prec); ++i) rl[i] = 0; return; }
    std::vector<uint32_t> full(N), tmp(n);
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < K; ++j) full[lane * K + j] = r.m[j].v[lane];
    for (int i = 0; i < N - n; ++i) if (full[i] != 0u) { *rs = 99; return; }        // bits below the precision must be zero
    for (int i = 0; i < n; ++i) tmp[i] = full[N - n + i];
    sig32_to_sig64(tmp.data(), n, prec, rl);
    *rs = r.s ? -1 : 1;
    *re = r.e;
}

template <int K>
static int binop(int op, long prec, const uint64_t* al, int as, long ae, const uint64_t* bl, int bs, long be,
                 uint64_t* rl, int* rs, long* re)
{
    const CoopCfg cfg = make_coop_cfg<K>((int)prec);
    std::vector<uint32_t> scr(CoopScratchWords<K>::value, 0xdeadbeefu);
    coop_scratch_init<K>(scr.data());
    CNum<K> a, b, r;
    load<K>(a, al, as, ae, prec); load<K>(b, bl, bs, be, prec);
    switch (op) {
    case 0: cmul<K>(a, b, r, cfg); break;
    case 1: cmul<K>(a, a, r, cfg); r.s = 0; break;
    case 2: cadd<K, MODE_GENERIC>(a, b, r, cfg, scr.data()); break;
    case 3: b.s ^= 1u; cadd<K, MODE_GENERIC>(a, b, r, cfg, scr.data()); break;
    case 4: cadd<K, MODE_SUB_POS>(a, b, r, cfg, scr.data()); break;
    case 5: cadd<K, MODE_ADD_POS>(a, b, r, cfg, scr.data()); break;
    case 6: *rs = cgreater_than_4<K>(a) ? 1 : 0; return 1;
    case 11: *rs = cescaped<K>(a, b, cfg, scr.data()) ? 1 : 0; return 1;
    default: return 0;
    }
    store<K>(r, rl, rs, re, prec);
    return 1;
}

extern "C" int coop_binop(int K, int op, long prec, const uint64_t* al, int as, long ae, const uint64_t* bl, int bs, long be,
                          uint64_t* rl, int* rs, long* re)
{
    switch (K) {
    case 2: return binop<2>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 4: return binop<4>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 6: return binop<6>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    case 8: return binop<8>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
    default: return 0;
    }
}

template <int K>
static long pixel(long prec, int fractal, long depth, const uint64_t* const* l, const int* sg, const long* ex)
{
    const CoopCfg cfg = make_coop_cfg<K>((int)prec);
    std::vector<uint32_t> scr(CoopScratchWords<K>::value, 0u);
    coop_scratch_init<K>(scr.data());
    CNum<K> v[4];
    for (int k = 0; k < 4; ++k) load<K>(v[k], l[k], sg[k], ex[k], prec);
    CPixel<K> st;
    cpixel_init<K>(st, v[0], v[1], v[2], v[3], cfg);
    const bool abs_im = fractal == 1;
    const int abs_re = fractal == 2 ? 1 : fractal == 3 ? 2 : 0;
    while (st.iter < depth)
        if (cpixel_step<K>(st, cfg, scr.data(), abs_im, abs_re)) return st.iter;
    return 0;
}

extern "C" long coop_pixel(int K, long prec, int fractal, long depth,
                           const uint64_t* xl, int xs, long xe, const uint64_t* yl, int ys, long ye,
                           const uint64_t* cxl, int cxs, long cxe, const uint64_t* cyl, int cys, long cye)
{
    const uint64_t* l[4] = {xl, yl, cxl, cyl};
    const int sg[4] = {xs, ys, cxs, cys};
    const long ex[4] = {xe, ye, cxe, cye};
    switch (K) {
    case 2: return pixel<2>(prec, fractal, depth, l, sg, ex);
    case 4: return pixel<4>(prec, fractal, depth, l, sg, ex);
    case 6: return pixel<6>(prec, fractal, depth, l, sg, ex);
    case 8: return pixel<8>(prec, fractal, depth, l, sg, ex);
    default: return -1;
    }
}
