// tests/host_emu/coop_gmp_emu.cpp -- TEST BUILD ONLY.
// Compiles the lane-group GMP mpf arithmetic (mdz_b200/csrc/coop_mpf.cuh) for the host with MDZ_HOST_EMU (a warp is
// a vector of 32 lanes in lock step, coop_ops.cuh).  tests/test_coop_vs_gmp.py compares every operation with
// libgmp, and whole pixels with the reference's frac_*_gmp functions.
#define MDZ_HOST_EMU 1
#include <vector>
#include "../../mdz_b200/csrc/coop_mpf.cuh"

using namespace mdz;

// nl limbs, least significant first, top aligned in the N = T K words
template <int K, int T>
static void gload(CNum<K, T>& a, const uint64_t* l, long e, int sgn, int nl)
{
    constexpr int N = T * K;
    if (sgn == 0) { cset_zero(a); return; }
    std::vector<uint32_t> full(N, 0u);
    for (int i = 0; i < nl; ++i) { full[N - 2 * nl + 2 * i] = (uint32_t)l[i]; full[N - 2 * nl + 2 * i + 1] = (uint32_t)(l[i] >> 32); }
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < K; ++j) a.m[j].v[lane] = full[(lane % T) * K + j];
    a.e = (int32_t)e; a.s = sgn < 0;
    cg_adopt<K, T>(a);
}

template <int K, int T>
static void gstore(const CNum<K, T>& r, uint64_t* rl, long* re, int* rs, int nl)
{
    constexpr int N = T * K;
    if (cis_zero(r)) {
        *rs = 0; *re = 0;
        for (int i = 0; i < nl; ++i) rl[i] = 0;
        for (int lane = 0; lane < 32; ++lane) for (int j = 0; j < K; ++j) if (r.m[j].v[lane] != 0u) *rs = 97;   // a zero with words set
        return;
    }
    std::vector<uint32_t> full(N);
    for (int lane = 0; lane < T; ++lane)
        for (int j = 0; j < K; ++j) full[lane * K + j] = r.m[j].v[lane];
    for (int lane = T; lane < 32; ++lane)
        for (int j = 0; j < K; ++j) if (r.m[j].v[lane] != r.m[j].v[lane % T]) { *rs = 98; return; }
    for (int i = 0; i < N - 2 * nl; ++i) if (full[i] != 0u) { *rs = 99; return; }         // words below the value's limbs must be zero
    for (int i = 0; i < nl; ++i) rl[i] = ((uint64_t)full[N - 2 * nl + 2 * i + 1] << 32) | full[N - 2 * nl + 2 * i];
    *rs = r.s ? -1 : 1;
    *re = r.e;
}

template <int K, int T>
static int gop(int op, int nl, const uint64_t* al, long ae, int as, const uint64_t* bl, long be, int bs,
               uint64_t* rl, long* re, int* rs)
{
    if (T * K / 2 < nl + 1) return 0;
    const CoopGCfg cfg = make_coop_gcfg<K, T>(nl);
    std::vector<uint32_t> scr(CoopScratchWords<K, T>::value, 0xdeadbeefu);
    coop_scratch_init<K, T>(scr.data());
    CNum<K, T> a, b, r;
    gload<K, T>(a, al, ae, as, nl); gload<K, T>(b, bl, be, bs, nl);
    switch (op) {
    case 0: cg_mul<K, T>(a, b, r, cfg); break;
    case 1: cg_mul2<K, T>(a, r, cfg); break;
    case 2: cg_add<K, T>(a, b, r, false, cfg, scr.data()); break;
    case 3: cg_add<K, T>(a, b, r, true, cfg, scr.data()); break;
    case 4: *rs = cg_gt4<K, T>(a) ? 1 : 0; return 1;
    default: return 0;
    }
    // the strip must be back in its resting state (zero margins) whatever path the operation took
    {
        constexpr int N = T * K;
        for (int i = 0; i < N + 1; ++i) if (scr[i] != 0u) { *rs = 96; return 1; }
        for (int i = 2 * N + 2; i < 3 * N + 4; ++i) if (scr[i] != 0u) { *rs = 96; return 1; }
    }
    gstore<K, T>(r, rl, re, rs, nl);
    return 1;
}

#define SHAPES(F, ...) \
    switch (K * 100 + T) { \
    case 416: return F<4, 16>(__VA_ARGS__); \
    case 816: return F<8, 16>(__VA_ARGS__); \
    case 632: return F<6, 32>(__VA_ARGS__); \
    case 832: return F<8, 32>(__VA_ARGS__); \
    case 432: return F<4, 32>(__VA_ARGS__); \
    case 808: return F<8, 8>(__VA_ARGS__); \
    case 608: return F<6, 8>(__VA_ARGS__); \
    default: return 0; \
    }

extern "C" int coop_gmp_op(int K, int T, int op, int nl, const uint64_t* al, long ae, int as,
                           const uint64_t* bl, long be, int bs, uint64_t* rl, long* re, int* rs)
{
    SHAPES(gop, op, nl, al, ae, as, bl, be, bs, rl, re, rs)
}

template <int K, int T>
static long gpixel(int nl, int fractal, long depth, const uint64_t* const* l, const long* ex, const int* sg)
{
    if (T * K / 2 < nl + 1) return -1;
    const CoopGCfg cfg = make_coop_gcfg<K, T>(nl);
    std::vector<uint32_t> scr(CoopScratchWords<K, T>::value, 0u);
    coop_scratch_init<K, T>(scr.data());
    CNum<K, T> v[4];
    for (int k = 0; k < 4; ++k) gload<K, T>(v[k], l[k], ex[k], sg[k], nl);
    CPixel<K, T> st;
    cgpixel_init<K, T>(st, v[0], v[1], v[2], v[3], cfg);
    const bool abs_im = fractal == 1;
    const int abs_re = fractal == 2 ? 1 : fractal == 3 ? 2 : 0;
    while (st.iter < depth)
        if (cgpixel_step<K, T>(st, cfg, scr.data(), abs_im, abs_re)) return st.iter;
    return 0;
}

extern "C" long coop_gmp_pixel(int K, int T, int nl, int fractal, long depth,
                               const uint64_t* xl, long xe, int xs, const uint64_t* yl, long ye, int ys,
                               const uint64_t* cxl, long cxe, int cxs, const uint64_t* cyl, long cye, int cys)
{
    const uint64_t* l[4] = {xl, yl, cxl, cyl};
    const long ex[4] = {xe, ye, cxe, cye};
    const int sg[4] = {xs, ys, cxs, cys};
    SHAPES(gpixel, nl, fractal, depth, l, ex, sg)
}

extern "C" long coop_gmp_close_calls() { return g_cg_close_calls; }
