// tests/host_emu/emu.cpp -- TEST BUILD ONLY.
// Compiles the device arithmetic headers (mdz_b200/csrc/limb_ops.cuh,
// mpfr_sf.cuh) for the host with MDZ_HOST_EMU, where every PTX carry-chain
// primitive is replaced by a C++ emulation with an explicit carry flag.  This
// lets the limb algorithms be differential-tested against libmpfr on a box
// without a GPU.  It is not a product path: libmdzcuda never links this file
// and has no CPU fallback.
#define MDZ_HOST_EMU 1
#include "../../mdz_b200/csrc/mpfr_sf.cuh"
#include "../../mdz_b200/csrc/mp_convert.h"

using namespace mdz;

template <int N>
static int binop(int op, long prec,
                 const uint64_t* al, int as, long ae,
                 const uint64_t* bl, int bs, long be,
                 uint64_t* rl, int* rs, long* re)
{
    RoundCfg rc = make_round_cfg(N, (int)prec);
    Num<N> a, b, r;
    if (as == 0) set_zero(a); else { sig64_to_sig32(al, prec, a.m, N); a.e = (int32_t)ae; a.s = as < 0; }
    if (bs == 0) set_zero(b); else { sig64_to_sig32(bl, prec, b.m, N); b.e = (int32_t)be; b.s = bs < 0; }
    switch (op) {
    case 0: fmul<N>(a, b, r, rc); break;
    case 1: fsqr<N>(a, r, rc); break;
    case 2: fadd<N, MODE_GENERIC>(a, b, r, rc); break;
    case 3: b.s ^= 1u; fadd<N, MODE_GENERIC>(a, b, r, rc); break;
    case 4: fadd<N, MODE_SUB_POS>(a, b, r, rc); break;
    case 5: fadd<N, MODE_ADD_POS>(a, b, r, rc); break;
    case 6: *rs = greater_than_4<N>(a) ? 1 : 0; return 1;
    default: return 0;
    }
    if (is_zero(r)) { *rs = 0; *re = 0; for (int i = 0; i < limbs64_for_prec(prec); ++i) rl[i] = 0; return 1; }
    sig32_to_sig64(r.m, N, prec, rl);
    *rs = r.s ? -1 : 1;
    *re = r.e;
    return 1;
}

#define CASE(n) case n: return binop<n>(op, prec, al, as, ae, bl, bs, be, rl, rs, re);
extern "C" int emu_binop(int op, long prec,
                         const uint64_t* al, int as, long ae,
                         const uint64_t* bl, int bs, long be,
                         uint64_t* rl, int* rs, long* re)
{
    switch (limbs32_for_prec(prec)) {
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
    CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(20) CASE(24) CASE(32)
    default: return 0;
    }
}
