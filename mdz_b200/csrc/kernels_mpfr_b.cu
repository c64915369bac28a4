// kernels_mpfr_b.cu -- MPFR / long double escape-time kernels for 9..12 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_b_kernel(int n, int cyc)
{
    switch (n) {
    case 9: return cyc ? escape_mpfr_kernel<9, true> : escape_mpfr_kernel<9, false>;
    case 10: return cyc ? escape_mpfr_kernel<10, true> : escape_mpfr_kernel<10, false>;
    case 11: return cyc ? escape_mpfr_kernel<11, true> : escape_mpfr_kernel<11, false>;
    case 12: return cyc ? escape_mpfr_kernel<12, true> : escape_mpfr_kernel<12, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_b_smem(int n)
{
    switch (n) {
    case 9: return SmemWords<9>::value;
    case 10: return SmemWords<10>::value;
    case 11: return SmemWords<11>::value;
    case 12: return SmemWords<12>::value;
    default: return 0;
    }
}
