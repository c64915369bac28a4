// kernels_mpfr_c.cu -- MPFR / long double escape-time kernels for 13..16 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_c_kernel(int n, int cyc)
{
    switch (n) {
    case 13: return cyc ? escape_mpfr_kernel<13, true> : escape_mpfr_kernel<13, false>;
    case 14: return cyc ? escape_mpfr_kernel<14, true> : escape_mpfr_kernel<14, false>;
    case 15: return cyc ? escape_mpfr_kernel<15, true> : escape_mpfr_kernel<15, false>;
    case 16: return cyc ? escape_mpfr_kernel<16, true> : escape_mpfr_kernel<16, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_c_smem(int n)
{
    switch (n) {
    case 13: return SmemWords<13>::value;
    case 14: return SmemWords<14>::value;
    case 15: return SmemWords<15>::value;
    case 16: return SmemWords<16>::value;
    default: return 0;
    }
}
