// kernels_mpfr_d.cu -- MPFR / long double escape-time kernels for 17..21 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_d_kernel(int n, int cyc)
{
    switch (n) {
    case 17: return cyc ? escape_mpfr_kernel<17, true> : escape_mpfr_kernel<17, false>;
    case 18: return cyc ? escape_mpfr_kernel<18, true> : escape_mpfr_kernel<18, false>;
    case 19: return cyc ? escape_mpfr_kernel<19, true> : escape_mpfr_kernel<19, false>;
    case 20: return cyc ? escape_mpfr_kernel<20, true> : escape_mpfr_kernel<20, false>;
    case 21: return cyc ? escape_mpfr_kernel<21, true> : escape_mpfr_kernel<21, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_d_smem(int n)
{
    switch (n) {
    case 17: return SmemWords<17>::value;
    case 18: return SmemWords<18>::value;
    case 19: return SmemWords<19>::value;
    case 20: return SmemWords<20>::value;
    case 21: return SmemWords<21>::value;
    default: return 0;
    }
}
