// kernels_mpfr_e.cu -- MPFR / long double escape-time kernels for 22..26 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_e_kernel(int n, int cyc)
{
    switch (n) {
    case 22: return cyc ? escape_mpfr_kernel<22, true> : escape_mpfr_kernel<22, false>;
    case 23: return cyc ? escape_mpfr_kernel<23, true> : escape_mpfr_kernel<23, false>;
    case 24: return cyc ? escape_mpfr_kernel<24, true> : escape_mpfr_kernel<24, false>;
    case 25: return cyc ? escape_mpfr_kernel<25, true> : escape_mpfr_kernel<25, false>;
    case 26: return cyc ? escape_mpfr_kernel<26, true> : escape_mpfr_kernel<26, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_e_smem(int n)
{
    switch (n) {
    case 22: return SmemWords<22>::value;
    case 23: return SmemWords<23>::value;
    case 24: return SmemWords<24>::value;
    case 25: return SmemWords<25>::value;
    case 26: return SmemWords<26>::value;
    default: return 0;
    }
}
