// kernels_mpfr_f.cu -- MPFR / long double escape-time kernels for 27..29 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_f_kernel(int n, int cyc)
{
    switch (n) {
    case 27: return cyc ? escape_mpfr_kernel<27, true> : escape_mpfr_kernel<27, false>;
    case 28: return cyc ? escape_mpfr_kernel<28, true> : escape_mpfr_kernel<28, false>;
    case 29: return cyc ? escape_mpfr_kernel<29, true> : escape_mpfr_kernel<29, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_f_smem(int n)
{
    switch (n) {
    case 27: return SmemWords<27>::value;
    case 28: return SmemWords<28>::value;
    case 29: return SmemWords<29>::value;
    default: return 0;
    }
}
