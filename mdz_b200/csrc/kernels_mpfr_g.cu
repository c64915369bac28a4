// kernels_mpfr_g.cu -- MPFR / long double escape-time kernels for 30..32 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_g_kernel(int n, int cyc)
{
    switch (n) {
    case 30: return cyc ? escape_mpfr_kernel<30, true> : escape_mpfr_kernel<30, false>;
    case 31: return cyc ? escape_mpfr_kernel<31, true> : escape_mpfr_kernel<31, false>;
    case 32: return cyc ? escape_mpfr_kernel<32, true> : escape_mpfr_kernel<32, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_g_smem(int n)
{
    switch (n) {
    case 30: return SmemWords<30>::value;
    case 31: return SmemWords<31>::value;
    case 32: return SmemWords<32>::value;
    default: return 0;
    }
}
