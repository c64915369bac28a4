// kernels_mpfr_a.cu -- MPFR / long double escape-time kernels for 2..8 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_a_kernel(int n, int cyc)
{
    switch (n) {
    case 2: return cyc ? escape_mpfr_kernel<2, true> : escape_mpfr_kernel<2, false>;
    case 3: return cyc ? escape_mpfr_kernel<3, true> : escape_mpfr_kernel<3, false>;
    case 4: return cyc ? escape_mpfr_kernel<4, true> : escape_mpfr_kernel<4, false>;
    case 5: return cyc ? escape_mpfr_kernel<5, true> : escape_mpfr_kernel<5, false>;
    case 6: return cyc ? escape_mpfr_kernel<6, true> : escape_mpfr_kernel<6, false>;
    case 7: return cyc ? escape_mpfr_kernel<7, true> : escape_mpfr_kernel<7, false>;
    case 8: return cyc ? escape_mpfr_kernel<8, true> : escape_mpfr_kernel<8, false>;
    default: return nullptr;
    }
}
int kernels_mpfr_a_smem(int n)
{
    switch (n) {
    case 2: return SmemWords<2>::value;
    case 3: return SmemWords<3>::value;
    case 4: return SmemWords<4>::value;
    case 5: return SmemWords<5>::value;
    case 6: return SmemWords<6>::value;
    case 7: return SmemWords<7>::value;
    case 8: return SmemWords<8>::value;
    default: return 0;
    }
}
