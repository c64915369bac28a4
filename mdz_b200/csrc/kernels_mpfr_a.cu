// kernels_mpfr_a.cu -- MPFR / long double escape-time kernels for 2..8 words (generated list; see
// mdzcuda.cu "kernels are instantiated in separate translation units").
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_a_kernel(int n)
{
    switch (n) {
    case 2: return escape_mpfr_kernel<2>;
    case 3: return escape_mpfr_kernel<3>;
    case 4: return escape_mpfr_kernel<4>;
    case 5: return escape_mpfr_kernel<5>;
    case 6: return escape_mpfr_kernel<6>;
    case 7: return escape_mpfr_kernel<7>;
    case 8: return escape_mpfr_kernel<8>;
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
