// kernels_gmpf_a.cu -- GMP mpf mode, fast implementation (mpf_fast.cuh), NL = 4..7 limbs.
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_gmpf_a_kernel(int nl)
{
    switch (nl) {
    case 4: return escape_gmpf_kernel<8>;
    case 5: return escape_gmpf_kernel<10>;
    case 6: return escape_gmpf_kernel<12>;
    case 7: return escape_gmpf_kernel<14>;
    default: return nullptr;
    }
}
int kernels_gmpf_a_smem(int nl)
{
    switch (nl) {
    case 4: return GSmemWords<8>::value;
    case 5: return GSmemWords<10>::value;
    case 6: return GSmemWords<12>::value;
    case 7: return GSmemWords<14>::value;
    default: return 0;
    }
}
