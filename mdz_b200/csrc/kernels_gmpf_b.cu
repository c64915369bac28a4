// kernels_gmpf_b.cu -- GMP mpf mode, fast implementation (mpf_fast.cuh), NL = 8..10 limbs.
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_gmpf_b_kernel(int nl)
{
    switch (nl) {
    case 8: return escape_gmpf_kernel<16>;
    case 9: return escape_gmpf_kernel<18>;
    case 10: return escape_gmpf_kernel<20>;
    default: return nullptr;
    }
}
int kernels_gmpf_b_smem(int nl)
{
    switch (nl) {
    case 8: return GSmemWords<16>::value;
    case 9: return GSmemWords<18>::value;
    case 10: return GSmemWords<20>::value;
    default: return 0;
    }
}
