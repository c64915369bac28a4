// kernels_gmpf_c.cu -- GMP mpf mode, fast implementation (mpf_fast.cuh), NL = 11..13 limbs (513..640 bits).
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_gmpf_c_kernel(int nl)
{
    switch (nl) {
    case 11: return escape_gmpf_kernel<22>;
    case 12: return escape_gmpf_kernel<24>;
    case 13: return escape_gmpf_kernel<26>;
    default: return nullptr;
    }
}
int kernels_gmpf_c_smem(int nl)
{
    switch (nl) {
    case 11: return GSmemWords<22>::value;
    case 12: return GSmemWords<24>::value;
    case 13: return GSmemWords<26>::value;
    default: return 0;
    }
}
