// kernels_gmpf_d.cu -- GMP mpf mode, fast implementation (mpf_fast.cuh), NL = 14..16 limbs (641..832 bits).
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_gmpf_d_kernel(int nl)
{
    switch (nl) {
    case 14: return escape_gmpf_kernel<28>;
    case 15: return escape_gmpf_kernel<30>;
    case 16: return escape_gmpf_kernel<32>;
    default: return nullptr;
    }
}
int kernels_gmpf_d_smem(int nl)
{
    switch (nl) {
    case 14: return GSmemWords<28>::value;
    case 15: return GSmemWords<30>::value;
    case 16: return GSmemWords<32>::value;
    default: return 0;
    }
}
