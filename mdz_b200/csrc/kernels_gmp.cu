// kernels_gmp.cu -- GMP mpf mode, clear implementation (mpf_sf.cuh), NL = 3..10 limbs.
#include "escape_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn mdz_kernel_gmp_clear(int nl)
{
    switch (nl) {
    case 3: return escape_gmp_kernel<3>;  case 4: return escape_gmp_kernel<4>;  case 5: return escape_gmp_kernel<5>;
    case 6: return escape_gmp_kernel<6>;  case 7: return escape_gmp_kernel<7>;  case 8: return escape_gmp_kernel<8>;
    case 9: return escape_gmp_kernel<9>;  case 10: return escape_gmp_kernel<10>;
    default: return nullptr;
    }
}
