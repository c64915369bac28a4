// kernels_coop.cu -- warp-per-pixel MPFR kernels (coop_kernel.cuh): K limbs per lane, 32 K limbs per value.
#include "coop_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn mdz_kernel_coop(int k)
{
    switch (k) {
    case 2: return escape_coop_kernel<2>;
    case 4: return escape_coop_kernel<4>;
    case 6: return escape_coop_kernel<6>;
    case 8: return escape_coop_kernel<8>;
    default: return nullptr;
    }
}
int mdz_smem_words_coop(int k)      // per block
{
    switch (k) {
    case 2: return CoopSmemWords<2>::value;
    case 4: return CoopSmemWords<4>::value;
    case 6: return CoopSmemWords<6>::value;
    case 8: return CoopSmemWords<8>::value;
    default: return 0;
    }
}
