// kernels_coop_gmp.cu -- lane-group-per-pixel GMP mpf kernels (coop_kernel.cuh with coop_mpf.cuh's arithmetic).
#include "coop_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn mdz_kernel_coop_gmp(int k, int t)
{
    switch (k * 100 + t) {
    case 608: return escape_coop_kernel<6, 8, true>;
    case 808: return escape_coop_kernel<8, 8, true>;
    case 416: return escape_coop_kernel<4, 16, true>;
    case 816: return escape_coop_kernel<8, 16, true>;
    case 632: return escape_coop_kernel<6, 32, true>;
    case 832: return escape_coop_kernel<8, 32, true>;
    default: return nullptr;
    }
}
