// kernels_coop.cu -- lane-group-per-pixel MPFR kernels (coop_kernel.cuh): T lanes x K limbs per value.
#include "coop_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
// the shape for a significand of n32 limbs (33 .. 256): 16 x 4, 16 x 8, 32 x 6, 32 x 8
void mdz_coop_shape(int n32, int* k, int* t)
{
    if (n32 <= 64) { *k = 4; *t = 16; }
    else if (n32 <= 128) { *k = 8; *t = 16; }
    else if (n32 <= 192) { *k = 6; *t = 32; }
    else { *k = 8; *t = 32; }
}
kernel_fn mdz_kernel_coop(int k, int t)
{
    switch (k * 100 + t) {
    case 416: return escape_coop_kernel<4, 16, false>;
    case 816: return escape_coop_kernel<8, 16, false>;
    case 632: return escape_coop_kernel<6, 32, false>;
    case 832: return escape_coop_kernel<8, 32, false>;
    default: return nullptr;
    }
}
int mdz_smem_words_coop(int k, int t)      // per block
{
    switch (k * 100 + t) {
    case 416: return CoopSmemWords<4, 16>::value;
    case 816: return CoopSmemWords<8, 16>::value;
    case 632: return CoopSmemWords<6, 32>::value;
    case 832: return CoopSmemWords<8, 32>::value;
    default: return 0;
    }
}
