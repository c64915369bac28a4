// kernels_coop.cu -- lane-group-per-pixel MPFR kernels (coop_kernel.cuh): T lanes x K limbs per value.
#include <stdlib.h>
#include "coop_kernel.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
// the shape for a significand of n32 limbs (33 .. 256): 8 x 6 and 8 x 8 (four pixels per warp), 16 x 8, 32 x 6, 32 x 8
// (16 x 4 for the first 64 limbs measured 12 % slower than 8 x 8: MDZCUDA_COOP_SHAPE=416 brings it back for A/B runs)
void mdz_coop_shape(int n32, int* k, int* t)
{
    if (const char* e = getenv("MDZCUDA_COOP_SHAPE")) {        // A/B measurements: K * 100 + T
        const int v = atoi(e);
        if (v > 0 && n32 <= (v / 100) * (v % 100)) { *k = v / 100; *t = v % 100; return; }
    }
    if (n32 <= 48) { *k = 6; *t = 8; }
    else if (n32 <= 64) { *k = 8; *t = 8; }
    else if (n32 <= 128) { *k = 8; *t = 16; }
    else if (n32 <= 192) { *k = 6; *t = 32; }
    else { *k = 8; *t = 32; }
}
kernel_fn mdz_kernel_coop(int k, int t)
{
    switch (k * 100 + t) {
    case 608: return escape_coop_kernel<6, 8, false>;
    case 808: return escape_coop_kernel<8, 8, false>;
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
    case 608: return CoopSmemWords<6, 8>::value;
    case 808: return CoopSmemWords<8, 8>::value;
    case 416: return CoopSmemWords<4, 16>::value;
    case 816: return CoopSmemWords<8, 16>::value;
    case 632: return CoopSmemWords<6, 32>::value;
    case 832: return CoopSmemWords<8, 32>::value;
    default: return 0;
    }
}
