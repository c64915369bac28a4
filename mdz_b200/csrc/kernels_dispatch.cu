// kernels_dispatch.cu -- limb count -> kernel, over the per-range translation units.
#include "escape_params.cuh"
using namespace mdz;
typedef void (*kernel_fn)(const EscapeParams);
kernel_fn kernels_mpfr_a_kernel(int n, int cyc); int kernels_mpfr_a_smem(int n);
kernel_fn kernels_mpfr_b_kernel(int n, int cyc); int kernels_mpfr_b_smem(int n);
kernel_fn kernels_mpfr_c_kernel(int n, int cyc); int kernels_mpfr_c_smem(int n);
kernel_fn kernels_mpfr_d_kernel(int n, int cyc); int kernels_mpfr_d_smem(int n);
kernel_fn kernels_mpfr_e_kernel(int n, int cyc); int kernels_mpfr_e_smem(int n);
kernel_fn kernels_mpfr_f_kernel(int n, int cyc); int kernels_mpfr_f_smem(int n);
kernel_fn kernels_mpfr_g_kernel(int n, int cyc); int kernels_mpfr_g_smem(int n);
kernel_fn mdz_kernel_mpfr(int n, int cyc)
{
    kernel_fn f = nullptr;
    if (!f) f = kernels_mpfr_a_kernel(n, cyc);
    if (!f) f = kernels_mpfr_b_kernel(n, cyc);
    if (!f) f = kernels_mpfr_c_kernel(n, cyc);
    if (!f) f = kernels_mpfr_d_kernel(n, cyc);
    if (!f) f = kernels_mpfr_e_kernel(n, cyc);
    if (!f) f = kernels_mpfr_f_kernel(n, cyc);
    if (!f) f = kernels_mpfr_g_kernel(n, cyc);
    return f;
}
int mdz_smem_words_mpfr(int n)
{
    int w = 0;
    if (!w) w = kernels_mpfr_a_smem(n);
    if (!w) w = kernels_mpfr_b_smem(n);
    if (!w) w = kernels_mpfr_c_smem(n);
    if (!w) w = kernels_mpfr_d_smem(n);
    if (!w) w = kernels_mpfr_e_smem(n);
    if (!w) w = kernels_mpfr_f_smem(n);
    if (!w) w = kernels_mpfr_g_smem(n);
    return w;
}

kernel_fn kernels_gmpf_a_kernel(int nl); int kernels_gmpf_a_smem(int nl);
kernel_fn kernels_gmpf_b_kernel(int nl); int kernels_gmpf_b_smem(int nl);
kernel_fn kernels_gmpf_c_kernel(int nl); int kernels_gmpf_c_smem(int nl);
kernel_fn kernels_gmpf_d_kernel(int nl); int kernels_gmpf_d_smem(int nl);
kernel_fn mdz_kernel_gmp_fast(int nl)
{
    kernel_fn f = kernels_gmpf_a_kernel(nl);
    if (!f) f = kernels_gmpf_b_kernel(nl);
    if (!f) f = kernels_gmpf_c_kernel(nl);
    return f ? f : kernels_gmpf_d_kernel(nl);
}
int mdz_smem_words_gmp_fast(int nl)
{
    int w = kernels_gmpf_a_smem(nl);
    if (!w) w = kernels_gmpf_b_smem(nl);
    if (!w) w = kernels_gmpf_c_smem(nl);
    return w ? w : kernels_gmpf_d_smem(nl);
}
