/*
 * mdz_mp_abi.h -- hand-declared slice of the public GMP / MPFR C ABI.
 *
 * This image ships libgmp.so.10 (GMP 6.3.0) and libmpfr.so.6 (MPFR 4.2.1) as
 * runtime libraries only: there is no gmp.h / mpfr.h.  MDZ's host API passes
 * `mpfr_t` and `mpf_t` values across the render boundary
 * (reference src/image_info.h:65-66), so both libmdzcuda (host prologue) and
 * the oracle build (oracle/shim/) need the struct layouts and a handful of
 * prototypes.  Everything below is the documented public ABI of those
 * libraries (struct layouts are unchanged since GMP 4 / MPFR 2.x); nothing here
 * is taken from the reference.
 */
#ifndef MDZ_MP_ABI_H
#define MDZ_MP_ABI_H

#include <stdio.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- GMP ------------------------------------------------------------- */
typedef unsigned long mp_limb_t;
typedef long          mp_limb_signed_t;
typedef long          mp_exp_t;
typedef unsigned long mp_bitcnt_t;
typedef long          mp_size_t;

typedef struct {
    int        _mp_prec;   /* limbs of guaranteed precision; alloc is prec+1 */
    int        _mp_size;   /* signed limb count (sign = sign of value)       */
    mp_exp_t   _mp_exp;    /* exponent in limbs                              */
    mp_limb_t *_mp_d;
} __mpf_struct;
typedef __mpf_struct        mpf_t[1];
typedef __mpf_struct       *mpf_ptr;
typedef const __mpf_struct *mpf_srcptr;

#define __GNU_MP_VERSION 6
#define __GNU_MP_VERSION_MINOR 3

#define mpf_init2    __gmpf_init2
#define mpf_init     __gmpf_init
#define mpf_clear    __gmpf_clear
#define mpf_set      __gmpf_set
#define mpf_set_si   __gmpf_set_si
#define mpf_set_ui   __gmpf_set_ui
#define mpf_set_d    __gmpf_set_d
#define mpf_set_str  __gmpf_set_str
#define mpf_get_str  __gmpf_get_str
#define mpf_get_d    __gmpf_get_d
#define mpf_add      __gmpf_add
#define mpf_sub      __gmpf_sub
#define mpf_mul      __gmpf_mul
#define mpf_mul_ui   __gmpf_mul_ui
#define mpf_div      __gmpf_div
#define mpf_ui_div   __gmpf_ui_div
#define mpf_abs      __gmpf_abs
#define mpf_neg      __gmpf_neg
#define mpf_cmp      __gmpf_cmp
#define mpf_cmp_ui   __gmpf_cmp_ui
#define mpf_cmp_si   __gmpf_cmp_si
#define mpf_get_prec __gmpf_get_prec
#define mpf_set_prec __gmpf_set_prec

void   __gmpf_init2(mpf_ptr, mp_bitcnt_t);
void   __gmpf_init(mpf_ptr);
void   __gmpf_clear(mpf_ptr);
void   __gmpf_set(mpf_ptr, mpf_srcptr);
void   __gmpf_set_si(mpf_ptr, long);
void   __gmpf_set_ui(mpf_ptr, unsigned long);
void   __gmpf_set_d(mpf_ptr, double);
int    __gmpf_set_str(mpf_ptr, const char *, int);
char  *__gmpf_get_str(char *, mp_exp_t *, int, size_t, mpf_srcptr);
double __gmpf_get_d(mpf_srcptr);
void   __gmpf_add(mpf_ptr, mpf_srcptr, mpf_srcptr);
void   __gmpf_sub(mpf_ptr, mpf_srcptr, mpf_srcptr);
void   __gmpf_mul(mpf_ptr, mpf_srcptr, mpf_srcptr);
void   __gmpf_mul_ui(mpf_ptr, mpf_srcptr, unsigned long);
void   __gmpf_div(mpf_ptr, mpf_srcptr, mpf_srcptr);
void   __gmpf_ui_div(mpf_ptr, unsigned long, mpf_srcptr);
void   __gmpf_abs(mpf_ptr, mpf_srcptr);
void   __gmpf_neg(mpf_ptr, mpf_srcptr);
int    __gmpf_cmp(mpf_srcptr, mpf_srcptr);
int    __gmpf_cmp_ui(mpf_srcptr, unsigned long);
int    __gmpf_cmp_si(mpf_srcptr, long);
mp_bitcnt_t __gmpf_get_prec(mpf_srcptr);
void   __gmpf_set_prec(mpf_ptr, mp_bitcnt_t);

/* ---- MPFR ------------------------------------------------------------ */
#define MPFR_VERSION_MAJOR 4
#define MPFR_VERSION_MINOR 2
#define MPFR_VERSION_PATCHLEVEL 1

typedef long mpfr_prec_t;
typedef long mpfr_exp_t;
typedef int  mpfr_sign_t;
#define mp_prec_t mpfr_prec_t
#define mp_rnd_t  mpfr_rnd_t

typedef enum {
    MPFR_RNDN = 0, MPFR_RNDZ, MPFR_RNDU, MPFR_RNDD, MPFR_RNDA, MPFR_RNDF,
    MPFR_RNDNA = -1
} mpfr_rnd_t;
#define GMP_RNDN MPFR_RNDN
#define GMP_RNDZ MPFR_RNDZ
#define GMP_RNDU MPFR_RNDU
#define GMP_RNDD MPFR_RNDD

typedef struct {
    mpfr_prec_t _mpfr_prec;
    mpfr_sign_t _mpfr_sign;   /* +1 / -1 */
    mpfr_exp_t  _mpfr_exp;    /* value = 0.1xxx * 2^exp; special values below */
    mp_limb_t  *_mpfr_d;      /* ceil(prec/64) limbs, least significant first  */
} __mpfr_struct;
typedef __mpfr_struct        mpfr_t[1];
typedef __mpfr_struct       *mpfr_ptr;
typedef const __mpfr_struct *mpfr_srcptr;

/* special exponents (MPFR 3.x/4.x): LONG_MIN+1 zero, +2 NaN, +3 Inf */
#define MDZ_MPFR_EXP_ZERO (-0x7fffffffffffffffL)
#define MDZ_MPFR_EXP_NAN  (-0x7ffffffffffffffeL)
#define MDZ_MPFR_EXP_INF  (-0x7ffffffffffffffdL)

void   mpfr_init2(mpfr_ptr, mpfr_prec_t);
void   mpfr_init(mpfr_ptr);
void   mpfr_clear(mpfr_ptr);
void   mpfr_set_prec(mpfr_ptr, mpfr_prec_t);
mpfr_prec_t mpfr_get_prec(mpfr_srcptr);
int    mpfr_set(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t);
int    mpfr_set_ui(mpfr_ptr, unsigned long, mpfr_rnd_t);
int    mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int    mpfr_set_ld(mpfr_ptr, long double, mpfr_rnd_t);
int    mpfr_set_str(mpfr_ptr, const char *, int, mpfr_rnd_t);
int    mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_sqr(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_div(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_abs(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_mul_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int    mpfr_mul_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int    mpfr_mul_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int    mpfr_mul_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int    mpfr_div_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int    mpfr_div_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int    mpfr_div_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int    mpfr_si_div(mpfr_ptr, long, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_log2(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int    mpfr_greater_p(mpfr_srcptr, mpfr_srcptr);
int    mpfr_cmp(mpfr_srcptr, mpfr_srcptr);
int    mpfr_cmp_si(mpfr_srcptr, long);
int    mpfr_nan_p(mpfr_srcptr);
int    mpfr_inf_p(mpfr_srcptr);
int    mpfr_zero_p(mpfr_srcptr);
int    mpfr_sgn(mpfr_srcptr);
long double mpfr_get_ld(mpfr_srcptr, mpfr_rnd_t);
double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
long   mpfr_get_si(mpfr_srcptr, mpfr_rnd_t);
mpfr_exp_t mpfr_get_exp(mpfr_srcptr);
char  *mpfr_get_str(char *, mpfr_exp_t *, int, size_t, mpfr_srcptr, mpfr_rnd_t);
void   mpfr_free_str(char *);
void   mpfr_free_cache(void);
const char *mpfr_get_version(void);
int    mpfr_printf(const char *, ...);
int    mpfr_snprintf(char *, size_t, const char *, ...);
int    mpfr_sprintf(char *, const char *, ...);
/* mpfr.h maps these two names onto internal symbols when <stdio.h> is seen */
size_t __gmpfr_out_str(FILE *, int, size_t, mpfr_srcptr, mpfr_rnd_t);
int    __gmpfr_fprintf(FILE *, const char *, ...);
#define mpfr_out_str __gmpfr_out_str
#define mpfr_fprintf __gmpfr_fprintf

#ifdef __cplusplus
}
#endif
#endif /* MDZ_MP_ABI_H */
