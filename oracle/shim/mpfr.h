/* oracle/shim/mpfr.h -- lets the reference sources `#include <mpfr.h>` on a box
 * that has libmpfr.so.6 but no headers.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_MPFR_H
#define ORACLE_SHIM_MPFR_H
#include "../../include/mdz_mp_abi.h"
#endif
