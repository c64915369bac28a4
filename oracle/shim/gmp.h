/* oracle/shim/gmp.h -- lets the reference sources `#include <gmp.h>` on a box
 * that has libgmp.so.10 but no headers.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_GMP_H
#define ORACLE_SHIM_GMP_H
#include "../../include/mdz_mp_abi.h"
#endif
