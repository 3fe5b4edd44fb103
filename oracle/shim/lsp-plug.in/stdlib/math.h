/*
 * oracle/shim -- TEST INFRASTRUCTURE.  Stand-in for lsp-common-lib's <lsp-plug.in/stdlib/math.h>
 * (not present offline): the C math library, which is all the reference sources compiled for the
 * oracle (src/main/misc/windows.cpp) take from it.
 */
#ifndef ORACLE_SHIM_STDLIB_MATH_H_
#define ORACLE_SHIM_STDLIB_MATH_H_

#include <math.h>

#endif /* ORACLE_SHIM_STDLIB_MATH_H_ */
