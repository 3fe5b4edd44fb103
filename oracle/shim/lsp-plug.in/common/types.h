/*
 * oracle/shim -- TEST INFRASTRUCTURE.  Stand-in for lsp-common-lib's
 * <lsp-plug.in/common/types.h> (lsp-common-lib 1.0.47 is not present offline,
 * reference modules.mk:23).  Only what the reference's Convolver.cpp,
 * Convolver.h, IStateDumper.h and IStateDumper.cpp use.
 */
#ifndef ORACLE_SHIM_COMMON_TYPES_H_
#define ORACLE_SHIM_COMMON_TYPES_H_

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#define LSP_EXPORT_MODIFIER     __attribute__((visibility("default")))
#define LSP_IMPORT_MODIFIER

namespace lsp
{
    /* Same-type pairs only, as at Convolver.cpp:92,141,169,186,275,290 */
    template <class T> inline T lsp_min(T a, T b)   { return (a < b) ? a : b; }
    template <class T> inline T lsp_max(T a, T b)   { return (a > b) ? a : b; }

    /* lsp_limit(ssize_t, int, int) at Convolver.cpp:87 */
    template <class A, class B, class C>
    inline A lsp_limit(A value, B lo, C hi)
    {
        return (value < A(lo)) ? A(lo) : (value > A(hi)) ? A(hi) : value;
    }
}

#endif /* ORACLE_SHIM_COMMON_TYPES_H_ */
