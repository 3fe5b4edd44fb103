/*
 * oracle/shim -- TEST INFRASTRUCTURE.  Stand-in for lsp-common-lib's
 * <lsp-plug.in/common/status.h> (lsp-common-lib 1.0.47 is not present offline, reference
 * modules.mk:23).  Only the codes that SpectralSplitter.cpp returns; apart from STATUS_OK = 0 the
 * numeric values are stand-ins (the checker only tells success from failure).
 */
#ifndef ORACLE_SHIM_COMMON_STATUS_H_
#define ORACLE_SHIM_COMMON_STATUS_H_

#include <lsp-plug.in/common/types.h>

namespace lsp
{
    typedef int status_t;

    enum status_codes
    {
        STATUS_OK               = 0,
        STATUS_INVALID_VALUE    = 1001,
        STATUS_OVERFLOW         = 1002,
        STATUS_NOT_BOUND        = 1003
    };
}

#endif /* ORACLE_SHIM_COMMON_STATUS_H_ */
