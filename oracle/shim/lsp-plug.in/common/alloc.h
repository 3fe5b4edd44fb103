/*
 * oracle/shim -- TEST INFRASTRUCTURE.  Stand-in for lsp-common-lib's
 * <lsp-plug.in/common/alloc.h>: alloc_aligned / free_aligned as used at
 * reference Convolver.cpp:73,104, SpectralProcessor.cpp:72,80 and SpectralSplitter.cpp:90-100.
 */
#ifndef ORACLE_SHIM_COMMON_ALLOC_H_
#define ORACLE_SHIM_COMMON_ALLOC_H_

#include <lsp-plug.in/common/types.h>
#include <stdlib.h>

#ifndef DEFAULT_ALIGN
    #define DEFAULT_ALIGN       0x10        /* lsp-common-lib's default (used at SpectralProcessor.cpp:72) */
#endif

namespace lsp
{
    template <class T>
    inline T *alloc_aligned(uint8_t * &raw, size_t count, size_t align)
    {
        raw             = static_cast<uint8_t *>(::malloc(count * sizeof(T) + align));
        if (raw == NULL)
            return NULL;
        uintptr_t p     = reinterpret_cast<uintptr_t>(raw);
        uintptr_t rem   = p % align;
        if (rem != 0)
            p              += align - rem;
        return reinterpret_cast<T *>(p);
    }

    /* rounds `size` up to a multiple of `align` (SpectralSplitter.cpp:90) */
    inline size_t align_size(size_t size, size_t align)
    {
        size_t rem      = size % align;
        return (rem == 0) ? size : size + align - rem;
    }

    inline void free_aligned(uint8_t * &raw)
    {
        if (raw != NULL)
            ::free(raw);
        raw             = NULL;
    }
}

#endif /* ORACLE_SHIM_COMMON_ALLOC_H_ */
