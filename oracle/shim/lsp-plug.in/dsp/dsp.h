/*
 * oracle/shim -- TEST INFRASTRUCTURE.  Stand-in for lsp-dsp-lib's
 * <lsp-plug.in/dsp/dsp.h> (lsp-dsp-lib 1.0.36 is not present offline, reference
 * modules.mk:29-33): the eight lsp::dsp:: entry points on the Convolver path,
 * forwarded to the restated scalar kernels in oracle/dsp_restated.c.
 */
#ifndef ORACLE_SHIM_DSP_DSP_H_
#define ORACLE_SHIM_DSP_DSP_H_

#include <lsp-plug.in/common/types.h>
#include "dsp_restated.h"

namespace lsp
{
    namespace dsp
    {
        inline void init()  { rs_dsp_init(); }

        inline void fill_zero(float *dst, size_t count)                 { rs_fill_zero(dst, count); }
        inline void copy(float *dst, const float *src, size_t count)    { rs_copy(dst, src, count); }
        inline void move(float *dst, const float *src, size_t count)    { rs_move(dst, src, count); }

        inline void convolve(float *dst, const float *src, const float *conv, size_t length, size_t count)
            { rs_convolve(dst, src, conv, length, count); }

        inline void fastconv_parse(float *dst, const float *src, size_t rank)
            { rs_fastconv_parse(dst, src, rank); }
        inline void fastconv_apply(float *dst, float *tmp, const float *c1, const float *c2, size_t rank)
            { rs_fastconv_apply(dst, tmp, c1, c2, rank); }
        inline void fastconv_parse_apply(float *dst, float *tmp, const float *c, const float *src, size_t rank)
            { rs_fastconv_parse_apply(dst, tmp, c, src, rank); }
        inline void fastconv_restore(float *dst, float *src, size_t rank)
            { rs_fastconv_restore(dst, src, rank); }

        /* SpectralProcessor.cpp:163-183 */
        inline void mul3(float *dst, const float *a, const float *b, size_t count)     { rs_mul3(dst, a, b, count); }
        inline void fmadd3(float *dst, const float *a, const float *b, size_t count)   { rs_fmadd3(dst, a, b, count); }
        inline void pcomplex_r2c(float *dst, const float *src, size_t count)           { rs_pcomplex_r2c(dst, src, count); }
        inline void pcomplex_c2r(float *dst, const float *src, size_t count)           { rs_pcomplex_c2r(dst, src, count); }
        inline void packed_direct_fft(float *dst, const float *src, size_t rank)       { rs_packed_direct_fft(dst, src, rank); }
        inline void packed_reverse_fft(float *dst, const float *src, size_t rank)      { rs_packed_reverse_fft(dst, src, rank); }
    }
}

#endif /* ORACLE_SHIM_DSP_DSP_H_ */
