/*
 * oracle/equalizer_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the DATA PATH of lsp::dspu::Equalizer in its EQM_FIR / EQM_FFT modes
 * (scope-table row f2): one-partition overlap-add convolution with a block of latency and a
 * cross-fade when the kernel is replaced "smoothly".
 *
 *   buffers              reference src/main/filters/Equalizer.cpp:96-121 (init)
 *   kernel hand-over     Equalizer.cpp:336-345 (reconfigure: fastconv_parse into vConv, or into
 *                        vNewConv + EF_XFADE when EF_SMOOTH is set)
 *   clear                Equalizer.cpp:273-278 (EF_CLEAR)
 *   process              Equalizer.cpp:474-518 (EQM_FIR / EQM_FFT case)
 *
 * The filter DESIGN that produces the impulse response (FilterBank, Filter::freq_chart, windows)
 * is control-plane work outside the row; the oracle starts where the reference calls
 * fastconv_parse on the finished nFirSize-tap impulse response.
 *
 * Parity status: Equalizer.cpp cannot be compiled from its own few sources (it pulls in
 * Filter / FilterBank and ~40 more lsp-dsp-lib functions), so this is a restatement only.  It is
 * pinned by identity where the domain offers one -- without a cross-fade the output is the direct
 * convolution delayed by nFirSize samples, and the reference's own unit test
 * (src/test/utest/filters/equalizer.cpp:34-84) pins "index of the response peak == latency" --
 * while the cross-fade arithmetic (Equalizer.cpp:486-501) and the semantics of the absent
 * dsp::lramp1 / dsp::lramp_add2 (taken from their names and argument lists:
 * dst[i] *= v1 + (v2-v1)*i/count, dst[i] += src[i] * (v1 + (v2-v1)*i/count)) are PARITY UNPINNED.
 *
 * Only tests/ may call into this file.
 */
#ifndef ORACLE_EQUALIZER_ORACLE_H_
#define ORACLE_EQUALIZER_ORACLE_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_eq orc_eq_t;

orc_eq_t   *orc_eq_create(size_t fir_rank);                 /* nFirSize = 1 << fir_rank; NULL on OOM */
void        orc_eq_free(orc_eq_t *e);
/* ir: nFirSize taps.  smooth != 0: cross-fade to it at the next block boundary */
void        orc_eq_set_kernel(orc_eq_t *e, const float *ir, int smooth);
void        orc_eq_clear(orc_eq_t *e);
void        orc_eq_process(orc_eq_t *e, float *out, const float *in, size_t samples);
size_t      orc_eq_fir_size(const orc_eq_t *e);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_EQUALIZER_ORACLE_H_ */
