/*
 * oracle/dsp_restated.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the eight lsp::dsp:: primitives that the reference
 * dspu::Convolver calls.  The real implementations live in lsp-dsp-lib 1.0.36
 * (reference modules.mk:29-33), which is fetched by `git clone` at build time
 * and is NOT present under /root/reference, so the contract below is derived
 * from the reference's call sites and its own unit test:
 *
 *   fill_zero / copy / move   src/main/util/Convolver.cpp:110,156-158,291,296,308-310
 *   convolve                  src/main/util/Convolver.cpp:295,
 *                             src/test/utest/util/convolver.cpp:32-40 (identical helper)
 *   fastconv_parse            src/main/util/Convolver.cpp:159,174,191,270
 *   fastconv_parse_apply      src/main/util/Convolver.cpp:256,293
 *   fastconv_apply            src/main/util/Convolver.cpp:282
 *   fastconv_restore          no call site in the reference tree
 *
 * The "image" produced by fastconv_parse is opaque to every caller (they only
 * multiply two images and restore), so any self-consistent layout is valid.
 * Layout used here: 2^rank real parts followed by 2^rank imaginary parts of the
 * full complex FFT of size 2^rank, in bit-reversed bin order (forward transform
 * is decimation-in-frequency without the final reorder, inverse transform is
 * decimation-in-time consuming that order) -- the same "no reorder" idea the
 * reference library is known to use, at the same image size 2^(rank+1) floats
 * (Convolver.cpp:91,99; Equalizer.cpp:99-121).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call into this file.
 */
#ifndef ORACLE_DSP_RESTATED_H_
#define ORACLE_DSP_RESTATED_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Builds the per-rank twiddle tables (ranks 1..16).  Thread-safe, idempotent. */
void rs_dsp_init(void);

void rs_fill_zero(float *dst, size_t count);
void rs_copy(float *dst, const float *src, size_t count);
void rs_move(float *dst, const float *src, size_t count);

/* dst[i+j] += src[i] * conv[j]   for i < count, j < length */
void rs_convolve(float *dst, const float *src, const float *conv, size_t length, size_t count);

/* dst[2^(rank+1)] = image of FFT_{2^rank}([src[0 .. 2^(rank-1)), zeros]) */
void rs_fastconv_parse(float *dst, const float *src, size_t rank);

/* tmp = c1 (*) c2 ; dst[0 .. 2^rank) += Re(IFFT(tmp)) / 2^rank ; tmp clobbered */
void rs_fastconv_apply(float *dst, float *tmp, const float *c1, const float *c2, size_t rank);

/* tmp = parse(src) ; then as rs_fastconv_apply(dst, tmp, c, tmp) */
void rs_fastconv_parse_apply(float *dst, float *tmp, const float *c, const float *src, size_t rank);

/* dst[0 .. 2^rank) = Re(IFFT(src)) / 2^rank ; src clobbered (store, not add) */
void rs_fastconv_restore(float *dst, float *src, size_t rank);

/* ---- the primitives of lsp::dspu::SpectralProcessor (SpectralProcessor.cpp:163-183) ---------- */
void rs_mul3(float *dst, const float *a, const float *b, size_t count);            /* dst = a * b   */
void rs_fmadd3(float *dst, const float *a, const float *b, size_t count);          /* dst += a * b  */
void rs_pcomplex_r2c(float *dst, const float *src, size_t count);                  /* (re, 0) pairs */
void rs_pcomplex_c2r(float *dst, const float *src, size_t count);                  /* real parts    */
/* complex FFT of 2^rank points, interleaved re / im, natural order in and out; dst may equal src;
 * the reverse transform is scaled by 1 / 2^rank */
void rs_packed_direct_fft(float *dst, const float *src, size_t rank);
void rs_packed_reverse_fft(float *dst, const float *src, size_t rank);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_DSP_RESTATED_H_ */
