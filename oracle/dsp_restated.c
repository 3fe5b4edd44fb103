/*
 * oracle/dsp_restated.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See dsp_restated.h for the contract and the reference call sites it follows.
 */
#include "dsp_restated.h"

#include <math.h>
#include <pthread.h>
#include <string.h>

#define RS_MAX_RANK     16
#define RS_MAX_N        (1u << RS_MAX_RANK)

/* Twiddles shared by all ranks: for a butterfly half-span h (power of two) the
 * factor exp(-2*pi*i*j/(2h)), j < h, sits at index h + j. */
static float rs_tw_re[RS_MAX_N];
static float rs_tw_im[RS_MAX_N];
static pthread_once_t rs_once = PTHREAD_ONCE_INIT;

static void rs_build_tables(void)
{
    for (size_t h = 1; h < RS_MAX_N; h <<= 1)
    {
        for (size_t j = 0; j < h; ++j)
        {
            double a        = -M_PI * (double)j / (double)h;
            rs_tw_re[h + j] = (float)cos(a);
            rs_tw_im[h + j] = (float)sin(a);
        }
    }
}

void rs_dsp_init(void)
{
    pthread_once(&rs_once, rs_build_tables);
}

void rs_fill_zero(float *dst, size_t count)
{
    memset(dst, 0, count * sizeof(float));
}

void rs_copy(float *dst, const float *src, size_t count)
{
    if (dst != src)
        memmove(dst, src, count * sizeof(float));
}

void rs_move(float *dst, const float *src, size_t count)
{
    memmove(dst, src, count * sizeof(float));
}

void rs_convolve(float *dst, const float *src, const float *conv, size_t length, size_t count)
{
    for (size_t i = 0; i < count; ++i)
    {
        float k     = src[i];
        float *d    = &dst[i];
        for (size_t j = 0; j < length; ++j)
            d[j]       += k * conv[j];
    }
}

/* One decimation-in-frequency pass with half-span h over n points. */
static void rs_dif_pass(float *re, float *im, size_t n, size_t h)
{
    if (h == 1)
    {
        for (size_t b = 0; b < n; b += 2)
        {
            float ar = re[b], ai = im[b], cr = re[b+1], ci = im[b+1];
            re[b]   = ar + cr;  im[b]   = ai + ci;
            re[b+1] = ar - cr;  im[b+1] = ai - ci;
        }
        return;
    }

    const float *wr = &rs_tw_re[h];
    const float *wi = &rs_tw_im[h];
    for (size_t b = 0; b < n; b += 2*h)
    {
        float *r0 = &re[b], *i0 = &im[b], *r1 = &re[b+h], *i1 = &im[b+h];
        for (size_t j = 0; j < h; ++j)
        {
            float dr = r0[j] - r1[j];
            float di = i0[j] - i1[j];
            r0[j]   += r1[j];
            i0[j]   += i1[j];
            r1[j]    = dr * wr[j] - di * wi[j];
            i1[j]    = dr * wi[j] + di * wr[j];
        }
    }
}

/* One decimation-in-time pass (inverse transform, conjugated twiddles). */
static void rs_dit_pass(float *re, float *im, size_t n, size_t h)
{
    if (h == 1)
    {
        for (size_t b = 0; b < n; b += 2)
        {
            float ar = re[b], ai = im[b], cr = re[b+1], ci = im[b+1];
            re[b]   = ar + cr;  im[b]   = ai + ci;
            re[b+1] = ar - cr;  im[b+1] = ai - ci;
        }
        return;
    }

    const float *wr = &rs_tw_re[h];
    const float *wi = &rs_tw_im[h];
    for (size_t b = 0; b < n; b += 2*h)
    {
        float *r0 = &re[b], *i0 = &im[b], *r1 = &re[b+h], *i1 = &im[b+h];
        for (size_t j = 0; j < h; ++j)
        {
            /* c = x1 * conj(w) */
            float cr = r1[j] * wr[j] + i1[j] * wi[j];
            float ci = i1[j] * wr[j] - r1[j] * wi[j];
            r1[j]    = r0[j] - cr;
            i1[j]    = i0[j] - ci;
            r0[j]   += cr;
            i0[j]   += ci;
        }
    }
}

void rs_fastconv_parse(float *dst, const float *src, size_t rank)
{
    rs_dsp_init();

    size_t n    = (size_t)1 << rank;
    size_t h    = n >> 1;
    float *re   = dst;
    float *im   = &dst[n];

    /* First pass: the upper half of the input is zero padding and the input is
     * real, so x[j] stays and x[j+h] = x[j] * w^j. */
    if (h == 0)
    {
        re[0] = src[0]; im[0] = 0.0f;
        return;
    }
    if (h == 1)
    {
        re[0] = src[0]; im[0] = 0.0f;
        re[1] = src[0]; im[1] = 0.0f;
        return;
    }
    {
        const float *wr = &rs_tw_re[h];
        const float *wi = &rs_tw_im[h];
        for (size_t j = 0; j < h; ++j)
        {
            float a     = src[j];
            re[j]       = a;
            im[j]       = 0.0f;
            re[j+h]     = a * wr[j];
            im[j+h]     = a * wi[j];
        }
    }

    for (h >>= 1; h > 0; h >>= 1)
        rs_dif_pass(re, im, n, h);
}

static void rs_inverse(float *re, float *im, size_t n)
{
    for (size_t h = 1; h < n; h <<= 1)
        rs_dit_pass(re, im, n, h);
}

void rs_fastconv_apply(float *dst, float *tmp, const float *c1, const float *c2, size_t rank)
{
    rs_dsp_init();

    size_t n        = (size_t)1 << rank;
    const float *ar = c1, *ai = &c1[n];
    const float *br = c2, *bi = &c2[n];
    float *re       = tmp, *im = &tmp[n];

    for (size_t i = 0; i < n; ++i)
    {
        float xr    = ar[i] * br[i] - ai[i] * bi[i];
        float xi    = ar[i] * bi[i] + ai[i] * br[i];
        re[i]       = xr;
        im[i]       = xi;
    }

    rs_inverse(re, im, n);

    float k         = 1.0f / (float)n;
    for (size_t i = 0; i < n; ++i)
        dst[i]         += re[i] * k;
}

void rs_fastconv_parse_apply(float *dst, float *tmp, const float *c, const float *src, size_t rank)
{
    rs_fastconv_parse(tmp, src, rank);
    rs_fastconv_apply(dst, tmp, c, tmp, rank);
}

void rs_fastconv_restore(float *dst, float *src, size_t rank)
{
    rs_dsp_init();

    size_t n    = (size_t)1 << rank;
    rs_inverse(src, &src[n], n);

    float k     = 1.0f / (float)n;
    for (size_t i = 0; i < n; ++i)
        dst[i]      = src[i] * k;
}

/* ------------------------------------------------------------------------------------------- */
/* The primitives lsp::dspu::SpectralProcessor calls (SpectralProcessor.cpp:163-183): element-wise
 * products, real <-> packed complex, and the packed (interleaved re / im, natural bin order)
 * complex FFT pair.  The reverse transform is normalised by 1 / N, as the reference relies on
 * (window -> FFT -> IFFT -> window -> overlap-add reproduces the input when the callback leaves
 * the spectrum alone). */

void rs_mul3(float *dst, const float *a, const float *b, size_t count)
{
    for (size_t i = 0; i < count; ++i)
        dst[i]  = a[i] * b[i];
}

void rs_fmadd3(float *dst, const float *a, const float *b, size_t count)
{
    for (size_t i = 0; i < count; ++i)
        dst[i] += a[i] * b[i];
}

void rs_pcomplex_r2c(float *dst, const float *src, size_t count)
{
    /* forwards: the reference converts in place with src = dst + count (SpectralProcessor.cpp:164),
     * where the write index 2 i + 1 never passes the read index count + i */
    for (size_t i = 0; i < count; ++i)
    {
        float v     = src[i];
        dst[2*i]    = v;
        dst[2*i+1]  = 0.0f;
    }
}

void rs_pcomplex_c2r(float *dst, const float *src, size_t count)
{
    for (size_t i = 0; i < count; ++i)      /* forwards: dst may alias src */
        dst[i]      = src[2*i];
}

static void rs_packed_fft(float *dst, const float *src, size_t rank, int inverse)
{
    rs_dsp_init();
    const size_t n = (size_t)1 << rank;
    /* bit-reversal copy (in place when dst == src) */
    if (dst != src)
    {
        for (size_t i = 0; i < n; ++i)
        {
            size_t j = 0;
            for (size_t b = 0; b < rank; ++b)
                j  |= ((i >> b) & 1) << (rank - 1 - b);
            dst[2*j]    = src[2*i];
            dst[2*j+1]  = src[2*i+1];
        }
    }
    else
    {
        for (size_t i = 0; i < n; ++i)
        {
            size_t j = 0;
            for (size_t b = 0; b < rank; ++b)
                j  |= ((i >> b) & 1) << (rank - 1 - b);
            if (j > i)
            {
                float tr = dst[2*i], ti = dst[2*i+1];
                dst[2*i] = dst[2*j]; dst[2*i+1] = dst[2*j+1];
                dst[2*j] = tr;       dst[2*j+1] = ti;
            }
        }
    }
    /* decimation in time, natural order out */
    for (size_t h = 1; h < n; h <<= 1)
    {
        const float *wr = &rs_tw_re[h], *wi = &rs_tw_im[h];
        for (size_t b = 0; b < n; b += 2*h)
            for (size_t j = 0; j < h; ++j)
            {
                float c = wr[j], s = inverse ? -wi[j] : wi[j];
                float *p = &dst[2*(b + j)], *q = &dst[2*(b + j + h)];
                float tr = q[0] * c - q[1] * s;
                float ti = q[0] * s + q[1] * c;
                q[0] = p[0] - tr;  q[1] = p[1] - ti;
                p[0] += tr;        p[1] += ti;
            }
    }
    if (inverse)
    {
        const float k = 1.0f / (float)n;
        for (size_t i = 0; i < 2*n; ++i)
            dst[i]     *= k;
    }
}

void rs_packed_direct_fft(float *dst, const float *src, size_t rank)   { rs_packed_fft(dst, src, rank, 0); }
void rs_packed_reverse_fft(float *dst, const float *src, size_t rank)  { rs_packed_fft(dst, src, rank, 1); }
