/*
 * oracle/equalizer_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See equalizer_oracle.h.
 */
#include "equalizer_oracle.h"
#include "dsp_restated.h"

#include <stdlib.h>
#include <string.h>

struct orc_eq
{
    size_t  nFirSize, nFirRank, nBufSize;
    int     xfade;
    float  *vInBuffer;      /* nFirSize * 2 */
    float  *vOutBuffer;     /* nFirSize * 2 */
    float  *vConv;          /* nFirSize * 4 */
    float  *vNewConv;       /* nFirSize * 4 */
    float  *vFft;           /* nFirSize * 4 */
    float  *vTemp;          /* nFirSize * 4 */
    float  *slab;
};

/* dsp::lramp1 / dsp::lramp_add2 of lsp-dsp-lib (absent): linear ramp from v1 towards v2 over
 * `count` samples, end point excluded */
static void lramp1(float *dst, float v1, float v2, size_t count)
{
    float delta = (v2 - v1) / (float)count;
    for (size_t i = 0; i < count; ++i)
        dst[i] = dst[i] * (v1 + delta * (float)i);
}

static void lramp_add2(float *dst, const float *src, float v1, float v2, size_t count)
{
    float delta = (v2 - v1) / (float)count;
    for (size_t i = 0; i < count; ++i)
        dst[i] = dst[i] + src[i] * (v1 + delta * (float)i);
}

/* Equalizer.cpp:96-121 */
orc_eq_t *orc_eq_create(size_t fir_rank)
{
    rs_dsp_init();
    orc_eq_t *e = (orc_eq_t *)calloc(1, sizeof(orc_eq_t));
    if (e == NULL)
        return NULL;
    e->nFirRank     = fir_rank;
    e->nFirSize     = (size_t)1 << fir_rank;
    size_t fft_size = e->nFirSize << 1, conv_size = e->nFirSize << 2;
    e->slab         = (float *)calloc(fft_size * 2 + conv_size * 4, sizeof(float));
    if (e->slab == NULL)
    {
        free(e);
        return NULL;
    }
    float *p        = e->slab;
    e->vInBuffer    = p; p += fft_size;
    e->vOutBuffer   = p; p += fft_size;
    e->vConv        = p; p += conv_size;
    e->vNewConv     = p; p += conv_size;
    e->vFft         = p; p += conv_size;
    e->vTemp        = p;
    return e;
}

void orc_eq_free(orc_eq_t *e)
{
    if (e == NULL)
        return;
    free(e->slab);
    free(e);
}

size_t orc_eq_fir_size(const orc_eq_t *e)
{
    return e->nFirSize;
}

/* Equalizer.cpp:336-345 */
void orc_eq_set_kernel(orc_eq_t *e, const float *ir, int smooth)
{
    if (smooth)
    {
        e->xfade    = 1;
        rs_fastconv_parse(e->vNewConv, ir, e->nFirRank + 1);
    }
    else
        rs_fastconv_parse(e->vConv, ir, e->nFirRank + 1);
}

/* Equalizer.cpp:273-278 */
void orc_eq_clear(orc_eq_t *e)
{
    rs_fill_zero(e->vInBuffer, e->nFirSize << 1);
    rs_fill_zero(e->vOutBuffer, e->nFirSize << 1);
    e->nBufSize = 0;
}

/* Equalizer.cpp:474-518 */
void orc_eq_process(orc_eq_t *e, float *out, const float *in, size_t samples)
{
    const size_t nFirSize = e->nFirSize, conv_rank = e->nFirRank + 1;

    while (samples > 0)
    {
        if (e->nBufSize >= nFirSize)
        {
            rs_move(e->vOutBuffer, &e->vOutBuffer[nFirSize], nFirSize);
            rs_fill_zero(&e->vOutBuffer[nFirSize], nFirSize);
            rs_fastconv_parse_apply(e->vOutBuffer, e->vTemp, e->vConv, e->vInBuffer, conv_rank);

            if (e->xfade)
            {
                size_t half = nFirSize >> 1;

                rs_fill_zero(e->vFft, nFirSize * 2);
                rs_copy(e->vConv, e->vNewConv, nFirSize * 4);
                rs_fastconv_parse_apply(e->vFft, e->vTemp, e->vConv, e->vInBuffer, conv_rank);

                lramp1(&e->vOutBuffer[half], 1.0f, 0.0f, nFirSize);
                lramp_add2(&e->vOutBuffer[half], &e->vFft[half], 0.0f, 1.0f, nFirSize);
                rs_copy(&e->vOutBuffer[nFirSize + half], &e->vFft[nFirSize + half], half);

                e->xfade    = 0;
            }
            e->nBufSize = 0;
        }

        size_t to_process = samples < nFirSize - e->nBufSize ? samples : nFirSize - e->nBufSize;
        rs_copy(&e->vInBuffer[e->nBufSize], in, to_process);
        rs_copy(out, &e->vOutBuffer[e->nBufSize], to_process);

        e->nBufSize    += to_process;
        out            += to_process;
        in             += to_process;
        samples        -= to_process;
    }
}
