/*
 * oracle/chirp_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE (see chirp_oracle.h).
 *
 * Restatement of SyncChirpProcessor's O(P^2) partitioned linear convolution.  Line tags refer to
 * /root/reference/src/main/util/SyncChirpProcessor.cpp.
 */
#include "chirp_oracle.h"
#include "dsp_restated.h"

#include <stdlib.h>
#include <sys/types.h>

void orc_chirp_plan(orc_chirp_plan_t *plan, size_t *partitions, size_t *padded, size_t *prepends,
                    size_t *conv_lengths, size_t *align_offsets, const size_t *in_len,
                    size_t nchannels, size_t inverse_len, size_t part_size_limit)
{
    /* :1226-1227 */
    part_size_limit = (part_size_limit < ORC_CHIRP_MAX_PART_SIZE) ? part_size_limit : ORC_CHIRP_MAX_PART_SIZE;
    part_size_limit = (part_size_limit == 0) ? ORC_CHIRP_MAX_PART_SIZE : part_size_limit;

    /* :1230-1240 : power of two, rank of each partition convolution */
    size_t part = 1, exponent = 0;
    while (part < part_size_limit)
    {
        part      <<= 1;
        ++exponent;
    }
    plan->partition_size    = part;
    plan->conv_rank         = exponent + 1;
    plan->image             = (size_t)1 << (plan->conv_rank + 1);

    /* :1301-1324 */
    plan->allocation_size   = 0;
    for (size_t ch = 0; ch < nchannels; ++ch)
    {
        size_t n_in     = in_len[ch];
        size_t n_max    = (n_in > inverse_len) ? n_in : inverse_len;
        partitions[ch]  = (n_max / part) + 1;
        padded[ch]      = partitions[ch] * part;
        prepends[ch]    = padded[ch] - inverse_len;
        conv_lengths[ch] = 2 * padded[ch];
        if (conv_lengths[ch] > plan->allocation_size)
            plan->allocation_size = conv_lengths[ch];
    }

    /* :1327-1330 */
    size_t middle = (plan->allocation_size / 2) - 1;
    for (size_t ch = 0; ch < nchannels; ++ch)
        align_offsets[ch] = middle - (conv_lengths[ch] / 2) + 1;
}

/* do_linear_convolution, :1406-1508 */
static void chirp_channel(float *result, const float *input, size_t n_input, const float *inverse,
                          size_t partitions, size_t prepend, size_t align, size_t conv_length,
                          const orc_chirp_plan_t *plan, float *in_part, float *inv_part,
                          float *in_image, float *inv_image, float *temp, float scale)
{
    const size_t P = plan->partition_size, rank = plan->conv_rank;

    rs_fill_zero(in_part, P);                       /* :1414-1418 */
    rs_fill_zero(inv_part, P);
    rs_fill_zero(in_image, plan->image);
    rs_fill_zero(inv_image, plan->image);
    rs_fill_zero(temp, plan->image);

    int null_in = 0, null_inv = 0;                  /* :1434-1435 */
    for (size_t inp = 0; inp < partitions; ++inp)   /* :1437 */
    {
        size_t input_head   = inp * P;
        ssize_t ahead       = (ssize_t)n_input - (ssize_t)input_head;       /* :1443 */

        if (ahead > (ssize_t)P)                     /* :1445-1449 */
        {
            null_in         = 0;
            rs_fastconv_parse(in_image, &input[input_head], rank);
        }
        else if (ahead > 0)                         /* :1450-1456 */
        {
            null_in         = 0;
            rs_copy(in_part, &input[input_head], (size_t)ahead);
            rs_fill_zero(&in_part[ahead], P - (size_t)ahead);
            rs_fastconv_parse(in_image, in_part, rank);
        }
        else                                        /* :1457-1460 */
            null_in         = 1;

        size_t inverse_head = 0;                    /* :1465 */
        for (size_t invp = 0; invp < partitions; ++invp)    /* :1467 */
        {
            size_t virtual_head = invp * P;
            ssize_t pad_ahead   = (ssize_t)prepend - (ssize_t)virtual_head;     /* :1474 */

            if (pad_ahead > (ssize_t)P)             /* :1476-1479 */
                null_inv        = 1;
            else if (pad_ahead > 0)                 /* :1480-1488 */
            {
                null_inv        = 0;
                size_t to_copy  = P - (size_t)pad_ahead;
                rs_fill_zero(inv_part, (size_t)pad_ahead);
                rs_copy(&inv_part[pad_ahead], &inverse[inverse_head], to_copy);
                rs_fastconv_parse(inv_image, inv_part, rank);
                inverse_head   += to_copy;
            }
            else                                    /* :1489-1494 */
            {
                null_inv        = 0;
                rs_fastconv_parse(inv_image, &inverse[inverse_head], rank);
                inverse_head   += P;
            }

            if (null_in || null_inv)                /* :1496-1497 */
                continue;

            rs_fastconv_apply(&result[P * (inp + invp) + align], temp, in_image, inv_image, rank);  /* :1499-1504 */
        }
    }

    /* :1508 : dsp::mul_k2(vResult, k, vConvLengths[channel]) -- from index 0, not from the align offset */
    for (size_t i = 0; i < conv_length; ++i)
        result[i]          *= scale;
}

int orc_chirp_linear_convolutions(float *result, const float *const *inputs, const size_t *in_len,
                                  size_t nchannels, const float *inverse, size_t inverse_len,
                                  size_t part_size_limit, float scale)
{
    if ((result == NULL) || (inputs == NULL) || (in_len == NULL) || (nchannels == 0) || (inverse == NULL))
        return -1;                                  /* :1376-1377 */
    rs_dsp_init();

    size_t *tab = (size_t *)malloc(5 * nchannels * sizeof(size_t));
    if (tab == NULL)
        return -1;
    size_t *partitions = tab, *padded = tab + nchannels, *prepends = tab + 2 * nchannels,
           *conv_lengths = tab + 3 * nchannels, *aligns = tab + 4 * nchannels;
    orc_chirp_plan_t plan;
    orc_chirp_plan(&plan, partitions, padded, prepends, conv_lengths, aligns, in_len, nchannels,
                   inverse_len, part_size_limit);   /* :1379,1385 */

    /* allocateConvolutionTempArrays, :1341-1358 */
    size_t samples  = 2 * plan.partition_size + 3 * plan.image;
    float *tmp      = (float *)calloc(samples, sizeof(float));
    if (tmp == NULL)
    {
        free(tab);
        return -1;
    }
    float *in_part  = tmp, *inv_part = in_part + plan.partition_size;
    float *in_image = inv_part + plan.partition_size, *inv_image = in_image + plan.image;
    float *temp     = inv_image + plan.image;

    rs_fill_zero(result, nchannels * plan.allocation_size);     /* allocateConvolutionResult, :1387 */
    for (size_t ch = 0; ch < nchannels; ++ch)       /* :1395-1401 */
        chirp_channel(result + ch * plan.allocation_size, inputs[ch], in_len[ch], inverse, partitions[ch],
                      prepends[ch], aligns[ch], conv_lengths[ch], &plan, in_part, inv_part, in_image,
                      inv_image, temp, scale);
    free(tmp);
    free(tab);
    return 0;
}
