/*
 * oracle/chirp_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the offline partitioned linear convolution of
 * lsp::dspu::SyncChirpProcessor (reference src/main/util/SyncChirpProcessor.cpp):
 *   calculateConvolutionPartitionSize   :1224-1250
 *   calculateConvolutionParameters      :1299-1331
 *   do_linear_convolutions              :1374-1404
 *   do_linear_convolution               :1406-1508
 * over the restated dsp:: kernels (dsp_restated.h).  dspu::Sample is replaced by plain arrays:
 * `inputs[ch]` = data[ch]->channel(0, offset[ch]), `in_len[ch]` = data[ch]->length() - offset[ch].
 *
 * PARITY UNPINNED against a build of the reference: SyncChirpProcessor.cpp needs dspu::Sample,
 * lsp-runtime-lib and more of lsp-dsp-lib than can be compiled from a few files, and the reference
 * holds no golden vectors for this operator (only a manual test, src/test/mtest/util/sync_chirp.cpp).
 * The restatement is pinned by identity instead (tests/test_oracle_chirp.py): the result is the
 * float64 linear convolution of the zero-padded input with the prepend-padded inverse filter, at
 * the align offsets, scaled over the first vConvLengths samples.
 *
 * Only tests/ may call into this file.
 */
#ifndef ORACLE_CHIRP_ORACLE_H_
#define ORACLE_CHIRP_ORACLE_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CHIRP_MAX_PART_SIZE     32768       /* MAX_PART_SIZE, SyncChirpProcessor.cpp:43 */

typedef struct orc_chirp_plan
{
    size_t  partition_size;     /* sConvParams.nPartitionSize */
    size_t  conv_rank;          /* sConvParams.nConvRank      */
    size_t  image;              /* sConvParams.nImage         */
    size_t  allocation_size;    /* sConvParams.nAllocationSize: samples per result channel */
} orc_chirp_plan_t;

/* Per channel (arrays of nchannels entries, caller-allocated): vPartitions, vPaddedLengths,
 * vInversePrepends, vConvLengths, vAlignOffsets. */
void    orc_chirp_plan(orc_chirp_plan_t *plan, size_t *partitions, size_t *padded, size_t *prepends,
                       size_t *conv_lengths, size_t *align_offsets, const size_t *in_len,
                       size_t nchannels, size_t inverse_len, size_t part_size_limit);

/* do_linear_convolutions: result is [nchannels][allocation_size] floats, zeroed here
 * (allocateConvolutionResult); scale = fConvScale / (nSampleRate * nSampleRate).
 * Returns 0 on success, -1 on bad arguments / allocation failure. */
int     orc_chirp_linear_convolutions(float *result, const float *const *inputs, const size_t *in_len,
                                      size_t nchannels, const float *inverse, size_t inverse_len,
                                      size_t part_size_limit, float scale);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_CHIRP_ORACLE_H_ */
