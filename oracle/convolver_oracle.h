/*
 * oracle/convolver_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference zero-latency partitioned convolver
 * lsp::dspu::Convolver (reference include/lsp-plug.in/dsp-units/util/Convolver.h:35-114,
 * src/main/util/Convolver.cpp:36-340) on top of the restated dsp primitives of
 * dsp_restated.h.
 *
 * PINNING.  The reference holds no golden vectors for this path; its own unit
 * test (src/test/utest/util/convolver.cpp) pins the convolver against naive
 * direct convolution.  This restatement is pinned
 *   (1) against the same identities at the same shapes and tolerances
 *       (tests/test_oracle.py: test_small / test_large / test_collisions),
 *   (2) bit-for-bit against the reference's own Convolver.cpp compiled verbatim
 *       from /root/reference over the same primitives (oracle/_ref, built by
 *       oracle/Makefile; tests/test_oracle.py::test_restatement_matches_verbatim_reference),
 *   (3) against committed fixtures generated from that verbatim build
 *       (tests/golden/, script tests/golden/make_golden.py).
 * What is NOT pinned: bit patterns of lsp-dsp-lib's SIMD fastconv_* kernels
 * (that library is absent offline); parity is numerical, <= 1e-5 of peak.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call into this file.
 */
#ifndef ORACLE_CONVOLVER_ORACLE_H_
#define ORACLE_CONVOLVER_ORACLE_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_RANK_MIN    8       /* Convolver.h:28 */
#define ORC_RANK_MAX    16      /* Convolver.h:29 */

typedef struct orc_convolver orc_convolver_t;

/* Convolver::Convolver / construct  (Convolver.cpp:36-69) */
orc_convolver_t *orc_create(void);
/* Convolver::~Convolver             (Convolver.cpp:41-44)  */
void    orc_free(orc_convolver_t *c);
/* Convolver::destroy                (Convolver.cpp:71-75)  */
void    orc_destroy(orc_convolver_t *c);
/* Convolver::init                   (Convolver.cpp:77-215); returns 1 = true, 0 = false */
int     orc_init(orc_convolver_t *c, const float *data, size_t count, size_t rank, float phase);
/* Convolver::process                (Convolver.cpp:217-313) */
void    orc_process(orc_convolver_t *c, float *dst, const float *src, size_t count);
/* Convolver::data_size / rank       (Convolver.h:101,107)  */
size_t  orc_data_size(const orc_convolver_t *c);
size_t  orc_rank(const orc_convolver_t *c);

/* Scheduler state, for tests of the frame bookkeeping (Convolver.h:45-55). */
typedef struct orc_state
{
    size_t  data_buffer_size, direct_size, frame_size, frame_off, conv_size;
    size_t  levels, blocks, blocks_done, rank, blk_init;
    float   blk_coef;
} orc_state_t;
void    orc_get_state(const orc_convolver_t *c, orc_state_t *st);

/*
 * CPU baseline driver: `instances` independent convolvers, each with a
 * `taps`-long synthetic exponentially-decaying noise IR, fed
 * `blocks` calls of `block` white-noise samples, spread over `threads` host
 * threads (instances partitioned contiguously).  `warm_blocks` calls are made
 * before the clock starts.  Returns elapsed seconds of the timed part (max
 * over threads), or a negative value on allocation failure.  *checksum gets a
 * sum of all outputs so the work cannot be optimised away.
 */
double  orc_bench(size_t instances, size_t taps, size_t rank, size_t block,
                  size_t warm_blocks, size_t blocks, size_t threads, double *checksum);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_CONVOLVER_ORACLE_H_ */
