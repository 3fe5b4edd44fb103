/*
 * oracle/convolver_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Plain-C restatement of lsp::dspu::Convolver; see convolver_oracle.h.
 * Every function cites the reference lines (src/main/util/Convolver.cpp) it follows.
 */
#include "convolver_oracle.h"
#include "dsp_restated.h"

#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SUB_FRAME       ((size_t)1 << (ORC_RANK_MIN - 1))   /* 128: Convolver.cpp:27 */
#define SUB_IMAGE       ((size_t)1 << (ORC_RANK_MIN + 1))   /* 512: Convolver.cpp:28 */
#define SLAB_ALIGN      0x40                                /* Convolver.cpp:30      */

struct orc_convolver
{
    /* Views into the slab (Convolver.h:38-43) */
    float      *tail;           /* vDataBuffer : output accumulator, index 0 = start of current frame */
    float      *frame;          /* vFrame      : current input frame; frame - F is the previous one   */
    float      *scratch;        /* vConvBuffer */
    float      *task;           /* vTaskData   : image of the previous complete frame                 */
    float      *images;         /* vConvData   : IR images                                            */
    float      *direct;         /* vDirectData : first <=128 raw taps                                 */

    orc_state_t s;              /* Convolver.h:45-55 */
    void       *slab;           /* vData */
};

static void orc_reset_fields(orc_convolver_t *c)        /* Convolver.cpp:46-69 */
{
    memset(c, 0, sizeof(*c));
}

orc_convolver_t *orc_create(void)
{
    orc_convolver_t *c = (orc_convolver_t *)malloc(sizeof(orc_convolver_t));
    if (c != NULL)
        orc_reset_fields(c);
    rs_dsp_init();
    return c;
}

void orc_destroy(orc_convolver_t *c)                    /* Convolver.cpp:71-75 */
{
    free(c->slab);
    orc_reset_fields(c);
}

void orc_free(orc_convolver_t *c)
{
    if (c == NULL)
        return;
    orc_destroy(c);
    free(c);
}

size_t orc_data_size(const orc_convolver_t *c)  { return c->s.conv_size; }
size_t orc_rank(const orc_convolver_t *c)       { return c->s.rank;      }
void orc_get_state(const orc_convolver_t *c, orc_state_t *st) { *st = c->s; }

static size_t min_sz(size_t a, size_t b)        { return (a < b) ? a : b; }

/* Transform one zero-padded IR piece into its image (the repeated
 * fill_zero / copy / fastconv_parse triple of Convolver.cpp:157-159,172-174,189-191). */
static void orc_parse_piece(orc_convolver_t *c, float *image, const float *taps, size_t n,
                            size_t piece_rank, size_t scratch_len)
{
    rs_fill_zero(c->scratch, scratch_len);
    rs_copy(c->scratch, taps, n);
    rs_fastconv_parse(image, c->scratch, piece_rank);
}

int orc_init(orc_convolver_t *c, const float *data, size_t count, size_t rank, float phase)
{
    if (count <= 0)                                     /* Convolver.cpp:80-84 */
    {
        orc_destroy(c);
        return 1;
    }

    /* Convolver.cpp:87 : clamp through a signed value */
    {
        long r = (long)rank;
        if (r < ORC_RANK_MIN) r = ORC_RANK_MIN;
        if (r > ORC_RANK_MAX) r = ORC_RANK_MAX;
        rank = (size_t)r;
    }

    /* Convolver.cpp:90-100 */
    size_t F        = (size_t)1 << (rank - 1);
    size_t image    = (size_t)1 << (rank + 1);
    size_t dlen     = SUB_FRAME;                        /* max(128, 64/4) */
    size_t bins     = (count + F - 1) >> (rank - 1);
    size_t total    = (bins + 1) * F + 2 * F + 2 * image + bins * image + dlen;

    /* Convolver.cpp:103-110 : allocate first, old state survives a failure */
    void *slab      = NULL;
    if (posix_memalign(&slab, SLAB_ALIGN, total * sizeof(float)) != 0)
        return 0;
    orc_destroy(c);
    c->slab         = slab;
    float *p        = (float *)slab;
    rs_fill_zero(p, total);

    /* Convolver.cpp:113-135 */
    c->tail         = p;    p += (bins + 1) * F;
    p              += F;                                /* previous input frame */
    c->frame        = p;    p += F;
    c->scratch      = p;    p += image;
    c->task         = p;    p += image;
    c->images       = p;    p += bins * image;
    c->direct       = p;

    /* Convolver.cpp:138-142 */
    c->s.data_buffer_size   = (bins + 1) * F;
    c->s.frame_size         = F;
    c->s.frame_off          = (size_t)(phase * F) % F;
    c->s.direct_size        = min_sz(count, SUB_FRAME);
    c->s.conv_size          = count;

    /* Head piece: taps [0,128) at the minimum rank (Convolver.cpp:152-163) */
    float *img      = c->images;
    size_t prank    = ORC_RANK_MIN;
    rs_copy(c->direct, data, c->s.direct_size);
    orc_parse_piece(c, img, data, c->s.direct_size, prank, image);
    data           += c->s.direct_size;
    count          -= c->s.direct_size;
    img            += (size_t)1 << (prank + 1);

    /* Raising levels: 2^(prank-1) taps at rank prank (Convolver.cpp:165-180) */
    c->s.levels     = 0;
    while ((count > 0) && (prank < rank))
    {
        size_t n        = min_sz(count, (size_t)1 << (prank - 1));
        orc_parse_piece(c, img, data, n, prank, image);
        data           += n;
        count          -= n;
        img            += (size_t)1 << (prank + 1);
        ++prank;
        ++c->s.levels;
    }

    /* Uniform blocks of F taps at the full rank (Convolver.cpp:182-197) */
    c->s.blocks     = 0;
    while (count > 0)
    {
        size_t n        = min_sz(count, F);
        orc_parse_piece(c, img, data, n, rank, image);
        data           += n;
        count          -= n;
        img            += image;
        ++c->s.blocks;
    }

    /* Load-spreading constants (Convolver.cpp:199-210) */
    c->s.blocks_done = c->s.blocks;
    long steps      = (long)(F >> (ORC_RANK_MIN - 1));
    if (steps <= 1)
    {
        c->s.blk_init   = c->s.blocks;
        c->s.blk_coef   = 0.0f;
    }
    else
    {
        c->s.blk_init   = 1;
        c->s.blk_coef   = ((float)c->s.blocks + 1e-3f) / ((float)steps - 1.0f);
    }

    c->s.rank       = rank;                             /* Convolver.cpp:212 */
    return 1;
}

/* Work done when nFrameOff sits on a 128-sample boundary (Convolver.cpp:230-287). */
static void orc_boundary(orc_convolver_t *c)
{
    size_t off      = c->s.frame_off;
    size_t sub_id   = off >> (ORC_RANK_MIN - 1);                /* :245 */
    size_t mask     = (sub_id - 1) ^ sub_id;                    /* :246 */
    size_t lrank    = ORC_RANK_MIN;
    const float *img = &c->images[SUB_IMAGE];                   /* :248 */

    /* Raising levels (:251-262): level i convolves the previous 2^(7+i) samples */
    for (size_t i = 0; i < c->s.levels; ++i)
    {
        if (mask & 1)
        {
            const float *in = c->frame + off - ((size_t)1 << (lrank - 1));
            rs_fastconv_parse_apply(&c->tail[off], c->scratch, img, in, lrank);
        }
        ++lrank;
        img            += (size_t)1 << lrank;
        mask          >>= 1;
    }

    if (c->s.blocks == 0)                                       /* :265 */
        return;

    if (mask & 1)                                               /* :268-272 frame start */
    {
        rs_fastconv_parse(c->task, c->frame - c->s.frame_size, c->s.rank);
        c->s.blocks_done = 0;
    }

    /* :275 : fp32 arithmetic, truncated */
    size_t target   = (size_t)((float)c->s.blk_init + c->s.blk_coef * (float)sub_id);
    target          = min_sz(c->s.blocks, target);
    size_t image    = (size_t)1 << (c->s.rank + 1);
    img             = &c->images[(c->s.blocks_done + 1) * image];       /* :277 */
    float *out      = &c->tail[c->s.blocks_done << (c->s.rank - 1)];    /* :278 */

    for ( ; c->s.blocks_done < target; ++c->s.blocks_done)              /* :280-285 */
    {
        rs_fastconv_apply(out, c->scratch, img, c->task, lrank);
        out            += image >> 2;
        img            += image;
    }
}

void orc_process(orc_convolver_t *c, float *dst, const float *src, size_t count)
{
    if (c->slab == NULL)                                        /* Convolver.cpp:219-223 */
    {
        rs_fill_zero(dst, count);
        return;
    }

    while (count > 0)                                           /* :225 */
    {
        size_t sub_off  = c->s.frame_off & (SUB_FRAME - 1);     /* :227 */
        if (sub_off == 0)
            orc_boundary(c);

        /* Head: the samples just received against taps [0,128) (:289-296) */
        size_t n        = min_sz(count, SUB_FRAME - sub_off);
        float *acc      = &c->tail[c->s.frame_off];
        rs_copy(&c->frame[c->s.frame_off], src, n);
        if (n == SUB_FRAME)
            rs_fastconv_parse_apply(acc, c->scratch, c->images, src, ORC_RANK_MIN);
        else
            rs_convolve(acc, src, c->direct, c->s.direct_size, n);
        rs_copy(dst, acc, n);

        c->s.frame_off += n;                                    /* :298-302 */
        src            += n;
        dst            += n;
        count          -= n;

        /* Frame roll (:304-311) */
        if (c->s.frame_off >= c->s.frame_size)
        {
            size_t F        = c->s.frame_size;
            size_t len      = c->s.data_buffer_size;
            c->s.frame_off -= F;
            rs_move(c->frame - F, c->frame, F);
            rs_move(c->tail, &c->tail[F], len - F);
            rs_fill_zero(&c->tail[len - F], F);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* CPU baseline driver                                                       */

typedef struct bench_job
{
    size_t      first, last;            /* instance range */
    size_t      taps, rank, block, warm_blocks, blocks;
    double      seconds, checksum;
    int         failed;
} bench_job_t;

static inline float bench_rand(uint64_t *st)    /* xorshift64*, uniform [-1,1) */
{
    uint64_t x  = *st;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *st         = x;
    uint32_t r  = (uint32_t)((x * 0x2545F4914F6CDD1DULL) >> 40);   /* 24 bits */
    return (float)r * (2.0f / 16777216.0f) - 1.0f;
}

static double bench_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *bench_thread(void *arg)
{
    bench_job_t *j  = (bench_job_t *)arg;
    size_t n        = j->last - j->first;
    orc_convolver_t **cv = (orc_convolver_t **)calloc(n ? n : 1, sizeof(*cv));
    float *ir       = (float *)malloc(j->taps * sizeof(float));
    float *in       = (float *)malloc(j->block * sizeof(float));
    float *out      = (float *)malloc(j->block * sizeof(float));
    j->failed       = (cv == NULL) || (ir == NULL) || (in == NULL) || (out == NULL);

    for (size_t i = 0; (i < n) && (!j->failed); ++i)
    {
        uint64_t st     = 0x1A000000ULL + j->first + i + 1;
        double tau      = (double)j->taps / log(1000.0), e = 0.0;
        for (size_t k = 0; k < j->taps; ++k)
        {
            ir[k]           = bench_rand(&st) * (float)exp(-(double)k / tau);
            e              += (double)ir[k] * ir[k];
        }
        float g         = (float)(1.0 / sqrt(e));
        for (size_t k = 0; k < j->taps; ++k)
            ir[k]          *= g;
        cv[i]           = orc_create();
        if ((cv[i] == NULL) || (!orc_init(cv[i], ir, j->taps, j->rank, 0.0f)))
            j->failed       = 1;
    }

    uint64_t st     = 0x5EED0000ULL + j->first + 1;
    double sum      = 0.0, t0 = 0.0;
    for (size_t b = 0; (b < j->warm_blocks + j->blocks) && (!j->failed); ++b)
    {
        if (b == j->warm_blocks)
            t0              = bench_now();
        for (size_t i = 0; i < n; ++i)
        {
            for (size_t k = 0; k < j->block; ++k)
                in[k]           = bench_rand(&st);
            orc_process(cv[i], out, in, j->block);
            sum            += out[j->block - 1];
        }
    }
    j->seconds      = bench_now() - t0;
    j->checksum     = sum;

    for (size_t i = 0; (cv != NULL) && (i < n); ++i)
        orc_free(cv[i]);
    free(cv); free(ir); free(in); free(out);
    return NULL;
}

double orc_bench(size_t instances, size_t taps, size_t rank, size_t block,
                 size_t warm_blocks, size_t blocks, size_t threads, double *checksum)
{
    if (threads < 1)            threads = 1;
    if (threads > instances)    threads = instances;
    if (threads < 1)            return -1.0;

    rs_dsp_init();
    bench_job_t *jobs   = (bench_job_t *)calloc(threads, sizeof(*jobs));
    pthread_t *tid      = (pthread_t *)calloc(threads, sizeof(*tid));
    if ((jobs == NULL) || (tid == NULL))
    {
        free(jobs); free(tid);
        return -1.0;
    }

    for (size_t t = 0; t < threads; ++t)
    {
        jobs[t].first       = instances * t / threads;
        jobs[t].last        = instances * (t + 1) / threads;
        jobs[t].taps        = taps;
        jobs[t].rank        = rank;
        jobs[t].block       = block;
        jobs[t].warm_blocks = warm_blocks;
        jobs[t].blocks      = blocks;
        pthread_create(&tid[t], NULL, bench_thread, &jobs[t]);
    }

    double worst = 0.0, sum = 0.0;
    int failed = 0;
    for (size_t t = 0; t < threads; ++t)
    {
        pthread_join(tid[t], NULL);
        if (jobs[t].seconds > worst)
            worst               = jobs[t].seconds;
        sum                += jobs[t].checksum;
        failed             |= jobs[t].failed;
    }
    free(jobs); free(tid);
    if (checksum != NULL)
        *checksum           = sum;
    return failed ? -1.0 : worst;
}
