/*
 * bench_facade -- what a host application sees through the reference-facing call.
 *
 *   (a) 64 x lsp::dspu::Convolver::process(float *dst, const float *src, size_t count) -- the
 *       reference's own contract (include/lsp-plug.in/dsp-units/util/Convolver.h:95): one object per
 *       (input, IR) pair, called per 1024-sample block in a serial loop on caller-owned PAGEABLE
 *       buffers.  Every call is a launch + a synchronisation of its own; nothing can be coalesced
 *       behind this signature because each call must return its block before the next one is made.
 *   (b) the explicit coalescing API: b200conv_process (include/b200conv.h), one call per block for
 *       all 64 instances with a table of the same per-instance PAGEABLE pointers.
 *
 * Prints one JSON object (BASELINE config 3 geometry unless overridden):
 *     bench_facade [instances=64] [taps=480000] [rank=11] [block=1024] [blocks=200]
 * Also reports what Convolver::init costs per instance (IR upload + partition transforms).
 */
#include <lsp-plug.in/dsp-units/util/Convolver.h>
#include <b200conv.h>
#include "ConvolverBatch.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static float lcg(unsigned long long &st)
{
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    return float(double(st >> 40) / double(1ull << 23) - 1.0);
}

int main(int argc, char **argv)
{
    size_t n        = (argc > 1) ? size_t(atol(argv[1])) : 64;
    size_t taps     = (argc > 2) ? size_t(atol(argv[2])) : 480000;
    size_t rank     = (argc > 3) ? size_t(atol(argv[3])) : 11;
    size_t block    = (argc > 4) ? size_t(atol(argv[4])) : 1024;
    size_t blocks   = (argc > 5) ? size_t(atol(argv[5])) : 200;

    std::vector<float> ir(taps);
    unsigned long long st = 0x1A000000ull;
    double tau = double(taps) / std::log(1000.0), e = 0.0;
    for (size_t k = 0; k < taps; ++k)
    {
        ir[k]   = lcg(st) * float(std::exp(-double(k) / tau));
        e      += double(ir[k]) * ir[k];
    }
    for (size_t k = 0; k < taps; ++k)
        ir[k]  *= float(1.0 / std::sqrt(e));

    std::vector<std::vector<float> > in(n, std::vector<float>(block)), out(n, std::vector<float>(block));
    for (size_t c = 0; c < n; ++c)
        for (size_t k = 0; k < block; ++k)
            in[c][k]    = lcg(st);

    /* ---- (a) the facade, one object per channel ------------------------------------------------ */
    std::vector<lsp::dspu::Convolver> conv(n);
    double t0 = now_s();
    for (size_t c = 0; c < n; ++c)
        if (!conv[c].init(ir.data(), taps, rank, 0.0f))
        {
            printf("{\"error\": \"Convolver::init failed: %s\"}\n", b200conv_last_error());
            return 1;
        }
    double facade_init_ms = (now_s() - t0) * 1e3;

    for (size_t w = 0; w < 8; ++w)
        for (size_t c = 0; c < n; ++c)
            conv[c].process(out[c].data(), in[c].data(), block);
    t0 = now_s();
    for (size_t b = 0; b < blocks; ++b)
        for (size_t c = 0; c < n; ++c)
            conv[c].process(out[c].data(), in[c].data(), block);
    double facade_s = now_s() - t0;
    double chk_a = 0.0;
    for (size_t c = 0; c < n; ++c)
        chk_a += out[c][block - 1];
    for (size_t c = 0; c < n; ++c)
        conv[c].destroy();

    /* ---- (b) one batched call per block, the same pageable per-channel buffers ------------------- */
    b200conv::ConvolverBatch batch(n);
    if (!batch.valid())
    {
        printf("{\"error\": \"b200conv_create failed: %s\"}\n", b200conv_last_error());
        return 1;
    }
    t0 = now_s();
    for (size_t c = 0; c < n; ++c)
        if (!batch.init(c, ir.data(), taps, rank, 0.0f))
        {
            printf("{\"error\": \"b200conv_init failed: %s\"}\n", b200conv_last_error());
            return 1;
        }
    double batch_init_ms = (now_s() - t0) * 1e3;
    std::vector<float *> dptr(n);
    std::vector<const float *> sptr(n);
    for (size_t c = 0; c < n; ++c)
    {
        dptr[c] = out[c].data();
        sptr[c] = in[c].data();
    }
    for (size_t w = 0; w < 8; ++w)
        batch.process(dptr.data(), sptr.data(), block);
    t0 = now_s();
    for (size_t b = 0; b < blocks; ++b)
        batch.process(dptr.data(), sptr.data(), block);
    double table_s = now_s() - t0;
    double chk_b = 0.0;
    for (size_t c = 0; c < n; ++c)
        chk_b += out[c][block - 1];

    const double samples = double(blocks) * double(n) * double(block);
    printf("{\"instances\": %zu, \"taps\": %zu, \"rank\": %zu, \"block\": %zu, \"blocks\": %zu, "
           "\"facade_64x_process\": {\"value\": %.6g, \"unit\": \"samples/s\", \"us_per_block\": %.3f, "
           "\"us_per_call\": %.3f, \"api\": \"N x lsp::dspu::Convolver::process on pageable buffers, serial loop "
           "(one launch + one synchronisation per call)\", \"init_ms_per_instance\": %.3f}, "
           "\"pointer_table_pageable\": {\"value\": %.6g, \"unit\": \"samples/s\", \"us_per_block\": %.3f, "
           "\"api\": \"b200conv_process: one call per block for all instances, table of pageable per-instance "
           "pointers (gather, one H2D, one launch, one D2H, scatter)\", \"init_ms_per_instance\": %.3f}, "
           "\"checksums\": [%.6g, %.6g]}\n",
           n, taps, rank, block, blocks,
           samples / facade_s, facade_s * 1e6 / double(blocks), facade_s * 1e6 / double(blocks * n), facade_init_ms / double(n),
           samples / table_s, table_s * 1e6 / double(blocks), batch_init_ms / double(n), chk_a, chk_b);
    return 0;
}
