/*
 * oracle/ref_wrap_splitter.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * extern "C" handle API around the reference's own lsp::dspu::SpectralSplitter
 * (src/main/util/SpectralSplitter.cpp, compiled VERBATIM by path by oracle/Makefile into
 * oracle/_ref/libref_convolver.so) over the restated dsp:: kernels.
 *
 * The reference hands every handler the spectrum through a host callback
 * (spectral_splitter_func_t, SpectralSplitter.h:44) and the processed samples through another
 * (spectral_splitter_sink_t, :60).  The functions offered here are the spectral operations the
 * B200 engine implements on the device (include/b200conv.h, b200conv_ss_*):
 *     kind 1: out[k] = in[k] * H[k]        (packed complex table of 2^rank bins)
 *     kind 2: out[k] = in[k] * g[k]        (real gain per bin, 2^rank values -- what
 *                                           FFTCrossover::spectral_func does, FFTCrossover.cpp:124-140)
 *     kind 3: no function, sink only       (SpectralSplitter.cpp:327-328: the input frame itself)
 * and the sink stores the samples of handler h at out[h][first .. first + count).
 */
#include <lsp-plug.in/dsp-units/util/SpectralSplitter.h>
#include <lsp-plug.in/dsp/dsp.h>

#include <vector>

using lsp::dspu::SpectralSplitter;

namespace
{
    struct RefSs;
    struct Band
    {
        RefSs              *owner   = nullptr;
        size_t              id      = 0;
        int                 kind    = 0;
        std::vector<float>  table;
    };

    struct RefSs
    {
        SpectralSplitter    ss;
        std::vector<Band>   bands;
        float              *out     = nullptr;      /* [handlers][out_stride] of the call in progress */
        size_t              out_stride = 0;
    };

    void func(void *, void *subject, float *out, const float *in, size_t rank)
    {
        Band *b         = static_cast<Band *>(subject);
        const size_t n  = size_t(1) << rank;
        if (b->kind == 1)
        {
            for (size_t k = 0; k < n; ++k)
            {
                float re = in[2*k], im = in[2*k+1];
                float hr = b->table[2*k], hi = b->table[2*k+1];
                out[2*k]    = re * hr - im * hi;
                out[2*k+1]  = re * hi + im * hr;
            }
        }
        else
        {
            for (size_t k = 0; k < n; ++k)
            {
                out[2*k]    = in[2*k] * b->table[k];
                out[2*k+1]  = in[2*k+1] * b->table[k];
            }
        }
    }

    void sink(void *object, void *subject, const float *samples, size_t first, size_t count)
    {
        RefSs *r        = static_cast<RefSs *>(object);
        Band *b         = static_cast<Band *>(subject);
        if (r->out == nullptr)
            return;
        float *dst      = r->out + b->id * r->out_stride + first;
        for (size_t i = 0; i < count; ++i)
            dst[i]          = samples[i];
    }
}

extern "C"
{
    void *refss_create(size_t max_rank, size_t handlers)
    {
        lsp::dsp::init();
        RefSs *r = new RefSs();
        if (r->ss.init(max_rank, handlers) != lsp::STATUS_OK)
        {
            delete r;
            return nullptr;
        }
        r->bands.resize(handlers);
        for (size_t i = 0; i < handlers; ++i)
        {
            r->bands[i].owner   = r;
            r->bands[i].id      = i;
        }
        return r;
    }
    void refss_free(void *h)                            { delete static_cast<RefSs *>(h); }
    void refss_set_rank(void *h, size_t rank)           { static_cast<RefSs *>(h)->ss.set_rank(rank); }
    void refss_set_chunk_rank(void *h, long rank)       { static_cast<RefSs *>(h)->ss.set_chunk_rank(rank); }
    void refss_set_phase(void *h, float phase)          { static_cast<RefSs *>(h)->ss.set_phase(phase); }
    size_t refss_rank(void *h)                          { return static_cast<RefSs *>(h)->ss.rank(); }
    long refss_chunk_rank(void *h)                      { return static_cast<RefSs *>(h)->ss.chunk_rank(); }
    size_t refss_latency(void *h)                       { return static_cast<RefSs *>(h)->ss.latency(); }
    size_t refss_bindings(void *h)                      { return static_cast<RefSs *>(h)->ss.bindings(); }
    void refss_clear(void *h)                           { static_cast<RefSs *>(h)->ss.clear(); }
    void refss_update_settings(void *h)                 { static_cast<RefSs *>(h)->ss.update_settings(); }

    /* kind 0: unbind; 1: complex table (2^(rank+1) floats); 2: real gains (2^rank floats); 3: sink only */
    int refss_bind(void *h, size_t id, int kind, const float *table, size_t floats)
    {
        RefSs *r = static_cast<RefSs *>(h);
        if (id >= r->bands.size())
            return -1;
        Band *b  = &r->bands[id];
        if (kind == 0)
        {
            b->kind     = 0;
            return int(r->ss.unbind(id));
        }
        b->kind  = kind;
        if (kind != 3)
            b->table.assign(table, table + floats);
        return int(r->ss.bind(id, r, b, (kind == 3) ? nullptr : func, sink));
    }

    /* out: [handlers][out_stride]; rows of handlers without a sink are left untouched */
    void refss_process(void *h, float *out, size_t out_stride, const float *src, size_t count)
    {
        RefSs *r        = static_cast<RefSs *>(h);
        r->out          = out;
        r->out_stride   = out_stride;
        r->ss.process(src, count);
        r->out          = nullptr;
    }
}

/* ---- FFTCrossover's band curves ---------------------------------------------------------------
 * lsp::dspu::FFTCrossover (src/main/util/FFTCrossover.cpp) is a SpectralSplitter whose band function
 * multiplies the spectrum by a real curve vFFT (:124-140).  The class itself cannot be compiled from
 * a few files (FFTCrossover.h includes Crossover.h -> Filter.h / FilterBank.h, which need lsp-dsp-lib's
 * biquad types), but the curve functions can: src/main/misc/fft_crossover.cpp is compiled VERBATIM,
 * and refx_band_curve below restates the dozen lines of FFTCrossover::update_band (:458-480) that
 * combine them. */
#include <lsp-plug.in/dsp-units/misc/fft_crossover.h>

extern "C" void refx_band_curve(float *curve, size_t rank, float sample_rate, int hpf, float hpf_freq, float hpf_slope,
                                int lpf, float lpf_freq, float lpf_slope, float gain, float flatten)
{
    using namespace lsp::dspu;
    const size_t bins = size_t(1) << rank;
    if (hpf)
    {
        crossover::hipass_fft_set(curve, hpf_freq, hpf_slope, sample_rate, rank);           /* :468 */
        if (lpf)
            crossover::lopass_fft_apply(curve, lpf_freq, lpf_slope, sample_rate, rank);     /* :470 */
    }
    else if (lpf)
        crossover::lopass_fft_set(curve, lpf_freq, lpf_slope, sample_rate, rank);           /* :477 */
    else
    {
        for (size_t i = 0; i < bins; ++i)                                                   /* :482 dsp::fill */
            curve[i]    = flatten * gain;
        return;
    }
    for (size_t i = 0; i < bins; ++i)                                                       /* :472-473,478-479 */
    {
        /* dsp::limit1(vFFT, 0.0f, fFlatten, bins): clamp to [0, flatten]; dsp::mul_k2(vFFT, fGain, bins) */
        float v     = curve[i];
        v           = (v < 0.0f) ? 0.0f : ((v > flatten) ? flatten : v);
        curve[i]    = v * gain;
    }
}
