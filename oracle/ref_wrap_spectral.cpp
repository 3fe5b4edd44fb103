/*
 * oracle/ref_wrap_spectral.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * extern "C" handle API around the reference's own lsp::dspu::SpectralProcessor
 * (src/main/util/SpectralProcessor.cpp + src/main/misc/windows.cpp, compiled VERBATIM by path by
 * oracle/Makefile into oracle/_ref/libref_convolver.so) over the restated dsp:: kernels.
 *
 * The reference hands the spectrum to a host callback (spectral_processor_func_t,
 * SpectralProcessor.h:38).  The callbacks offered here are the two spectral operations the
 * B200 engine implements on the device (include/b200conv.h, b200conv_sp_*):
 *     kind 1: spectrum[k] *= H[k]          (packed complex table of 2^rank bins)
 *     kind 2: spectrum[k] *= g[k]          (real gain per bin, 2^rank values)
 */
#include <lsp-plug.in/dsp-units/util/SpectralProcessor.h>
#include <lsp-plug.in/dsp/dsp.h>

#include <vector>

using lsp::dspu::SpectralProcessor;

namespace
{
    struct RefSp
    {
        SpectralProcessor   sp;
        std::vector<float>  table;
        int                 kind = 0;
    };

    void hook(void *object, void *, float *spectrum, size_t rank)
    {
        RefSp *r        = static_cast<RefSp *>(object);
        const size_t n  = size_t(1) << rank;
        if (r->kind == 1)
        {
            for (size_t k = 0; k < n; ++k)
            {
                float re = spectrum[2*k], im = spectrum[2*k+1];
                float hr = r->table[2*k], hi = r->table[2*k+1];
                spectrum[2*k]   = re * hr - im * hi;
                spectrum[2*k+1] = re * hi + im * hr;
            }
        }
        else if (r->kind == 2)
        {
            for (size_t k = 0; k < n; ++k)
            {
                spectrum[2*k]  *= r->table[k];
                spectrum[2*k+1] *= r->table[k];
            }
        }
    }
}

extern "C"
{
    void *refsp_create(size_t max_rank)
    {
        lsp::dsp::init();
        RefSp *r = new RefSp();
        r->sp.init(max_rank);
        return r;
    }
    void refsp_free(void *h)                            { delete static_cast<RefSp *>(h); }
    void refsp_set_rank(void *h, size_t rank)           { static_cast<RefSp *>(h)->sp.set_rank(rank); }
    void refsp_set_phase(void *h, float phase)          { static_cast<RefSp *>(h)->sp.set_phase(phase); }
    size_t refsp_rank(void *h)                          { return static_cast<RefSp *>(h)->sp.get_rank(); }
    size_t refsp_latency(void *h)                       { return static_cast<RefSp *>(h)->sp.latency(); }
    size_t refsp_remaining(void *h)                     { return static_cast<RefSp *>(h)->sp.remaining(); }
    void refsp_reset(void *h)                           { static_cast<RefSp *>(h)->sp.reset(); }
    void refsp_update_settings(void *h)                 { static_cast<RefSp *>(h)->sp.update_settings(); }

    /* kind 0: unbind; 1: complex table of 2^rank bins (2^(rank+1) floats); 2: real gains (2^rank floats) */
    void refsp_bind(void *h, int kind, const float *table, size_t floats)
    {
        RefSp *r = static_cast<RefSp *>(h);
        r->kind  = kind;
        if (kind == 0)
        {
            r->sp.unbind();
            return;
        }
        r->table.assign(table, table + floats);
        r->sp.bind(hook, r, NULL);
    }

    void refsp_process(void *h, float *dst, const float *src, size_t count)
    {
        static_cast<RefSp *>(h)->sp.process(dst, src, count);
    }
}
