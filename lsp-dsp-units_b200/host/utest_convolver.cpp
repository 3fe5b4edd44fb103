/*
 * Host-side unit test of the lsp::dspu::Convolver facade, in the spirit of the reference's
 * src/test/utest/util/convolver.cpp: same shapes (31-tap ramp / rank 9 / calls of 31;
 * 8192 random taps / rank 10 / calls of 31; a stereo pair with phases through the batch class),
 * checked against naive direct convolution computed here in double precision.
 * Exit code 0 = all passed.  Needs a CUDA device.
 */
#include <lsp-plug.in/dsp-units/util/Convolver.h>
#include "ConvolverBatch.h"
#include "EqualizerBatch.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

using lsp::dspu::Convolver;

static std::vector<double> direct(const std::vector<float> &x, const std::vector<float> &h)
{
    std::vector<double> y(x.size(), 0.0);
    for (size_t i = 0; i < x.size(); ++i)
    {
        if (x[i] == 0.0f)
            continue;
        for (size_t j = 0; (j < h.size()) && (i + j < x.size()); ++j)
            y[i + j] += double(x[i]) * double(h[j]);
    }
    return y;
}

static float urand(uint64_t &s)
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return float(uint32_t(s >> 40)) * (2.0f / 16777216.0f) - 1.0f;
}

static void feed(Convolver &c, std::vector<float> &dst, const std::vector<float> &src, size_t step)
{
    for (size_t i = 0; i < src.size(); i += step)
    {
        size_t n = (src.size() - i < step) ? src.size() - i : step;
        c.process(&dst[i], &src[i], n);
    }
}

static int check(const char *name, const std::vector<float> &got, const std::vector<double> &want, double tol)
{
    double peak = 0.0, worst = 0.0;
    for (size_t i = 0; i < want.size(); ++i)
    {
        peak    = fmax(peak, fabs(want[i]));
        worst   = fmax(worst, fabs(double(got[i]) - want[i]));
    }
    int ok = (worst <= tol * peak);
    printf("%-28s max|err|/peak = %.3e  %s\n", name, worst / peak, ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
}

int main()
{
    int failures = 0;

    {   // not initialised: zeros; rank clamp; count == 0
        Convolver c;
        std::vector<float> x(100, 1.0f), y(100, 5.0f);
        c.process(y.data(), x.data(), x.size());
        int ok = 1;
        for (float v : y) ok &= (v == 0.0f);
        float one = 1.0f;
        ok &= c.init(&one, 1, 3, 0.0f)  && (c.rank() == CONVOLVER_RANK_MIN) && (c.data_size() == 1);
        ok &= c.init(&one, 1, 20, 0.0f) && (c.rank() == CONVOLVER_RANK_MAX);
        ok &= c.init(&one, 0, 10, 0.0f) && (c.rank() == 0) && (c.data_size() == 0);
        c.destroy(); c.destroy();
        printf("%-28s %s\n", "api contract", ok ? "ok" : "FAILED");
        failures += !ok;
    }

    {   // small
        std::vector<float> h(31), x(0x2000 + 31, 0.0f);
        for (size_t i = 0; i < h.size(); ++i) h[i] = float(i + 1);
        for (size_t i = 0, j = 0; i < 0x2000; i += 5, ++j)
            x[i] = ((j % 3) == 0) ? 1.0f : ((j % 3) == 1) ? 0.1f : 0.01f;
        std::vector<float> y(x.size(), 0.0f);
        Convolver c;
        if (!c.init(h.data(), h.size(), 9, 0.0f)) { printf("init failed: %s\n", b200conv_last_error()); return 2; }
        feed(c, y, x, 31);
        failures += check("small (31 taps, rank 9)", y, direct(x, h), 1e-5);
    }

    {   // large
        uint64_t seed = 7;
        std::vector<float> h(0x2000), x(0x20 + 0x2000, 0.0f);
        for (float &v : h) v = urand(seed);
        for (size_t i = 0; i < 0x20; ++i) x[i] = urand(seed);
        std::vector<float> y(x.size(), 0.0f);
        Convolver c;
        if (!c.init(h.data(), h.size(), 10, 0.0f)) return 2;
        feed(c, y, x, 31);
        failures += check("large (8192 taps, rank 10)", y, direct(x, h), 1e-5);
        // in place
        std::vector<float> z(x);
        if (!c.init(h.data(), h.size(), 10, 0.0f)) return 2;
        feed(c, z, z, 100);
        failures += check("large, dst == src", z, direct(x, h), 1e-5);
    }

    {   // stereo pair through the batch class, phases 0 and 0.5, 256-sample blocks
        uint64_t seed = 99;
        const size_t n = 256 * 64;
        std::vector<float> h0(5000), h1(3000), x0(n), x1(n), y0(n), y1(n);
        for (float &v : h0) v = urand(seed) * 0.05f;
        for (float &v : h1) v = urand(seed) * 0.05f;
        for (float &v : x0) v = urand(seed);
        for (float &v : x1) v = urand(seed);
        b200conv::ConvolverBatch b(2);
        if (!b.valid() || !b.init(0, h0.data(), h0.size(), 9, 0.0f) || !b.init(1, h1.data(), h1.size(), 9, 0.5f))
            { printf("batch init failed: %s\n", b.error()); return 2; }
        for (size_t i = 0; i < n; i += 256)
        {
            const float *src[2] = { &x0[i], &x1[i] };
            float *dst[2]       = { &y0[i], &y1[i] };
            if (!b.process(dst, src, 256)) { printf("process failed: %s\n", b.error()); return 2; }
        }
        failures += check("batch ch0 (phase 0)", y0, direct(x0, h0), 1e-5);
        failures += check("batch ch1 (phase 0.5)", y1, direct(x1, h1), 1e-5);
    }

    {   // Equalizer FIR data path, in the spirit of src/test/utest/filters/equalizer.cpp:34-84:
        // impulse in, the peak of the response sits at latency + (kernel peak = nFirSize / 2);
        // plus the delayed-direct-convolution identity on a second instance
        const size_t fir_rank = 10, F = size_t(1) << fir_rank, n = 6 * F;
        uint64_t seed = 5;
        std::vector<float> k0(F), k1(F), x0(n, 0.0f), x1(n), y0(n), y1(n);
        for (size_t i = 0; i < F; ++i)
        {
            double t = (double(i) - double(F / 2)) * 0.35, w = 0.5 - 0.5 * cos(2.0 * M_PI * double(i) / double(F - 1));
            k0[i] = float(((t == 0.0) ? 1.0 : sin(t) / t) * w);      // windowed sinc, peak at F / 2
            k1[i] = urand(seed) * 0.05f;
        }
        x0[0] = 1.0f;
        for (float &v : x1) v = urand(seed);
        b200conv::EqualizerBatch e(2, fir_rank);
        if (!e.valid() || !e.set_kernel(0, k0.data()) || !e.set_kernel(1, k1.data()))
            { printf("equalizer setup failed: %s\n", e.error()); return 2; }
        for (size_t i = 0; i < n; i += 300)
        {
            size_t m = (n - i < 300) ? n - i : 300;
            const float *src[2] = { &x0[i], &x1[i] };
            float *dst[2]       = { &y0[i], &y1[i] };
            if (!e.process(dst, src, m)) { printf("equalizer process failed: %s\n", e.error()); return 2; }
        }
        size_t peak = 0;
        for (size_t i = 0; i < n; ++i)
            if (fabsf(y0[i]) > fabsf(y0[peak])) peak = i;
        int ok = (peak == e.latency() + F / 2) && (e.latency() == F);
        printf("%-28s peak at %zu, latency %zu + %zu  %s\n", "equalizer latency", peak, e.latency(), F / 2, ok ? "ok" : "FAILED");
        failures += !ok;
        std::vector<double> want(n, 0.0), full = direct(x1, k1);
        for (size_t i = F; i < n; ++i) want[i] = full[i - F];
        failures += check("equalizer = delayed direct", y1, want, 1e-5);
    }

    printf("%s\n", failures ? "FAILED" : "ALL PASSED");
    return failures ? 1 : 0;
}
