/*
 * oracle/ref_wrap.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * extern "C" handle API around the reference's own lsp::dspu::Convolver, whose
 * translation units (src/main/util/Convolver.cpp, src/main/iface/IStateDumper.cpp)
 * are compiled VERBATIM, by path, from /root/reference by oracle/Makefile into
 * oracle/_ref/libref_convolver.so.  Nothing from the reference is copied into
 * this repository.  The arithmetic below the class (lsp::dsp::) comes from the
 * restated scalar kernels (oracle/dsp_restated.c) because lsp-dsp-lib is not
 * available offline.
 */
#include <lsp-plug.in/dsp-units/util/Convolver.h>
#include <lsp-plug.in/dsp/dsp.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using lsp::dspu::Convolver;

namespace
{
    /* Collects the field names Convolver::dump emits (Convolver.cpp:315-337). */
    class NameDumper: public lsp::dspu::IStateDumper
    {
        public:
            std::string names;
            void add(const char *name)      { if (!names.empty()) names += ','; names += name; }

            virtual void write(const char *name, const void *) override         { add(name); }
            virtual void write(const char *name, unsigned long) override        { add(name); }
            virtual void write(const char *name, unsigned long long) override   { add(name); }
            virtual void write(const char *name, unsigned int) override         { add(name); }
            virtual void write(const char *name, float) override                { add(name); }
    };

    inline float bench_rand(uint64_t &st)
    {
        uint64_t x  = st;
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
        st          = x;
        uint32_t r  = uint32_t((x * 0x2545F4914F6CDD1DULL) >> 40);
        return float(r) * (2.0f / 16777216.0f) - 1.0f;
    }

    struct BenchJob
    {
        size_t first, last, taps, rank, block, warm_blocks, blocks;
        double seconds = 0.0, checksum = 0.0;
        bool failed = false;
    };

    void bench_thread(BenchJob *j)
    {
        size_t n = j->last - j->first;
        std::vector<Convolver *> cv(n, NULL);
        std::vector<float> ir(j->taps), in(j->block), out(j->block);

        for (size_t i = 0; i < n; ++i)
        {
            uint64_t st = 0x1A000000ULL + j->first + i + 1;
            double tau  = double(j->taps) / std::log(1000.0), e = 0.0;
            for (size_t k = 0; k < j->taps; ++k)
            {
                ir[k]       = bench_rand(st) * float(std::exp(-double(k) / tau));
                e          += double(ir[k]) * ir[k];
            }
            float g     = float(1.0 / std::sqrt(e));
            for (size_t k = 0; k < j->taps; ++k)
                ir[k]      *= g;
            cv[i]       = new Convolver();
            if (!cv[i]->init(ir.data(), j->taps, j->rank, 0.0f))
                j->failed   = true;
        }

        uint64_t st = 0x5EED0000ULL + j->first + 1;
        double sum  = 0.0;
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        for (size_t b = 0; (b < j->warm_blocks + j->blocks) && (!j->failed); ++b)
        {
            if (b == j->warm_blocks)
                t0          = std::chrono::steady_clock::now();
            for (size_t i = 0; i < n; ++i)
            {
                for (size_t k = 0; k < j->block; ++k)
                    in[k]       = bench_rand(st);
                cv[i]->process(out.data(), in.data(), j->block);
                sum        += out[j->block - 1];
            }
        }
        j->seconds  = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        j->checksum = sum;
        for (size_t i = 0; i < n; ++i)
            delete cv[i];
    }
}

extern "C"
{
    void *refconv_create()                      { lsp::dsp::init(); return new Convolver(); }
    void refconv_free(void *h)                  { delete static_cast<Convolver *>(h); }
    void refconv_destroy(void *h)               { static_cast<Convolver *>(h)->destroy(); }
    int refconv_init(void *h, const float *data, size_t count, size_t rank, float phase)
        { return static_cast<Convolver *>(h)->init(data, count, rank, phase) ? 1 : 0; }
    void refconv_process(void *h, float *dst, const float *src, size_t count)
        { static_cast<Convolver *>(h)->process(dst, src, count); }
    size_t refconv_data_size(void *h)           { return static_cast<Convolver *>(h)->data_size(); }
    size_t refconv_rank(void *h)                { return static_cast<Convolver *>(h)->rank(); }

    /* Writes the comma-separated dump() field names into buf; returns their count. */
    size_t refconv_dump_names(void *h, char *buf, size_t cap)
    {
        NameDumper d;
        static_cast<Convolver *>(h)->dump(&d);
        size_t n = 0;
        for (size_t i = 0; i < d.names.size(); ++i)
            n += (d.names[i] == ',');
        if (!d.names.empty())
            ++n;
        if (cap > 0)
        {
            ::strncpy(buf, d.names.c_str(), cap - 1);
            buf[cap - 1] = '\0';
        }
        return n;
    }

    /* Same contract as orc_bench (convolver_oracle.h), on the reference class. */
    double refconv_bench(size_t instances, size_t taps, size_t rank, size_t block,
                         size_t warm_blocks, size_t blocks, size_t threads, double *checksum)
    {
        if (threads < 1)            threads = 1;
        if (threads > instances)    threads = instances;
        if (threads < 1)            return -1.0;
        lsp::dsp::init();

        std::vector<BenchJob> jobs(threads);
        std::vector<std::thread> pool;
        for (size_t t = 0; t < threads; ++t)
        {
            jobs[t].first       = instances * t / threads;
            jobs[t].last        = instances * (t + 1) / threads;
            jobs[t].taps        = taps;
            jobs[t].rank        = rank;
            jobs[t].block       = block;
            jobs[t].warm_blocks = warm_blocks;
            jobs[t].blocks      = blocks;
            pool.emplace_back(bench_thread, &jobs[t]);
        }

        double worst = 0.0, sum = 0.0;
        bool failed = false;
        for (size_t t = 0; t < threads; ++t)
        {
            pool[t].join();
            worst   = (jobs[t].seconds > worst) ? jobs[t].seconds : worst;
            sum    += jobs[t].checksum;
            failed |= jobs[t].failed;
        }
        if (checksum != NULL)
            *checksum = sum;
        return failed ? -1.0 : worst;
    }
}
