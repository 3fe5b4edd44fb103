/*
 * lsp::dspu::Convolver -- B200 drop-in facade.
 *
 * Same class name, include path and public section as the reference unit
 * (lsp-dsp-units include/lsp-plug.in/dsp-units/util/Convolver.h:59-113), so code written
 * against the reference compiles and links against this one unchanged.  The object is a thin
 * handle: all arithmetic runs in hand-written sm_100a CUDA behind the C ABI of
 * include/b200conv.h (one engine batch of size one).  There is no CPU fallback -- init()
 * fails (returns false) when no CUDA device is available.
 *
 * Callers with many channels should use b200conv::ConvolverBatch (ConvolverBatch.h) instead of
 * N facades: one kernel sequence per audio block for all instances x partitions.
 */
#ifndef LSP_PLUG_IN_DSP_UNITS_UTIL_CONVOLVER_H_
#define LSP_PLUG_IN_DSP_UNITS_UTIL_CONVOLVER_H_

#include <stddef.h>

#ifndef LSP_DSP_UNITS_PUBLIC
    #define LSP_DSP_UNITS_PUBLIC    __attribute__((visibility("default")))
#endif

#define CONVOLVER_RANK_MIN          8       /* frame of 128 samples   */
#define CONVOLVER_RANK_MAX          16      /* frame of 32768 samples */

struct b200conv_batch;

namespace lsp
{
    namespace dspu
    {
        class IStateDumper;

        class LSP_DSP_UNITS_PUBLIC Convolver
        {
            private:
                b200conv_batch     *pEngine;        // device engine, batch of one; NULL = not initialised
                int                 nDevice;        // CUDA device (-1: current), from $B200CONV_DEVICE

            public:
                explicit Convolver();
                Convolver(const Convolver &) = delete;
                Convolver(Convolver &&) = delete;
                ~Convolver();

                Convolver & operator = (const Convolver &) = delete;
                Convolver & operator = (Convolver &&) = delete;

                /** Put the object into the empty state (no resources are released) */
                void construct();

                /** Release the engine; the object can be initialised again */
                void destroy();

            public:
                /** Load an impulse response.
                 * @param data  impulse response samples (host memory, borrowed for the call)
                 * @param count number of samples; 0 destroys the convolver and succeeds
                 * @param rank  convolution rank, clamped to [CONVOLVER_RANK_MIN, CONVOLVER_RANK_MAX];
                 *              frames are 2^(rank-1) samples
                 * @param phase position inside the frame at which processing starts, [0, 1)
                 * @return false only if resources could not be obtained; the previous state is kept
                 */
                bool init(const float *data, size_t count, size_t rank, float phase);

                /** Convolve count samples with zero latency; dst may equal src.
                 * Emits zeros while the convolver is not initialised. */
                void process(float *dst, const float *src, size_t count);

                /** Number of taps passed to init(), 0 when not initialised */
                size_t data_size() const;

                /** Effective (clamped) rank, 0 when not initialised */
                size_t rank() const;

                /** Dump the internal state (available when built with B200CONV_WITH_STATE_DUMPER
                 *  against the lsp-dsp-units headers; a no-op otherwise) */
                void dump(IStateDumper *v) const;
        };

    } /* namespace dspu */
} /* namespace lsp */

#endif /* LSP_PLUG_IN_DSP_UNITS_UTIL_CONVOLVER_H_ */
