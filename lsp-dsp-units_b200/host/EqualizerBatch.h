/*
 * b200conv::EqualizerBatch -- C++ RAII view of b200conv_eq_* (include/b200conv.h): the data path
 * of lsp::dspu::Equalizer in its EQM_FIR / EQM_FFT modes (reference
 * src/main/filters/Equalizer.cpp:474-518) for many equalizers advanced together.  The caller
 * designs the nFirSize-tap impulse responses (Equalizer::reconfigure up to :335) and hands them
 * over; process() has nFirSize samples of latency.  Header only.
 */
#ifndef B200CONV_EQUALIZER_BATCH_H_
#define B200CONV_EQUALIZER_BATCH_H_

#include <b200conv.h>
#include <stddef.h>

namespace b200conv
{
    class EqualizerBatch
    {
        private:
            b200conv_eq_t  *pEq;

        public:
            EqualizerBatch(size_t instances, size_t fir_rank, int device = -1): pEq(NULL)
            {
                b200conv_eq_create(&pEq, device, instances, fir_rank);
            }
            EqualizerBatch(const EqualizerBatch &) = delete;
            EqualizerBatch & operator = (const EqualizerBatch &) = delete;
            ~EqualizerBatch()                       { b200conv_eq_free(pEq); }

            bool valid() const                      { return pEq != NULL; }
            const char *error() const               { return b200conv_last_error(); }
            b200conv_eq_t *handle()                 { return pEq; }

            /** nFirSize taps; smooth: cross-fade to them at the next block boundary (EF_SMOOTH) */
            bool set_kernel(size_t idx, const float *ir, bool smooth = false)
                { return b200conv_eq_set_kernel(pEq, idx, ir, smooth ? 1 : 0) == B200CONV_OK; }

            /** EF_CLEAR */
            bool clear()                            { return b200conv_eq_clear(pEq) == B200CONV_OK; }

            /** N x Equalizer::process with host buffers, synchronous */
            bool process(float * const *dst, const float * const *src, size_t samples)
                { return b200conv_eq_process(pEq, dst, src, samples) == B200CONV_OK; }

            /** Same with one planar host matrix [instances][stride] each way */
            bool process_planar(float *dst, const float *src, size_t stride, size_t samples)
                { return b200conv_eq_process_planar(pEq, dst, src, stride, samples) == B200CONV_OK; }

            /** Same with device matrices, asynchronous on stream */
            bool process_device(float *dst, size_t dst_stride, const float *src, size_t src_stride, size_t samples,
                                void *stream = NULL)
                { return b200conv_eq_process_device(pEq, dst, dst_stride, src, src_stride, samples, stream) == B200CONV_OK; }

            bool sync()                             { return b200conv_eq_sync(pEq) == B200CONV_OK; }
            size_t fir_size() const                 { return b200conv_eq_fir_size(pEq); }
            size_t latency() const                  { return b200conv_eq_latency(pEq); }
            size_t instances() const                { return b200conv_eq_instances(pEq); }
    };
}

#endif /* B200CONV_EQUALIZER_BATCH_H_ */
