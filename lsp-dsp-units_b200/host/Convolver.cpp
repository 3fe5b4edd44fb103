/*
 * lsp::dspu::Convolver facade over the B200 engine (see the header).  Semantics follow the
 * reference implementation, lsp-dsp-units src/main/util/Convolver.cpp:
 *   construct/destroy :46-75, init :77-215, process :217-313, dump :315-337.
 */
#ifdef B200CONV_WITH_STATE_DUMPER
    #include <lsp-plug.in/dsp-units/iface/IStateDumper.h>
#endif
#include <lsp-plug.in/dsp-units/util/Convolver.h>
#include <b200conv.h>

#include <stdlib.h>
#include <string.h>

namespace lsp
{
    namespace dspu
    {
        Convolver::Convolver()
        {
            construct();
        }

        Convolver::~Convolver()
        {
            destroy();
        }

        void Convolver::construct()
        {
            pEngine         = NULL;
            const char *dev = ::getenv("B200CONV_DEVICE");
            nDevice         = (dev != NULL) ? ::atoi(dev) : -1;
        }

        void Convolver::destroy()
        {
            b200conv_free(pEngine);         // NULL is fine
            construct();
        }

        bool Convolver::init(const float *data, size_t count, size_t rank, float phase)
        {
            if (count == 0)                 // reference :80-84
            {
                destroy();
                return true;
            }

            // Build the new engine first: a failure leaves the old one untouched (reference :103-108)
            b200conv_batch *fresh = NULL;
            if (b200conv_create(&fresh, nDevice, 1) != B200CONV_OK)
                return false;
            if (b200conv_init(fresh, 0, data, count, rank, phase) != B200CONV_OK)
            {
                b200conv_free(fresh);
                return false;
            }

            b200conv_free(pEngine);
            pEngine         = fresh;
            return true;
        }

        void Convolver::process(float *dst, const float *src, size_t count)
        {
            if (pEngine == NULL)            // reference :219-223
            {
                ::memset(dst, 0, count * sizeof(float));
                return;
            }
            if (count == 0)
                return;

            // A batch of one.  The caller's buffers are ordinary pageable memory (Convolver.h:95):
            // the pointer-table call gathers the block into page-locked memory which the kernels
            // then read and write in place across PCIe -- no copy operations in the stream.  The
            // reference has no error channel here; a device failure yields silence.
            float *d = dst;
            const float *s = src;
            if (b200conv_process(pEngine, &d, &s, count) != B200CONV_OK)
                ::memset(dst, 0, count * sizeof(float));
        }

        size_t Convolver::data_size() const
        {
            return b200conv_data_size(pEngine, 0);
        }

        size_t Convolver::rank() const
        {
            return b200conv_rank(pEngine, 0);
        }

        void Convolver::dump(IStateDumper *v) const
        {
        #ifdef B200CONV_WITH_STATE_DUMPER
            // The reference's 18 fields, same names and order (reference :315-337), with the values
            // the reference object would hold after the same history (b200conv_get_dump); pointer
            // fields report the device buffer that plays the same part.  Engine extras follow.
            b200conv_dump_t d;
            b200conv_state_t st;
            ::memset(&d, 0, sizeof(d));
            ::memset(&st, 0, sizeof(st));
            if (pEngine != NULL)
            {
                b200conv_get_dump(pEngine, 0, &d);
                b200conv_get_state(pEngine, 0, &st);
            }

            v->write("pDataBuffer", d.vDataBuffer);
            v->write("vFrame", d.vFrame);
            v->write("vConvBuffer", d.vConvBuffer);
            v->write("vTaskData", d.vTaskData);
            v->write("vConvData", d.vConvData);
            v->write("vDirectData", d.vDirectData);

            v->write("nDataBufferSize", d.nDataBufferSize);
            v->write("nDirectSize", d.nDirectSize);
            v->write("nFrameSize", d.nFrameSize);
            v->write("nFrameOff", d.nFrameOff);
            v->write("nConvSize", d.nConvSize);
            v->write("nLevels", d.nLevels);
            v->write("nBlocks", d.nBlocks);
            v->write("nBlocksDone", d.nBlocksDone);
            v->write("nRank", d.nRank);
            v->write("nBlkInit", d.nBlkInit);
            v->write("fBlkCoef", d.fBlkCoef);

            v->write("vData", d.vData);

            v->write("pEngine", static_cast<const void *>(pEngine));
            v->write("nDevice", nDevice);
            v->write("nPartitions", st.partitions);
            v->write("nPartOffset", st.part_offset);
            v->write("nFrames", static_cast<unsigned long long>(st.frames));
        #else
            (void)v;
        #endif
        }

    } /* namespace dspu */
} /* namespace lsp */
