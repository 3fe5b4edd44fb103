/*
 * equalizer.cuh -- data path of lsp::dspu::Equalizer in its EQM_FIR / EQM_FFT modes, batched over
 * many equalizer instances (scope-table row f2; included by engine.cu).
 *
 * Reference: src/main/filters/Equalizer.cpp:474-518.  Every nFirSize input samples the reference
 * shifts its 2 * nFirSize output buffer by nFirSize, adds the convolution of the finished input
 * block with the kernel (fastconv_parse_apply at rank nFirRank + 1), and -- when a new kernel was
 * handed over "smoothly" -- cross-fades towards the same block convolved with the new kernel
 * (:486-501).  Samples leave the output buffer while the next block is collected, hence a block
 * of latency.
 *
 * Here: one launch per block boundary for the whole batch (k_eq).  A CTA iteration owns one
 * instance: forward transform of the finished block, product with the kernel spectrum,
 * full-length inverse transform, overlap-add / cross-fade, and the sample exchange (input
 * segment -> vInBuffer, vOutBuffer -> output segment) of the call that triggered it.  On the
 * ping-pong ranks (transform rank <= 12) the real-FFT split, the product and the merge are fused
 * in registers, the operands arrive through TMA bulk copies a phase ahead, and a call that takes
 * the whole block gets its outputs straight from the last inverse pass; the other ranks and the
 * (rare) hand-over blocks go through fwd_body / inv_body and global scratch.
 * The kernel spectra use the engine's packed half-spectrum layout (bin 0 = (DC, Nyquist)).
 * Measured progression and ncu summary: profiles/README.md; design notes: DESIGN.md 4.3.
 */
#ifndef B200CONV_EQUALIZER_CUH_
#define B200CONV_EQUALIZER_CUH_

struct EqArgs
{
    const float2   *tw;             /* twiddle table of rank nFirRank + 1                          */
    const float    *src;            /* [instances][stride_src]: the input segment of this step      */
    float          *dst;            /* [instances][stride_dst]: the output segment                  */
    uint64_t        stride_src, stride_dst;
    float          *inbuf;          /* [instances][F]      vInBuffer  (Equalizer.cpp:115)           */
    float          *outbuf;         /* [instances][2F]     vOutBuffer (:116)                        */
    float2         *spec;           /* [instances][2][M]   block spectrum, product (scratch)        */
    float          *conv;           /* [instances][2][2F]  block (*) current kernel, (*) new kernel  */
    const float2   *kern;           /* [instances][2][M]   kernel spectra: vConv / vNewConv          */
    const uint8_t  *state;          /* [instances] bit 0: slot of the current kernel, bit 1: cross-fade pending */
    uint32_t        n_inst;
    uint32_t        off, n;         /* segment: n samples at offset off of the block (nBufSize)     */
    uint32_t        do_block;       /* a block boundary precedes the segment                        */
};

__device__ __forceinline__ float ld_cg_f1(const float *p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

/* The block of a ping-pong rank, without the packed spectrum ever being materialised:
 *   eq_forward          z[m] = x[2m] + i x[2m+1] and z[m] w_M^m  ->  two P-point forward transforms
 *   eq_middle_inverse   per bin pair (k, M-k): real-FFT split (fwd_body's post-pass), product with
 *                       the kernel spectrum, real-FFT merge (inv_body's pre-pass) -- all in
 *                       registers, from one work buffer into the other -- then two inverse
 *                       transforms and the combine pass, which also shifts and overlap-adds
 *                       vOutBuffer (Equalizer.cpp:482-484).
 * Same arithmetic as fwd_body -> multiply -> inv_body, three shared-memory sweeps fewer. */
template <int RANK>
__device__ __forceinline__ void eq_forward(float2 *A, float2 *B, const float *x, const float2 *tw, int tid)
{
    using C = FftCfg<RANK>;
    constexpr int P = C::P, T = C::T;
    for (int m = tid; m < P; m += T)
    {
        float2 z    = reinterpret_cast<const float2 *>(x)[m];
        A[m]        = z;
        A[P + m]    = cmul(z, tw[C::TW_PRE + m]);
    }
    __syncthreads();
}

template <int RANK>
__device__ __forceinline__ void eq_middle_inverse(float2 *A, float2 *B, float2 *H, float *ob, float *emit,
                                                  uint64_t *bar_h, uint32_t &ph_h, const float2 *tw, int tid)
{
    using C = FftCfg<RANK>;
    constexpr int P = C::P, M = C::M, N = C::N, T = C::T;
    float2 *R   = fft_smem<RANK, false, true, 0, true, true>(A, B, tw, tid); /* Z even | Z odd */
    float2 *D   = (R == A) ? B : A;

    for (int k = tid; k < M / 2; k += T)
    {
        if (k == 0)
        {
            float2 z0   = R[0], h0 = H[0];                              /* bin 0 = (DC, Nyquist) */
            float2 y0   = make_float2((z0.x + z0.y) * h0.x, (z0.x - z0.y) * h0.y);
            D[0]        = make_float2(y0.x + y0.y, y0.x - y0.y);
            float2 zh   = R[P / 2];                                     /* bin M/2 pairs with itself */
            float2 yh   = cmul(make_float2(zh.x, -zh.y), H[M / 2]);
            D[P / 2]    = make_float2(2.0f * yh.x, -2.0f * yh.y);
            continue;
        }
        const int par   = k & 1, ik = k >> 1;
        const int im    = par ? (P - 1 - ik) : (P - ik);
        const float2 w  = tw[C::TW_POST + k];
        float2 zk = R[par * P + ik], zm = R[par * P + im];
        /* split: X[k] = e - i w o, X[M-k] = conj(e) - i conj(w o) */
        float2 e    = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        float2 o    = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
        float2 wo   = cmul(w, o);
        float2 yk   = cmul(make_float2(e.x + wo.y, e.y - wo.x), H[k]);
        float2 ym   = cmul(make_float2(e.x - wo.y, -e.y - wo.x), H[M - k]);
        /* merge: Z[k] = e' + i o', Z[M-k] = conj(e') + i conj(o'), o' = conj(w) (yk - conj(ym)) */
        float2 e2   = make_float2(yk.x + ym.x, yk.y - ym.y);
        float2 o2   = cmulc(make_float2(yk.x - ym.x, yk.y + ym.y), w);
        D[par * P + ik] = make_float2(e2.x - o2.y, e2.y + o2.x);
        D[par * P + im] = make_float2(e2.x + o2.y, o2.x - e2.y);
    }
    __syncthreads();

    /* the kernel spectrum has been consumed: its buffer now receives the overlap tail
     * vOutBuffer[F, 2F) while the inverse transforms run */
    const float2 *tail_s = reinterpret_cast<const float2 *>(H);
    if (tid == 0)
    {
        mbar_expect_tx(bar_h, M * sizeof(float));
        bulk_g2s_plain(H, ob + M, M * sizeof(float), bar_h);
    }

    float2 *Q   = fft_smem<RANK, true, true, 0, true, true>(D, R, tw, tid);

    mbar_wait(bar_h, ph_h);
    ph_h ^= 1u;

    /* y[2m], y[2m+1] = (Qe[m] + conj(w_M^m) Qo[m]) / N ; the same with a minus sign F samples later.
     * emit != NULL: the call takes the whole block, so the first half goes straight to the
     * caller (vOutBuffer[0, F) would be dead: the next block boundary overwrites it unread). */
    const float scale = 1.0f / float(N);
    for (int m = tid; m < P; m += T)
    {
        float2 av   = Q[m];
        float2 bv   = cmulc(Q[P + m], tw[C::TW_PRE + m]);
        float2 tail = tail_s[m];
        float2 lo   = make_float2(tail.x + (av.x + bv.x) * scale, tail.y + (av.y + bv.y) * scale);
        if (emit != nullptr)
            reinterpret_cast<float2 *>(emit)[m]    = lo;
        else
            reinterpret_cast<float2 *>(ob)[m]      = lo;
        reinterpret_cast<float2 *>(ob)[m + P]  = make_float2((av.x - bv.x) * scale, (av.y - bv.y) * scale);
    }
}

/* A CTA walks through load -> transform -> load -> transform phases separated by barriers, so
 * every global load it waits for is exposed.  On the ping-pong ranks the two block-sized operands
 * therefore arrive through TMA bulk copies issued a phase ahead by one elected thread:
 *   the kernel spectrum  -> Hs   (issued at the top of the instance, consumed after the forward pass)
 *   the NEXT instance's finished input block -> ibs (issued once the forward pass has read ibs)
 * and the call's input segment is fetched into registers before the transforms start. */
template <int RANK>
struct EqCfg
{
    using C = FftCfg<RANK>;
    static constexpr bool   PIPE    = C::PP;
    static constexpr size_t OFF_H   = (C::SMEM + 127) & ~size_t(127);
    static constexpr size_t OFF_IB  = OFF_H + (PIPE ? C::M * sizeof(float2) : 0);
    static constexpr size_t OFF_BAR = OFF_IB + (PIPE ? C::M * sizeof(float) : 0);
    static constexpr size_t SMEM    = PIPE ? OFF_BAR + 2 * sizeof(uint64_t) : C::SMEM;
    static constexpr int    XPT     = PIPE ? C::M / C::T : 1;           /* input samples per thread */
    static constexpr int    MINB    = (RANK == 11) ? 5 : (RANK == 12) ? 2 : (RANK == 13) ? 2 : (RANK == 10) ? 1 : 0;     /* measured (profiles/); 0: left to the compiler.
                                                                           rank 13: two 512-thread CTAs per SM need <= 64 registers (the batched loads of the transform bodies would take 94) */
};

template <int RANK>
__global__ void __launch_bounds__(FftCfg<RANK>::T, EqCfg<RANK>::MINB)
k_eq(const EqArgs a)
{
    using C = FftCfg<RANK>;
    using E = EqCfg<RANK>;
    constexpr int F = C::M, M = C::M, N = C::N, T = C::T;
    extern __shared__ __align__(128) unsigned char eq_sm[];
    float2 *A               = reinterpret_cast<float2 *>(eq_sm);
    float2 *B               = C::PP ? A + C::WORK : nullptr;
    float2 *Hs              = reinterpret_cast<float2 *>(eq_sm + E::OFF_H);
    float *ibs              = reinterpret_cast<float *>(eq_sm + E::OFF_IB);
    uint64_t *bars          = reinterpret_cast<uint64_t *>(eq_sm + E::OFF_BAR);     /* [0] input block, [1] kernel */
    const float2 *tw        = a.tw;
    const int tid           = threadIdx.x;
    const bool pipe         = E::PIPE && (a.do_block != 0);
    if (C::TWS && a.do_block)
    {
        float2 *tws         = A + C::WORK * (C::PP ? 2 : 1);
        for (int i = tid; i < C::TW_TOTAL; i += T)
            tws[i]              = a.tw[i];
        tw                  = tws;
    }
    const float2 *twg       = tw;           /* the shared-memory copy serves every twiddle read */
    if ((!C::TWS) && a.do_block)
    {
        /* ranks >= 13: the compact pass table (one factor per butterfly) next to the work buffer */
        float2 *twc         = A + C::WORK;
        stage_compact_twiddles<C>(twc, a.tw, tid);
        tw                  = twc;
        twg                 = a.tw;
    }
    if (pipe && (tid == 0))
    {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (blockIdx.x < a.n_inst)
        {
            mbar_expect_tx(&bars[0], F * sizeof(float));
            bulk_g2s_plain(ibs, a.inbuf + uint64_t(blockIdx.x) * F, F * sizeof(float), &bars[0]);
        }
    }
    __syncthreads();
    uint32_t ph_ib = 0, ph_h = 0;
    uint32_t st_next        = (a.do_block && (blockIdx.x < a.n_inst)) ? a.state[blockIdx.x] : 0u;

    for (uint32_t inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x)
    {
        float *ib               = a.inbuf + uint64_t(inst) * F;
        float *ob               = a.outbuf + uint64_t(inst) * N;
        const float *s          = a.src + uint64_t(inst) * a.stride_src;
        float *d                = a.dst + uint64_t(inst) * a.stride_dst;

        float xr[E::XPT];
        if (E::PIPE)
        {
            #pragma unroll
            for (int j = 0; j < E::XPT; ++j)
            {
                uint32_t i      = uint32_t(tid + j * T);
                xr[j]           = (i < a.n) ? ld_cg_f1(s + i) : 0.0f;
            }
        }

        bool emitted            = false;
        if (a.do_block)
        {
            float2 *xs              = a.spec + uint64_t(inst) * 2 * M;
            float2 *ys              = xs + M;
            float *c0               = a.conv + uint64_t(inst) * 2 * N;
            float *c1               = c0 + N;
            const uint32_t st       = st_next;
            const uint32_t cur      = st & 1u;
            const bool xfade        = (st & 2u) != 0;
            const float2 *H         = a.kern + (uint64_t(inst) * 2 + cur) * M;
            const bool more         = (inst + gridDim.x < a.n_inst);
            if (more)
                st_next                 = a.state[inst + gridDim.x];

            if (E::PIPE)
            {
                if ((tid == 0) && !xfade)
                {
                    mbar_expect_tx(&bars[1], M * sizeof(float2));
                    bulk_g2s_plain(Hs, H, M * sizeof(float2), &bars[1]);
                }
                mbar_wait(&bars[0], ph_ib);
                ph_ib ^= 1u;
            }

            if (!xfade)
            {
                /* the common block: spectrum and product stay in shared memory on the ping-pong
                 * ranks, and the inverse transform's last pass does the shift + overlap-add
                 * (:482-484) on its way out */
                if constexpr (E::PIPE)
                {
                    eq_forward<RANK>(A, B, ibs, tw, tid);       /* ends with a barrier: ibs is free */
                    if ((tid == 0) && more)                     /* fetch the next instance's block */
                    {
                        mbar_expect_tx(&bars[0], F * sizeof(float));
                        bulk_g2s_plain(ibs, ib + uint64_t(gridDim.x) * F, F * sizeof(float), &bars[0]);
                    }
                    mbar_wait(&bars[1], ph_h);
                    ph_h ^= 1u;
                    /* whole-block call on an 8-byte aligned row: the result goes straight out */
                    emitted                 = (a.off == 0) && (a.n == uint32_t(F)) &&
                                              ((reinterpret_cast<uintptr_t>(d) & 7) == 0);
                    eq_middle_inverse<RANK>(A, B, Hs, ob, emitted ? d : nullptr, &bars[1], ph_h, tw, tid);
                }
                else
                {
                    fwd_body<RANK, C::PP, 0, false, (RANK >= 12), 0, !C::TWS>(A, B, ib, xs, twg, tw, tid);
                    __syncthreads();
                    for (int k = tid; k < M; k += T)
                    {
                        float2 x    = xs[k], h = H[k];
                        ys[k]       = (k == 0) ? make_float2(x.x * h.x, x.y * h.y) : cmul(x, h);
                    }
                    __syncthreads();
                    inv_body<RANK, C::PP, 8, 0, INV_OLA, (RANK >= 12), 0, !C::TWS>(A, B, ys, 1, ob, twg, tw, true, tid);
                }
            }
            else
            {
                /* hand-over block (:486-501): both results in full, then the ramps.  Positions
                 * [F/2, F/2 + F) fade from the old to the new result, the rest of the tail is new. */
                fwd_body<RANK, C::PP, 0, false, (RANK >= 12), 0, !C::TWS>(A, B, E::PIPE ? ibs : ib, xs, twg, tw, tid);
                __syncthreads();
                if (E::PIPE && (tid == 0) && more)
                {
                    mbar_expect_tx(&bars[0], F * sizeof(float));
                    bulk_g2s_plain(ibs, ib + uint64_t(gridDim.x) * F, F * sizeof(float), &bars[0]);
                }
                for (uint32_t v = 0; v < 2; ++v)
                {
                    const float2 *Hv        = a.kern + (uint64_t(inst) * 2 + (cur ^ v)) * M;
                    for (int k = tid; k < M; k += T)
                    {
                        float2 x    = xs[k], h = Hv[k];
                        ys[k]       = (k == 0) ? make_float2(x.x * h.x, x.y * h.y) : cmul(x, h);
                    }
                    __syncthreads();
                    inv_body<RANK, C::PP, 8, 0, 0, (RANK >= 12), 0, !C::TWS>(A, B, ys, 1, v ? c1 : c0, twg, tw, true, tid);
                    __syncthreads();
                }
                constexpr int half      = F / 2;
                const float delta       = 1.0f / float(F);
                for (int j = tid; j < F; j += T)
                {
                    float lo    = ob[j + F] + c0[j];
                    float hi    = c0[j + F];
                    if (j >= half)
                    {
                        float i     = float(j - half);
                        lo          = lo * (1.0f - delta * i) + c1[j] * (delta * i);
                        hi          = c1[j + F];
                    }
                    else
                    {
                        float i     = float(j + F - half);
                        hi          = hi * (1.0f - delta * i) + c1[j + F] * (delta * i);
                    }
                    ob[j]       = lo;
                    ob[j + F]   = hi;
                }
            }
            __syncthreads();
        }

        /* sample exchange (:510-511); src is read before dst is written, so dst == src is fine */
        if (E::PIPE)
        {
            #pragma unroll
            for (int j = 0; j < E::XPT; ++j)
            {
                uint32_t i      = uint32_t(tid + j * T);
                if (i < a.n)
                {
                    if (!emitted)
                    {
                        float y         = ob[a.off + i];
                        d[i]            = y;
                    }
                    ib[a.off + i]   = xr[j];
                }
            }
        }
        else
        {
            for (uint32_t i = tid; i < a.n; i += T)
            {
                float x         = s[i];
                float y         = ob[a.off + i];
                ib[a.off + i]   = x;
                d[i]            = y;
            }
        }
        __syncthreads();
    }
}

template <int RANK>
static cudaError_t launch_eq_r(const EqArgs &a, cudaStream_t st)
{
    using C = FftCfg<RANK>;
    static bool attr_set[MAX_DEVICES] = { false };
    int dev = current_device();
    constexpr size_t SMEM = EqCfg<RANK>::SMEM;
    if ((!attr_set[dev]) && (SMEM > 48 * 1024))
    {
        cudaError_t e = cudaFuncSetAttribute(k_eq<RANK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM));
        if (e != cudaSuccess)
            return e;
    }
    attr_set[dev] = true;
    k_eq<RANK><<<resident_grid(a.n_inst, C::T, SMEM), C::T, SMEM, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t launch_eq(uint32_t rank, const EqArgs &a, cudaStream_t st)
{
    RANK_SWITCH(launch_eq_r, rank, a, st)
}

/* ------------------------------------------------------------------------------------------- */
/* host side                                                                                    */

struct b200conv_eq
{
    int             device      = 0;
    size_t          instances   = 0;
    size_t          fir_rank    = 0;        /* nFirRank */
    size_t          F           = 0;        /* nFirSize */
    size_t          buf_size    = 0;        /* nBufSize, common to the batch */
    cudaStream_t    stream      = nullptr;
    cudaEvent_t     ev_in       = nullptr, ev_out = nullptr;
    float2         *tw          = nullptr;
    float          *inbuf       = nullptr, *outbuf = nullptr, *conv = nullptr, *taps = nullptr;
    float2         *spec        = nullptr, *kern = nullptr;
    uint8_t        *d_state     = nullptr;
    Job            *d_jobs      = nullptr;
    float          *io_in       = nullptr, *io_out = nullptr;     /* staging for host callers */
    size_t          io_cap      = 0;                                /* samples per instance row */
    std::vector<uint8_t> state;             /* bit 0: current slot, bit 1: cross-fade pending */
    size_t          xfades      = 0;        /* instances with a pending cross-fade */
    uint64_t        launches    = 0;
};

static void eq_release(b200conv_eq *e)
{
    if (e == nullptr)
        return;
    DeviceScope scope(e->device);
    if (e->stream)  cudaStreamSynchronize(e->stream);
    cudaFree(e->tw); cudaFree(e->inbuf); cudaFree(e->outbuf); cudaFree(e->conv); cudaFree(e->taps);
    cudaFree(e->spec); cudaFree(e->kern); cudaFree(e->d_state); cudaFree(e->d_jobs);
    cudaFree(e->io_in); cudaFree(e->io_out);
    if (e->ev_in)   cudaEventDestroy(e->ev_in);
    if (e->ev_out)  cudaEventDestroy(e->ev_out);
    if (e->stream)  cudaStreamDestroy(e->stream);
    delete e;
}

#define ENTER_EQ(e)                                                                         \
    if ((e) == nullptr)                                                                     \
        return fail(B200CONV_ERR_ARG, "null equalizer handle");                             \
    DeviceScope device_scope_((e)->device);                                                 \
    if (device_scope_.error() != cudaSuccess)                                               \
        return fail(B200CONV_ERR_CUDA, "cannot select device %d: %s", (e)->device,          \
                    cudaGetErrorString(device_scope_.error()))

static int eq_create_impl(b200conv_eq **out, int device, size_t instances, size_t fir_rank)
{
    if (out == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_eq_create: null output pointer");
    *out = nullptr;
    if ((instances == 0) || (instances > (size_t(1) << 24)))
        return fail(B200CONV_ERR_ARG, "b200conv_eq_create: %zu instances not supported", instances);
    if ((fir_rank + 1 < B200CONV_RANK_MIN) || (fir_rank + 1 > B200CONV_RANK_MAX))
        return fail(B200CONV_ERR_ARG, "b200conv_eq_create: fir_rank %zu outside [%d, %d]", fir_rank,
                    B200CONV_RANK_MIN - 1, B200CONV_RANK_MAX - 1);
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if ((ce != cudaSuccess) || (n == 0))
        return fail(B200CONV_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(ce));
    if (device < 0)
        cudaGetDevice(&device);
    if ((device >= n) || (device >= MAX_DEVICES))
        return fail(B200CONV_ERR_ARG, "device %d out of range", device);

    b200conv_eq *e  = new b200conv_eq();
    e->device       = device;
    e->instances    = instances;
    e->fir_rank     = fir_rank;
    e->F            = size_t(1) << fir_rank;
    e->state.assign(instances, 0);
    const size_t F  = e->F, N = 2 * F, M = F;

    DeviceScope scope(device);
    int rc = B200CONV_OK;
    do {
        #define EQ_BRK(call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail((e_ == cudaErrorMemoryAllocation) ? B200CONV_ERR_NOMEM : B200CONV_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); break; } }
        if (scope.error() != cudaSuccess) { rc = fail(B200CONV_ERR_CUDA, "cannot select device %d", device); break; }
        EQ_BRK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        EQ_BRK(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
        EQ_BRK(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
        if ((rc = make_twiddles(uint32_t(fir_rank + 1), &e->tw)) != B200CONV_OK) break;
        EQ_BRK(cudaMalloc(&e->inbuf, instances * F * sizeof(float)));
        EQ_BRK(cudaMalloc(&e->outbuf, instances * N * sizeof(float)));
        EQ_BRK(cudaMalloc(&e->conv, instances * 2 * N * sizeof(float)));
        EQ_BRK(cudaMalloc(&e->taps, instances * F * sizeof(float)));
        EQ_BRK(cudaMalloc(&e->spec, instances * 2 * M * sizeof(float2)));
        EQ_BRK(cudaMalloc(&e->kern, instances * 2 * M * sizeof(float2)));
        EQ_BRK(cudaMalloc(&e->d_state, instances));
        EQ_BRK(cudaMalloc(&e->d_jobs, instances * sizeof(Job)));
        EQ_BRK(cudaMemsetAsync(e->inbuf, 0, instances * F * sizeof(float), e->stream));
        EQ_BRK(cudaMemsetAsync(e->outbuf, 0, instances * N * sizeof(float), e->stream));
        EQ_BRK(cudaMemsetAsync(e->kern, 0, instances * 2 * M * sizeof(float2), e->stream));
        EQ_BRK(cudaMemsetAsync(e->d_state, 0, instances, e->stream));
        EQ_BRK(cudaStreamSynchronize(e->stream));
        #undef EQ_BRK
    } while (0);
    if (rc != B200CONV_OK)
    {
        std::string keep = g_last_error;
        eq_release(e);
        g_last_error = keep;
        return rc;
    }
    *out = e;
    return B200CONV_OK;
}

extern "C" int b200conv_eq_create(b200conv_eq **out, int device, size_t instances, size_t fir_rank)
{
    try { return eq_create_impl(out, device, instances, fir_rank); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

extern "C" void b200conv_eq_free(b200conv_eq *e)
{
    eq_release(e);
}

extern "C" size_t b200conv_eq_fir_size(const b200conv_eq *e)    { return (e != nullptr) ? e->F : 0; }
extern "C" size_t b200conv_eq_latency(const b200conv_eq *e)     { return (e != nullptr) ? e->F : 0; }
extern "C" size_t b200conv_eq_instances(const b200conv_eq *e)   { return (e != nullptr) ? e->instances : 0; }

/* Equalizer.cpp:336-345: the finished impulse response goes through fastconv_parse into vConv,
 * or into vNewConv with EF_XFADE raised when EF_SMOOTH is set */
extern "C" int b200conv_eq_set_kernel(b200conv_eq *e, size_t idx, const float *ir, int smooth)
{
    ENTER_EQ(e);
    if ((idx >= e->instances) || (ir == nullptr))
        return fail(B200CONV_ERR_ARG, "b200conv_eq_set_kernel: bad arguments");
    const size_t F  = e->F, M = F;
    const uint32_t cur  = e->state[idx] & 1u;
    const uint32_t slot = smooth ? (cur ^ 1u) : cur;

    Job j;
    memset(&j, 0, sizeof(j));
    j.src   = e->taps + idx * F;
    j.spec  = e->kern + (idx * 2 + slot) * M;
    /* pageable sources: both copies are staged by the runtime before the call returns */
    CU(cudaMemcpyAsync(e->taps + idx * F, ir, F * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_jobs + idx, &j, sizeof(Job), cudaMemcpyHostToDevice, e->stream));
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.jobs      = e->d_jobs + idx;
    a.tw        = e->tw;
    a.rank      = uint32_t(e->fir_rank + 1);
    a.splits    = 1;
    CU(launch_fwd(a, 1, e->stream));
    e->launches++;
    if (smooth && !(e->state[idx] & 2u))
    {
        e->state[idx]  |= 2u;
        e->xfades++;
        CU(cudaMemcpyAsync(e->d_state + idx, &e->state[idx], 1, cudaMemcpyHostToDevice, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));       /* `ir` and `j` are the caller's / the stack's */
    return B200CONV_OK;
}

/* Equalizer.cpp:273-278 (EF_CLEAR) */
extern "C" int b200conv_eq_clear(b200conv_eq *e)
{
    ENTER_EQ(e);
    CU(cudaMemsetAsync(e->inbuf, 0, e->instances * e->F * sizeof(float), e->stream));
    CU(cudaMemsetAsync(e->outbuf, 0, e->instances * 2 * e->F * sizeof(float), e->stream));
    e->buf_size = 0;
    return B200CONV_OK;
}

/* the loop of Equalizer.cpp:477-518 on e->stream; src / dst are device matrices */
static int eq_run(b200conv_eq *e, float *dst, size_t dst_stride, const float *src, size_t src_stride,
                  size_t samples)
{
    const size_t F  = e->F;
    size_t pos      = 0;
    while (samples > 0)
    {
        EqArgs a;
        memset(&a, 0, sizeof(a));
        a.do_block      = (e->buf_size >= F) ? 1u : 0u;
        if (a.do_block)
            e->buf_size     = 0;
        size_t n        = (samples < F - e->buf_size) ? samples : F - e->buf_size;
        a.tw            = e->tw;
        a.src           = src + pos;
        a.dst           = dst + pos;
        a.stride_src    = src_stride;
        a.stride_dst    = dst_stride;
        a.inbuf         = e->inbuf;
        a.outbuf        = e->outbuf;
        a.spec          = e->spec;
        a.conv          = e->conv;
        a.kern          = e->kern;
        a.state         = e->d_state;
        a.n_inst        = uint32_t(e->instances);
        a.off           = uint32_t(e->buf_size);
        a.n             = uint32_t(n);
        CU(launch_eq(uint32_t(e->fir_rank + 1), a, e->stream));
        e->launches++;

        if (a.do_block && (e->xfades > 0))
        {
            /* the block just launched consumed every pending hand-over: vConv = vNewConv (:491) */
            for (size_t i = 0; i < e->instances; ++i)
                if (e->state[i] & 2u)
                    e->state[i]     = uint8_t((e->state[i] ^ 1u) & 1u);
            e->xfades       = 0;
            CU(cudaMemcpyAsync(e->d_state, e->state.data(), e->instances, cudaMemcpyHostToDevice, e->stream));
        }
        e->buf_size    += n;
        pos            += n;
        samples        -= n;
    }
    return B200CONV_OK;
}

extern "C" int b200conv_eq_process_device(b200conv_eq *e, float *dst, size_t dst_stride, const float *src,
                                          size_t src_stride, size_t samples, void *stream)
{
    ENTER_EQ(e);
    if (samples == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr) || (dst_stride < samples) || (src_stride < samples))
        return fail(B200CONV_ERR_ARG, "b200conv_eq_process_device: bad arguments");
    cudaStream_t caller = cudaStream_t(stream);
    const bool foreign  = (caller != nullptr) && (caller != e->stream);
    if (foreign)
    {
        CU(cudaEventRecord(e->ev_in, caller));
        CU(cudaStreamWaitEvent(e->stream, e->ev_in, 0));
    }
    TRY(eq_run(e, dst, dst_stride, src, src_stride, samples));
    if (foreign)
    {
        CU(cudaEventRecord(e->ev_out, e->stream));
        CU(cudaStreamWaitEvent(caller, e->ev_out, 0));
    }
    return B200CONV_OK;
}

static int eq_stage(b200conv_eq *e, size_t samples)
{
    if (samples <= e->io_cap)
        return B200CONV_OK;
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(e->io_in);  e->io_in = nullptr;
    cudaFree(e->io_out); e->io_out = nullptr;
    e->io_cap = 0;
    CU(cudaMalloc(&e->io_in, e->instances * samples * sizeof(float)));
    CU(cudaMalloc(&e->io_out, e->instances * samples * sizeof(float)));
    e->io_cap = samples;
    return B200CONV_OK;
}

static size_t eq_chunk(const b200conv_eq *e, size_t samples)
{
    size_t cap = (e->F * 16 > 65536) ? e->F * 16 : 65536;
    return (samples < cap) ? samples : cap;
}

/* Equalizer::process (EQM_FIR / EQM_FFT) for every instance, one planar HOST matrix each way */
extern "C" int b200conv_eq_process_planar(b200conv_eq *e, float *dst, const float *src, size_t stride,
                                          size_t samples)
{
    ENTER_EQ(e);
    if (samples == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr) || (stride < samples))
        return fail(B200CONV_ERR_ARG, "b200conv_eq_process_planar: bad arguments");
    const size_t chunk = eq_chunk(e, samples);
    TRY(eq_stage(e, chunk));
    for (size_t pos = 0; pos < samples; pos += chunk)
    {
        size_t n = (samples - pos < chunk) ? samples - pos : chunk;
        CU(cudaMemcpy2DAsync(e->io_in, e->io_cap * sizeof(float), src + pos, stride * sizeof(float),
                             n * sizeof(float), e->instances, cudaMemcpyHostToDevice, e->stream));
        TRY(eq_run(e, e->io_out, e->io_cap, e->io_in, e->io_cap, n));
        CU(cudaMemcpy2DAsync(dst + pos, stride * sizeof(float), e->io_out, e->io_cap * sizeof(float),
                             n * sizeof(float), e->instances, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return B200CONV_OK;
}

/* the same with one host pointer per instance (the facade's shape: Equalizer::process(out, in, n)) */
extern "C" int b200conv_eq_process(b200conv_eq *e, float *const *dst, const float *const *src, size_t samples)
{
    ENTER_EQ(e);
    if (samples == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr))
        return fail(B200CONV_ERR_ARG, "b200conv_eq_process: bad arguments");
    for (size_t i = 0; i < e->instances; ++i)
        if ((dst[i] == nullptr) || (src[i] == nullptr))
            return fail(B200CONV_ERR_ARG, "b200conv_eq_process: null buffer for instance %zu", i);
    const size_t chunk = eq_chunk(e, samples);
    TRY(eq_stage(e, chunk));
    for (size_t pos = 0; pos < samples; pos += chunk)
    {
        size_t n = (samples - pos < chunk) ? samples - pos : chunk;
        for (size_t i = 0; i < e->instances; ++i)
            CU(cudaMemcpyAsync(e->io_in + i * e->io_cap, src[i] + pos, n * sizeof(float),
                               cudaMemcpyHostToDevice, e->stream));
        TRY(eq_run(e, e->io_out, e->io_cap, e->io_in, e->io_cap, n));
        for (size_t i = 0; i < e->instances; ++i)
            CU(cudaMemcpyAsync(dst[i] + pos, e->io_out + i * e->io_cap, n * sizeof(float),
                               cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return B200CONV_OK;
}

extern "C" int b200conv_eq_sync(b200conv_eq *e)
{
    ENTER_EQ(e);
    CU(cudaStreamSynchronize(e->stream));
    return B200CONV_OK;
}

extern "C" void *b200conv_eq_stream(b200conv_eq *e)
{
    return (e != nullptr) ? e->stream : nullptr;
}

#endif /* B200CONV_EQUALIZER_CUH_ */
