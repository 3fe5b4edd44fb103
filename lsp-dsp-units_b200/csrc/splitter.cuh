/*
 * splitter.cuh -- lsp::dspu::SpectralSplitter, batched over many instances (scope-table row f4, second
 * sibling; kernel + host side + C ABI b200conv_ss_*; included by engine.cu after spectral_host.cuh).
 *
 * Reference: src/main/util/SpectralSplitter.cpp.  One forward transform of the last N = 2^rank input
 * samples every F = 2^(chunk_rank - 1) samples (:313-314, no window in front); every bound handler
 * ("band") gets the spectrum through its own function, transforms back, keeps the LAST 2 F samples
 * (:318-324) -- a handler without a function keeps the FIRST 2 F samples of the input window instead
 * (:327) --, multiplies them by sin^2 (:244, misc/windows.cpp:249-260) and overlap-adds them with hop
 * F into its own output stream, which a sink callback reads F samples per hop (:331-340,365-371):
 * latency 2 F = 2^chunk_rank (:289).  lsp::dspu::FFTCrossover is this class with a real gain curve per
 * band as the function (FFTCrossover.cpp:124-140: pcomplex_r2c_mul2 with vFFT, built on the host by
 * crossover::hipass_fft_set / lopass_fft_apply, :458-480).
 *
 * Here: ONE launch per process call for the whole batch (k_ss).  A CTA owns one instance and walks
 * the call like the reference's loop (:305-382).  The host callbacks cannot run on the device, so a
 * handler is one of
 *     table       function out[k] = in[k] * H[k] (complex table, or real gains = FFTCrossover's band)
 *                 + sink: the table is folded at bind time to the N / 2 + 1 unique bins exactly as for
 *                 the SpectralProcessor (spectral.cuh), the forward transform is shared by the bands;
 *     sink only   no function (:327);
 * and the sink of handler h is row h of the output: dst[h * band_stride + instance * dst_stride + i].
 * The real N-point transforms are N / 2-point complex transforms plus a split / merge pass, both work
 * buffers resident: ranks 7 .. 14.
 */
#ifndef B200CONV_SPLITTER_CUH_
#define B200CONV_SPLITTER_CUH_

struct SsArgs
{
    const float2   *tw;             /* twiddle table of transform rank `rank + 1` (FftCfg<rank + 1>)     */
    const float    *wnd;            /* [2 F] sin^2 window                                                */
    const float    *src;            /* [instances][stride_src]                                           */
    float          *dst;            /* [handlers][band_stride]: row h = [instances][stride_dst]           */
    uint64_t        stride_src, stride_dst, band_stride;
    float          *inbuf;          /* [instances][N]: the last N input samples, newest last               */
    float          *outbuf;         /* [instances][handlers][N]: the first 2 F floats are the overlap buffer */
    const float2   *table;          /* [instances][handlers][N/2 + 1] folded tables                        */
    const uint8_t  *kind;           /* [instances][handlers] 0: unbound, 1: table + sink, 2: sink only      */
    uint32_t       *fill;           /* [instances] nFrameSize                                              */
    uint32_t        n_inst, handlers;
    uint32_t        frame;          /* F                                                                   */
    uint32_t        count;          /* samples of this call                                                */
};

/* RANKP = rank + 1: FftCfg<RANKP, 0, 1> is ONE sequence of P = N / 2 complex points. */
template <int RANKP>
struct SsCfg
{
    using C = FftCfg<RANKP, 0, 1>;
    static constexpr int  P     = C::P;
    static constexpr int  N     = 2 * P;
    static constexpr bool PPF   = (RANKP <= 13);    /* the forward transform may ping-pong between the two buffers */
    static constexpr bool TWS   = (RANKP <= 12);    /* whole twiddle table in shared memory                        */
    static constexpr int  TWN   = TWS ? C::TW_TOTAL : C::TWC_N;
    static constexpr size_t SMEM = (size_t(P) * 2 + TWN) * sizeof(float2);
};

template <int RANKP>
__global__ void __launch_bounds__(FftCfg<RANKP, 0, 1>::T)
k_ss(const SsArgs a)
{
    using S = SsCfg<RANKP>;
    using C = typename S::C;
    constexpr int P = S::P, N = S::N, T = C::T;
    extern __shared__ float2 ss_sm[];
    float2 *A               = ss_sm;
    float2 *B               = ss_sm + P;
    float2 *tws             = ss_sm + 2 * P;
    const int tid           = threadIdx.x;
    const uint32_t F        = a.frame;

    if (S::TWS)
        for (int i = tid; i < C::TW_TOTAL; i += T)
            tws[i]              = a.tw[i];
    else
        stage_compact_twiddles<C>(tws, a.tw, tid);
    __syncthreads();
    const float2 *wN        = (S::TWS ? tws : a.tw) + C::TW_PRE;    /* exp(-2 pi i k / N), k < N / 2 */

    for (uint32_t inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x)
    {
        const uint8_t *kind     = a.kind + uint64_t(inst) * a.handlers;
        bool any = false, any_func = false;
        for (uint32_t h = 0; h < a.handlers; ++h)
        {
            any                    |= (kind[h] != 0);
            any_func               |= (kind[h] == 1);
        }
        if (!any)
            continue;                                               /* :301-302: nothing bound, nothing happens */

        float *inb              = a.inbuf + uint64_t(inst) * N;
        const float *src        = a.src + uint64_t(inst) * a.stride_src;
        uint32_t fill           = a.fill[inst];
        uint32_t pos            = 0;

        while (pos < a.count)                                       /* :305 */
        {
            if (fill >= F)                                          /* :308 : a frame boundary */
            {
                float2 *Z = A, *D = B;
                if (any_func)
                {
                    /* :313-314 the window as N / 2 complex points, forward transform */
                    for (int m = tid; m < P; m += T)
                        A[m]            = reinterpret_cast<const float2 *>(inb)[m];
                    __syncthreads();
                    Z                   = fft_smem<RANKP, false, S::PPF, 0, true, true, 1, !S::TWS>(A, S::PPF ? B : nullptr, tws, tid);
                    D                   = (Z == A) ? B : A;
                }
                for (uint32_t h = 0; h < a.handlers; ++h)
                {
                    if (kind[h] == 0)
                        continue;
                    float *outb         = a.outbuf + (uint64_t(inst) * a.handlers + h) * N;
                    if (kind[h] == 1)
                    {
                        const float2 *H     = a.table + (uint64_t(inst) * a.handlers + h) * (P + 1);
                        /* split -> X[k], :320 the band's function X[k] * H[k], merge -> the packed
                         * spectrum of the band in D (Z stays for the next band) */
                        for (int k = tid; k <= P / 2; k += T)
                        {
                            const int km    = P - k;
                            if (k == 0)
                            {
                                const float2 z0 = Z[0];
                                const float x0  = (z0.x + z0.y) * H[0].x;
                                const float xn  = (z0.x - z0.y) * H[P].x;
                                D[0]            = make_float2(0.5f * (x0 + xn), 0.5f * (x0 - xn));
                                continue;
                            }
                            const float2 zk = Z[k], zm = Z[km];
                            const float2 w  = wN[k];
                            const float2 E  = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                            const float2 O  = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
                            const float2 wO = cmul(w, O);
                            float2 Xk       = cmul(make_float2(E.x + wO.x, E.y + wO.y), H[k]);
                            float2 Xm       = cmul(make_float2(E.x - wO.x, -E.y + wO.y), H[km]);
                            const float2 E2 = make_float2(0.5f * (Xk.x + Xm.x), 0.5f * (Xk.y - Xm.y));
                            const float2 O2 = cmulc(make_float2(0.5f * (Xk.x - Xm.x), 0.5f * (Xk.y + Xm.y)), w);
                            D[k]            = make_float2(E2.x - O2.y, E2.y + O2.x);
                            if (km != k)
                                D[km]           = make_float2(E2.x + O2.y, O2.x - E2.y);
                        }
                        __syncthreads();
                        /* :321 reverse transform, in place (1 / P: the half-size transform carries the whole scale) */
                        float2 *Y           = fft_smem<RANKP, true, false, 0, true, true, 1, !S::TWS>(D, nullptr, tws, tid);
                        const float *y      = reinterpret_cast<const float *>(Y) + (N - 2 * F);     /* :322 the last 2 F samples */
                        const float scale   = 1.0f / float(P);
                        /* :331-340 shift the band's overlap buffer by F, clear its tail, add the frame times the window */
                        for (uint32_t i = tid; i < F; i += T)
                        {
                            const float lo  = outb[i + F] + y[i] * scale * a.wnd[i];
                            const float hi  = y[i + F] * scale * a.wnd[i + F];
                            outb[i]         = lo;
                            outb[i + F]     = hi;
                        }
                        __syncthreads();            /* D is rewritten by the next band */
                    }
                    else
                    {
                        /* :327 no function: the first 2 F samples of the window */
                        for (uint32_t i = tid; i < F; i += T)
                        {
                            const float lo  = outb[i + F] + inb[i] * a.wnd[i];
                            const float hi  = inb[i + F] * a.wnd[i + F];
                            outb[i]         = lo;
                            outb[i + F]     = hi;
                        }
                    }
                }
                __syncthreads();
                /* :344-351 the window slides by F (staged through shared memory: the ranges overlap) */
                float *stage        = reinterpret_cast<float *>(A);
                for (uint32_t i = tid; i < uint32_t(N) - F; i += T)
                    stage[i]            = inb[i + F];
                __syncthreads();
                for (uint32_t i = tid; i < uint32_t(N) - F; i += T)
                    inb[i]              = stage[i];
                fill                = 0;
                __syncthreads();
            }

            /* :358-371 take the samples up to the next frame boundary, hand every sink its share */
            const uint32_t n    = min(F - fill, a.count - pos);
            for (uint32_t i = tid; i < n; i += T)
                inb[uint32_t(N) - F + fill + i] = src[pos + i];
            for (uint32_t h = 0; h < a.handlers; ++h)
            {
                if (kind[h] == 0)
                    continue;
                const float *outb   = a.outbuf + (uint64_t(inst) * a.handlers + h) * N;
                float *dsth         = a.dst + uint64_t(h) * a.band_stride + uint64_t(inst) * a.stride_dst;
                for (uint32_t i = tid; i < n; i += T)
                    dsth[pos + i]       = outb[fill + i];
            }
            fill               += n;
            pos                += n;
            __syncthreads();
        }
        if (tid == 0)
            a.fill[inst]        = fill;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* host side                                                                                    */

static const size_t SS_RANK_MIN = 7, SS_RANK_MAX = 14;      /* transform rank + 1 in 8..15: two work buffers resident */

struct b200conv_ss
{
    int                     device      = 0;
    int                     sm_count    = 148;
    size_t                  n           = 0, handlers = 0;
    size_t                  max_rank    = 0, rank = 0;
    long                    user_chunk  = 0;                /* nUserChunkRank (construct(): 0)        */
    size_t                  chunk_rank  = 0;                /* nChunkRank, valid once committed       */
    bool                    update      = true;             /* rank / chunk rank changed: every instance restarts */
    std::vector<uint8_t>    dirty;                          /* bUpdate per instance                    */
    cudaStream_t            stream      = nullptr;
    float2                 *tw[B200CONV_RANK_MAX + 1] = { nullptr };
    float                  *d_wnd       = nullptr;          /* [2^max_rank]                            */
    float                  *d_in        = nullptr;          /* [n][2^max_rank]                         */
    float                  *d_out       = nullptr;          /* [n][handlers][2^max_rank]               */
    float2                 *d_table     = nullptr;          /* [n][handlers][2^(max_rank-1) + 1]       */
    uint8_t                *d_kind      = nullptr;
    uint32_t               *d_fill      = nullptr;
    std::vector<float>      phase;
    std::vector<uint8_t>    kind;
    bool                    kind_dirty  = true;
    float                  *sd_in = nullptr, *sd_out = nullptr;     /* staging of the host entry point */
    size_t                  stage_in = 0, stage_out = 0;
};

typedef b200conv_ss Ss;

template <int RANKP>
static cudaError_t launch_ss_r(const SsArgs &a, uint32_t grid, cudaStream_t st)
{
    using S = SsCfg<RANKP>;
    static bool attr_set[MAX_DEVICES] = { false };
    int dev = current_device();
    if ((!attr_set[dev]) && (S::SMEM > 48 * 1024))
    {
        cudaError_t e = cudaFuncSetAttribute(k_ss<RANKP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::SMEM));
        if (e != cudaSuccess)
            return e;
    }
    attr_set[dev] = true;
    k_ss<RANKP><<<grid, S::C::T, S::SMEM, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t launch_ss(const SsArgs &a, size_t rank, uint32_t grid, cudaStream_t st)
{
    switch (rank + 1)
    {
        case 8:  return launch_ss_r<8>(a, grid, st);
        case 9:  return launch_ss_r<9>(a, grid, st);
        case 10: return launch_ss_r<10>(a, grid, st);
        case 11: return launch_ss_r<11>(a, grid, st);
        case 12: return launch_ss_r<12>(a, grid, st);
        case 13: return launch_ss_r<13>(a, grid, st);
        case 14: return launch_ss_r<14>(a, grid, st);
        case 15: return launch_ss_r<15>(a, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

extern "C" void b200conv_ss_free(b200conv_ss_t *s)
{
    if (s == nullptr)
        return;
    DeviceScope device_scope_(s->device);
    if (s->stream)  cudaStreamSynchronize(s->stream);
    for (float2 *t : s->tw)
        if (t) cudaFree(t);
    if (s->d_wnd)   cudaFree(s->d_wnd);
    if (s->d_in)    cudaFree(s->d_in);
    if (s->d_out)   cudaFree(s->d_out);
    if (s->d_table) cudaFree(s->d_table);
    if (s->d_kind)  cudaFree(s->d_kind);
    if (s->d_fill)  cudaFree(s->d_fill);
    if (s->sd_in)   cudaFree(s->sd_in);
    if (s->sd_out)  cudaFree(s->sd_out);
    if (s->stream)  cudaStreamDestroy(s->stream);
    delete s;
}

static int ss_create_impl(b200conv_ss_t **out, int device, size_t instances, size_t max_rank, size_t handlers)
{
    if ((out == nullptr) || (instances == 0) || (instances > (size_t(1) << 20)) || (handlers == 0) || (handlers > 64) ||
        (max_rank < SS_RANK_MIN) || (max_rank > SS_RANK_MAX))
        return fail(B200CONV_ERR_ARG, "b200conv_ss_create: bad arguments (ranks %zu..%zu, 1..64 handlers)", SS_RANK_MIN, SS_RANK_MAX);
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if ((e != cudaSuccess) || (count == 0))
        return fail(B200CONV_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0)
        CU(cudaGetDevice(&device));
    if (device >= count)
        return fail(B200CONV_ERR_ARG, "device %d out of range (%d devices)", device, count);

    Ss *s = new (std::nothrow) Ss();
    if (s == nullptr)
        return fail(B200CONV_ERR_NOMEM, "out of host memory");
    s->device   = device;
    s->n        = instances;
    s->handlers = handlers;
    s->max_rank = s->rank = max_rank;               /* SpectralSplitter.cpp:68-69 */
    s->phase.assign(instances, 0.0f);
    s->dirty.assign(instances, 1);                  /* bUpdate = true (:79) */
    s->kind.assign(instances * handlers, 0);

    ENTER_DEVICE(s);
    const size_t N = size_t(1) << max_rank;
    int rc = B200CONV_OK;
    do
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
            s->sm_count = prop.multiProcessorCount;
        #define CU_BRK(call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail((e_ == cudaErrorMemoryAllocation) ? B200CONV_ERR_NOMEM : B200CONV_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); break; } }
        CU_BRK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CU_BRK(cudaMalloc(&s->d_wnd, N * sizeof(float)));
        CU_BRK(cudaMalloc(&s->d_in, instances * N * sizeof(float)));
        CU_BRK(cudaMalloc(&s->d_out, instances * handlers * N * sizeof(float)));
        CU_BRK(cudaMalloc(&s->d_table, instances * handlers * (N / 2 + 1) * sizeof(float2)));
        CU_BRK(cudaMalloc(&s->d_kind, instances * handlers));
        CU_BRK(cudaMalloc(&s->d_fill, instances * sizeof(uint32_t)));
        CU_BRK(cudaMemset(s->d_kind, 0, instances * handlers));
        #undef CU_BRK
    } while (false);
    if (rc != B200CONV_OK)
    {
        std::string keep = g_last_error;
        b200conv_ss_free(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return B200CONV_OK;
}

extern "C" int b200conv_ss_create(b200conv_ss_t **out, int device, size_t instances, size_t max_rank, size_t handlers)
{
    try { return ss_create_impl(out, device, instances, max_rank, handlers); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

/* nChunkRank as update_settings computes it (SpectralSplitter.cpp:232-233) */
static size_t ss_chunk_rank(const Ss *s)
{
    if (s->user_chunk <= 0)
        return s->rank;
    long r = s->user_chunk;
    if (r < 5)              r = 5;
    if (r > long(s->rank))  r = long(s->rank);
    return size_t(r);
}

extern "C" int b200conv_ss_set_rank(b200conv_ss_t *s, size_t rank)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_set_rank: NULL handle");
    if ((rank == s->rank) || (rank > s->max_rank))          /* :273-274 */
        return B200CONV_OK;
    if (rank < SS_RANK_MIN)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_set_rank: ranks below %zu are not supported", SS_RANK_MIN);
    s->rank         = rank;
    s->update       = true;
    /* the tables are rank-specific (the reference's callbacks are not): bind again */
    std::fill(s->kind.begin(), s->kind.end(), uint8_t(0));
    s->kind_dirty   = true;
    return B200CONV_OK;
}

extern "C" int b200conv_ss_set_chunk_rank(b200conv_ss_t *s, long rank)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_set_chunk_rank: NULL handle");
    if (rank == s->user_chunk)                              /* :282-283 */
        return B200CONV_OK;
    s->user_chunk   = rank;
    s->update       = true;
    return B200CONV_OK;
}

extern "C" int b200conv_ss_set_phase(b200conv_ss_t *s, size_t idx, float phase)
{
    if ((s == nullptr) || (idx >= s->n))
        return fail(B200CONV_ERR_ARG, "b200conv_ss_set_phase: bad handle or index");
    s->phase[idx]   = (phase < 0.0f) ? 0.0f : ((phase > 1.0f) ? 1.0f : phase);      /* :266 */
    s->dirty[idx]   = 1;                                    /* bUpdate = true (:267) */
    return B200CONV_OK;
}

/* kind 0: unbind; 1: complex table of 2^rank bins; 2: 2^rank real gains; 3: sink only */
static int ss_bind(b200conv_ss_t *s, size_t idx, size_t handler, int kind, const float *table)
{
    if ((s == nullptr) || (idx >= s->n) || (handler >= s->handlers) || (((kind == 1) || (kind == 2)) && (table == nullptr)))
        return fail(B200CONV_ERR_ARG, "b200conv_ss_bind: bad arguments");
    ENTER_DEVICE(s);
    uint8_t &slot = s->kind[idx * s->handlers + handler];
    if (kind == 0)
    {
        if (slot == 0)
            return fail(B200CONV_ERR_STATE, "b200conv_ss_unbind: the handler is not bound");   /* STATUS_NOT_BOUND, :190-191 */
        slot            = 0;
        s->kind_dirty   = true;
        return B200CONV_OK;
    }
    /* the state arrays are allocated for max_rank and indexed with the CURRENT rank's N (every rank
     * change unbinds and clears everything, so the packing is consistent) */
    const size_t N = size_t(1) << s->rank, P = N / 2;
    CU(cudaStreamSynchronize(s->stream));
    if (kind != 3)
    {
        /* Re(IFFT(X H)) for real input = the half-spectrum product with (H[k] + conj(H[N - k])) / 2 */
        std::vector<float2> folded(P + 1);
        for (size_t k = 0; k <= P; ++k)
        {
            const size_t km = (N - k) % N;
            if (kind == 1)
                folded[k]   = make_float2(0.5f * (table[2 * k] + table[2 * km]), 0.5f * (table[2 * k + 1] - table[2 * km + 1]));
            else
                folded[k]   = make_float2(0.5f * (table[k] + table[km]), 0.0f);
        }
        CU(cudaMemcpy(s->d_table + (idx * s->handlers + handler) * (P + 1), folded.data(), (P + 1) * sizeof(float2),
                      cudaMemcpyHostToDevice));
    }
    /* bind() clears the handler's output buffer (:175-176) */
    CU(cudaMemset(s->d_out + (idx * s->handlers + handler) * N, 0, N * sizeof(float)));
    slot            = (kind == 3) ? 2 : 1;
    s->kind_dirty   = true;
    return B200CONV_OK;
}

extern "C" int b200conv_ss_bind_complex(b200conv_ss_t *s, size_t idx, size_t handler, const float *table)
{
    try { return ss_bind(s, idx, handler, 1, table); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

extern "C" int b200conv_ss_bind_gain(b200conv_ss_t *s, size_t idx, size_t handler, const float *gain)
{
    try { return ss_bind(s, idx, handler, 2, gain); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

extern "C" int b200conv_ss_bind_sink(b200conv_ss_t *s, size_t idx, size_t handler)
{
    return ss_bind(s, idx, handler, 3, nullptr);
}

extern "C" int b200conv_ss_unbind(b200conv_ss_t *s, size_t idx, size_t handler)
{
    return ss_bind(s, idx, handler, 0, nullptr);
}

extern "C" int b200conv_ss_unbind_all(b200conv_ss_t *s, size_t idx)
{
    if ((s == nullptr) || (idx >= s->n))
        return fail(B200CONV_ERR_ARG, "b200conv_ss_unbind_all: bad handle or index");
    for (size_t h = 0; h < s->handlers; ++h)
        s->kind[idx * s->handlers + h] = 0;
    s->kind_dirty   = true;
    return B200CONV_OK;
}

extern "C" size_t b200conv_ss_bindings(const b200conv_ss_t *s, size_t idx)
{
    if ((s == nullptr) || (idx >= s->n))
        return 0;
    size_t c = 0;
    for (size_t h = 0; h < s->handlers; ++h)
        c              += (s->kind[idx * s->handlers + h] != 0) ? 1 : 0;
    return c;
}

/* SpectralSplitter::clear (:247-260) */
static int ss_clear(Ss *s, cudaStream_t st)
{
    const size_t Nmax = size_t(1) << s->max_rank;
    CU(cudaMemsetAsync(s->d_in, 0, s->n * Nmax * sizeof(float), st));
    CU(cudaMemsetAsync(s->d_out, 0, s->n * s->handlers * Nmax * sizeof(float), st));
    return B200CONV_OK;
}

extern "C" int b200conv_ss_clear(b200conv_ss_t *s)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_clear: NULL handle");
    ENTER_DEVICE(s);
    return ss_clear(s, s->stream);
}

/* SpectralSplitter::update_settings (:227-245), then the tables the kernel reads */
static int ss_commit(Ss *s, cudaStream_t st)
{
    bool any_dirty = s->update;
    for (size_t i = 0; i < s->n; ++i)
        any_dirty       = any_dirty || (s->dirty[i] != 0);
    if (any_dirty)
    {
        const bool all  = s->update;
        s->chunk_rank   = ss_chunk_rank(s);
        const size_t N  = size_t(1) << s->rank;
        const size_t F  = size_t(1) << (s->chunk_rank - 1);
        /* windows::sqr_cosine (misc/windows.cpp:249-260): f = M_PI / n in fp32, a = sinf(f * i), a * a */
        std::vector<float> w(2 * F);
        const float f   = float(M_PI / double(2 * F));
        for (size_t i = 0; i < 2 * F; ++i)
        {
            const float a   = sinf(f * float(i));
            w[i]            = a * a;
        }
        if (all)
            TRY(ss_clear(s, st));
        CU(cudaMemcpyAsync(s->d_wnd, w.data(), 2 * F * sizeof(float), cudaMemcpyHostToDevice, st));
        for (size_t i = 0; i < s->n; ++i)
        {
            if ((!all) && (!s->dirty[i]))
                continue;
            const uint32_t fill = uint32_t(size_t(float(F) * (s->phase[i] * 0.5f)));        /* :241, fp32 */
            if (!all)
            {
                CU(cudaMemsetAsync(s->d_in + i * N, 0, N * sizeof(float), st));            /* clear(), :239 */
                CU(cudaMemsetAsync(s->d_out + i * s->handlers * N, 0, s->handlers * N * sizeof(float), st));
            }
            CU(cudaMemcpyAsync(s->d_fill + i, &fill, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            s->dirty[i]     = 0;
        }
        CU(cudaStreamSynchronize(st));          /* pageable sources about to go out of scope */
        s->update       = false;
    }
    if (s->kind_dirty)
    {
        CU(cudaMemcpyAsync(s->d_kind, s->kind.data(), s->kind.size(), cudaMemcpyHostToDevice, st));
        s->kind_dirty   = false;
    }
    if (s->tw[s->rank + 1] == nullptr)
        TRY(make_twiddles(uint32_t(s->rank + 1), &s->tw[s->rank + 1]));
    return B200CONV_OK;
}

static int ss_process_device_impl(b200conv_ss_t *s, float *dst, size_t band_stride, size_t dst_stride, const float *src,
                                  size_t src_stride, size_t count, void *stream)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_process_device: NULL handle");
    if (count == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr) || (count >= (size_t(1) << 32)) ||
        ((s->n > 1) && ((src_stride < count) || (dst_stride < count))) ||
        ((s->handlers > 1) && (band_stride < (s->n - 1) * dst_stride + count)))
        return fail(B200CONV_ERR_ARG, "b200conv_ss_process_device: bad buffers");
    ENTER_DEVICE(s);
    cudaStream_t st = (stream != nullptr) ? cudaStream_t(stream) : s->stream;
    TRY(ss_commit(s, st));

    SsArgs a;
    memset(&a, 0, sizeof(a));
    a.tw            = s->tw[s->rank + 1];
    a.wnd           = s->d_wnd;
    a.src           = src;
    a.dst           = dst;
    a.stride_src    = src_stride;
    a.stride_dst    = dst_stride;
    a.band_stride   = band_stride;
    a.inbuf         = s->d_in;
    a.outbuf        = s->d_out;
    a.table         = s->d_table;
    a.kind          = s->d_kind;
    a.fill          = s->d_fill;
    a.n_inst        = uint32_t(s->n);
    a.handlers      = uint32_t(s->handlers);
    a.frame         = uint32_t(1) << (s->chunk_rank - 1);
    a.count         = uint32_t(count);
    uint32_t grid   = uint32_t((s->n < size_t(8 * s->sm_count)) ? s->n : size_t(8 * s->sm_count));
    CU(launch_ss(a, s->rank, grid, st));
    return B200CONV_OK;
}

extern "C" int b200conv_ss_process_device(b200conv_ss_t *s, float *dst, size_t band_stride, size_t dst_stride,
                                          const float *src, size_t src_stride, size_t count, void *stream)
{
    try { return ss_process_device_impl(s, dst, band_stride, dst_stride, src, src_stride, count, stream); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

/* Host matrices: src [instances][stride], dst [handlers][instances][stride]; synchronous.  Rows of
 * handlers that are not bound are left untouched. */
extern "C" int b200conv_ss_process_planar(b200conv_ss_t *s, float *dst, const float *src, size_t stride, size_t count)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_process_planar: NULL handle");
    if (count == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr) || (stride < count))
        return fail(B200CONV_ERR_ARG, "b200conv_ss_process_planar: bad buffers");
    ENTER_DEVICE(s);
    size_t cap = (size_t(1) << 24) / (s->n * s->handlers);
    if (cap > (size_t(1) << 20))    cap = size_t(1) << 20;
    if (cap < 1)                    cap = 1;
    for (size_t done = 0; done < count; )
    {
        const size_t c = (count - done < cap) ? count - done : cap;
        if ((s->n * c > s->stage_in) || (s->n * s->handlers * c > s->stage_out))
        {
            CU(cudaStreamSynchronize(s->stream));
            if (s->sd_in)   cudaFree(s->sd_in);
            if (s->sd_out)  cudaFree(s->sd_out);
            s->sd_in = s->sd_out = nullptr;
            s->stage_in = s->stage_out = 0;
            CU(cudaMalloc(&s->sd_in, s->n * c * sizeof(float)));
            CU(cudaMalloc(&s->sd_out, s->n * s->handlers * c * sizeof(float)));
            s->stage_in     = s->n * c;
            s->stage_out    = s->n * s->handlers * c;
        }
        CU(cudaMemcpy2DAsync(s->sd_in, c * sizeof(float), src + done, stride * sizeof(float), c * sizeof(float), s->n,
                             cudaMemcpyHostToDevice, s->stream));
        TRY(b200conv_ss_process_device(s, s->sd_out, s->n * c, c, s->sd_in, c, c, s->stream));
        for (size_t h = 0; h < s->handlers; ++h)
        {
            /* only the rows of bound handlers come back */
            for (size_t i = 0; i < s->n; ++i)
                if (s->kind[i * s->handlers + h] != 0)
                    CU(cudaMemcpyAsync(dst + (h * s->n + i) * stride + done, s->sd_out + (h * s->n + i) * c, c * sizeof(float),
                                       cudaMemcpyDeviceToHost, s->stream));
        }
        CU(cudaStreamSynchronize(s->stream));
        done += c;
    }
    return B200CONV_OK;
}

extern "C" int b200conv_ss_sync(b200conv_ss_t *s)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_ss_sync: NULL handle");
    ENTER_DEVICE(s);
    CU(cudaStreamSynchronize(s->stream));
    return B200CONV_OK;
}

extern "C" void *b200conv_ss_stream(b200conv_ss_t *s)              { return (s != nullptr) ? (void *)s->stream : nullptr; }
extern "C" size_t b200conv_ss_rank(const b200conv_ss_t *s)          { return (s != nullptr) ? s->rank : 0; }
extern "C" size_t b200conv_ss_chunk_rank(const b200conv_ss_t *s)    { return (s != nullptr) ? ss_chunk_rank(s) : 0; }
extern "C" size_t b200conv_ss_latency(const b200conv_ss_t *s)       { return (s != nullptr) ? (size_t(1) << ss_chunk_rank(s)) : 0; }
extern "C" size_t b200conv_ss_instances(const b200conv_ss_t *s)     { return (s != nullptr) ? s->n : 0; }
extern "C" size_t b200conv_ss_handlers(const b200conv_ss_t *s)      { return (s != nullptr) ? s->handlers : 0; }

#endif /* B200CONV_SPLITTER_CUH_ */
