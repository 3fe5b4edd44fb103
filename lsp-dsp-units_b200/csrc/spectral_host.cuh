/*
 * spectral_host.cuh -- host side and C ABI (b200conv_sp_*) of the batched SpectralProcessor
 * (scope-table row f4; included by engine.cu after spectral.cuh).
 *
 * Bookkeeping follows lsp::dspu::SpectralProcessor (reference src/main/util/SpectralProcessor.cpp):
 * init :58-75, update_settings :107-125 (buffers cleared, nOffset = size_t(N * (fPhase * 0.5f))),
 * set_phase :127-131 (clamped to [0, 1]), set_rank :133-141 (ignored when equal or above the
 * maximum), process :143-199, remaining :241-245, reset :247-257 (clears the buffers, keeps nOffset,
 * does nothing while an update is pending), latency() = 2^rank (SpectralProcessor.h).
 */
#ifndef B200CONV_SPECTRAL_HOST_CUH_
#define B200CONV_SPECTRAL_HOST_CUH_

static const size_t SP_RANK_MIN = 7, SP_RANK_MAX = 15;      /* transform rank + 1 must be one of the engine's 8..16 */

struct b200conv_sp
{
    int                     device      = 0;
    int                     sm_count    = 148;
    size_t                  n           = 0;
    size_t                  max_rank    = 0, rank = 0;
    cudaStream_t            stream      = nullptr;
    float2                 *tw[B200CONV_RANK_MAX + 1] = { nullptr };    /* by transform rank (= rank + 1) */
    float                  *d_wnd       = nullptr;      /* [2^max_rank] window of the CURRENT rank            */
    float                  *d_in = nullptr, *d_out = nullptr;           /* [n][2^max_rank] each               */
    float2                 *d_table     = nullptr;      /* [n][2^(max_rank-1) + 1] folded tables              */
    uint8_t                *d_bound     = nullptr;
    uint32_t               *d_off       = nullptr;
    std::vector<uint32_t>   h_off;
    std::vector<float>      phase;
    std::vector<uint8_t>    dirty, bound;               /* bUpdate per instance; a table is bound             */
    bool                    wnd_dirty   = true, bound_dirty = true;
    float                  *s_in = nullptr, *s_out = nullptr, *sd_in = nullptr, *sd_out = nullptr;   /* host staging */
    size_t                  stage_floats = 0;
};

typedef b200conv_sp Sp;

template <int RANKP>
static cudaError_t launch_sp_r(const SpArgs &a, uint32_t grid, cudaStream_t st)
{
    using S = SpCfg<RANKP>;
    static bool attr_set[MAX_DEVICES] = { false };
    int dev = current_device();
    if ((!attr_set[dev]) && (S::SMEM > 48 * 1024))
    {
        cudaError_t e = cudaFuncSetAttribute(k_sp<RANKP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::SMEM));
        if (e != cudaSuccess)
            return e;
    }
    attr_set[dev] = true;
    k_sp<RANKP><<<grid, S::C::T, S::SMEM, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t launch_sp(const SpArgs &a, size_t rank, uint32_t grid, cudaStream_t st)
{
    switch (rank + 1)
    {
        case 8:  return launch_sp_r<8>(a, grid, st);
        case 9:  return launch_sp_r<9>(a, grid, st);
        case 10: return launch_sp_r<10>(a, grid, st);
        case 11: return launch_sp_r<11>(a, grid, st);
        case 12: return launch_sp_r<12>(a, grid, st);
        case 13: return launch_sp_r<13>(a, grid, st);
        case 14: return launch_sp_r<14>(a, grid, st);
        case 15: return launch_sp_r<15>(a, grid, st);
        case 16: return launch_sp_r<16>(a, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

extern "C" void b200conv_sp_free(b200conv_sp_t *s)
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
    if (s->d_bound) cudaFree(s->d_bound);
    if (s->d_off)   cudaFree(s->d_off);
    if (s->s_in)    cudaFreeHost(s->s_in);
    if (s->s_out)   cudaFreeHost(s->s_out);
    if (s->sd_in)   cudaFree(s->sd_in);
    if (s->sd_out)  cudaFree(s->sd_out);
    if (s->stream)  cudaStreamDestroy(s->stream);
    delete s;
}

static int sp_create_impl(b200conv_sp_t **out, int device, size_t instances, size_t max_rank)
{
    if ((out == nullptr) || (instances == 0) || (instances > (size_t(1) << 20)) ||
        (max_rank < SP_RANK_MIN) || (max_rank > SP_RANK_MAX))
        return fail(B200CONV_ERR_ARG, "b200conv_sp_create: bad arguments (ranks %zu..%zu)", SP_RANK_MIN, SP_RANK_MAX);
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if ((e != cudaSuccess) || (count == 0))
        return fail(B200CONV_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0)
        CU(cudaGetDevice(&device));
    if (device >= count)
        return fail(B200CONV_ERR_ARG, "device %d out of range (%d devices)", device, count);

    Sp *s = new (std::nothrow) Sp();
    if (s == nullptr)
        return fail(B200CONV_ERR_NOMEM, "out of host memory");
    s->device   = device;
    s->n        = instances;
    s->max_rank = s->rank = max_rank;               /* SpectralProcessor.cpp:60-61 */
    s->h_off.assign(instances, 0);
    s->phase.assign(instances, 0.0f);
    s->dirty.assign(instances, 1);                  /* bUpdate = true (:63) */
    s->bound.assign(instances, 0);

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
        CU_BRK(cudaMalloc(&s->d_out, instances * N * sizeof(float)));
        CU_BRK(cudaMalloc(&s->d_table, instances * (N / 2 + 1) * sizeof(float2)));
        CU_BRK(cudaMalloc(&s->d_bound, instances));
        CU_BRK(cudaMalloc(&s->d_off, instances * sizeof(uint32_t)));
        CU_BRK(cudaMemset(s->d_bound, 0, instances));
        #undef CU_BRK
    } while (false);
    if (rc != B200CONV_OK)
    {
        std::string keep = g_last_error;
        b200conv_sp_free(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return B200CONV_OK;
}

extern "C" int b200conv_sp_create(b200conv_sp_t **out, int device, size_t instances, size_t max_rank)
{
    try { return sp_create_impl(out, device, instances, max_rank); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

extern "C" int b200conv_sp_set_rank(b200conv_sp_t *s, size_t rank)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_sp_set_rank: NULL handle");
    if ((rank == s->rank) || (rank > s->max_rank))          /* SpectralProcessor.cpp:135-136 */
        return B200CONV_OK;
    if (rank < SP_RANK_MIN)
        return fail(B200CONV_ERR_ARG, "b200conv_sp_set_rank: ranks below %zu are not supported", SP_RANK_MIN);
    s->rank         = rank;
    s->wnd_dirty    = true;
    for (size_t i = 0; i < s->n; ++i)
    {
        s->dirty[i]     = 1;                                /* bUpdate = true (:139) */
        s->bound[i]     = 0;                                /* tables are rank-specific: bind again */
    }
    s->bound_dirty  = true;
    return B200CONV_OK;
}

extern "C" int b200conv_sp_set_phase(b200conv_sp_t *s, size_t idx, float phase)
{
    if ((s == nullptr) || (idx >= s->n))
        return fail(B200CONV_ERR_ARG, "b200conv_sp_set_phase: bad handle or index");
    s->phase[idx]   = (phase < 0.0f) ? 0.0f : ((phase > 1.0f) ? 1.0f : phase);      /* lsp_limit, :129 */
    s->dirty[idx]   = 1;
    return B200CONV_OK;
}

/* kind 0: unbind; 1: `table` = 2^rank packed complex bins; 2: `table` = 2^rank real gains */
static int sp_bind(b200conv_sp_t *s, size_t idx, int kind, const float *table)
{
    if ((s == nullptr) || (idx >= s->n) || ((kind != 0) && (table == nullptr)))
        return fail(B200CONV_ERR_ARG, "b200conv_sp_bind: bad arguments");
    ENTER_DEVICE(s);
    if (kind == 0)
    {
        s->bound[idx]   = 0;
        s->bound_dirty  = true;
        return B200CONV_OK;
    }
    /* Re(IFFT(X H)) for real input = the half-spectrum product with (H[k] + conj(H[N - k])) / 2 */
    const size_t N = size_t(1) << s->rank, P = N / 2;
    std::vector<float2> folded(P + 1);
    for (size_t k = 0; k <= P; ++k)
    {
        const size_t km = (N - k) % N;
        if (kind == 1)
            folded[k]   = make_float2(0.5f * (table[2 * k] + table[2 * km]), 0.5f * (table[2 * k + 1] - table[2 * km + 1]));
        else
            folded[k]   = make_float2(0.5f * (table[k] + table[km]), 0.0f);
    }
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(s->d_table + idx * (P + 1), folded.data(), (P + 1) * sizeof(float2), cudaMemcpyHostToDevice));
    s->bound[idx]   = 1;
    s->bound_dirty  = true;
    return B200CONV_OK;
}

extern "C" int b200conv_sp_bind_complex(b200conv_sp_t *s, size_t idx, const float *table)
{
    try { return sp_bind(s, idx, 1, table); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

extern "C" int b200conv_sp_bind_gain(b200conv_sp_t *s, size_t idx, const float *gain)
{
    try { return sp_bind(s, idx, 2, gain); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

extern "C" int b200conv_sp_unbind(b200conv_sp_t *s, size_t idx)
{
    return sp_bind(s, idx, 0, nullptr);
}

/* SpectralProcessor::update_settings for the instances that asked for it (:107-125), then the
 * tables the kernel reads. */
static int sp_commit(Sp *s, cudaStream_t st)
{
    const size_t N = size_t(1) << s->rank;
    bool any = false;
    for (size_t i = 0; i < s->n; ++i)
    {
        if (!s->dirty[i])
            continue;
        any             = true;
        /* the buffers of the instance are laid out with the CURRENT rank's pitch: a rank change marks
         * every instance, so the whole arrays are cleared below */
        s->h_off[i]     = uint32_t(size_t(float(N) * (s->phase[i] * 0.5f)));        /* :123, fp32 */
    }
    if (any)
    {
        bool all = true;
        for (size_t i = 0; i < s->n; ++i)
            all             = all && (s->dirty[i] != 0);
        if (all)
        {
            CU(cudaMemsetAsync(s->d_in, 0, s->n * N * sizeof(float), st));
            CU(cudaMemsetAsync(s->d_out, 0, s->n * N * sizeof(float), st));
        }
        else
            for (size_t i = 0; i < s->n; ++i)
                if (s->dirty[i])
                {
                    CU(cudaMemsetAsync(s->d_in + i * N, 0, N * sizeof(float), st));
                    CU(cudaMemsetAsync(s->d_out + i * N, 0, N * sizeof(float), st));
                }
        CU(cudaMemcpyAsync(s->d_off, s->h_off.data(), s->n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        for (size_t i = 0; i < s->n; ++i)
            s->dirty[i]     = 0;
    }
    if (s->wnd_dirty)
    {
        /* windows::cosine (reference src/main/misc/windows.cpp:238-246): f = M_PI / n in fp32, sinf(f * i) */
        std::vector<float> w(N);
        const float f   = float(M_PI / double(N));
        for (size_t i = 0; i < N; ++i)
            w[i]            = sinf(f * float(i));
        CU(cudaMemcpyAsync(s->d_wnd, w.data(), N * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));          /* `w` is pageable and about to go out of scope */
        s->wnd_dirty    = false;
    }
    if (s->bound_dirty)
    {
        CU(cudaMemcpyAsync(s->d_bound, s->bound.data(), s->n, cudaMemcpyHostToDevice, st));
        s->bound_dirty  = false;
    }
    if (s->tw[s->rank + 1] == nullptr)
        TRY(make_twiddles(uint32_t(s->rank + 1), &s->tw[s->rank + 1]));
    return B200CONV_OK;
}

static int sp_process_device_impl(b200conv_sp_t *s, float *dst, size_t dst_stride, const float *src, size_t src_stride,
                                  size_t count, void *stream)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_sp_process_device: NULL handle");
    if (count == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr) || (count >= (size_t(1) << 32)) ||
        ((s->n > 1) && ((src_stride < count) || (dst_stride < count))))
        return fail(B200CONV_ERR_ARG, "b200conv_sp_process_device: bad buffers");
    ENTER_DEVICE(s);
    cudaStream_t st = (stream != nullptr) ? cudaStream_t(stream) : s->stream;
    TRY(sp_commit(s, st));

    SpArgs a;
    memset(&a, 0, sizeof(a));
    a.tw            = s->tw[s->rank + 1];
    a.wnd           = s->d_wnd;
    a.src           = src;
    a.dst           = dst;
    a.stride_src    = src_stride;
    a.stride_dst    = dst_stride;
    a.inbuf         = s->d_in;
    a.outbuf        = s->d_out;
    a.table         = s->d_table;
    a.bound         = s->d_bound;
    a.off           = s->d_off;
    a.n_inst        = uint32_t(s->n);
    a.count         = uint32_t(count);
    /* one CTA per instance, as many resident CTAs as the device holds */
    uint32_t grid   = uint32_t((s->n < size_t(8 * s->sm_count)) ? s->n : size_t(8 * s->sm_count));
    CU(launch_sp(a, s->rank, grid, st));

    /* the host mirror of nOffset (remaining()), same arithmetic as the kernel's loop (:154-198) */
    const uint32_t F = uint32_t(1) << (s->rank - 1);
    for (size_t i = 0; i < s->n; ++i)
    {
        uint64_t total  = uint64_t(s->h_off[i]) + count;
        /* every time the offset reaches F with samples left it restarts at 0 */
        if (s->h_off[i] >= F)
            total           = count;            /* a transform precedes the first sample */
        uint32_t off    = (total <= F) ? uint32_t(total) : uint32_t((total - 1) % F + 1);
        s->h_off[i]     = off;
    }
    return B200CONV_OK;
}

extern "C" int b200conv_sp_process_device(b200conv_sp_t *s, float *dst, size_t dst_stride, const float *src,
                                          size_t src_stride, size_t count, void *stream)
{
    try { return sp_process_device_impl(s, dst, dst_stride, src, src_stride, count, stream); }
    catch (const std::bad_alloc &) { return fail(B200CONV_ERR_NOMEM, "out of host memory"); }
}

/* Host matrix [instances][stride] floats; synchronous (the reference's process() is). */
extern "C" int b200conv_sp_process_planar(b200conv_sp_t *s, float *dst, const float *src, size_t stride, size_t count)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_sp_process_planar: NULL handle");
    if (count == 0)
        return B200CONV_OK;
    if ((dst == nullptr) || (src == nullptr) || (stride < count))
        return fail(B200CONV_ERR_ARG, "b200conv_sp_process_planar: bad buffers");
    ENTER_DEVICE(s);
    size_t cap = (size_t(1) << 24) / s->n;
    if (cap > (size_t(1) << 20))    cap = size_t(1) << 20;
    if (cap < 1)                    cap = 1;
    for (size_t done = 0; done < count; )
    {
        const size_t c = (count - done < cap) ? count - done : cap;
        if (s->n * c > s->stage_floats)
        {
            CU(cudaStreamSynchronize(s->stream));
            if (s->sd_in)   cudaFree(s->sd_in);
            if (s->sd_out)  cudaFree(s->sd_out);
            s->sd_in = s->sd_out = nullptr;
            s->stage_floats = 0;
            CU(cudaMalloc(&s->sd_in, s->n * c * sizeof(float)));
            CU(cudaMalloc(&s->sd_out, s->n * c * sizeof(float)));
            s->stage_floats = s->n * c;
        }
        CU(cudaMemcpy2DAsync(s->sd_in, c * sizeof(float), src + done, stride * sizeof(float), c * sizeof(float), s->n,
                             cudaMemcpyHostToDevice, s->stream));
        TRY(b200conv_sp_process_device(s, s->sd_out, c, s->sd_in, c, c, s->stream));
        CU(cudaMemcpy2DAsync(dst + done, stride * sizeof(float), s->sd_out, c * sizeof(float), c * sizeof(float), s->n,
                             cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        done += c;
    }
    return B200CONV_OK;
}

extern "C" int b200conv_sp_reset(b200conv_sp_t *s)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_sp_reset: NULL handle");
    ENTER_DEVICE(s);
    const size_t N = size_t(1) << s->rank;
    for (size_t i = 0; i < s->n; ++i)
    {
        if (s->dirty[i])                            /* update_settings() will clear the buffers (:249-250) */
            continue;
        CU(cudaMemsetAsync(s->d_in + i * N, 0, N * sizeof(float), s->stream));      /* :256: pOutBuf, 2 N floats */
        CU(cudaMemsetAsync(s->d_out + i * N, 0, N * sizeof(float), s->stream));
    }
    return B200CONV_OK;
}

extern "C" int b200conv_sp_sync(b200conv_sp_t *s)
{
    if (s == nullptr)
        return fail(B200CONV_ERR_ARG, "b200conv_sp_sync: NULL handle");
    ENTER_DEVICE(s);
    CU(cudaStreamSynchronize(s->stream));
    return B200CONV_OK;
}

extern "C" void *b200conv_sp_stream(b200conv_sp_t *s)              { return (s != nullptr) ? (void *)s->stream : nullptr; }
extern "C" size_t b200conv_sp_rank(const b200conv_sp_t *s)          { return (s != nullptr) ? s->rank : 0; }
extern "C" size_t b200conv_sp_latency(const b200conv_sp_t *s)       { return (s != nullptr) ? (size_t(1) << s->rank) : 0; }
extern "C" size_t b200conv_sp_instances(const b200conv_sp_t *s)     { return (s != nullptr) ? s->n : 0; }

/* SpectralProcessor::remaining (:241-245) */
extern "C" size_t b200conv_sp_remaining(const b200conv_sp_t *s, size_t idx)
{
    if ((s == nullptr) || (idx >= s->n))
        return 0;
    const size_t F = size_t(1) << (s->rank - 1);
    return (s->h_off[idx] <= F) ? F - s->h_off[idx] : 0;
}

#endif /* B200CONV_SPECTRAL_HOST_CUH_ */
