/*
 * ab_shuffle.cu -- A/B for the transform's last five radix-2 stages (span 32): warp-shuffle
 * butterflies (BASELINE.json north_star: "radix-2/4 ... with warp-shuffle butterflies") against
 * the shared-memory exchange the product's fft_smem uses (radix-4 passes through shared memory).
 *
 * Both kernels take the same input -- `groups` independent 32-point complex sequences resident in
 * shared memory, as they are between two passes of fft_smem -- apply a decimation-in-frequency
 * 32-point transform to each, and leave the result in shared memory:
 *   k_shfl : one element per lane in registers, 5 stages of __shfl_xor_sync butterflies
 *            (2 shuffles per complex value per stage) + twiddle multiplies;
 *   k_smem : the product's way: radix-4 butterflies, operands fetched with LDS.64 and scattered
 *            back with STS.64 (two radix-4 passes + one radix-2 pass cover span 32), one
 *            __syncwarp per pass.
 * A CTA holds 2048 complex points (16 KiB, the work buffer of a rank-11/12 frame) and repeats the
 * stage set `reps` times so that the launch is long enough to time; the figure of merit is
 * complex points per second through the five stages.
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ab/ab_shuffle tools/ab_shuffle.cu
 *   tools/ab/ab_shuffle            (prints both rates and the checksum difference)
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>

#define POINTS  2048
#define THREADS 256

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__constant__ float2 c_tw[32];       /* exp(-2 pi i j / 32), j < 16 used */

/* element e of a group lives in lane e; DIF: stage s pairs lanes at distance 16 >> s */
__global__ void k_shfl(float2 *data, int reps)
{
    __shared__ float2 sm[POINTS];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < POINTS; i += THREADS)
        sm[i] = data[blockIdx.x * POINTS + i];
    __syncthreads();
    for (int r = 0; r < reps; ++r)
    {
        #pragma unroll
        for (int g = 0; g < POINTS / THREADS; ++g)          /* 8 groups per warp */
        {
            const int base = (tid >> 5) * (POINTS / (THREADS / 32)) + g * 32;
            float2 v = sm[base + lane];
            #pragma unroll
            for (int s = 0; s < 5; ++s)
            {
                const int h = 16 >> s;
                float2 o;
                o.x = __shfl_xor_sync(0xffffffffu, v.x, h);
                o.y = __shfl_xor_sync(0xffffffffu, v.y, h);
                const bool upper = (lane & h) != 0;
                float2 sum = make_float2(upper ? o.x + v.x : v.x + o.x, upper ? o.y + v.y : v.y + o.y);
                float2 dif = make_float2(upper ? o.x - v.x : v.x - o.x, upper ? o.y - v.y : v.y - o.y);
                /* lower lane keeps the sum, upper lane keeps (lower - upper) * w^{j}, j = lane % h scaled */
                const float2 w = c_tw[(lane & (h - 1)) << s];
                v = upper ? cmul(dif, w) : sum;
            }
            sm[base + lane] = v;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = tid; i < POINTS; i += THREADS)
        data[blockIdx.x * POINTS + i] = sm[i];
}

/* the same 32-point DIF through shared memory: radix-4 (h = 16,8), radix-4 (h = 4,2), radix-2 (h = 1) */
__global__ void k_smem(float2 *data, int reps)
{
    __shared__ float2 sm[POINTS];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < POINTS; i += THREADS)
        sm[i] = data[blockIdx.x * POINTS + i];
    __syncthreads();
    float2 *wsm = sm + (tid >> 5) * (POINTS / (THREADS / 32));      /* this warp's 8 groups = 256 points */
    for (int r = 0; r < reps; ++r)
    {
        /* pass A: radix-4 over distances 16 and 8: butterfly b of group g takes elements j, j+8, j+16, j+24 */
        #pragma unroll
        for (int it = 0; it < 2; ++it)
        {
            const int b = lane + 32 * it;               /* 64 butterflies = 8 groups x 8 */
            float2 *p = wsm + (b >> 3) * 32 + (b & 7);
            float2 x0 = p[0], x1 = p[8], x2 = p[16], x3 = p[24];
            /* stage h = 16 */
            float2 a0 = make_float2(x0.x + x2.x, x0.y + x2.y), a2 = cmul(make_float2(x0.x - x2.x, x0.y - x2.y), c_tw[(b & 7)]);
            float2 a1 = make_float2(x1.x + x3.x, x1.y + x3.y), a3 = cmul(make_float2(x1.x - x3.x, x1.y - x3.y), c_tw[(b & 7) + 8]);
            /* stage h = 8 */
            const float2 w8 = c_tw[((b & 7)) << 1];
            p[0]  = make_float2(a0.x + a1.x, a0.y + a1.y);
            p[8]  = cmul(make_float2(a0.x - a1.x, a0.y - a1.y), w8);
            p[16] = make_float2(a2.x + a3.x, a2.y + a3.y);
            p[24] = cmul(make_float2(a2.x - a3.x, a2.y - a3.y), w8);
        }
        __syncwarp();
        /* pass B: radix-4 over distances 4 and 2 inside each octet */
        #pragma unroll
        for (int it = 0; it < 2; ++it)
        {
            const int b = lane + 32 * it;               /* 64 butterflies: (group, octet, j<2) */
            float2 *p = wsm + (b >> 1) * 8 + (b & 1);   /* 32 octets x ... : (b>>1) indexes octets 0..31 */
            float2 x0 = p[0], x1 = p[2], x2 = p[4], x3 = p[6];
            float2 a0 = make_float2(x0.x + x2.x, x0.y + x2.y), a2 = cmul(make_float2(x0.x - x2.x, x0.y - x2.y), c_tw[(b & 1) << 2]);
            float2 a1 = make_float2(x1.x + x3.x, x1.y + x3.y), a3 = cmul(make_float2(x1.x - x3.x, x1.y - x3.y), c_tw[((b & 1) + 2) << 2]);
            const float2 w2 = c_tw[(b & 1) << 3];
            p[0] = make_float2(a0.x + a1.x, a0.y + a1.y);
            p[2] = cmul(make_float2(a0.x - a1.x, a0.y - a1.y), w2);
            p[4] = make_float2(a2.x + a3.x, a2.y + a3.y);
            p[6] = cmul(make_float2(a2.x - a3.x, a2.y - a3.y), w2);
        }
        __syncwarp();
        /* pass C: radix-2 over distance 1, two neighbours per 16-byte access */
        #pragma unroll
        for (int it = 0; it < 4; ++it)
        {
            float4 *p = reinterpret_cast<float4 *>(wsm) + lane + 32 * it;
            float4 v = *p;
            *p = make_float4(v.x + v.z, v.y + v.w, v.x - v.z, v.y - v.w);
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = tid; i < POINTS; i += THREADS)
        data[blockIdx.x * POINTS + i] = sm[i];
}

int main()
{
    float2 tw[32];
    for (int j = 0; j < 32; ++j)
        tw[j] = make_float2(float(cos(-2.0 * M_PI * j / 32.0)), float(sin(-2.0 * M_PI * j / 32.0)));
    cudaMemcpyToSymbol(c_tw, tw, sizeof(tw));
    const int ctas = 148 * 8, reps = 2000;
    std::vector<float2> h(size_t(ctas) * POINTS);
    for (size_t i = 0; i < h.size(); ++i)
        h[i] = make_float2(float((i * 2654435761u) % 1000) * 1e-3f - 0.5f, float((i * 40503u) % 1000) * 1e-3f - 0.5f);
    float2 *d[2];
    for (int k = 0; k < 2; ++k)
    {
        cudaMalloc(&d[k], h.size() * sizeof(float2));
        cudaMemcpy(d[k], h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
    }
    /* correctness: one application each on the same data */
    k_shfl<<<ctas, THREADS>>>(d[0], 1);
    k_smem<<<ctas, THREADS>>>(d[1], 1);
    std::vector<float2> r0(h.size()), r1(h.size());
    cudaMemcpy(r0.data(), d[0], h.size() * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaMemcpy(r1.data(), d[1], h.size() * sizeof(float2), cudaMemcpyDeviceToHost);
    double worst = 0.0;
    for (size_t i = 0; i < h.size(); ++i)
        worst = fmax(worst, fmax(fabs(double(r0[i].x) - r1[i].x), fabs(double(r0[i].y) - r1[i].y)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[2];
    for (int k = 0; k < 2; ++k)
    {
        cudaMemset(d[k], 0, h.size() * sizeof(float2));     /* zeros stay zeros: no overflow over many reps */
        for (int w = 0; w < 2; ++w)
            (k == 0) ? k_shfl<<<ctas, THREADS>>>(d[k], reps) : k_smem<<<ctas, THREADS>>>(d[k], reps);
        cudaEventRecord(e0);
        (k == 0) ? k_shfl<<<ctas, THREADS>>>(d[k], reps) : k_smem<<<ctas, THREADS>>>(d[k], reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[k], e0, e1);
    }
    const double pts = double(ctas) * POINTS * reps;
    printf("{\"points_per_cta\": %d, \"ctas\": %d, \"reps\": %d, \"shuffle_ms\": %.3f, \"smem_ms\": %.3f, "
           "\"shuffle_gpoints_per_s\": %.1f, \"smem_gpoints_per_s\": %.1f, \"smem_over_shuffle\": %.2f, "
           "\"max_abs_difference_one_pass\": %.3g, \"error\": \"%s\"}\n",
           POINTS, ctas, reps, ms[0], ms[1], pts / ms[0] * 1e-6, pts / ms[1] * 1e-6, ms[0] / ms[1], worst,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
