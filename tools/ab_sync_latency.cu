// Developer A/B: how long does a synchronous host round trip take on this box?
//   (1) kernel + cudaEventRecord + cudaEventSynchronize        (what finish_sync_call does)
//   (2) kernel that stores a sequence number into host-mapped memory + host spin on it
//   (3) as (1) with cudaStreamSynchronize
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ab/ab_sync_latency tools/ab_sync_latency.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <immintrin.h>

__global__ void k_work(const float *src, float *dst, volatile unsigned *flag, unsigned seq, int spin_ns)
{
    // read 4 KiB from (possibly host-mapped) src, write 4 KiB to dst, then optionally ring the flag
    float v = src[threadIdx.x];
    if (spin_ns > 0)
    {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < (unsigned long long)spin_ns);
    }
    dst[threadIdx.x] = v + 1.0f;
    if (flag != nullptr)
    {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0)
            *flag = seq;
    }
}

static double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main()
{
    float *hs, *hd; unsigned *hf;
    cudaHostAlloc(&hs, 4096, cudaHostAllocMapped); cudaHostAlloc(&hd, 4096, cudaHostAllocMapped); cudaHostAlloc(&hf, 64, cudaHostAllocMapped);
    float *ds, *dd; unsigned *df;
    cudaHostGetDevicePointer(&ds, hs, 0); cudaHostGetDevicePointer(&dd, hd, 0); cudaHostGetDevicePointer(&df, hf, 0);
    for (int i = 0; i < 1024; ++i) hs[i] = float(i);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    const int N = 3000;
    for (int spin : {0, 10000})
    {
        for (int mode = 1; mode <= 3; ++mode)
        {
            std::vector<double> t(N);
            unsigned seq = 0;
            *hf = 0;
            for (int i = 0; i < N + 100; ++i)
            {
                hs[5] = float(i);
                double t0 = now_us();
                ++seq;
                if (mode == 2)
                {
                    k_work<<<1, 1024, 0, st>>>(ds, dd, df, seq, spin);
                    while (*(volatile unsigned *)hf != seq) _mm_pause();
                }
                else
                {
                    k_work<<<1, 1024, 0, st>>>(ds, dd, nullptr, seq, spin);
                    if (mode == 1) { cudaEventRecord(ev, st); cudaEventSynchronize(ev); }
                    else           cudaStreamSynchronize(st);
                }
                double t1 = now_us();
                if (hd[5] != float(i) + 1.0f) { printf("wrong result\n"); return 1; }
                if (i >= 100) t[i - 100] = t1 - t0;
            }
            std::sort(t.begin(), t.end());
            printf("{\"kernel_spin_us\": %d, \"mode\": \"%s\", \"median_us\": %.2f, \"p10_us\": %.2f, \"p99_us\": %.2f}\n", spin / 1000,
                   mode == 1 ? "event record + event synchronize" : mode == 2 ? "flag in host-mapped memory + host spin" : "stream synchronize",
                   t[N / 2], t[N / 10], t[N * 99 / 100]);
        }
    }
    return 0;
}
