// How fast does an SM retire SMALL cp.async.bulk shared->global copies (the band runs of a bin tile: ~128 copies of
// ~256 B per 2048-point tile)?  Persistent CTAs; per iteration `nthr` threads issue one bulk store of `bytes` each from a
// 32 KB stage to distinct destinations, then wait (read) before the next iteration.  Compared with the same bytes moved by
// 16-B st.global of all threads.   Build: nvcc -arch=sm_100a -O3 -cudart=shared -o tools/_build/tma_small_probe tools/tma_small_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: one bulk copy per band by threads 0..127; 1: plain 16-B stores by all threads
__global__ void __launch_bounds__(512, 3) probe(uint4* __restrict__ dst, int iters, int bytes_per_run, size_t run_stride16) {
    __shared__ __align__(128) uint4 stage[2048];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += 512) stage[i] = make_uint4(i, tid, blockIdx.x, 7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int recs = bytes_per_run / 16;
    for (int it = 0; it < iters; ++it) {
        uint4* base = dst + ((size_t)blockIdx.x * iters + it) % 4096 * 2048;   // a moving 128 MB window
        if (MODE == 0) {
            if (tid < 128) {
                uint4* g = base + (size_t)tid * run_stride16;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(stage + tid * recs)), "r"(bytes_per_run) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
            for (int s = tid; s < 128 * recs; s += 512) {
                const int b = s / recs;
                base[(size_t)b * run_stride16 + (s - b * recs)] = stage[s];
            }
        }
        __syncthreads();
    }
    if (MODE == 0 && tid < 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    uint4* dst;
    cudaMalloc(&dst, (size_t)4096 * 2048 * 16 + (1 << 20));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 200, ctas = 148 * 3;
    for (int bytes : {64, 128, 256, 512}) {
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) probe<0><<<ctas, 512>>>(dst, iters, bytes, 16);
                else probe<1><<<ctas, 512>>>(dst, iters, bytes, 16);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double runs = (double)ctas * iters * 128;
            printf("%s  %4d B per run: %8.3f ms  %7.1f M runs/s  %7.1f GB/s  %6.1f cycles per run per SM (at 1.965 GHz)\n",
                   mode == 0 ? "bulk copy per run" : "16-B stores      ", bytes, best, runs / best / 1e3, runs * bytes / best / 1e6,
                   best * 1e-3 * 1.965e9 / (runs / 148));
        }
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
