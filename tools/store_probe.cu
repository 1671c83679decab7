// Micro-benchmark: how fast can persistent CTAs drain `item_bytes` blocks from shared memory to HBM?
// Variants: plain / streaming 16-B stores from all threads, or TMA bulk stores issued by one thread
// (waiting for the read-out before the next item, like bev_band), at several CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/store_probe tools/store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// PLANES = 1: the three pieces of item i go where bev_band writes them (frame = i / 128, band = i % 128; plane k of the
// frame at ((frame * 3 + k) * 128 + band) * piece_bytes) instead of back to back.
__device__ int g_planes = 0;
__device__ __forceinline__ unsigned char* piece_ptr(float* out, int item, int k, int item_bytes, int pieces) {
    const size_t pb = (size_t)item_bytes / pieces;
    if (!g_planes) return reinterpret_cast<unsigned char*>(out) + (size_t)item * item_bytes + k * pb;
    const size_t frame = item / 128, band = item % 128;
    return reinterpret_cast<unsigned char*>(out) + ((frame * pieces + k) * 128 + band) * pb;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int n_items, int item_bytes, int pieces) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < item_bytes / 16; i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, 1, 2, 3);
    __syncthreads();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pb16 = item_bytes / pieces / 16;
        if (MODE == 0 || MODE == 1) {
            for (int i = tid; i < item_bytes / 16; i += 256) {
                uint4 v = reinterpret_cast<uint4*>(sm)[i];
                unsigned char* dst = piece_ptr(out, item, i / pb16, item_bytes, pieces) - (size_t)(i / pb16) * pb16 * 16;
                if (MODE == 0) reinterpret_cast<uint4*>(dst)[i] = v;
                else asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 16 * (size_t)i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            __syncthreads();
        } else {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                const int pb = item_bytes / pieces;
                for (int k = 0; k < pieces; ++k)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(piece_ptr(out, item, k, item_bytes, pieces)), "r"(smem_u32(sm + (size_t)k * pb)), "r"(pb) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (MODE == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // MODE 3: one item in flight, no smem reuse hazard in this probe
            }
            __syncthreads();
        }
    }
    if (MODE >= 2 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
float run(float* out, int n_items, int item_bytes, int ctas_per_sm, int pieces) {
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, item_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<148 * ctas_per_sm, 256, item_bytes>>>(out, n_items, item_bytes, pieces);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t total = 1ull << 30;
    float* out; cudaMalloc(&out, total);
    const char* names[] = {"st.v4", "st.cs.v4", "tma wait_read0", "tma 1 in flight"};
    for (int item_bytes : {35328, 61440}) {
        const int n_items = (int)(total / item_bytes);
        for (int cps : {1, 2, 3, 4, 6}) {
            if ((size_t)item_bytes * cps > 220 * 1024) continue;
            float t0 = run<0>(out, n_items, item_bytes, cps, 3);
            float t1 = run<1>(out, n_items, item_bytes, cps, 3);
            float t2 = run<2>(out, n_items, item_bytes, cps, 3);
            float t3 = run<3>(out, n_items, item_bytes, cps, 3);
            float t4 = run<2>(out, n_items, item_bytes, cps, 1);
            printf("item %6d B  %d CTAs/SM : %s %.0f GB/s | %s %.0f | %s %.0f | %s %.0f | tma 1 piece %.0f\n", item_bytes, cps,
                   names[0], total / t0 * 1e-6, names[1], total / t1 * 1e-6, names[2], total / t2 * 1e-6, names[3], total / t3 * 1e-6, total / t4 * 1e-6);
        }
    }
    // bev_band's layout: 3 planes of 128 bands x 11552 B per frame, 4 CTAs per SM
    {
        const int one = 1;
        cudaMemcpyToSymbol(g_planes, &one, sizeof(int));
        const int item_bytes = 3 * 11552;
        const int n_items = (int)(total / item_bytes) / 128 * 128;
        const double bytes = (double)n_items * item_bytes;
        for (int cps : {1, 2, 4}) {
            float t0 = run<0>(out, n_items, item_bytes, cps, 3), t1 = run<1>(out, n_items, item_bytes, cps, 3);
            float t2 = run<2>(out, n_items, item_bytes, cps, 3);
            printf("plane layout (3 x 11552 B per item, planes 1.48 MB apart)  %d CTAs/SM : st.v4 %.0f GB/s | st.cs.v4 %.0f | tma wait_read0 %.0f\n",
                   cps, bytes / t0 * 1e-6, bytes / t1 * 1e-6, bytes / t2 * 1e-6);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(err));
    return 0;
}
