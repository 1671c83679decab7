// Stand-alone get_filtered_lidar (reference: data_process/kitti_data_utils.py:228-241) on B200:
// order-preserving stream compaction of the points inside the inclusive boundary box, with the
// reference's `z -= minZ` applied to the survivors.  (The batched BEV path fuses this filter into
// the rasteriser and never materialises the filtered sweep; this entry point exists for callers
// that want the filtered sweep itself, e.g. `get_filtered_lidar` used on its own.)
//
// Three small launches: per-CTA survivor counts -> single-CTA exclusive scan -> ordered scatter.
#include "sfa_common.cuh"

namespace sfa {
namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 4;                    // CONSECUTIVE points per thread keeps order trivial
constexpr int kPerCta = kThreads * kPerThread;

struct Box { float min_x, max_x, min_y, max_y, min_z, max_z; };

__device__ __forceinline__ bool inside(const float4& p, const Box& b) {
    // kitti_data_utils.py:237-239 — inclusive on both ends; NaN fails
    return (p.x >= b.min_x) & (p.x <= b.max_x) & (p.y >= b.min_y) & (p.y <= b.max_y) & (p.z >= b.min_z) &
           (p.z <= b.max_z);
}

__global__ void __launch_bounds__(kThreads)
filter_count_kernel(const float4* __restrict__ pts, int64_t n, Box box, uint32_t* __restrict__ block_counts) {
    __shared__ uint32_t wsum[kThreads / 32];
    int64_t i0 = (int64_t)blockIdx.x * kPerCta + (int64_t)threadIdx.x * kPerThread;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j)
        if (i0 + j < n) c += inside(pts[i0 + j], box) ? 1u : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int q = 0; q < kThreads / 32; ++q) t += wsum[q];
        block_counts[blockIdx.x] = t;
    }
}

// in-place exclusive scan of block_counts by ONE CTA; total -> *out_count
__global__ void __launch_bounds__(1024)
filter_scan_kernel(uint32_t* __restrict__ block_counts, int64_t n_blocks, int64_t* __restrict__ out_count) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        int64_t i = base + threadIdx.x;
        uint32_t v = i < n_blocks ? block_counts[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = wsum[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += t;
            }
            wsum[lane] = wi - w;  // exclusive warp offsets
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + wsum[warp] + incl - v;
        if (i < n_blocks) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = (int64_t)carry_s;
}

__global__ void __launch_bounds__(kThreads)
filter_scatter_kernel(const float4* __restrict__ pts, int64_t n, Box box, const uint32_t* __restrict__ block_offsets,
                      float4* __restrict__ out) {
    __shared__ uint32_t wsum[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t i0 = (int64_t)blockIdx.x * kPerCta + (int64_t)threadIdx.x * kPerThread;
    float4 p[kPerThread];
    bool k[kPerThread];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
        k[j] = false;
        if (i0 + j < n) {
            p[j] = pts[i0 + j];
            k[j] = inside(p[j], box);
        }
        c += k[j] ? 1u : 0u;
    }
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t pos = block_offsets[blockIdx.x] + incl - c;
    for (int q = 0; q < warp; ++q) pos += wsum[q];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
        if (k[j]) {
            p[j].z = __fsub_rn(p[j].z, box.min_z);  // kitti_data_utils.py:241
            out[pos++] = p[j];
        }
    }
}

}  // namespace
}  // namespace sfa

using namespace sfa;

extern "C" size_t sfa_filter_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    int64_t blocks = (n + kPerCta - 1) / kPerCta;
    return (size_t)(blocks > 0 ? blocks : 1) * sizeof(uint32_t);
}

extern "C" int sfa_filter_lidar(const float* pts, int64_t n, const SfaBevParams* p, float* out_pts, int64_t* out_count,
                                void* workspace, size_t workspace_bytes, sfa_stream_t stream_) {
    SFA_REQUIRE(p != nullptr && out_count != nullptr, "NULL pointer argument");
    SFA_REQUIRE(n >= 0 && n <= 0xFFFFFFFFll, "n=%lld out of range", (long long)n);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) {
        SFA_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(int64_t), stream));
        return SFA_OK;
    }
    SFA_REQUIRE(pts && out_pts && workspace, "NULL pointer argument");
    SFA_REQUIRE(((reinterpret_cast<uintptr_t>(pts) | reinterpret_cast<uintptr_t>(out_pts)) & 15) == 0,
                "pts/out_pts need 16-B alignment");
    if (workspace_bytes < sfa_filter_workspace_bytes(n)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, sfa_filter_workspace_bytes(n));
        return SFA_ERR_WORKSPACE_TOO_SMALL;
    }
    Box box = {p->min_x, p->max_x, p->min_y, p->max_y, p->min_z, p->max_z};
    int64_t blocks = (n + kPerCta - 1) / kPerCta;
    uint32_t* counts = static_cast<uint32_t*>(workspace);
    const float4* in = reinterpret_cast<const float4*>(pts);
    SFA_LAUNCH("filter_count", stream, filter_count_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(in, n, box, counts));
    SFA_LAUNCH("filter_scan", stream, filter_scan_kernel<<<1, 1024, 0, stream>>>(counts, blocks, out_count));
    SFA_LAUNCH("filter_scatter", stream, filter_scatter_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(in, n, box, counts, reinterpret_cast<float4*>(out_pts)));
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}
