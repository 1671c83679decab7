// Stage B of the SFA3D hot path on B200: heat-map peak decode.
//
// Replaces (reference, read-only at /root/reference), utils/evaluation_utils.py:
//   _nms :21-26     3x3 max-pool (stride 1, -inf padding), keep where pooled == value, heat * keep
//   _topk :47-62    per-class top-K then top-K of the C*K candidates (== global top-K of C*h*w)
//   _transpose_and_gather_feat :40-44 x4   NHWC copies of every head just to fetch K rows
//   decode :77-105  assembly of [B, K, 10]
//   post_processing :112-163 ("evaluation_utils copy.py":112-143 per-sample semantics), dense form
//
// One kernel, one thread-block CLUSTER of kSlabs CTAs per frame:
//   - each CTA owns a slab of rows of all C classes: it stages slab + halo in shared memory with
//     coalesced loads, applies the 3x3 peak-keep there, and turns every value into an orderable
//     32-bit key;
//   - exact top-K of the slab by an 8-bit MSB-first radix select over shared memory (4 passes,
//     warp-aggregated histogram updates), ties at the K-th value resolved toward the lower index;
//   - the K survivors of every CTA go to the cluster leader's shared memory through DSMEM as
//     64-bit (key << 32 | ~linear_index) words; the leader bitonic-sorts the kSlabs*K candidates
//     and gathers the 8 regression values per detection straight from the NCHW heads (one 32-B
//     sector each — no NHWC transpose).
// The heat map is read from HBM exactly once; nothing intermediate is written to HBM.
//
// Tie rule (torch.topk leaves it implementation-defined): equal scores are ordered by lower class,
// then lower y*w+x — i.e. descending 64-bit composite key, which is a strict total order.
#include "sfa_common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace sfa {
namespace {

constexpr int kSlabs = 8;          // CTAs per cluster (portable maximum)
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 128;
constexpr uint32_t kNanKey = 0xFFFFFFFFu;  // torch.topk ranks NaN above everything

struct DecodeArgs {
    const float* hm;      // [B,C,h,w]
    const float* off;     // [B,2,h,w] or null
    const float* dir;     // [B,2,h,w]
    const float* zc;      // [B,1,h,w]
    const float* dim;     // [B,3,h,w]
    int B, C, h, w, K;
    int rows_per_slab;
    int do_nms;
    float* det;           // [B,K,10] or null
    int64_t* inds;        // [B,K] or null
    // _topk outputs (all null for decode)
    float* tk_score; int32_t* tk_cls; float* tk_ys; float* tk_xs;
};

__device__ __forceinline__ float nanmax(float a, float b) {
    // max_pool2d propagates NaN (ATen: `val > max || isnan(val)`)
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}

// EPT = elements per thread held in registers between the NMS read phase and the in-place key write.
template <int EPT>
__global__ void __launch_bounds__(kThreads)
decode_kernel(DecodeArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long cand[kSlabs * kMaxK];  // only the leader's copy is used
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sel_prefix, sel_need, n_cand, eq_running;
    __shared__ unsigned int warp_sums[kWarps];

    const int slab = cluster.block_rank();
    const int b = blockIdx.y;
    const int C = a.C, h = a.h, w = a.w, K = a.K;
    const int hw = h * w;
    const int r0 = slab * a.rows_per_slab;
    const int rows = max(0, min(a.rows_per_slab, h - r0));
    const int n = C * rows * w;                  // elements this CTA selects from
    const int trows = a.rows_per_slab + 2;       // tile rows incl. halo
    float* tile = reinterpret_cast<float*>(smem_raw);          // [C][trows][w] raw heat
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);    // [n] after the in-place rewrite
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // every CTA of the cluster must be running before anyone writes into the leader's shared memory
    cluster.barrier_arrive();

    // ---- stage slab + halo (rows r0-1 .. r0+rows) -------------------------------------------------
    const float* hmb = a.hm + (size_t)b * C * hw;
    const float ninf = __int_as_float(0xFF800000);
    if (rows > 0) {
        const int tile_n = C * trows * w;
        for (int e = tid; e < tile_n; e += kThreads) {
            int c = e / (trows * w);
            int rem = e - c * trows * w;
            int tr = rem / w;
            int x = rem - tr * w;
            int y = r0 - 1 + tr;
            float v = ninf;
            if (y >= 0 && y < h && tr < rows + 2) v = __ldg(hmb + (size_t)c * hw + (size_t)y * w + x);
            tile[e] = v;
        }
    }
    __syncthreads();

    // ---- 3x3 peak keep -> orderable keys (held in registers until the whole tile has been read) ---
    uint32_t kreg[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        int e = tid + j * kThreads;
        kreg[j] = 0;
        if (e < n) {
            int c = e / (rows * w);
            int rem = e - c * rows * w;
            int r = rem / w;
            int x = rem - r * w;
            const float* t = tile + (c * trows + r + 1) * w + x;
            float v = t[0];
            float val = v;
            if (a.do_nms) {
                float m = v;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const float* tr_ = t + dy * w;
                    if (x > 0) m = nanmax(m, tr_[-1]);
                    m = nanmax(m, tr_[0]);
                    if (x < w - 1) m = nanmax(m, tr_[1]);
                }
                float keep = (m == v) ? 1.0f : 0.0f;
                val = __fmul_rn(v, keep);  // heat * keep, evaluation_utils.py:26
            }
            kreg[j] = orderable_u32(val, kNanKey);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        int e = tid + j * kThreads;
        if (e < n) keys[e] = kreg[j];
    }
    if (tid == 0) { sel_prefix = 0; sel_need = (unsigned)min(K, n); n_cand = 0; eq_running = 0; }
    __syncthreads();

    // ---- radix select: key of the K-th largest element of this slab --------------------------------
    const int kk = min(K, n);
    unsigned int eq_total = 0;
    if (kk > 0) {
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
            for (int i = tid; i < 256; i += kThreads) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = sel_prefix;
            for (int base = 0; base < n; base += kThreads) {
                int e = base + tid;
                uint32_t bin = 256;  // sentinel: not counted
                if (e < n) {
                    uint32_t k = keys[e];
                    if (((k ^ prefix) & himask) == 0) bin = (k >> shift) & 255u;
                }
                unsigned m = __match_any_sync(0xFFFFFFFFu, bin);
                if (bin < 256 && lane == (__ffs(m) - 1)) atomicAdd(&hist[bin], (unsigned)__popc(m));
            }
            __syncthreads();
            if (warp == 0) {
                // lane l owns bins 255-8l .. 248-8l (descending); find the bin holding the need-th element
                const unsigned need = sel_need;
                unsigned hloc[8];
                unsigned s = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { hloc[q] = hist[255 - 8 * lane - q]; s += hloc[q]; }
                unsigned incl = s;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += t;
                }
                unsigned excl = incl - s;
                if (excl < need && need <= incl) {
                    unsigned cum = excl;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (cum < need && need <= cum + hloc[q]) {
                            sel_prefix = prefix | ((uint32_t)(255 - 8 * lane - q) << shift);
                            sel_need = need - cum;
                            hist[0] = hloc[q];  // population of the chosen bin (read below after the last pass)
                            break;
                        }
                        cum += hloc[q];
                    }
                }
            }
            __syncthreads();
        }
        eq_total = hist[0];  // elements equal to the threshold key
    }
    const uint32_t thr = sel_prefix;
    const unsigned need_eq = sel_need;  // how many of the == thr elements belong to the top K
    __syncthreads();

    // ---- collect survivors into the leader's candidate table ---------------------------------------
    cluster.barrier_wait();  // pairs with the arrive at the top: all CTAs of the cluster are alive
    unsigned long long* lead_cand = cluster.map_shared_rank(cand, 0) + slab * kMaxK;
    auto composite = [&](uint32_t k, int e) -> unsigned long long {
        int c = e / (rows * w);
        int rem = e - c * rows * w;
        uint32_t lin = (uint32_t)(c * hw + r0 * w + rem);
        return ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
    };
    if (kk > 0) {
        const bool take_all_eq = (eq_total == need_eq);
        for (int e = tid; e < n; e += kThreads) {
            uint32_t k = keys[e];
            if (k > thr || (take_all_eq && k == thr)) {
                unsigned pos = atomicAdd(&n_cand, 1u);
                lead_cand[pos] = composite(k, e);
            }
        }
        if (!take_all_eq) {
            // more elements tie at the threshold than fit: take the need_eq lowest indices, in order
            for (int base = 0; base < n; base += kThreads) {
                __syncthreads();
                unsigned running = eq_running;
                if (running >= need_eq) break;
                int e = base + tid;
                bool eq = (e < n) && (keys[e] == thr);
                unsigned bal = __ballot_sync(0xFFFFFFFFu, eq);
                if (lane == 0) warp_sums[warp] = __popc(bal);
                __syncthreads();
                unsigned before = running;
                for (int q = 0; q < warp; ++q) before += warp_sums[q];
                unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
                if (eq && rank < need_eq) {
                    unsigned pos = atomicAdd(&n_cand, 1u);
                    lead_cand[pos] = composite(thr, e);
                }
                if (tid == 0) {
                    unsigned tot = 0;
                    for (int q = 0; q < kWarps; ++q) tot += warp_sums[q];
                    eq_running = running + tot;
                }
            }
        }
    }
    __syncthreads();
    // unused tail of this CTA's K slots: composite 0 sorts below every real candidate
    for (int i = kk + tid; i < K; i += kThreads) lead_cand[i] = 0ull;
    cluster.sync();  // DSMEM writes are visible to the leader; non-leaders may now exit
    if (slab != 0) return;

    // ---- leader: sort kSlabs*K candidates (descending) and emit the frame's top K -------------------
    // compact [slab][kMaxK] -> dense [kSlabs*K] at the front of the dynamic shared memory, padded to 2^m
    unsigned long long* sorted = reinterpret_cast<unsigned long long*>(smem_raw);
    const int total = kSlabs * K;
    int P = 1;
    while (P < total) P <<= 1;
    for (int i = tid; i < P; i += kThreads) {
        unsigned long long v = 0ull;
        if (i < total) v = cand[(i / K) * kMaxK + (i % K)];
        sorted[i] = v;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (P >> 1); i += kThreads) {
                int lo = ((i / stride) * (stride << 1)) + (i % stride);
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                unsigned long long x = sorted[lo], y = sorted[hi];
                if ((x < y) == desc) { sorted[lo] = y; sorted[hi] = x; }
            }
            __syncthreads();
        }
    }
    for (int k = tid; k < K; k += kThreads) {
        unsigned long long v = sorted[k];
        uint32_t key = (uint32_t)(v >> 32);
        uint32_t lin = 0xFFFFFFFFu - (uint32_t)v;
        int c = lin / hw;
        int sp = lin - c * hw;
        int y = sp / w;
        int x = sp - y * w;
        float score = orderable_to_float(key);
        size_t o = (size_t)b * K + k;
        if (a.inds) a.inds[o] = sp;
        if (a.tk_score) {
            a.tk_score[o] = score;
            a.tk_cls[o] = c;
            a.tk_ys[o] = (float)y;   // floor_divide(ind, w).float(), evaluation_utils.py:53
            a.tk_xs[o] = (float)x;   // (ind % w).int().float(), :54
        }
        if (a.det) {
            float xs, ys;
            if (a.off) {
                const float* ob = a.off + (size_t)b * 2 * hw;
                xs = __fadd_rn((float)x, ob[sp]);        // :85
                ys = __fadd_rn((float)y, ob[hw + sp]);   // :86
            } else {
                xs = __fadd_rn((float)x, 0.5f);          // :88-89
                ys = __fadd_rn((float)y, 0.5f);
            }
            const float* db = a.dir + (size_t)b * 2 * hw;
            const float* zb = a.zc + (size_t)b * hw;
            const float* mb = a.dim + (size_t)b * 3 * hw;
            float* d = a.det + o * 10;                   // :103 column order
            d[0] = score; d[1] = xs; d[2] = ys; d[3] = zb[sp];
            d[4] = mb[sp]; d[5] = mb[hw + sp]; d[6] = mb[2 * hw + sp];
            d[7] = db[sp]; d[8] = db[hw + sp];
            d[9] = (float)c;
        }
    }
}

__global__ void __launch_bounds__(256)
nms_kernel(const float* __restrict__ heat, int planes, int h, int w, float* __restrict__ out) {
    const size_t total = (size_t)planes * h * w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % w);
        int y = (int)((i / w) % h);
        const float* p = heat + i;
        float v = p[0], m = v;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = y + dy;
            if (yy < 0 || yy >= h) continue;
            const float* r = p + dy * w;
            if (x > 0) m = nanmax(m, r[-1]);
            m = nanmax(m, r[0]);
            if (x < w - 1) m = nanmax(m, r[1]);
        }
        out[i] = __fmul_rn(v, (m == v) ? 1.0f : 0.0f);
    }
}

__global__ void __launch_bounds__(128)
post_process_kernel(const float* __restrict__ det, int n, int num_classes, float down_ratio, float bsy, float bev_w,
                    float bsx, float bev_h, float thresh, float* __restrict__ out, int32_t* __restrict__ cls,
                    uint8_t* __restrict__ keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* d = det + (size_t)i * 10;
    float* o = out + (size_t)i * 8;
    float score = d[0];
    o[0] = score;
    o[1] = __fmul_rn(d[1], down_ratio);                      // evaluation_utils.py:138
    o[2] = __fmul_rn(d[2], down_ratio);                      // :139
    o[3] = d[3];
    o[4] = d[4];
    o[5] = __fmul_rn(__fdiv_rn(d[5], bsy), bev_w);           // :142  (divide, then multiply)
    o[6] = __fmul_rn(__fdiv_rn(d[6], bsx), bev_h);           // :143
    o[7] = atan2f(d[7], d[8]);                               // :108-109, :144
    float cf = d[9];
    int c = (cf >= 0.0f && cf < (float)num_classes && cf == floorf(cf)) ? (int)cf : -1;
    cls[i] = c;
    keep[i] = (c >= 0 && score > thresh) ? 1 : 0;            // :134, :152
}

int launch_decode(DecodeArgs a, cudaStream_t stream) {
    SFA_REQUIRE(a.B >= 0 && a.C > 0 && a.h > 0 && a.w > 0, "bad head shape B=%d C=%d h=%d w=%d", a.B, a.C, a.h, a.w);
    SFA_REQUIRE(a.K > 0 && a.K <= kMaxK, "K=%d unsupported (1..%d)", a.K, kMaxK);
    // torch.topk(scores.view(B, C, -1), K) raises when K > h*w (evaluation_utils.py:50)
    SFA_REQUIRE((long long)a.K <= (long long)a.h * a.w, "K=%d exceeds h*w=%d (the reference's topk raises)", a.K, a.h * a.w);
    SFA_REQUIRE((long long)a.C * a.h * a.w < 0xFFFFFFFFll, "head too large");
    if (a.B == 0) return SFA_OK;
    a.rows_per_slab = (a.h + kSlabs - 1) / kSlabs;
    const long long n_max = (long long)a.C * a.rows_per_slab * a.w;
    size_t tile_bytes = (size_t)a.C * (a.rows_per_slab + 2) * a.w * sizeof(float);
    int P = 1;
    while (P < kSlabs * a.K) P <<= 1;
    size_t smem = tile_bytes > (size_t)P * 8 ? tile_bytes : (size_t)P * 8;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kSlabs, a.B, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kSlabs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;

    auto launch = [&](auto kernel) -> int {
        if (smem > 48 * 1024)
            SFA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaError_t err = cudaSuccess;
        SFA_LAUNCH("peak_decode", stream, err = cudaLaunchKernelEx(&cfg, kernel, a));
        SFA_CUDA_TRY(err);
        return SFA_OK;
    };
    if (smem > 200 * 1024 || n_max > 96ll * kThreads) {
        set_error("head %dx%dx%d too large for the fused decode (slab of %lld elements)", a.C, a.h, a.w, n_max);
        return SFA_ERR_UNSUPPORTED;
    }
    if (n_max <= 36ll * kThreads) return launch(decode_kernel<36>);
    if (n_max <= 64ll * kThreads) return launch(decode_kernel<64>);
    return launch(decode_kernel<96>);
}

}  // namespace
}  // namespace sfa

using namespace sfa;

extern "C" int sfa_decode(const float* hm, const float* cen_offset, const float* direction, const float* z_coor,
                          const float* dim, int32_t B, int32_t C, int32_t h, int32_t w, int32_t K, float* det,
                          int64_t* inds, sfa_stream_t stream) {
    SFA_REQUIRE(B == 0 || (hm && direction && z_coor && dim && det), "NULL pointer argument");
    DecodeArgs a = {};
    a.hm = hm; a.off = cen_offset; a.dir = direction; a.zc = z_coor; a.dim = dim;
    a.B = B; a.C = C; a.h = h; a.w = w; a.K = K;
    a.do_nms = 1;
    a.det = det; a.inds = inds;
    return launch_decode(a, (cudaStream_t)stream);
}

extern "C" int sfa_topk(const float* scores, int32_t B, int32_t C, int32_t h, int32_t w, int32_t K, float* score,
                        int64_t* inds, int32_t* clses, float* ys, float* xs, sfa_stream_t stream) {
    SFA_REQUIRE(B == 0 || (scores && score && inds && clses && ys && xs), "NULL pointer argument");
    DecodeArgs a = {};
    a.hm = scores;
    a.B = B; a.C = C; a.h = h; a.w = w; a.K = K;
    a.do_nms = 0;
    a.inds = inds; a.tk_score = score; a.tk_cls = clses; a.tk_ys = ys; a.tk_xs = xs;
    return launch_decode(a, (cudaStream_t)stream);
}

extern "C" int sfa_nms(const float* heat, int32_t planes, int32_t h, int32_t w, float* out, sfa_stream_t stream) {
    SFA_REQUIRE(planes >= 0 && h > 0 && w > 0, "bad shape planes=%d h=%d w=%d", planes, h, w);
    if (planes == 0) return SFA_OK;
    SFA_REQUIRE(heat && out, "NULL pointer argument");
    size_t total = (size_t)planes * h * w;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
    SFA_LAUNCH("nms", (cudaStream_t)stream, nms_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(heat, planes, h, w, out));
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}

extern "C" int sfa_post_process(const float* det, int32_t B, int32_t K, int32_t num_classes, float down_ratio,
                                float bound_size_y, float bev_width, float bound_size_x, float bev_height,
                                float peak_thresh, float* out, int32_t* cls, uint8_t* keep, sfa_stream_t stream) {
    SFA_REQUIRE(B >= 0 && K >= 0, "bad shape B=%d K=%d", B, K);
    int n = B * K;
    if (n == 0) return SFA_OK;
    SFA_REQUIRE(det && out && cls && keep, "NULL pointer argument");
    SFA_LAUNCH("post_process", (cudaStream_t)stream,
               post_process_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
                   det, n, num_classes, down_ratio, bound_size_y, bev_width, bound_size_x, bev_height, peak_thresh, out,
                   cls, keep));
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}
