// Stage B of the SFA3D hot path on B200: heat-map peak decode.
//
// Replaces (reference, read-only at /root/reference), utils/evaluation_utils.py:
//   _nms :21-26     3x3 max-pool (stride 1, -inf padding), keep where pooled == value, heat * keep
//   _topk :47-62    per-class top-K then top-K of the C*K candidates (== global top-K of C*h*w)
//   _transpose_and_gather_feat :40-44 x4   NHWC copies of every head just to fetch K rows
//   decode :77-105  assembly of [B, K, 10]
//   post_processing :112-163 ("evaluation_utils copy.py":112-143 per-sample semantics), dense form
//
// One kernel, one thread-block CLUSTER of S (<= 8) CTAs per frame:
//   - each CTA owns a slab of rows of all C classes: it stages slab + halo rows in shared memory
//     with 16-B coalesced loads, applies the 3x3 peak-keep there and stores every value as an
//     orderable 32-bit key in a second shared array;
//   - exact top-K of the slab by an 8-bit MSB-first radix select over those keys (4 passes; the
//     histogram updates peel the warp's two most common bins first, so the huge tie groups a
//     heat map has after NMS — suppressed cells are all 0, clamped sigmoids plateau at 1e-4 — cost
//     one shared atomic per warp instead of 32 serialised ones); ties at the K-th value resolve
//     toward the lower index;
//   - the slab's K survivors are ordered locally (rank by counting) and sent to the cluster
//     leader's shared memory through DSMEM as 64-bit (key << 32 | ~linear_index) words; the leader
//     merges the S sorted lists (own position + one binary search per other list) and gathers the
//     8 regression values per detection straight from the NCHW heads (one 32-B sector each — no
//     NHWC transpose).
// The heat map is read from HBM exactly once; nothing intermediate is written to HBM.
//
// Tie rule (torch.topk leaves it implementation-defined): equal scores are ordered by lower class,
// then lower y*w+x — i.e. descending 64-bit composite key, which is a strict total order.
#include "sfa_common.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace sfa {
namespace {

constexpr int kMaxSlabs = 8;       // CTAs per cluster (portable maximum)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 128;
constexpr int kListCap = 3072;             // compact list of positive cells per slab (u16 entries)
constexpr uint32_t kNanKey = 0xFFFFFFFFu;  // torch.topk ranks NaN above everything
constexpr size_t kSmemTarget = 80 * 1024;  // two CTAs per SM with room to spare
constexpr size_t kSmemLimit = 200 * 1024;

struct DecodeArgs {
    const float* hm;      // [B,C,h,w]
    const float* off;     // [B,2,h,w] or null
    const float* dir;     // [B,2,h,w]
    const float* zc;      // [B,1,h,w]
    const float* dim;     // [B,3,h,w]
    int B, C, h, w, K;
    int slabs, rows_per_slab;
    int do_nms;
    int vec4;             // hm is 16-B aligned and w % 4 == 0
    float* det;           // [B,K,10] or null
    int64_t* inds;        // [B,K] or null
    // _topk outputs (all null for decode)
    float* tk_score; int32_t* tk_cls; float* tk_ys; float* tk_xs;
};

__device__ __forceinline__ float nanmax(float a, float b) {
    // max_pool2d propagates NaN (ATen: `val > max || isnan(val)`)
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}

// hist[bin] += 1 for every active lane.  The warp's (up to) two most common bins are added by one
// lane each; whatever is left goes lane by lane.  Must be called by all 32 lanes.
__device__ __forceinline__ void hist_add(unsigned int* hist, bool active, uint32_t bin, int lane) {
    unsigned remaining = __ballot_sync(0xFFFFFFFFu, active);
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        if (remaining == 0) return;   // warp-uniform
        const int leader = __ffs(remaining) - 1;
        const uint32_t lb = __shfl_sync(0xFFFFFFFFu, bin, leader);
        const unsigned grp = __ballot_sync(0xFFFFFFFFu, active && bin == lb) & remaining;
        if (lane == leader) atomicAdd(&hist[lb], (unsigned)__popc(grp));
        remaining &= ~grp;
    }
    if (remaining & (1u << lane)) atomicAdd(&hist[bin], 1u);
}

__global__ void __launch_bounds__(kThreads)
decode_kernel(DecodeArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long cand[kMaxSlabs * kMaxK];  // leader: S sorted lists
    __shared__ unsigned long long surv[kMaxK];              // this slab's survivors, unordered
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sel_prefix, sel_need, sel_eqpop, n_surv, eq_running, n_pos;
    __shared__ unsigned int warp_sums[kWarps];

    const int slab = cluster.block_rank();
    const int S = a.slabs;
    const int b = blockIdx.y;
    const int C = a.C, h = a.h, w = a.w, K = a.K;
    const int hw = h * w;
    const int r0 = slab * a.rows_per_slab;
    const int rows = max(0, min(a.rows_per_slab, h - r0));
    const int n = C * rows * w;                  // elements this CTA selects from
    const int trows = a.rows_per_slab + 2;       // tile rows incl. halo
    const int tplane = trows * w;
    // shared layout: tkeys [C][trows][w] u32 | okeys [C][rows_per_slab][w] u32 | plist [kListCap] u16
    uint32_t* tkeys = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* okeys = tkeys + (size_t)C * tplane;
    unsigned short* plist = reinterpret_cast<unsigned short*>(okeys + (size_t)C * a.rows_per_slab * w);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // every CTA of the cluster must be running before anyone writes into the leader's shared memory
    cluster.barrier_arrive();

    // ---- stage slab + halo rows (r0-1 .. r0+rows) as ORDERABLE KEYS ---------------------------------
    // Per class the tile rows that exist in the map are one contiguous run of the NCHW plane, so the
    // copy is linear; rows outside the map get key 0, below every real value (max_pool2d pads with
    // -inf, whose key is 0x007FFFFF).  NaN maps to the largest key, so an integer max propagates it
    // exactly like ATen's max_pool2d does.
    if (tid == 0) { sel_prefix = 0; sel_need = (unsigned)min(K, n); sel_eqpop = 0; n_surv = 0; eq_running = 0; n_pos = 0; }
    if (rows > 0) {
        const float* hmb = a.hm + (size_t)b * C * hw;
        const int y_lo = r0 - 1;
        const int tr_first = y_lo < 0 ? 1 : 0;                      // first tile row inside the map
        const int tr_end = min(rows + 2, h - y_lo);                 // one past the last tile row inside the map
        if (a.vec4) {
            const int w4 = w >> 2, tplane4 = trows * w4;
            const int lo4 = tr_first * w4, hi4 = tr_end * w4;
            for (int c = 0; c < C; ++c) {
                const float4* src = reinterpret_cast<const float4*>(hmb + (size_t)c * hw) + (ptrdiff_t)y_lo * w4;
                uint4* dst = reinterpret_cast<uint4*>(tkeys + (size_t)c * tplane);
                for (int i = tid; i < tplane4; i += kThreads) {
                    uint4 k = make_uint4(0u, 0u, 0u, 0u);
                    if (i >= lo4 && i < hi4) {
                        const float4 v = __ldg(src + i);
                        k = make_uint4(orderable_u32(v.x, kNanKey), orderable_u32(v.y, kNanKey),
                                       orderable_u32(v.z, kNanKey), orderable_u32(v.w, kNanKey));
                    }
                    dst[i] = k;
                }
            }
        } else {
            const int lo = tr_first * w, hi = tr_end * w;
            for (int c = 0; c < C; ++c) {
                const float* src = hmb + (size_t)c * hw + (ptrdiff_t)y_lo * w;
                uint32_t* dst = tkeys + (size_t)c * tplane;
                for (int i = tid; i < tplane; i += kThreads)
                    dst[i] = (i >= lo && i < hi) ? orderable_u32(__ldg(src + i), kNanKey) : 0u;
            }
        }
    }
    __syncthreads();

    // ---- 3x3 peak keep in the key domain: one thread walks one (class, x) column down the slab ------
    // heat * (maxpool(heat) == heat), evaluation_utils.py:21-26:
    //   own is NaN             -> NaN           (NaN * keep)
    //   max of the 3x3 == own  -> own           (keep = 1)
    //   otherwise              -> own * 0 = 0, or NaN when own is +-inf
    // Cells with a value above zero are remembered in a per-thread row bitmask and, after one
    // block-wide scan of the per-thread counts, written to the compact list `plist`: in the common
    // case they are the only cells that can reach the top K, and the select runs on that list.
    {
        const uint32_t kZero = 0x80000000u, kPosInf = 0xFF800000u, kNegInf = 0x007FFFFFu;
        const int ncols = C * w;
        const bool listable = a.rows_per_slab <= 64 && n <= 65535;
        for (int col0 = 0; col0 < ncols; col0 += kThreads) {   // block-uniform trip count (barriers inside)
            const int col = col0 + tid;
            const bool valid = col < ncols && rows > 0;
            const int c = valid ? col / w : 0;
            const int x = valid ? col - c * w : 0;
            const uint32_t* t = tkeys + (size_t)c * tplane + x;   // tile row 0 (halo above the slab)
            const bool has_l = x > 0, has_r = x < w - 1;
            unsigned long long posmask = 0ull;
            if (valid) {
                uint32_t own = t[w];
                uint32_t h_prev = t[0], h_cur = own;
                if (has_l) { h_prev = max(h_prev, t[-1]); h_cur = max(h_cur, t[w - 1]); }
                if (has_r) { h_prev = max(h_prev, t[1]); h_cur = max(h_cur, t[w + 1]); }
                uint32_t* orow = okeys + (size_t)c * rows * w + x;
                for (int r = 0; r < rows; ++r) {
                    const uint32_t* nx = t + (r + 2) * w;
                    const uint32_t own_next = nx[0];
                    uint32_t h_next = own_next;
                    if (has_l) h_next = max(h_next, nx[-1]);
                    if (has_r) h_next = max(h_next, nx[1]);
                    uint32_t out = own;
                    if (a.do_nms && own != kNanKey) {
                        const uint32_t m = max(max(h_prev, h_cur), h_next);
                        if (m != own) out = (own == kPosInf || own == kNegInf) ? kNanKey : kZero;
                    }
                    orow[r * w] = out;
                    if (out > kZero) posmask |= 1ull << (r & 63);
                    h_prev = h_cur; h_cur = h_next; own = own_next;
                }
            }
            // exclusive scan of the per-thread counts -> list offsets (column-major, then row order)
            const unsigned cnt = listable ? (unsigned)__popcll(posmask) : 0u;
            unsigned incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            unsigned at = n_pos + incl - cnt;
            for (int q = 0; q < warp; ++q) at += warp_sums[q];
            unsigned block_total = 0;
            for (int q = 0; q < kWarps; ++q) block_total += warp_sums[q];
            while (posmask && listable) {
                const int r = __ffsll((long long)posmask) - 1;
                posmask &= posmask - 1;
                if (at < (unsigned)kListCap) plist[at] = (unsigned short)((c * rows + r) * w + x);
                ++at;
            }
            __syncthreads();
            if (tid == 0) n_pos = listable ? n_pos + block_total : 0xFFFFFFFFu;
            __syncthreads();
        }
    }

    // The select runs over the compact list of positive cells when it holds at least K of them and
    // did not overflow (then nothing <= 0 can be in the top K); otherwise over every cell of the slab.
    const int kk = min(K, n);
    const bool use_list = n_pos >= (unsigned)kk && n_pos <= (unsigned)kListCap && n <= 65535 && kk > 0;
    const int n_items = use_list ? (int)n_pos : n;
    auto item_elem = [&](int i) -> int { return use_list ? (int)plist[i] : i; };

    // ---- radix select: key of the K-th largest element of this slab --------------------------------
    unsigned int eq_total = 0;
    if (kk > 0) {
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
            for (int i = tid; i < 256; i += kThreads) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = sel_prefix;
            for (int base = 0; base < n_items; base += kThreads) {
                const int i = base + tid;
                bool active = false;
                uint32_t bin = 0;
                if (i < n_items) {
                    const uint32_t k = okeys[item_elem(i)];
                    active = ((k ^ prefix) & himask) == 0;
                    bin = (k >> shift) & 255u;
                }
                if (use_list) {   // positives are spread over many bins: plain shared atomics
                    if (active) atomicAdd(&hist[bin], 1u);
                } else {          // whole slab: huge tie groups (zeros, plateaus) -> peel them
                    hist_add(hist, active, bin, lane);
                }
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
                            sel_eqpop = hloc[q];  // population of the chosen bin
                            break;
                        }
                        cum += hloc[q];
                    }
                }
            }
            __syncthreads();
        }
        eq_total = sel_eqpop;  // after the last pass: elements equal to the threshold key
    }
    const uint32_t thr = sel_prefix;
    const unsigned need_eq = sel_need;  // how many of the == thr elements belong to the top K

    // ---- collect the slab's K survivors as composite words -----------------------------------------
    auto composite = [&](uint32_t k, int e) -> unsigned long long {
        const int c = e / (rows * w);
        const int rem = e - c * rows * w;
        const uint32_t lin = (uint32_t)(c * hw + r0 * w + rem);
        return ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
    };
    if (kk > 0) {
        const bool take_all_eq = (eq_total == need_eq);
        for (int i = tid; i < n_items; i += kThreads) {
            const int e = item_elem(i);
            const uint32_t k = okeys[e];
            if (k > thr || (take_all_eq && k == thr)) {
                const unsigned pos = atomicAdd(&n_surv, 1u);
                surv[pos] = composite(k, e);
            }
        }
        if (!take_all_eq) {
            // more elements tie at the threshold than fit: take the need_eq lowest indices, in order
            for (int base = 0; base < n; base += kThreads) {
                __syncthreads();
                const unsigned running = eq_running;
                if (running >= need_eq) break;
                const int e = base + tid;
                const bool eq = (e < n) && (okeys[e] == thr);
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, eq);
                if (lane == 0) warp_sums[warp] = __popc(bal);
                __syncthreads();
                unsigned before = running;
                for (int q = 0; q < warp; ++q) before += warp_sums[q];
                const unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
                if (eq && rank < need_eq) {
                    const unsigned pos = atomicAdd(&n_surv, 1u);
                    surv[pos] = composite(thr, e);
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

    // ---- order the survivors (rank by counting; composites are distinct) and ship them to the leader
    cluster.barrier_wait();  // pairs with the arrive at the top: all CTAs of the cluster are alive
    unsigned long long* lead = cluster.map_shared_rank(cand, 0) + slab * kMaxK;
    for (int i = tid; i < K; i += kThreads) {
        if (i < kk) {
            const unsigned long long v = surv[i];
            int rank = 0;
            for (int j = 0; j < kk; ++j) rank += (surv[j] > v) ? 1 : 0;
            lead[rank] = v;
        } else {
            lead[i] = 0ull;  // padding sorts below every real candidate
        }
    }
    cluster.sync();  // DSMEM writes are visible to the leader; non-leaders may now exit
    if (slab != 0) return;

    // ---- leader: merge S descending lists; rank = position in own list + #greater in every other ---
    for (int i = tid; i < S * K; i += kThreads) {
        const int s = i / K;
        const int j = i - s * K;
        const unsigned long long v = cand[s * kMaxK + j];
        if (v == 0ull) continue;
        int rank = j;
        for (int t = 0; t < S; ++t) {
            if (t == s) continue;
            const unsigned long long* lst = cand + t * kMaxK;
            int lo = 0, hi = K;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lst[mid] > v) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank >= K) continue;
        const int k = rank;
        const uint32_t key = (uint32_t)(v >> 32);
        const uint32_t lin = 0xFFFFFFFFu - (uint32_t)v;
        const int c = lin / hw;
        const int sp = lin - c * hw;
        const int y = sp / w;
        const int x = sp - y * w;
        const float score = orderable_to_float(key);
        const size_t o = (size_t)b * K + k;
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

size_t slab_smem_bytes(int C, int h, int w, int slabs) {
    const int rps = (h + slabs - 1) / slabs;
    return (size_t)C * (rps + 2) * w * sizeof(float) + (size_t)C * rps * w * sizeof(uint32_t) +
           (size_t)kListCap * sizeof(unsigned short);
}

int launch_decode(DecodeArgs a, cudaStream_t stream) {
    SFA_REQUIRE(a.B >= 0 && a.C > 0 && a.h > 0 && a.w > 0, "bad head shape B=%d C=%d h=%d w=%d", a.B, a.C, a.h, a.w);
    SFA_REQUIRE(a.K > 0 && a.K <= kMaxK, "K=%d unsupported (1..%d)", a.K, kMaxK);
    // torch.topk(scores.view(B, C, -1), K) raises when K > h*w (evaluation_utils.py:50)
    SFA_REQUIRE((long long)a.K <= (long long)a.h * a.w, "K=%d exceeds h*w=%d (the reference's topk raises)", a.K, a.h * a.w);
    SFA_REQUIRE((long long)a.C * a.h * a.w < 0x7FFFFFFFll, "head too large");
    if (a.B == 0) return SFA_OK;
    // slabs per frame: as few as keep a CTA's tile + keys within ~74 KB (3 CTAs / SM), more when the
    // batch alone cannot fill the 148 SMs; at most the portable cluster size
    int slabs = 1;
    while (slabs < kMaxSlabs && slabs < a.h &&
           (slab_smem_bytes(a.C, a.h, a.w, slabs) > kSmemTarget || (long long)a.B * slabs < 2 * kNumSMs))
        ++slabs;
    if (const char* e = getenv("SFA_DECODE_SLABS")) {   // tuning aid
        int v = atoi(e);
        if (v >= 1 && v <= kMaxSlabs && v <= a.h) slabs = v;
    }
    a.slabs = slabs;
    a.rows_per_slab = (a.h + slabs - 1) / slabs;
    const size_t smem = slab_smem_bytes(a.C, a.h, a.w, slabs);
    if (smem > kSmemLimit) {
        set_error("head %dx%dx%d too large for the fused decode (%zu B of shared memory per slab)", a.C, a.h, a.w, smem);
        return SFA_ERR_UNSUPPORTED;
    }
    a.vec4 = ((a.w & 3) == 0 && (reinterpret_cast<uintptr_t>(a.hm) & 15) == 0) ? 1 : 0;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(slabs, a.B, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = slabs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SFA_CUDA_TRY(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    cudaError_t err = cudaSuccess;
    SFA_LAUNCH("peak_decode", stream, err = cudaLaunchKernelEx(&cfg, decode_kernel, a));
    SFA_CUDA_TRY(err);
    return SFA_OK;
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
