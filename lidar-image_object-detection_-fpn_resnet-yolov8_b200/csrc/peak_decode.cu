// Stage B of the SFA3D hot path on B200: heat-map peak decode.
//
// Replaces (reference, read-only at /root/reference), utils/evaluation_utils.py:
//   _nms :21-26     3x3 max-pool (stride 1, -inf padding), keep where pooled == value, heat * keep
//   _topk :47-62    per-class top-K then top-K of the C*K candidates (== global top-K of C*h*w)
//   _transpose_and_gather_feat :40-44 x4   NHWC copies of every head just to fetch K rows
//   decode :77-105  assembly of [B, K, 10]
//   post_processing :112-163 ("evaluation_utils copy.py":112-143 per-sample semantics), dense form
//
// Two kernels per batch, nothing but the candidate list in between (it lives in L2):
//
//   peak_candidates : one CTA per (frame, slab of 16 rows, all classes).  Stages slab + halo rows in
//     shared memory as ORDERABLE 32-bit keys (16-B coalesced loads), applies the 3x3 peak-keep in the
//     key domain — a thread takes 4 columns x 4 rows (16-B shared loads, 3-input max), or, in tiles
//     that hold NaN / inf, walks one (class, x) column with the full heat * keep semantics — and
//     appends every cell whose kept value is > 0 to the frame's candidate list as a 64-bit word
//     (key << 32 | ~linear_index): one block-wide scan and ONE global atomicAdd per CTA.  In any
//     frame with at least K positive peaks these are the only cells that can reach the top K.
//   peak_select : one CTA per frame.  Exact top-K of the candidate list by MSB-first radix select
//     on the score key (4 passes of 8 bits) and, only if equal scores straddle the K-th place, 4 more
//     passes on the index word — so ties resolve toward the lower (class, y, x), a strict total
//     order.  The K survivors are ranked by counting and each detection gathers its 8 regression
//     values straight from the NCHW heads (one 32-B sector each — no NHWC transpose).
//     Lists that fit (<= 12288 entries) are selected in shared memory, longer ones (clamped
//     sigmoids plateau at 1e-4 and every plateau cell is a peak) from L2.  A frame with fewer than
//     K positive cells (tiny maps, all-negative inputs) takes a slow exact path that re-derives the
//     kept value of every cell from the heat map.
// The heat map is read from HBM exactly once; nothing but the candidate words is written.
//
// Tie rule (torch.topk leaves it implementation-defined): equal scores are ordered by lower class,
// then lower y*w+x — i.e. descending 64-bit composite key, which is a strict total order.
#include "sfa_common.cuh"

#include <stdlib.h>

namespace sfa {
namespace {

constexpr int kCandThreads = 512;
constexpr int kCandWarps = kCandThreads / 32;
#ifndef SFA_CAND_ROWS
#define SFA_CAND_ROWS 19
#endif
// rows of a slab.  19: the 152-row KITTI head splits into exactly 8 slabs, 512 CTAs for a batch of 64 = one wave of the
// 592 resident CTAs (16 rows: 10 slabs, 640 CTAs, a second wave of 48)
constexpr int kCandRows = SFA_CAND_ROWS;
constexpr int kQuadRows = (kCandRows + 3) / 4;   // slab rows per thread of the quad walker (4 row groups x kQuadRows >= kCandRows)
static_assert(4 * kQuadRows <= 32 && kCandRows <= 32, "a thread's cells are the bits of one 32-bit mask");
constexpr int kCapClasses = 8;             // plateau cap: classes a column group may span ...
constexpr int kCapWords = 16;              // ... and 32-bit words per map row (w <= 512)
#ifndef SFA_SEL_THREADS
#define SFA_SEL_THREADS 1024
#endif
#ifndef SFA_SEL_SMEM_ITEMS
#define SFA_SEL_SMEM_ITEMS 12288
#endif
constexpr int kSelThreads = SFA_SEL_THREADS;
constexpr int kSelSmemItems = SFA_SEL_SMEM_ITEMS;   // candidate words selected from shared memory (96 KB)
constexpr int kSelUnroll = 4;              // list entries a select thread loads before it processes them
constexpr int kSelTieWindow = 8192;        // tie-break shortcut: tied cells with a linear index below this
constexpr int kMaxK = 128;
constexpr uint32_t kNanKey = 0xFFFFFFFFu;  // torch.topk ranks NaN above everything
constexpr uint32_t kZeroKey = 0x80000000u, kPosInfKey = 0xFF800000u, kNegInfKey = 0x007FFFFFu;
constexpr size_t kDecodeHeaderBytes = 256;

struct DecodeArgs {
    const float* hm;      // [B,C,h,w]
    const float* off;     // [B,2,h,w] or null
    const float* dir;     // [B,2,h,w]
    const float* zc;      // [B,1,h,w]
    const float* dim;     // [B,3,h,w]
    int B, C, h, w, K;
    int do_nms;
    int apply_sigmoid;    // hm and cen_offset are raw logits: apply _sigmoid (utils/torch_utils.py:44-45) on load
    int vec4;             // hm is 16-B aligned and w % 4 == 0
    uint32_t* counts;               // [B] candidates per frame (workspace; zero between calls)
    unsigned long long* cands;      // [B][C*h*w] candidate words (workspace)
    float* det;           // [B,K,10] or null
    int64_t* inds;        // [B,K] or null
    // _topk outputs (all null for decode)
    float* tk_score; int32_t* tk_cls; float* tk_ys; float* tk_xs;
    // dense post_processing fused behind the decode (sfa_decode_post; pp_rows null = off)
    float* pp_rows; int32_t* pp_cls; uint8_t* pp_keep; float* pp_real;
    int pp_num_classes;
    float pp_down_ratio, pp_bsy, pp_bev_w, pp_bsx, pp_bev_h, pp_thresh, pp_min_x, pp_min_y, pp_min_z;
};

// _sigmoid of utils/torch_utils.py:44-45: clamp(sigmoid(x), 1e-4, 1 - 1e-4).  The clamp bounds are the
// float32 roundings torch uses; sigmoid itself agrees with torch's to the last ulp or two, which is
// why the fused path is tolerance-checked while the un-fused one is bit-exact.
__device__ __forceinline__ float sigmoid_clamped(float x) {
    const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
    return fminf(fmaxf(s, 1e-4f), 1.0f - 1e-4f);
}
__device__ __forceinline__ float act(float x, bool apply_sigmoid) { return apply_sigmoid ? sigmoid_clamped(x) : x; }

// heat * (maxpool3x3(heat) == heat) in the key domain (evaluation_utils.py:21-26):
//   own is NaN            -> NaN          (NaN * keep)
//   max of the 3x3 == own -> own          (keep = 1)
//   otherwise             -> own * 0 = 0, or NaN when own is +-inf
__device__ __forceinline__ uint32_t kept_key(uint32_t own, uint32_t m, bool do_nms) {
    if (!do_nms || own == kNanKey || m == own) return own;
    return (own == kPosInfKey || own == kNegInfKey) ? kNanKey : kZeroKey;
}

#ifndef SFA_CAND_MIN_CTAS
#define SFA_CAND_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kCandThreads, SFA_CAND_MIN_CTAS)
peak_candidates_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int warp_sums[kCandWarps];
    __shared__ unsigned int list_base, grp_min, grp_max, grp_nmin;
    __shared__ uint32_t capbits[kCapClasses * kCandRows * kCapWords];   // bitmap of the group's lowest-key cells
    __shared__ uint32_t caprow[kCapClasses * kCandRows];                // per (class, row): cells before it
    uint32_t* tkeys = reinterpret_cast<uint32_t*>(smem_raw);   // [C][kCandRows + 2][w]
    const int b = blockIdx.y;
    const int C = a.C, h = a.h, w = a.w;
    const int hw = h * w;
    const int r0 = blockIdx.x * kCandRows;
    const int rows = min(kCandRows, h - r0);
    constexpr int trows = kCandRows + 2;
    const int tplane = trows * w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(8) unsigned long long load_bar;   // vec4 staging: the slab's TMA bulk loads complete on it
    __shared__ unsigned int grp_raw;                        // ... and the largest raw bit pattern loaded
    if (tid == 0) {
        grp_min = 0xFFFFFFFFu; grp_max = 0; grp_nmin = 0; grp_raw = 0;
        if (a.vec4) mbar_init(&load_bar, 1);
    }
    fence_proxy_async_smem();   // (the mbarrier's init, before a bulk copy signals it)
    __syncthreads();   // before any warp's atomicMin on grp_min

    // ---- stage slab + halo rows (r0-1 .. r0+rows) as orderable keys ---------------------------------
    // Per class the tile rows that exist in the map are one contiguous run of the NCHW plane, so the
    // copy is linear; rows outside the map get key 0, below every real value (max_pool2d pads with
    // -inf, whose key is 0x007FFFFF).  NaN maps to the largest key, so an integer max propagates it
    // exactly like ATen's max_pool2d does.
    uint32_t tile_min = 0xFFFFFFFFu;   // lowest key among the tile's in-map cells (see "plateau cap" below)
    uint32_t tile_max = 0;             // highest one: >= the +inf key means an inf or NaN somewhere in the tile
    {
        const float* hmb = a.hm + (size_t)b * C * hw;
        const int y_lo = r0 - 1;
        const int tr_first = y_lo < 0 ? 1 : 0;          // first tile row inside the map
        const int tr_end = min(rows + 2, h - y_lo);     // one past the last tile row inside the map
        bool staged = false;
        if (a.vec4) {
            // Fast staging.  The in-map rows of a class are ONE contiguous run of the plane: thread 0 fetches the C runs with
            // TMA bulk loads straight into the tile (no registers, no per-thread address arithmetic, one latency for the whole
            // slab), then every thread turns its share of the floats into keys IN PLACE.  For positive finite values —
            // any sigmoid output — the key is just bits | 0x80000000; a tile holding anything else (sign bit, inf, NaN: the
            // largest RAW bit pattern is then >= 0x7F800000) is staged again by the general loop below.
            const int n_in = (tr_end - tr_first) * w;   // words per class
            if (tid == 0) {
                mbar_expect_tx(&load_bar, (uint32_t)(C * n_in) * 4u);
                for (int c = 0; c < C; ++c)
                    bulk_load_g2s(tkeys + (size_t)c * tplane + tr_first * w, hmb + (size_t)c * hw + (ptrdiff_t)(y_lo + tr_first) * w,
                                  (uint32_t)n_in * 4u, &load_bar);
            }
            for (int c = 0; c < C; ++c) {   // rows outside the map (first / last slab only): key 0
                uint32_t* tc = tkeys + (size_t)c * tplane;
                for (int i = tid; i < tr_first * w; i += kCandThreads) tc[i] = 0u;
                for (int i = tr_end * w + tid; i < tplane; i += kCandThreads) tc[i] = 0u;
            }
            mbar_wait(&load_bar, 0);
            const bool sg = a.apply_sigmoid != 0;
            uint32_t raw_max = 0, raw_min = 0xFFFFFFFFu;
            const int n4 = n_in >> 2;
            for (int c = 0; c < C; ++c) {
                uint4* d = reinterpret_cast<uint4*>(tkeys + (size_t)c * tplane + tr_first * w);
                for (int i = tid; i < n4; i += kCandThreads) {
                    uint4 bts = d[i];
                    if (sg) {
                        bts.x = __float_as_uint(sigmoid_clamped(__uint_as_float(bts.x))); bts.y = __float_as_uint(sigmoid_clamped(__uint_as_float(bts.y)));
                        bts.z = __float_as_uint(sigmoid_clamped(__uint_as_float(bts.z))); bts.w = __float_as_uint(sigmoid_clamped(__uint_as_float(bts.w)));
                    }
                    raw_max = max(raw_max, max(max(bts.x, bts.y), max(bts.z, bts.w)));
                    raw_min = min(raw_min, min(min(bts.x, bts.y), min(bts.z, bts.w)));
                    d[i] = make_uint4(bts.x | 0x80000000u, bts.y | 0x80000000u, bts.z | 0x80000000u, bts.w | 0x80000000u);
                }
            }
            raw_max = __reduce_max_sync(0xFFFFFFFFu, raw_max);
            if (lane == 0) atomicMax(&grp_raw, raw_max);
            __syncthreads();
            staged = grp_raw < 0x7F800000u;   // block-uniform
            if (staged) {
                tile_min = raw_min == 0xFFFFFFFFu ? raw_min : (raw_min | 0x80000000u);
                tile_max = raw_max | 0x80000000u;
            }
        }
        if (staged) {
            // done above
        } else if (a.vec4) {
            const int w4 = w >> 2, tplane4 = trows * w4;
            const int lo4 = tr_first * w4, hi4 = tr_end * w4;
            const int hw4 = hw >> 2;
            const float4* src0 = reinterpret_cast<const float4*>(hmb) + (ptrdiff_t)y_lo * w4;
            uint4* dst = reinterpret_cast<uint4*>(tkeys);
            int c = 0, t4 = tid;   // flat index i = c * tplane4 + t4, advanced without a division
            while (t4 >= tplane4 && c < C) { t4 -= tplane4; ++c; }
            for (int i = tid; i < C * tplane4; i += kCandThreads) {   // all classes in one flat loop
                uint4 k = make_uint4(0u, 0u, 0u, 0u);
                if (t4 >= lo4 && t4 < hi4) {
                    const float4 v = __ldg(src0 + (size_t)c * hw4 + t4);
                    const bool sg = a.apply_sigmoid != 0;
                    k = make_uint4(orderable_u32(act(v.x, sg), kNanKey), orderable_u32(act(v.y, sg), kNanKey),
                                   orderable_u32(act(v.z, sg), kNanKey), orderable_u32(act(v.w, sg), kNanKey));
                    tile_min = min(tile_min, min(min(k.x, k.y), min(k.z, k.w)));
                    tile_max = max(tile_max, max(max(k.x, k.y), max(k.z, k.w)));
                }
                dst[i] = k;
                t4 += kCandThreads;
                while (t4 >= tplane4) { t4 -= tplane4; ++c; }
            }
        } else {
            const int lo = tr_first * w, hi = tr_end * w;
            for (int c = 0; c < C; ++c) {
                const float* src = hmb + (size_t)c * hw + (ptrdiff_t)y_lo * w;
                uint32_t* dst = tkeys + (size_t)c * tplane;
                for (int i = tid; i < tplane; i += kCandThreads) {
                    const bool in = i >= lo && i < hi;
                    const uint32_t k = in ? orderable_u32(act(__ldg(src + i), a.apply_sigmoid != 0), kNanKey) : 0u;
                    if (in) tile_min = min(tile_min, k);
                    dst[i] = k;
                }
            }
        }
    }
    tile_min = __reduce_min_sync(0xFFFFFFFFu, tile_min);
    tile_max = __reduce_max_sync(0xFFFFFFFFu, tile_max);
    if (lane == 0) { atomicMin(&grp_min, tile_min); atomicMax(&grp_max, tile_max); }
    __syncthreads();
    const uint32_t vmin = grp_min;

    // ---- fast walker: 4 columns x 4 rows per thread on a separable 3x3 max ------------------------------
    // Block-uniform choice.  Without +-inf / NaN in the tile (vmin above the -inf key, the maximum below the
    // +inf key) the kept value of a cell is simply "own if own == max3x3 else 0", so a cell is listed iff
    // own == max3x3 and own > 0.  A thread takes a quad of columns and 4 slab rows: per tile row one 16-B
    // shared load plus the two neighbours left and right of the quad, and one 3-input max per cell each for
    // the horizontal and the vertical direction.
    const bool quad = a.vec4 && a.do_nms && vmin > kNegInfKey && grp_max < kPosInfKey;

    // ---- one thread walks one (class, x) column down the slab ---------------------------------------
    const size_t frame_cap = (size_t)C * hw;
    unsigned long long* list = a.cands + (size_t)b * frame_cap;
    const int ncols = C * w;
    // A thread's cells are the set bits of posmask / minmask.  Scalar walker: bit r = slab row r of column
    // (c, x).  Quad walker: bit 4*i + j = slab row rg*kQuadRows + i of column (c, x + j).
    for (int col0 = 0; col0 < ncols; col0 += kCandThreads) {   // block-uniform trip count (barriers inside)
        int c = 0, x = 0, rg = 0;
        unsigned posmask = 0, minmask = 0;
        if (quad) {
            const int q = (col0 >> 2) + (tid & (kCandThreads / 4 - 1));   // quad of columns; kCandThreads / 4 quads per pass
            rg = tid / (kCandThreads / 4);                                  // kQuadRows slab rows each
            const int w4 = w >> 2;
            if (q < (ncols >> 2) && rg * kQuadRows < rows) {
                c = q / w4;
                x = (q - c * w4) * 4;
                const uint32_t* tk = tkeys + (size_t)c * tplane + x;
                const bool has_l = x > 0, has_r = x + 4 < w;
                uint4 own;
                auto hrow = [&](int tr) -> uint4 {   // horizontal 3-max of the quad in tile row tr; own = the row's keys
                    own = *reinterpret_cast<const uint4*>(tk + tr * w);
                    const uint32_t l = has_l ? tk[tr * w - 1] : 0u;
                    const uint32_t r = has_r ? tk[tr * w + 4] : 0u;
                    return make_uint4(max(max(l, own.x), own.y), max(max(own.x, own.y), own.z),
                                      max(max(own.y, own.z), own.w), max(max(own.z, own.w), r));
                };
                uint4 h0 = hrow(rg * kQuadRows), h1 = hrow(rg * kQuadRows + 1);
                uint4 cur = own;   // keys of tile row rg*kQuadRows + 1 = slab row rg*kQuadRows
#pragma unroll
                for (int i = 0; i < kQuadRows; ++i) {
                    if (rg * kQuadRows + i >= kCandRows) break;   // (the last row group may be short: stay inside the tile)
                    const uint4 h2 = hrow(rg * kQuadRows + i + 2);
                    const uint4 nxt = own;
                    if (rg * kQuadRows + i < rows) {
                        const uint32_t m0 = max(max(h0.x, h1.x), h2.x), m1 = max(max(h0.y, h1.y), h2.y);
                        const uint32_t m2 = max(max(h0.z, h1.z), h2.z), m3 = max(max(h0.w, h1.w), h2.w);
                        const unsigned p = (m0 == cur.x && cur.x > kZeroKey ? 1u : 0u) | (m1 == cur.y && cur.y > kZeroKey ? 2u : 0u) |
                                           (m2 == cur.z && cur.z > kZeroKey ? 4u : 0u) | (m3 == cur.w && cur.w > kZeroKey ? 8u : 0u);
                        const unsigned e = (cur.x == vmin ? 1u : 0u) | (cur.y == vmin ? 2u : 0u) | (cur.z == vmin ? 4u : 0u) |
                                           (cur.w == vmin ? 8u : 0u);
                        posmask |= p << (4 * i);
                        minmask |= (p & e) << (4 * i);
                    }
                    h0 = h1; h1 = h2; cur = nxt;
                }
            }
        } else {
            const int col = col0 + tid;
            const bool valid = col < ncols;
            c = valid ? col / w : 0;
            x = valid ? col - c * w : 0;
            if (valid) {
                const uint32_t* t = tkeys + (size_t)c * tplane + x;   // tile row 0 (halo above the slab)
                const bool has_l = x > 0, has_r = x < w - 1;
                uint32_t own = t[w];
                uint32_t h_prev = t[0], h_cur = own;
                if (has_l) { h_prev = max(h_prev, t[-1]); h_cur = max(h_cur, t[w - 1]); }
                if (has_r) { h_prev = max(h_prev, t[1]); h_cur = max(h_cur, t[w + 1]); }
#pragma unroll
                for (int r = 0; r < kCandRows; ++r) {
                    if (r < rows) {
                        const uint32_t* nx = t + (r + 2) * w;
                        const uint32_t own_next = nx[0];
                        uint32_t h_next = own_next;
                        if (has_l) h_next = max(h_next, nx[-1]);
                        if (has_r) h_next = max(h_next, nx[1]);
                        const uint32_t kept = kept_key(own, max(max(h_prev, h_cur), h_next), a.do_nms != 0);
                        if (kept > kZeroKey) posmask |= 1u << r;
                        if (kept > kZeroKey && kept == vmin) minmask |= 1u << r;
                        h_prev = h_cur; h_cur = h_next; own = own_next;
                    }
                }
            }
        }
        auto cell_x = [&](int bit) -> int { return quad ? x + (bit & 3) : x; };
        auto cell_r = [&](int bit) -> int { return quad ? rg * kQuadRows + (bit >> 2) : bit; };
        // ---- plateau cap ---------------------------------------------------------------------------
        // Among cells with EQUAL keys the top-K takes the lowest linear indices, so of the cells of this
        // column group that share one key only the K lowest-index ones can ever be selected.  Applied to
        // the tile's LOWEST key (found for free while the tile is staged) — the floor a clamped sigmoid
        // plateaus at, or a constant map — this keeps such maps from listing every cell (69 k words per
        // frame) without changing any result; where the lowest cell is no kept peak (any ordinary map)
        // it costs one warp reduction and one barrier.
        const uint32_t* tcls = tkeys + (size_t)c * tplane;
        auto key_at = [&](int cx, int r) -> uint32_t {   // a positive kept cell keeps its own key, except -inf under a
            const uint32_t own = tcls[(r + 1) * w + cx];   // larger neighbour (-inf * 0 = NaN)
            return (a.do_nms && own == kNegInfKey) ? kNanKey : own;
        };
        {
            const unsigned wn = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(minmask));
            if (lane == 0 && wn) atomicAdd(&grp_nmin, wn);
        }
        __syncthreads();
        const int c_first = col0 / w;
        const int n_cls = min(ncols - 1, col0 + kCandThreads - 1) / w - c_first + 1;
        const int words = (w + 31) >> 5;
        if (grp_nmin > (unsigned)a.K && n_cls <= kCapClasses && words <= kCapWords) {   // block-uniform
            const int n_rows = n_cls * kCandRows;
            for (int i = tid; i < n_rows * kCapWords; i += kCandThreads) capbits[i] = 0;
            __syncthreads();
            const int row0 = (c - c_first) * kCandRows;
            for (unsigned m = minmask; m; m &= m - 1) {
                const int bit = __ffs(m) - 1, cx = cell_x(bit);
                atomicOr(&capbits[(row0 + cell_r(bit)) * kCapWords + (cx >> 5)], 1u << (cx & 31));
            }
            __syncthreads();
            if (tid < n_rows) {
                unsigned tot = 0;
                for (int j = 0; j < words; ++j) tot += __popc(capbits[tid * kCapWords + j]);
                caprow[tid] = tot;
            }
            __syncthreads();
            if (warp == 0) {   // exclusive scan over the (<= 128) rows in (class, row) order = linear-index order
                constexpr int per = (kCapClasses * kCandRows + 31) / 32;
                unsigned v[per], run = 0;
#pragma unroll
                for (int q = 0; q < per; ++q) { v[q] = lane * per + q < n_rows ? caprow[lane * per + q] : 0u; run += v[q]; }
                unsigned incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += t;
                }
                unsigned at = incl - run;
#pragma unroll
                for (int q = 0; q < per; ++q) { if (lane * per + q < n_rows) caprow[lane * per + q] = at; at += v[q]; }
            }
            __syncthreads();
            for (unsigned m = minmask; m; m &= m - 1) {
                const int bit = __ffs(m) - 1, cx = cell_x(bit), r = cell_r(bit);
                const uint32_t* rowbits = capbits + (row0 + r) * kCapWords;
                unsigned rank = caprow[row0 + r] + __popc(rowbits[cx >> 5] & ((1u << (cx & 31)) - 1u));
                for (int j = 0; j < (cx >> 5); ++j) rank += __popc(rowbits[j]);
                if (rank >= (unsigned)a.K) posmask &= ~(1u << bit);   // K lower-index cells of the same key exist
            }
        }

        // exclusive scan of the per-thread counts; one global atomicAdd reserves the CTA's run
        const unsigned cnt = (unsigned)__popc(posmask);
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned ws = lane < kCandWarps ? warp_sums[lane] : 0u, wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += v;
            }
            if (lane < kCandWarps) warp_sums[lane] = wi - ws;   // exclusive warp offsets
            if (lane == kCandWarps - 1) list_base = wi ? atomicAdd(a.counts + b, wi) : 0u;
        }
        __syncthreads();
        size_t at = (size_t)list_base + warp_sums[warp] + incl - cnt;
        while (posmask) {
            const int bit = __ffs(posmask) - 1;
            posmask &= posmask - 1;
            const int cx = cell_x(bit), r = cell_r(bit);
            const uint32_t lin = (uint32_t)(c * hw + (r0 + r) * w + cx);
            list[at++] = ((unsigned long long)key_at(cx, r) << 32) | (unsigned long long)(0xFFFFFFFFu - lin);
        }
        if (tid == 0) grp_nmin = 0;
        __syncthreads();   // warp_sums / list_base / grp_nmin are reused by the next column group
    }
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

struct SelectShared {
    unsigned int hist[256];
    unsigned int prefix, need, eqpop;
};

// MSB-first radix select (descending) over word(i) of the items i in [0, n) that pass filter(i):
// on return sh.prefix is the 32-bit value of the need-th largest word, sh.need how many of the words
// EQUAL to it belong to the top `need`, and sh.eqpop how many words equal it.  All threads call.
// PEEL: aggregate the warp's most common bins (huge tie groups) instead of one atomic per lane.
// UNROLL: list entries a thread loads before it processes them (memory-level parallelism for lists in L2).
template <bool PEEL, int UNROLL, class WordFn>
__device__ void block_radix_select(WordFn word, int n, unsigned need, SelectShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { sh.prefix = 0; sh.need = need; sh.eqpop = 0; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < 256; i += kSelThreads) sh.hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = sh.prefix;
        for (int base = 0; base < n; base += UNROLL * kSelThreads) {   // block-uniform trip count
            bool active[UNROLL];
            uint32_t k[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {   // UNROLL independent loads in flight per thread
                const int i = base + u * kSelThreads + tid;
                k[u] = 0;
                active[u] = (i < n) && word(i, k[u]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const bool a = active[u] && ((k[u] ^ prefix) & himask) == 0;
                // pass 0 bins by sign + 7 exponent bits: scores of one map share two or three of them, and same-address
                // shared atomics serialise — aggregate per warp there whatever the list looks like
                if (PEEL || pass == 0) hist_add(sh.hist, a, (k[u] >> shift) & 255u, lane);
                else if (a) atomicAdd(&sh.hist[(k[u] >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns bins 255-8l .. 248-8l (descending); find the bin holding the need-th word
            const unsigned want = sh.need;
            unsigned hloc[8];
            unsigned s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { hloc[q] = sh.hist[255 - 8 * lane - q]; s += hloc[q]; }
            unsigned incl = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            unsigned excl = incl - s;
            if (excl < want && want <= incl) {
                unsigned cum = excl;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (cum < want && want <= cum + hloc[q]) {
                        sh.prefix = prefix | ((uint32_t)(255 - 8 * lane - q) << shift);
                        sh.need = want - cum;
                        sh.eqpop = hloc[q];
                        break;
                    }
                    cum += hloc[q];
                }
            }
        }
        __syncthreads();
    }
}

// convert_det_to_real_values, evaluation_utils.py:177-193: one post_processing row (score, x, y, z, h, w, l, yaw
// in BEV pixels) -> metres in the lidar frame, [cls, x, y, z, h, w, l, yaw]; every step rounds to fp32 like
// numpy's float32 scalars
__device__ __forceinline__ void real_values_row(const float* __restrict__ o, float cls, float bsy, float bev_w, float bsx,
                                                float bev_h, float min_x, float min_y, float min_z, float* __restrict__ r) {
    r[0] = cls;
    r[1] = __fadd_rn(__fmul_rn(__fdiv_rn(o[2], bev_h), bsx), min_x);   // x <- y pixel, :185
    r[2] = __fadd_rn(__fmul_rn(__fdiv_rn(o[1], bev_w), bsy), min_y);   // y <- x pixel, :186
    r[3] = __fadd_rn(o[3], min_z);                                     // :187
    r[4] = o[4];
    r[5] = __fmul_rn(__fdiv_rn(o[5], bev_w), bsy);                     // :188
    r[6] = __fmul_rn(__fdiv_rn(o[6], bev_h), bsx);                     // :189
    r[7] = -o[7];                                                      // :184
}

// One dense post_processing row (utils/evaluation_utils.py:112-163) from one detection d[10].
__device__ __forceinline__ void post_row(const float* d, int num_classes, float down_ratio, float bsy, float bev_w, float bsx,
                                         float bev_h, float thresh, float min_x, float min_y, float min_z, float* __restrict__ o,
                                         int32_t* __restrict__ cls, uint8_t* __restrict__ keep, float* __restrict__ real) {
    const float score = d[0];
    o[0] = score;
    o[1] = __fmul_rn(d[1], down_ratio);                      // evaluation_utils.py:138
    o[2] = __fmul_rn(d[2], down_ratio);                      // :139
    o[3] = d[3];
    o[4] = d[4];
    o[5] = __fmul_rn(__fdiv_rn(d[5], bsy), bev_w);           // :142  (divide, then multiply)
    o[6] = __fmul_rn(__fdiv_rn(d[6], bsx), bev_h);           // :143
    o[7] = atan2f(d[7], d[8]);                               // :108-109, :144
    const float cf = d[9];
    const int c = (cf >= 0.0f && cf < (float)num_classes && cf == floorf(cf)) ? (int)cf : -1;
    *cls = c;
    *keep = (c >= 0 && score > thresh) ? 1 : 0;              // :134, :152
    if (real) real_values_row(o, cf, bsy, bev_w, bsx, bev_h, min_x, min_y, min_z, real);
}

__global__ void __launch_bounds__(kSelThreads)
peak_select_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SelectShared sel;
    __shared__ unsigned long long surv[kMaxK];
    __shared__ unsigned int n_surv, n_small;
    unsigned long long* scand = reinterpret_cast<unsigned long long*>(smem_raw);
    const int b = blockIdx.x;
    const int C = a.C, h = a.h, w = a.w, K = a.K;
    const int hw = h * w;
    const int tid = threadIdx.x;
    const size_t frame_cap = (size_t)C * hw;
    const unsigned long long* gcand = a.cands + (size_t)b * frame_cap;

    const unsigned n_list = a.counts[b];
    if (tid == 0) n_surv = 0;
    __syncthreads();
    if (tid == 0) a.counts[b] = 0;   // ready for the next call on this workspace

    // mode 0: list in shared memory, 1: list in L2 / global, 2: every cell of the map (kept value
    // re-derived from the heat map) — only when fewer than K cells are positive
    const int mode = n_list < (unsigned)K ? 2 : (n_list <= (unsigned)kSelSmemItems ? 0 : 1);
    const int n = mode == 2 ? (int)frame_cap : (int)n_list;
    if (mode == 0)
        for (int i = tid; i < n; i += kSelThreads) scand[i] = gcand[i];
    const float* hmb = a.hm + (size_t)b * C * hw;
    auto item = [&](int i) -> unsigned long long {
        if (mode == 0) return scand[i];
        if (mode == 1) return gcand[i];
        const int c = i / hw;
        const int sp = i - c * hw;
        const int y = sp / w;
        const int x = sp - y * w;
        const float* p = hmb + (size_t)c * hw + sp;
        const bool sg = a.apply_sigmoid != 0;
        const uint32_t own = orderable_u32(act(__ldg(p), sg), kNanKey);
        uint32_t m = own;
        if (a.do_nms) {
            for (int dy = -1; dy <= 1; ++dy) {
                if (y + dy < 0 || y + dy >= h) continue;
                for (int dx = -1; dx <= 1; ++dx)
                    if (x + dx >= 0 && x + dx < w) m = max(m, orderable_u32(act(__ldg(p + dy * w + dx), sg), kNanKey));
            }
        }
        return ((unsigned long long)kept_key(own, m, a.do_nms != 0) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    };
    __syncthreads();

    // K-th largest composite word: select on the score key, then (only when equal scores straddle
    // the K-th place) on the index word among the cells with exactly that score.  A list that fits
    // shared memory holds positive peaks whose scores are spread over many bins: plain shared
    // atomics; the long lists / whole maps are dominated by a few values: peel them.
    auto hi_word = [&](int i, uint32_t& k) { k = (uint32_t)(item(i) >> 32); return true; };
    if (mode == 0) block_radix_select<false, 1>(hi_word, n, (unsigned)K, sel);
    else           block_radix_select<true, kSelUnroll>(hi_word, n, (unsigned)K, sel);
    const uint32_t thr_hi = sel.prefix;
    uint32_t thr_lo = 0;
    const bool straddle = sel.eqpop > sel.need;
    const unsigned need_eq = sel.need;
    __syncthreads();
    if (straddle) {
        // Of the cells that tie at the K-th score the need_eq LOWEST linear indices win.  When the tie
        // group is huge (a plateau: most of the map) they all sit at the very start of the map, so one
        // pass first gathers the tied cells with an index below kSelTieWindow into shared memory; if
        // there are at least need_eq of them the 4-pass select runs on that short list, else on everything.
        bool done = false;
        if (mode != 0) {   // (mode 0: the whole list already sits in shared memory)
            if (tid == 0) n_small = 0;
            __syncthreads();
            const uint32_t lo_min = 0xFFFFFFFFu - (uint32_t)(kSelTieWindow - 1);   // lo word of index kSelTieWindow-1
            for (int base = 0; base < n; base += kSelUnroll * kSelThreads) {
                unsigned long long v[kSelUnroll];
#pragma unroll
                for (int u = 0; u < kSelUnroll; ++u) {
                    const int i = base + u * kSelThreads + tid;
                    v[u] = i < n ? item(i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < kSelUnroll; ++u)
                    if ((uint32_t)(v[u] >> 32) == thr_hi && (uint32_t)v[u] >= lo_min && v[u] != 0ull) {
                        const unsigned pos = atomicAdd(&n_small, 1u);
                        if (pos < (unsigned)kSelSmemItems) scand[pos] = v[u];
                    }
            }
            __syncthreads();
            const unsigned ns = n_small;
            if (ns >= need_eq && ns <= (unsigned)kSelSmemItems) {
                block_radix_select<true, 1>([&](int i, uint32_t& k) { k = (uint32_t)scand[i]; return true; }, (int)ns, need_eq, sel);
                thr_lo = sel.prefix;
                done = true;
                __syncthreads();
            }
        }
        if (!done) {
            auto lo_word = [&](int i, uint32_t& k) {
                const unsigned long long v = item(i);
                k = (uint32_t)v;
                return (uint32_t)(v >> 32) == thr_hi;
            };
            // index words are distinct but share their leading bytes (indices are < C*h*w): the first
            // passes put every tied cell into one bin, so aggregate
            if (mode == 0) block_radix_select<true, 1>(lo_word, n, need_eq, sel);
            else           block_radix_select<true, kSelUnroll>(lo_word, n, need_eq, sel);
            thr_lo = sel.prefix;
            __syncthreads();
        }
    }
    const unsigned long long thr = ((unsigned long long)thr_hi << 32) | thr_lo;
    if (mode == 0) {
        for (int i = tid; i < n; i += kSelThreads) {
            const unsigned long long v = scand[i];
            if (v >= thr) {
                const unsigned pos = atomicAdd(&n_surv, 1u);
                if (pos < (unsigned)kMaxK) surv[pos] = v;
            }
        }
    } else {
        for (int base = 0; base < n; base += kSelUnroll * kSelThreads) {
            unsigned long long v[kSelUnroll];
#pragma unroll
            for (int u = 0; u < kSelUnroll; ++u) {
                const int i = base + u * kSelThreads + tid;
                v[u] = i < n ? item(i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kSelUnroll; ++u)
                if (v[u] >= thr && v[u] != 0ull) {
                    const unsigned pos = atomicAdd(&n_surv, 1u);
                    if (pos < (unsigned)kMaxK) surv[pos] = v[u];
                }
        }
    }
    __syncthreads();

    // ---- rank the K survivors (composites are distinct) and emit the detections ---------------------
    for (int i = tid; i < K; i += kSelThreads) {
        const unsigned long long v = surv[i];
        int k = 0;
        for (int j = 0; j < K; ++j) k += (surv[j] > v) ? 1 : 0;
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
            // All eight gathers are issued before anything depends on them: written as `d[3] = zb[sp]; d[4] = ...` each
            // store waits for its own load and, issue being in order, holds back the next load — eight serialised trips
            // to HBM (the regression heads are read nowhere else) were 40 % of this kernel.
            const float* db = a.dir + (size_t)b * 2 * hw;
            const float* zb = a.zc + (size_t)b * hw;
            const float* mb = a.dim + (size_t)b * 3 * hw;
            float ox = 0.5f, oy = 0.5f;                  // :88-89 when there is no offset head
            if (a.off) {
                const float* ob = a.off + (size_t)b * 2 * hw;
                ox = __ldg(ob + sp);
                oy = __ldg(ob + hw + sp);
            }
            const float vz = __ldg(zb + sp);
            const float m0 = __ldg(mb + sp), m1 = __ldg(mb + hw + sp), m2 = __ldg(mb + 2 * hw + sp);
            const float d0 = __ldg(db + sp), d1 = __ldg(db + hw + sp);
            if (a.off) {
                ox = act(ox, a.apply_sigmoid != 0);
                oy = act(oy, a.apply_sigmoid != 0);
            }
            const float xs = __fadd_rn((float)x, ox);    // :85
            const float ys = __fadd_rn((float)y, oy);    // :86
            float* d = a.det + o * 10;                   // :103 column order
            d[0] = score; d[1] = xs; d[2] = ys; d[3] = vz;
            d[4] = m0; d[5] = m1; d[6] = m2;
            d[7] = d0; d[8] = d1;
            d[9] = (float)c;
            if (a.pp_rows) {   // dense post_processing of the row, from registers
                const float row[10] = {score, xs, ys, vz, m0, m1, m2, d0, d1, (float)c};
                post_row(row, a.pp_num_classes, a.pp_down_ratio, a.pp_bsy, a.pp_bev_w, a.pp_bsx, a.pp_bev_h, a.pp_thresh,
                         a.pp_min_x, a.pp_min_y, a.pp_min_z, a.pp_rows + o * 8, a.pp_cls + o, a.pp_keep + o,
                         a.pp_real ? a.pp_real + o * 8 : nullptr);
            }
        }
    }
}

__device__ __forceinline__ float nanmax(float a, float b) {
    // max_pool2d propagates NaN (ATen: `val > max || isnan(val)`)
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
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
real_values_kernel(const float* __restrict__ rows, const int32_t* __restrict__ cls, int n, float bsy, float bev_w, float bsx,
                   float bev_h, float min_x, float min_y, float min_z, float* __restrict__ real) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    real_values_row(rows + (size_t)i * 8, (float)cls[i], bsy, bev_w, bsx, bev_h, min_x, min_y, min_z, real + (size_t)i * 8);
}

__global__ void __launch_bounds__(128)
post_process_kernel(const float* __restrict__ det, int n, int num_classes, float down_ratio, float bsy, float bev_w,
                    float bsx, float bev_h, float thresh, float min_x, float min_y, float min_z,
                    float* __restrict__ out, int32_t* __restrict__ cls, uint8_t* __restrict__ keep,
                    float* __restrict__ real) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    post_row(det + (size_t)i * 10, num_classes, down_ratio, bsy, bev_w, bsx, bev_h, thresh, min_x, min_y, min_z,
             out + (size_t)i * 8, cls + i, keep + i, real ? real + (size_t)i * 8 : nullptr);
}

size_t decode_workspace_bytes(int B, int C, int h, int w) {
    return kDecodeHeaderBytes + (((size_t)(B > 0 ? B : 1) * sizeof(uint32_t) + 255) / 256) * 256 +
           (size_t)(B > 0 ? B : 1) * C * h * w * sizeof(unsigned long long);
}

int launch_decode(DecodeArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SFA_REQUIRE(a.B >= 0 && a.C > 0 && a.h > 0 && a.w > 0, "bad head shape B=%d C=%d h=%d w=%d", a.B, a.C, a.h, a.w);
    SFA_REQUIRE(a.K > 0 && a.K <= kMaxK, "K=%d unsupported (1..%d)", a.K, kMaxK);
    // torch.topk(scores.view(B, C, -1), K) raises when K > h*w (evaluation_utils.py:50)
    SFA_REQUIRE((long long)a.K <= (long long)a.h * a.w, "K=%d exceeds h*w=%d (the reference's topk raises)", a.K, a.h * a.w);
    SFA_REQUIRE((long long)a.C * a.h * a.w < 0x7FFFFFFFll, "head too large");
    if (a.B == 0) return SFA_OK;
    SFA_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "workspace must be a 256-B aligned device pointer");
    const size_t need = decode_workspace_bytes(a.B, a.C, a.h, a.w);
    if (workspace_bytes < need) {
        set_error("workspace too small: %zu < %zu (size it with sfa_decode_workspace_bytes)", workspace_bytes, need);
        return SFA_ERR_WORKSPACE_TOO_SMALL;
    }
    const size_t tile_smem = (size_t)a.C * (kCandRows + 2) * a.w * sizeof(uint32_t);
    if (tile_smem > 200 * 1024) {
        set_error("head rows of %d x %d values do not fit the peak-keep tile (%zu B of shared memory)", a.C, a.w, tile_smem);
        return SFA_ERR_UNSUPPORTED;
    }
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    a.counts = reinterpret_cast<uint32_t*>(ws + kDecodeHeaderBytes);
    a.cands = reinterpret_cast<unsigned long long*>(ws + kDecodeHeaderBytes + (((size_t)a.B * sizeof(uint32_t) + 255) / 256) * 256);
    a.vec4 = ((a.w & 3) == 0 && (reinterpret_cast<uintptr_t>(a.hm) & 15) == 0) ? 1 : 0;

    // the 48-KB default limit counts the kernel's static shared memory (plateau-cap bitmaps, ~11 KB) as well: opt in whenever
    // the sum could pass it, not only when the tile alone does
    if (tile_smem > 32 * 1024)
        SFA_CUDA_TRY(cudaFuncSetAttribute(peak_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
    dim3 cgrid((a.h + kCandRows - 1) / kCandRows, a.B);
    SFA_LAUNCH("peak_candidates", stream, peak_candidates_kernel<<<cgrid, kCandThreads, tile_smem, stream>>>(a));
    const size_t sel_smem = (size_t)kSelSmemItems * sizeof(unsigned long long);
    SFA_CUDA_TRY(cudaFuncSetAttribute(peak_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    SFA_LAUNCH("peak_select", stream, peak_select_kernel<<<a.B, kSelThreads, sel_smem, stream>>>(a));
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}

}  // namespace
}  // namespace sfa

using namespace sfa;

extern "C" size_t sfa_decode_workspace_bytes(int32_t B, int32_t C, int32_t h, int32_t w, int32_t K) {
    (void)K;
    if (B < 0 || C <= 0 || h <= 0 || w <= 0) {
        set_error("bad head shape B=%d C=%d h=%d w=%d", B, C, h, w);
        return 0;
    }
    return decode_workspace_bytes(B, C, h, w);
}

extern "C" int sfa_decode_workspace_init(void* workspace, size_t workspace_bytes, sfa_stream_t stream) {
    SFA_REQUIRE(workspace != nullptr || workspace_bytes == 0, "workspace is NULL");
    // only the per-frame candidate counters have to start at zero; clearing everything is simplest
    if (workspace_bytes) SFA_CUDA_TRY(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream));
    return SFA_OK;
}

extern "C" int sfa_decode(const float* hm, const float* cen_offset, const float* direction, const float* z_coor,
                          const float* dim, int32_t B, int32_t C, int32_t h, int32_t w, int32_t K, float* det,
                          int64_t* inds, int32_t apply_sigmoid, void* workspace, size_t workspace_bytes,
                          sfa_stream_t stream) {
    SFA_REQUIRE(B == 0 || (hm && direction && z_coor && dim && det), "NULL pointer argument");
    DecodeArgs a = {};
    a.hm = hm; a.off = cen_offset; a.dir = direction; a.zc = z_coor; a.dim = dim;
    a.B = B; a.C = C; a.h = h; a.w = w; a.K = K;
    a.do_nms = 1;
    a.apply_sigmoid = apply_sigmoid ? 1 : 0;
    a.det = det; a.inds = inds;
    return launch_decode(a, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sfa_decode_post(const float* hm, const float* cen_offset, const float* direction, const float* z_coor,
                               const float* dim, int32_t B, int32_t C, int32_t h, int32_t w, int32_t K, float* det,
                               int64_t* inds, int32_t apply_sigmoid, int32_t num_classes, float down_ratio,
                               float bound_size_y, float bev_width, float bound_size_x, float bev_height, float peak_thresh,
                               float min_x, float min_y, float min_z, float* rows, int32_t* cls, uint8_t* keep, float* real,
                               void* workspace, size_t workspace_bytes, sfa_stream_t stream) {
    SFA_REQUIRE(B == 0 || (hm && direction && z_coor && dim && det && rows && cls && keep), "NULL pointer argument");
    DecodeArgs a = {};
    a.hm = hm; a.off = cen_offset; a.dir = direction; a.zc = z_coor; a.dim = dim;
    a.B = B; a.C = C; a.h = h; a.w = w; a.K = K;
    a.do_nms = 1;
    a.apply_sigmoid = apply_sigmoid ? 1 : 0;
    a.det = det; a.inds = inds;
    a.pp_rows = rows; a.pp_cls = cls; a.pp_keep = keep; a.pp_real = real;
    a.pp_num_classes = num_classes; a.pp_down_ratio = down_ratio; a.pp_bsy = bound_size_y; a.pp_bev_w = bev_width;
    a.pp_bsx = bound_size_x; a.pp_bev_h = bev_height; a.pp_thresh = peak_thresh;
    a.pp_min_x = min_x; a.pp_min_y = min_y; a.pp_min_z = min_z;
    return launch_decode(a, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sfa_topk(const float* scores, int32_t B, int32_t C, int32_t h, int32_t w, int32_t K, float* score,
                        int64_t* inds, int32_t* clses, float* ys, float* xs, void* workspace, size_t workspace_bytes,
                        sfa_stream_t stream) {
    SFA_REQUIRE(B == 0 || (scores && score && inds && clses && ys && xs), "NULL pointer argument");
    DecodeArgs a = {};
    a.hm = scores;
    a.B = B; a.C = C; a.h = h; a.w = w; a.K = K;
    a.do_nms = 0;
    a.inds = inds; a.tk_score = score; a.tk_cls = clses; a.tk_ys = ys; a.tk_xs = xs;
    return launch_decode(a, workspace, workspace_bytes, (cudaStream_t)stream);
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
                                float peak_thresh, float min_x, float min_y, float min_z, float* out, int32_t* cls,
                                uint8_t* keep, float* real, sfa_stream_t stream) {
    SFA_REQUIRE(B >= 0 && K >= 0, "bad shape B=%d K=%d", B, K);
    int n = B * K;
    if (n == 0) return SFA_OK;
    SFA_REQUIRE(det && out && cls && keep, "NULL pointer argument");
    SFA_LAUNCH("post_process", (cudaStream_t)stream,
               post_process_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
                   det, n, num_classes, down_ratio, bound_size_y, bev_width, bound_size_x, bev_height, peak_thresh, min_x,
                   min_y, min_z, out, cls, keep, real));
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}

extern "C" int sfa_real_values(const float* rows, const int32_t* cls, int32_t n, float bound_size_y, float bev_width,
                               float bound_size_x, float bev_height, float min_x, float min_y, float min_z, float* real,
                               sfa_stream_t stream) {
    SFA_REQUIRE(n >= 0, "bad row count %d", n);
    if (n == 0) return SFA_OK;
    SFA_REQUIRE(rows && cls && real, "NULL pointer argument");
    SFA_LAUNCH("real_values", (cudaStream_t)stream,
               real_values_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
                   rows, cls, n, bound_size_y, bev_width, bound_size_x, bev_height, min_x, min_y, min_z, real));
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}
