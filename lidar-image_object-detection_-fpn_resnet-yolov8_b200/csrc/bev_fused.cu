// Stage A as ONE persistent kernel: sweep -> band buckets -> BEV planes, with the bucket hand-off inside L2.
//
// Replaces (reference, read-only at /root/reference):
//   get_filtered_lidar  data_process/kitti_data_utils.py:228-241
//   makeBEVMap          data_process/kitti_bev_utils.py:22-55
// Same arithmetic and the same two steps as the two-kernel tiled path of bev_rasterize.cu ("bin": multi-split
// of 1024-point tiles into the map's bands; "band": per-band reduction in shared memory whose three arrays
// leave as the fp32 planes through TMA bulk stores), but both steps are work items of one grid:
//
//   * 4 CTAs of 256 threads per SM pull TICKETS from a global counter.  Ticket order is
//       [bin tiles of frame s] [band items of frame s - LAG]   for s = 0, 1, ...
//     so reading sweeps (HBM read) and writing planes (HBM write) are in flight together on every SM, and
//     only RING (> LAG) frames of buckets are ever live: 8 x 1.9 MB for KITTI, written and re-read inside the
//     126 MB L2 and overwritten in place by the next frames — the records never need to reach HBM.
//   * a band item of frame f waits (thread 0, ld.acquire) until all of f's bin tiles have signalled
//     (threadfence + atomicAdd); a bin tile of frame f waits until the bands of frame f - RING are done with the
//     ring slot.  Every dependency points to a LOWER ticket and a CTA takes tickets only while it runs, so the
//     lowest unfinished ticket can always proceed: no deadlock whatever the number of resident CTAs.
//   * the next ticket is claimed one item ahead and its first loads (4 points or 4 records per thread) are
//     issued before the current item's tail, so the L2 / HBM latency of an item's head hides behind the
//     previous item's copy-out or TMA drain.
//   * the three planes are re-zeroed by the TMA as well (bulk copy of a zero block, mbarrier completion),
//     not by the threads.
#include "bev_common.cuh"

namespace sfa {
namespace {

constexpr int kFusedThreads = 256;
constexpr int kFusedPoints = 4;                               // points per thread of a bin tile
constexpr int kFusedTile = kFusedThreads * kFusedPoints;      // 1024
constexpr int kFusedBands = 128;
constexpr int kFusedCtasPerSm = 4;
constexpr int kFusedMaxRing = 32;
constexpr int kFusedRegRecords = 6;                           // records a thread keeps in registers across the phases
constexpr int kFusedSpecRecords = 4;                          // ... of which prefetched before the item starts
constexpr size_t kFusedStageBytes = (size_t)kFusedTile * sizeof(uint4);   // 16 KB
// control block inside the workspace header: every counter alone in a 128-B line
constexpr int kCtlLine = 32;                                   // uint32 per line
constexpr size_t kFusedCtlOffset = 256;
constexpr int kCtlTicket = 0;
constexpr int kCtlTilesDone = 1;                               // + ring slot
constexpr int kCtlBandsPre = 1 + kFusedMaxRing;
constexpr int kCtlBandsDone = 1 + 2 * kFusedMaxRing;
constexpr int kCtlOvf = 1 + 3 * kFusedMaxRing;
constexpr int kCtlTimeouts = 1 + 4 * kFusedMaxRing;
constexpr size_t kFusedCtlBytes = (size_t)(2 + 4 * kFusedMaxRing) * kCtlLine * sizeof(uint32_t);
constexpr size_t kFusedZerosOffset = 24576;
static_assert(kFusedCtlOffset + kFusedCtlBytes <= kFusedZerosOffset, "control block overlaps the zero block");
static_assert(kFusedZerosOffset + 3 * (size_t)kMaxCellsPerBand * sizeof(uint32_t) <= kHeaderBytes, "zero block does not fit the header");

struct FusedArgs {
    const float4* pts;
    const int64_t* offsets;
    int64_t max_points;
    int nf, tb, lag, ring;
    BevGeom g;
    BandPlan plan;
    uint32_t* ctl;
    uint32_t* cursors;        // [ring][nb] x kCursorStride
    BevRecord* buckets;       // [ring] slots of slot_recs records
    size_t slot_recs;
    uint32_t bucket_cap;
    const uint32_t* zeros;    // >= 3 * cpb zero words
    const float* lut;
    float* out;
    uint32_t* status;
    float inv_h;              // 1 / max_height (exact when max_height is a power of two: MUL_HEIGHT)
};

enum : int { kRoleBin = 0, kRoleBand = 1, kRoleDone = 2 };

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// Bounded: a dependency that does not resolve within ~1 s (a bug, never observed) is recorded in the control
// block's timeout word — sfa_bev_rasterize's caller sees wrong maps, not a hung GPU.
__device__ __forceinline__ void spin_until_ge(const uint32_t* p, uint32_t target, uint32_t* timeouts) {
    for (uint32_t spins = 0; ld_acquire_u32(p) < target; ++spins) {
        __nanosleep(128);
        if (spins > (1u << 22)) {
            atomicAdd(timeouts, 1u);
            return;
        }
    }
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared::cta bulk copy, completion (bytes) on an mbarrier of this CTA
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
}

// ticket -> (role, frame, index); see the file header for the order
__device__ __forceinline__ void decode_ticket(uint32_t t, const FusedArgs& a, int& role, int& frame, int& idx) {
    const uint32_t tb = (uint32_t)a.tb, nb = (uint32_t)a.plan.nb, nf = (uint32_t)a.nf;
    const uint32_t lag = min((uint32_t)a.lag, nf);
    const uint32_t n1 = lag * tb, per = tb + nb, n2 = (nf - lag) * per;
    if (t < n1) {
        role = kRoleBin; frame = (int)(t / tb); idx = (int)(t - (uint32_t)frame * tb);
    } else if (t < n1 + n2) {
        const uint32_t u = t - n1, s = u / per, r = u - s * per;
        if (r < tb) { role = kRoleBin; frame = (int)(lag + s); idx = (int)r; }
        else        { role = kRoleBand; frame = (int)s; idx = (int)(r - tb); }
    } else if (t < nf * per) {
        const uint32_t u = t - n1 - n2, s = u / nb;
        role = kRoleBand; frame = (int)(nf - lag + s); idx = (int)(u - s * nb);
    } else {
        role = kRoleDone; frame = 0; idx = 0;
    }
}

template <bool FILTER, bool RANGE_SAFE, bool MUL_HEIGHT>
__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm)
bev_fused_kernel(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(128) uint32_t fsm[];
    __shared__ uint32_t hist[kFusedBands];   // bin: points of this tile per band
    __shared__ uint32_t soff[kFusedBands];   // bin: exclusive scan of hist
    __shared__ uint32_t gpos[kFusedBands];   // bin: run's first position in the band's bucket minus its first stage slot
    __shared__ float lut[64];
    __shared__ int s_item[2][4];             // double-buffered work item: role, frame, idx, ready
    __shared__ __align__(8) unsigned long long zero_bar[2];   // [0] planes, [1] inv / stage region

    const int cpb = a.plan.cpb, nb = a.plan.nb;
    uint32_t* const inten = fsm;             // the three planes are contiguous: one zero-fill, and they leave as they are
    uint32_t* const zkey = fsm + cpb;        // phase 1: max orderable z -> final: height bits
    uint32_t* const cnt = fsm + 2 * cpb;     // phase 1: points in the cell -> final: density bits
    uint32_t* const inv = fsm + 3 * cpb;     // phase 2: max of ~index among the max-z points; idle state 0
    uint4* const stage = reinterpret_cast<uint4*>(inv);   // bin tile: records sorted by band (aliases inv)
    const int tid = threadIdx.x;
    auto height = [&](uint32_t zbits) -> float {   // kitti_bev_utils.py:44 (fp32 division)
        return MUL_HEIGHT ? __fmul_rn(__uint_as_float(zbits), a.inv_h) : __fdiv_rn(__uint_as_float(zbits), a.g.max_h);
    };
    auto ctl = [&](int what, int slot) { return a.ctl + (what + slot) * kCtlLine; };
    auto inv_words = [&]() { return (uint32_t)max((size_t)cpb, kFusedStageBytes / 4); };

    // ---- state carried from one loop iteration to the next ----
    // st bit 0: which s_item buffer holds the CURRENT item; 1: planes zero-fill not yet waited for; 2: same for the
    // inv / stage region; 3, 4: parity of the next completion of zero_bar[0] / [1]
    uint32_t st = 0;
    uint4 pre[kFusedSpecRecords];   // bin: the thread's 4 points; band: its first 4 records
    uint32_t aux = 0;               // bin: points in the tile; band: the cursor (record count), loaded with the records
    uint32_t n_oob_total = 0;

    // Thread 0: turn a claimed ticket into work item `buf` (+ whether a band item's records are already complete).
    auto publish = [&](uint32_t t, int buf) {
        int r, f, i, rdy = 1;
        decode_ticket(t, a, r, f, i);
        if (r == kRoleBand)
            rdy = ld_acquire_u32(ctl(kCtlTilesDone, f % a.ring)) >= (uint32_t)(f / a.ring + 1) * (uint32_t)a.tb;
        s_item[buf][0] = r; s_item[buf][1] = f; s_item[buf][2] = i; s_item[buf][3] = rdy;
    };
    auto bucket_of = [&](int f, int b) -> const BevRecord* {
        return a.buckets + (size_t)(f % a.ring) * a.slot_recs + (size_t)b * a.bucket_cap;
    };
    auto cursor_of = [&](int f, int b) -> uint32_t* {
        return a.cursors + ((size_t)(f % a.ring) * nb + b) * kCursorStride;
    };
    auto load_band_head = [&](int f, int b) {   // cursor + the first records of every thread (a bucket is bucket_cap records of mapped memory)
        const BevRecord* rec = bucket_of(f, b);
        aux = ld_cg_u32(cursor_of(f, b));
#pragma unroll
        for (int j = 0; j < kFusedSpecRecords; ++j) {
            const uint32_t i = tid + j * kFusedThreads;
            pre[j] = (i < a.bucket_cap) ? ld_record(rec + i) : make_uint4(0, 0, 0, 0);
        }
    };
    // Issue the first loads of work item `buf` (valid after a barrier that follows its publish).
    auto prefetch_item = [&](int buf) {
        const int r = s_item[buf][0], f = s_item[buf][1], i = s_item[buf][2];
        if (r == kRoleBin) {
            int64_t start, n;
            sweep_range(a.offsets, f, a.max_points, start, n);
            const int64_t first = (int64_t)i * kFusedTile;
            const int n_tile = (int)max((int64_t)0, min((int64_t)kFusedTile, n - first));
            aux = (uint32_t)n_tile;
            const float4* tile = a.pts + start + first;
            const unsigned long long pol = l2_evict_first_policy();
#pragma unroll
            for (int j = 0; j < kFusedPoints; ++j) {
                if (tid + kFusedThreads * j < n_tile) {
                    const float4 p = ld_stream_f4_evict_first(tile + tid + kFusedThreads * j, pol);
                    pre[j] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(p.w));
                }
            }
        } else if (r == kRoleBand && s_item[buf][3]) {
            load_band_head(f, i);
        }
    };
    auto wait_zero = [&](int which) {   // which: 0 planes, 1 inv / stage region
        if (st & (2u << which)) {
            mbar_wait(&zero_bar[which], (st >> (3 + which)) & 1u);
            st ^= 8u << which;
            st &= ~(2u << which);
        }
    };

    // ---- prologue ----
    if (tid == 0) {
        mbar_init(&zero_bar[0], 1);
        mbar_init(&zero_bar[1], 1);
        publish(atomicAdd(a.ctl + kCtlTicket * kCtlLine, 1u), 0);
    }
    if (tid < 64) lut[tid] = a.lut[tid];
    for (uint32_t i = tid; i < (3u * cpb + inv_words()) / 4; i += kFusedThreads) reinterpret_cast<uint4*>(fsm)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();   // mbarrier init + the generic-proxy zero fill, before any bulk copy touches them
    __syncthreads();
    prefetch_item(0);

    while (true) {
        const int cur = (int)(st & 1u);
        const int role = s_item[cur][0];
        if (role == kRoleDone) break;
        const int frame = s_item[cur][1], idx = s_item[cur][2];
        __syncthreads();   // the previous item's shared state is dead; s_item[cur ^ 1] may be overwritten
        uint32_t t_next = 0;
        if (tid == 0) t_next = atomicAdd(a.ctl + kCtlTicket * kCtlLine, 1u);   // consumed by publish() further down
        const int slot = frame % a.ring;
        const uint32_t use = (uint32_t)(frame / a.ring);   // how many frames used this ring slot before

        if (role == kRoleBin) {
            // =============================== bin tile (frame, idx) ===============================
            const int lane = tid & 31, warp = tid >> 5;
            const int n_tile = (int)aux;
            if (tid < kFusedBands) hist[tid] = 0;
            wait_zero(1);   // a zero-fill of the staging area may still be landing
            if (tid == 0 && use > 0) spin_until_ge(ctl(kCtlBandsDone, slot), use * (uint32_t)nb, ctl(kCtlTimeouts, 0));   // ring slot free again
            __syncthreads();
            // packed per point: band << 24 | rank-in-(tile, band);  0xFFFFFFFF = dropped
            uint32_t packed[kFusedPoints], local[kFusedPoints];
            {
                uint32_t n_oob = 0;
                const ExactDivisor dv = make_divisor(a.g.d);
#pragma unroll
                for (int j = 0; j < kFusedPoints; ++j) {
                    const float4 p = make_float4(__uint_as_float(pre[j].x), __uint_as_float(pre[j].y), __uint_as_float(pre[j].z),
                                                 __uint_as_float(pre[j].w));
                    float z;
                    bool oob = false;
                    int cell = point_to_cell_fast<FILTER, RANGE_SAFE>(p, a.g, dv, z, oob);
                    if (tid + kFusedThreads * j >= n_tile) { cell = -1; oob = false; }
                    n_oob += oob ? 1u : 0u;
                    const uint32_t b = band_of((uint32_t)max(cell, 0), a.plan);
                    local[j] = (b << 16) | ((uint32_t)max(cell, 0) - b * (uint32_t)cpb);
                    packed[j] = 0xFFFFFFFFu;
                    if (cell >= 0) packed[j] = (b << 24) | atomicAdd(&hist[b], 1u);
                    pre[j].z = __float_as_uint(z);
                }
                if (!RANGE_SAFE) n_oob_total += n_oob;
            }
            __syncthreads();
            // warp 0: exclusive scan over the bands -> stage slots, and the reservation of the global runs (atomics only
            // ISSUED here; their results are consumed after the staging)
            uint32_t res[kFusedBands / 32], slot0[kFusedBands / 32];
            if (warp == 0) {
                uint32_t c[kFusedBands / 32], run = 0;
#pragma unroll
                for (int q = 0; q < kFusedBands / 32; ++q) {   // lane owns bands 4*lane .. 4*lane+3
                    const int b = lane * (kFusedBands / 32) + q;
                    c[q] = b < nb ? hist[b] : 0u;
                    run += c[q];
                }
                uint32_t incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += v;
                }
                uint32_t at = incl - run;
                uint32_t* cursors = a.cursors + (size_t)slot * nb * kCursorStride;
#pragma unroll
                for (int q = 0; q < kFusedBands / 32; ++q) {
                    const int b = lane * (kFusedBands / 32) + q;
                    soff[b] = at;
                    slot0[q] = at;
                    at += c[q];
                    res[q] = c[q] ? atomicAdd(cursors + (size_t)b * kCursorStride, c[q]) : 0u;
                }
            }
            __syncthreads();
            {
                const uint32_t i0 = (uint32_t)idx * kFusedTile + tid;
#pragma unroll
                for (int j = 0; j < kFusedPoints; ++j) {
                    if (packed[j] != 0xFFFFFFFFu) {
                        const uint32_t s = soff[packed[j] >> 24] + (packed[j] & 0xFFFFFFu);
                        stage[s] = make_uint4(pre[j].z, pre[j].w, i0 + kFusedThreads * j, local[j]);
                    }
                }
            }
            if (warp == 0) {
#pragma unroll
                for (int q = 0; q < kFusedBands / 32; ++q) gpos[lane * (kFusedBands / 32) + q] = res[q] - slot0[q];
            }
            const int n_kept = (int)(soff[nb - 1] + hist[nb - 1]);   // both final since the barrier above
            if (tid == 0) publish(t_next, cur ^ 1);
            __syncthreads();
            prefetch_item(cur ^ 1);   // the next item's first loads fly during the copy-out
            // copy out in sorted order: stage slot s belongs to band (stage[s].w >> 16), record s - soff[band] of its run
            BevRecord* fb = a.buckets + (size_t)slot * a.slot_recs;
            for (int s0 = tid; s0 < n_kept; s0 += kFusedThreads) {
                uint4 r = stage[s0];
                const uint32_t b = r.w >> 16;
                const uint32_t pos = gpos[b] + (uint32_t)s0;
                if (pos < a.bucket_cap) {
                    r.w &= 0xFFFFu;
                    *reinterpret_cast<uint4*>(fb + (size_t)b * a.bucket_cap + pos) = r;
                } else {   // the band's bucket is full: frame overflow list, record keeps its band tag
                    BevRecord* ovf = fb + (size_t)nb * a.bucket_cap;
                    *reinterpret_cast<uint4*>(ovf + atomicAdd(ctl(kCtlOvf, slot), 1u)) = r;
                }
            }
            __syncthreads();   // all records of the tile are stored (and the staging area is free)
            const bool band_next = s_item[cur ^ 1][0] == kRoleBand;
            if (tid == 0) {
                __threadfence();                      // release: the CTA's record stores, then the signal
                atomicAdd(ctl(kCtlTilesDone, slot), 1u);
                if (band_next) {   // a band item needs `inv` idle (zero) again
                    fence_proxy_async_smem();
                    mbar_expect_tx(&zero_bar[1], inv_words() * 4u);
                    bulk_load_g2s(inv, a.zeros, inv_words() * 4u, &zero_bar[1]);
                }
            }
            if (band_next) st |= 4u;
        } else {
            // =============================== band item (frame, idx) ===============================
            const BevRecord* rec = bucket_of(frame, idx);
            const bool ready = s_item[cur][3] != 0;
            if (tid == 0 && !ready) spin_until_ge(ctl(kCtlTilesDone, slot), (use + 1u) * (uint32_t)a.tb, ctl(kCtlTimeouts, 0));
            wait_zero(0);
            wait_zero(1);
            if (!ready) {
                __syncthreads();
                load_band_head(frame, idx);
            }
            const uint32_t n_all = aux;                        // records of this band, in its bucket or overflowed
            const uint32_t n_rec = min(n_all, a.bucket_cap);   // ... of which in the bucket
            const bool overflowed = n_all > a.bucket_cap;

            if (!overflowed && n_rec <= (uint32_t)(kFusedRegRecords * kFusedThreads)) {
                uint4 r[kFusedRegRecords];
#pragma unroll
                for (int j = 0; j < kFusedSpecRecords; ++j) r[j] = pre[j];
#pragma unroll
                for (int j = kFusedSpecRecords; j < kFusedRegRecords; ++j) {
                    const uint32_t i = tid + j * kFusedThreads;
                    if (i < n_rec) r[j] = ld_record(rec + i);
                }
#pragma unroll
                for (int j = 0; j < kFusedRegRecords; ++j) {
                    const uint32_t i = tid + j * kFusedThreads;
                    if (i < n_rec) {
                        atomicMax(&zkey[r[j].w], orderable_u32(__uint_as_float(r[j].x), 0u));   // NaN z sorts last (key 0)
                        atomicAdd(&cnt[r[j].w], 1u);
                    }
                }
                __syncthreads();
                // A record alone in its cell is the winner: it writes the cell's final values at once.  Cells with
                // several records vote on the lowest index among their highest-z records.
                uint32_t multi = 0;
#pragma unroll
                for (int j = 0; j < kFusedRegRecords; ++j) {
                    const uint32_t i = tid + j * kFusedThreads;
                    if (i < n_rec) {
                        const uint32_t cell = r[j].w;
                        const uint32_t c = cnt[cell];
                        if (c == 1u) {
                            inten[cell] = r[j].y;                              // kitti_bev_utils.py:47
                            zkey[cell] = __float_as_uint(height(r[j].x));      // :44
                            cnt[cell] = __float_as_uint(lut[1]);               // :46,48
                        } else if (zkey[cell] == orderable_u32(__uint_as_float(r[j].x), 0u)) {
                            atomicMax(&inv[cell], 0xFFFFFFFFu - r[j].z);
                            multi |= 1u << j;
                        }
                    }
                }
                if (tid == 0) publish(t_next, cur ^ 1);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < kFusedRegRecords; ++j) {
                    if ((multi >> j) & 1u) {
                        const uint32_t cell = r[j].w;
                        if (inv[cell] == 0xFFFFFFFFu - r[j].z) {
                            inten[cell] = r[j].y;
                            zkey[cell] = __float_as_uint(height(r[j].x));
                            cnt[cell] = __float_as_uint(lut[min(cnt[cell], 63u)]);
                            inv[cell] = 0;   // back to its idle state
                        }
                    }
                }
            } else {
                // crowded band: records streamed from L2 once per phase (bucket, and the frame's overflow list if needed)
                const BevRecord* ovf = a.buckets + (size_t)slot * a.slot_recs + (size_t)nb * a.bucket_cap;
                const uint32_t n_ovf = overflowed ? ld_cg_u32(ctl(kCtlOvf, slot)) : 0u;
                if (tid == 0) publish(t_next, cur ^ 1);
                band_stream_reduce<MUL_HEIGHT>(zkey, inv, cnt, inten, lut, rec, n_rec, ovf, n_ovf, (uint32_t)idx, a.g.max_h);
            }
            fence_proxy_async_smem();   // this thread's st.shared / atom.shared -> visible to the async proxy (TMA) ...
            __syncthreads();            // ... and ordered before the bulk stores thread 0 issues below
            prefetch_item(cur ^ 1);     // lands during the drain
            if (tid == 0) {
                *reinterpret_cast<volatile uint32_t*>(cursor_of(frame, idx)) = 0;   // cursor ready for the slot's next frame
                // the last band of the frame to get here (all others have read the overflow counter already) resets it
                if (atomicAdd(ctl(kCtlBandsPre, slot), 1u) + 1u == (use + 1u) * (uint32_t)nb)
                    *reinterpret_cast<volatile uint32_t*>(ctl(kCtlOvf, slot)) = 0;
                __threadfence();
                atomicAdd(ctl(kCtlBandsDone, slot), 1u);
                // the three shared arrays ARE the band's planes (empty cells kept their zero fill; channel 0
                // intensity, 1 height, 2 density, kitti_bev_utils.py:50-53)
                const size_t cells = (size_t)a.g.H * a.g.W;
                const size_t cell0 = (size_t)idx * cpb;
                const uint32_t bytes = (uint32_t)(min((size_t)cpb, cells - cell0) * sizeof(float));
                float* const o = a.out + (size_t)frame * 3 * cells + cell0;
                const unsigned long long pol = l2_evict_first_policy();
                bulk_store_s2g_hint(o, inten, bytes, pol);
                bulk_store_s2g_hint(o + cells, zkey, bytes, pol);
                bulk_store_s2g_hint(o + 2 * cells, cnt, bytes, pol);
                bulk_commit_group();
                bulk_wait_group_read0();    // the planes have left shared memory ...
                mbar_expect_tx(&zero_bar[0], 3u * (uint32_t)cpb * 4u);
                bulk_load_g2s(inten, a.zeros, 3u * (uint32_t)cpb * 4u, &zero_bar[0]);   // ... and are zero-filled for the next band item
            }
            st |= 2u;
        }
        st ^= 1u;
    }
    if (tid == 0) bulk_wait_group0();   // all plane stores performed before the CTA retires
    // zero-fills still in flight target this CTA's shared memory: wait for them before it is released
    wait_zero(0);
    wait_zero(1);
    if (!RANGE_SAFE && a.status) {
        n_oob_total = __reduce_add_sync(0xFFFFFFFFu, n_oob_total);
        if ((tid & 31) == 0 && n_oob_total) atomicAdd(a.status, n_oob_total);
    }
}

inline int fused_env(const char* name, int dflt, int lo, int hi) { return env_int(name, dflt, lo, hi); }

}  // namespace

// Whether a geometry with this band plan can run on the fused kernel.
bool fused_supported(const BandPlan& plan) { return plan.nb <= kFusedBands && plan.cpb <= kMaxCellsPerBand; }
// What SFA_BEV_AUTO picks among the two tiled schedules (SFA_BEV_FUSED=0/1 overrides).
bool fused_is_default() {
    static const int enabled = env_int("SFA_BEV_FUSED", 0, 0, 1);
    return enabled != 0;
}

// Enqueue all B frames as ONE launch.  Workspace layout as the two-kernel tiled path (header | cursors | ring slots).
int fused_launch(const float* pts, const int64_t* offsets, int B, int64_t max_points, const SfaBevParams* p,
                 const BandPlan& plan, const float* lut, float* out, uint32_t* status, unsigned char* ws_base,
                 uint32_t* cursors, BevRecord* buckets, size_t slot_recs, uint32_t bucket_cap, int ring_avail,
                 cudaStream_t stream) {
    static const int ring_want = fused_env("SFA_BEV_FUSED_RING", 8, 2, kFusedMaxRing);
    static const int lag_want = fused_env("SFA_BEV_FUSED_LAG", 4, 1, kFusedMaxRing - 1);
    FusedArgs a;
    a.pts = reinterpret_cast<const float4*>(pts);
    a.offsets = offsets;
    a.max_points = max_points;
    a.nf = B;
    a.tb = (int)((max_points + kFusedTile - 1) / kFusedTile);
    a.ring = ring_want < ring_avail ? ring_want : ring_avail;
    if (a.ring > B) a.ring = B;
    a.lag = lag_want < a.ring ? lag_want : a.ring - 1;
    if (a.lag < 1) a.lag = 1;      // ring == 1 (B == 1): bands of frame 0 follow its tiles, nothing is reused
    a.g = make_geom(p);
    a.plan = plan;
    a.ctl = reinterpret_cast<uint32_t*>(ws_base + kFusedCtlOffset);
    a.cursors = cursors;
    a.buckets = buckets;
    a.slot_recs = slot_recs;
    a.bucket_cap = bucket_cap;
    a.zeros = reinterpret_cast<const uint32_t*>(ws_base + kFusedZerosOffset);
    a.lut = lut;
    a.out = out;
    a.status = status;
    a.inv_h = 1.0f / a.g.max_h;
    SFA_CUDA_TRY(cudaMemsetAsync(a.ctl, 0, kFusedCtlBytes, stream));
    const long long items = (long long)B * (a.tb + plan.nb);
    const int ctas = (int)(items < (long long)kFusedCtasPerSm * kNumSMs ? items : (long long)kFusedCtasPerSm * kNumSMs);
    const size_t inv_bytes = (size_t)plan.cpb * 4 > kFusedStageBytes ? (size_t)plan.cpb * 4 : kFusedStageBytes;
    const size_t smem = 3 * (size_t)plan.cpb * 4 + inv_bytes;
    const size_t max_smem = 3 * (size_t)kMaxCellsPerBand * 4 + kFusedStageBytes;
    int exp2 = 0;
    const float mant = frexpf(fabsf(a.g.max_h), &exp2);
    const bool mul_height = (mant == 0.5f) && exp2 > -120 && exp2 < 120 && a.g.max_h > 0.0f;
    const bool safe = p->apply_filter && filter_keeps_points_inside_map(a.g);
#define SFA_FUSED_LAUNCH(F, S, M)                                                                                     \
    do {                                                                                                              \
        SFA_CUDA_TRY(cudaFuncSetAttribute(bev_fused_kernel<F, S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem)); \
        SFA_LAUNCH("bev_fused", stream, bev_fused_kernel<F, S, M><<<ctas, kFusedThreads, smem, stream>>>(a));         \
    } while (0)
    if (p->apply_filter && safe) { if (mul_height) SFA_FUSED_LAUNCH(true, true, true); else SFA_FUSED_LAUNCH(true, true, false); }
    else if (p->apply_filter)    { if (mul_height) SFA_FUSED_LAUNCH(true, false, true); else SFA_FUSED_LAUNCH(true, false, false); }
    else                         { if (mul_height) SFA_FUSED_LAUNCH(false, false, true); else SFA_FUSED_LAUNCH(false, false, false); }
#undef SFA_FUSED_LAUNCH
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}

}  // namespace sfa
