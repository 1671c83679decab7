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

#ifndef SFA_FUSED_WORKERS
#define SFA_FUSED_WORKERS 224
#endif
constexpr int kFusedWorkers = SFA_FUSED_WORKERS;              // threads that do the arithmetic (multiple of 32, >= 128)
#ifndef SFA_FUSED_CTAS
#define SFA_FUSED_CTAS 4
#endif
#ifndef SFA_FUSED_NB
#define SFA_FUSED_NB 128
#endif
constexpr int kFusedThreads = kFusedWorkers + 32;             // + the service warp (one active thread)
constexpr int kFusedPlanBands = SFA_FUSED_NB;                 // bands per frame of the fused schedule (<= kFusedBands)
// cells per band the shared memory of kFusedCtasPerSm CTAs can hold next to the staging area
constexpr int kFusedMaxCellsPerBand = SFA_FUSED_CTAS >= 4 ? kMaxCellsPerBand : 2 * kMaxCellsPerBand;
constexpr int kFusedPoints = 4;                               // points per worker of a bin tile
constexpr int kFusedTile = kFusedWorkers * kFusedPoints;      // 1024
constexpr int kFusedBands = 128;
constexpr int kFusedCtasPerSm = SFA_FUSED_CTAS;
constexpr int kFusedMaxRing = 32;
constexpr int kFusedRegRecords = 6;                           // records a thread keeps in registers across the phases
constexpr int kFusedSpecRecords = 4;                          // ... of which prefetched before the item starts
constexpr size_t kFusedStageBytes = (size_t)kFusedTile * sizeof(uint4);   // 16 KB
// control block inside the workspace header: every counter alone in a 128-B line
constexpr int kCtlLine = 32;                                   // uint32 per line
constexpr size_t kFusedCtlOffset = 2048;   // behind the two-kernel schedule's overflow-counter banks
constexpr int kCtlTicket = 0;
constexpr int kCtlTilesDone = 1;                               // + ring slot
constexpr int kCtlBandsDone = 1 + kFusedMaxRing;
constexpr int kCtlOvf = 1 + 2 * kFusedMaxRing;                 // two per ring slot (+ kFusedMaxRing for odd uses)
constexpr int kCtlTimeouts = 1 + 4 * kFusedMaxRing;
constexpr size_t kFusedCtlBytes = (size_t)(2 + 4 * kFusedMaxRing) * kCtlLine * sizeof(uint32_t);
constexpr size_t kFusedZerosOffset = kZerosOffset;
static_assert(kFusedWorkers >= kFusedBands && kFusedWorkers % 32 == 0, "one scan thread per band");
static_assert(kFusedCtlOffset + kFusedCtlBytes <= kFusedZerosOffset, "control block overlaps the zero block");
static_assert(kFusedZerosOffset + 3 * (size_t)kFusedMaxCellsPerBand * sizeof(uint32_t) <= kHeaderBytes, "zero block does not fit the header");
static_assert(kFusedPlanBands <= kFusedBands, "one scan thread per band");

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
// Ticket t = (step s = t / P, k = t % P), P = max(tiles per frame, bands per frame): the bin tile k of frame s (if any)
// and then the band k of frame s - LAG (if any).  A CTA therefore alternates bin tile / band item: the TMA drain and
// zero-fill of a band's planes hide behind the next bin tile.
struct Work { int role, frame, idx; };
__device__ __forceinline__ int decode_ticket(uint32_t t, const FusedArgs& a, Work* w) {
    const uint32_t tb = (uint32_t)a.tb, nb = (uint32_t)a.plan.nb, nf = (uint32_t)a.nf;
    const uint32_t per = max(tb, nb), lag = (uint32_t)a.lag;
    const uint32_t s = t / per, k = t - s * per;
    int n = 0;
    if (s >= nf + lag) {
        w[0].role = kRoleDone; w[0].frame = 0; w[0].idx = 0;
        return 1;
    }
    if (s < nf && k < tb) { w[n].role = kRoleBin; w[n].frame = (int)s; w[n].idx = (int)k; ++n; }
    if (s >= lag && k < nb) { w[n].role = kRoleBand; w[n].frame = (int)(s - lag); w[n].idx = (int)k; ++n; }
    return n;   // may be 0: the ticket holds nothing (first / last LAG steps of the shorter kind)
}

struct __align__(16) FusedItem {   // what the service thread publishes for the workers
    int role;
    int idx;                    // bin: tile index inside its sweep; band: band index
    uint32_t aux;               // bin: points in the tile
    int slot;                   // ring slot of the item's frame
    unsigned long long ptr;     // bin: first point of the tile; band: the band's bucket (its cursor: slot, idx)
    unsigned long long out;     // the frame's overflow counter (control block)
};

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kFusedWorkers) : "memory"); }
// The kernel.  Threads 0 .. kFusedWorkers-1 are WORKERS (all the arithmetic); lane 0 of the last warp is the SERVICE
// thread: it claims tickets, waits for an item's dependencies, publishes the item (pointers, counts) one item ahead of
// the workers, and after the workers are done with an item performs its tail — release fence + completion signal,
// the TMA stores of a band's planes, the wait for the drain, and the TMA zero-fills — so none of those round trips
// sits between two items of the workers.
//   item_full[b]  service -> workers: s_item[b] is valid            (count 1)
//   tail_bar[b]   workers -> service: the item of buffer b is done   (count kFusedWorkers)
//   zero_bar      TMA -> workers: the planes are zero again            (transaction bytes)
template <bool FILTER, bool RANGE_SAFE, bool MUL_HEIGHT>
__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm)
bev_fused_kernel(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(128) uint32_t fsm[];
    __shared__ uint32_t hist[kFusedBands];   // bin: points of this tile per band (zero between tiles)
    __shared__ uint32_t soff[kFusedBands];   // bin: exclusive scan of hist
    __shared__ uint32_t gpos[kFusedBands];   // bin: run's first position in the band's bucket minus its first stage slot
    __shared__ float lut[64];
    __shared__ uint32_t s_nkept, wsum[kFusedBands / 32];
    __shared__ FusedItem s_item[2];
    __shared__ __align__(8) unsigned long long item_full[2], tail_bar[2], zero_bar;

    const int cpb = a.plan.cpb, nb = a.plan.nb;
    uint32_t* const inten = fsm;             // the three planes are contiguous: one zero-fill, and they leave as they are
    uint32_t* const zkey = fsm + cpb;        // phase 1: max orderable z -> final: height bits
    uint32_t* const cnt = fsm + 2 * cpb;     // phase 1: points in the cell -> final: density bits
    uint32_t* const inv = fsm + 3 * cpb;     // phase 2: max of ~index among the max-z points; idle state 0
    uint4* const stage = reinterpret_cast<uint4*>(inv);   // bin tile: records sorted by band (aliases inv)
    const int tid = threadIdx.x;
    const uint32_t inv_bytes = (uint32_t)max((size_t)cpb * 4, kFusedStageBytes);

    // ---- prologue (all threads) ----
    if (tid == 0) {
        mbar_init(&item_full[0], 1); mbar_init(&item_full[1], 1);
        mbar_init(&tail_bar[0], kFusedWorkers); mbar_init(&tail_bar[1], kFusedWorkers);
        mbar_init(&zero_bar, 1);
    }
    if (tid < 64) lut[tid] = a.lut[tid];
    if (tid < kFusedBands) hist[tid] = 0;
    for (uint32_t i = tid; i < (3u * cpb * 4u + inv_bytes) / 16; i += kFusedThreads) reinterpret_cast<uint4*>(fsm)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();   // mbarrier init + the generic-proxy zero fill, before any bulk copy touches them
    __syncthreads();

    if (tid >= kFusedWorkers) {
        // =========================================== SERVICE thread ===========================================
        if (tid != kFusedWorkers) return;
        auto ctl = [&](int what, int slot) { return a.ctl + (what + slot) * kCtlLine; };
        // tickets are claimed one ahead (the atomic's round trip overlaps the current iteration); a ticket yields 0-2 items
        Work fifo[2];
        int fifo_n = 0, fifo_at = 0;
        uint32_t t_ahead = atomicAdd(a.ctl + kCtlTicket * kCtlLine, 1u);
        auto claim = [&]() {
            while (fifo_at == fifo_n) {   // (an empty ticket: take the next one at once)
                fifo_n = decode_ticket(t_ahead, a, fifo);
                fifo_at = 0;
                if (fifo_n == 0) t_ahead = atomicAdd(a.ctl + kCtlTicket * kCtlLine, 1u);
            }
            const Work w = fifo[fifo_at++];
            // the next ticket is claimed when the last item of this one is handed out: its round trip hides behind that item
            if (fifo_at == fifo_n && w.role != kRoleDone) t_ahead = atomicAdd(a.ctl + kCtlTicket * kCtlLine, 1u);
            return w;
        };
        auto deps_ready = [&](const Work& w) -> bool {
            const int slot = w.frame % a.ring;
            const uint32_t use = (uint32_t)(w.frame / a.ring);
            if (w.role == kRoleBin) return use == 0 || ld_acquire_u32(ctl(kCtlBandsDone, slot)) >= use * (uint32_t)nb;
            if (w.role == kRoleBand) return ld_acquire_u32(ctl(kCtlTilesDone, slot)) >= (use + 1u) * (uint32_t)a.tb;
            return true;
        };
        auto wait_deps = [&](const Work& w) {
            for (uint32_t spins = 0; !deps_ready(w); ++spins) {
                __nanosleep(100);
                if (spins > (1u << 22)) {   // ~1 s: a bug, never observed; recorded instead of hanging the GPU
                    atomicAdd(ctl(kCtlTimeouts, 0), 1u);
                    return;
                }
            }
        };
        const size_t cells = (size_t)a.g.H * a.g.W;
        auto publish = [&](const Work& w, int buf) {
            FusedItem it;
            it.role = w.role; it.idx = w.idx; it.aux = 0; it.slot = w.frame % a.ring; it.ptr = 0; it.out = 0;
            if (w.role == kRoleBin) {
                // overflow counters alternate with the use parity of the ring slot; tile 0 of use u zeroes the one use
                // u + 1 will append to (its previous user, use u - 1, is done: that was this tile's dependency), and the
                // store is released by this tile's own completion signal, long before any tile of use u + 1 starts
                if (w.idx == 0) *reinterpret_cast<volatile uint32_t*>(ctl(kCtlOvf + kFusedMaxRing * (int)(((w.frame / a.ring) + 1) & 1), it.slot)) = 0;
                it.out = (unsigned long long)ctl(kCtlOvf + kFusedMaxRing * (int)((w.frame / a.ring) & 1), it.slot);
                int64_t start, n;
                sweep_range(a.offsets, w.frame, a.max_points, start, n);
                const int64_t first = (int64_t)w.idx * kFusedTile;
                it.aux = (uint32_t)max((int64_t)0, min((int64_t)kFusedTile, n - first));
                it.ptr = (unsigned long long)(a.pts + start + first);
            } else if (w.role == kRoleBand) {
                it.ptr = (unsigned long long)(a.buckets + (size_t)it.slot * a.slot_recs + (size_t)w.idx * a.bucket_cap);
                it.out = (unsigned long long)ctl(kCtlOvf + kFusedMaxRing * (int)((w.frame / a.ring) & 1), it.slot);
            }
            s_item[buf] = it;
            mbar_arrive(&item_full[buf]);   // release: the item, then the flag
        };
        uint32_t plane_fills = 0;
        auto tail = [&](const Work& w) {
            const int slot = w.frame % a.ring;
            if (w.role == kRoleBin) {
                __threadfence();                      // release: the workers' record stores, then the signal
                atomicAdd(ctl(kCtlTilesDone, slot), 1u);
            } else {
                // the three shared arrays ARE the band's planes (empty cells kept their zero fill; channel 0
                // intensity, 1 height, 2 density, kitti_bev_utils.py:50-53)
                const size_t cell0 = (size_t)w.idx * cpb;
                const uint32_t bytes = (uint32_t)(min((size_t)cpb, cells - cell0) * sizeof(float));
                float* const o = a.out + (size_t)w.frame * 3 * cells + cell0;
                const unsigned long long pol = l2_evict_first_policy();
                bulk_store_s2g_hint(o, inten, bytes, pol);
                bulk_store_s2g_hint(o + cells, zkey, bytes, pol);
                bulk_store_s2g_hint(o + 2 * cells, cnt, bytes, pol);
                bulk_commit_group();
                // while the TMA drains the planes: cursor back to zero for the slot's next frame, then the completion signal
                *reinterpret_cast<volatile uint32_t*>(a.cursors + ((size_t)slot * nb + w.idx) * kCursorStride) = 0;
                __threadfence();
                atomicAdd(ctl(kCtlBandsDone, slot), 1u);
                bulk_wait_group_read0();    // the planes have left shared memory ...
                mbar_expect_tx(&zero_bar, 3u * (uint32_t)cpb * 4u);
                bulk_load_g2s(inten, a.zeros, 3u * (uint32_t)cpb * 4u, &zero_bar);   // ... and are zero-filled for the next band item
                ++plane_fills;
            }
        };
        Work cur = claim();
        wait_deps(cur);
        publish(cur, 0);
        for (uint32_t i = 0; cur.role != kRoleDone; ++i) {
            const Work nxt = claim();
            // publish the next item EARLY when its dependencies already hold (buffer (i+1)&1 is free: the workers arrived
            // for item i-1 before the previous iteration's tail); never spin before this CTA's own tail is out
            bool published = false;
            if (deps_ready(nxt)) { publish(nxt, (int)((i + 1) & 1u)); published = true; }
            mbar_wait_relaxed(&tail_bar[i & 1u], (i >> 1) & 1u);
            tail(cur);
            if (!published) { wait_deps(nxt); publish(nxt, (int)((i + 1) & 1u)); }
            cur = nxt;
        }
        bulk_wait_group0();   // all plane stores performed before the CTA retires
        if (plane_fills) mbar_wait(&zero_bar, (plane_fills - 1u) & 1u);   // the last zero-fill targets this CTA's shared memory
        return;
    }

    // ================================================ WORKERS ================================================
    auto height = [&](uint32_t zbits) -> float {   // kitti_bev_utils.py:44 (fp32 division)
        return MUL_HEIGHT ? __fmul_rn(__uint_as_float(zbits), a.inv_h) : __fdiv_rn(__uint_as_float(zbits), a.g.max_h);
    };
    // st bit 1: a planes zero-fill has not been waited for; bit 3: parity of the next completion of zero_bar;
    // bit 5: this thread has issued the current item's first loads already
    uint32_t st = 0;
    uint4 pre[kFusedSpecRecords];   // bin: the thread's 4 points; band: its first 4 records
    uint32_t n_all_pre = 0;         // band: the record count, loaded with them
    uint32_t n_oob_total = 0;
    auto wait_zero = [&]() {
        if (st & 2u) {
            mbar_wait(&zero_bar, (st >> 3) & 1u);
            st ^= 8u;
            st &= ~2u;
        }
    };
    auto first_loads = [&](const FusedItem& it) {
        if (it.role == kRoleBin) {
            const float4* tile = reinterpret_cast<const float4*>(it.ptr);
            const unsigned long long pol = l2_evict_first_policy();
#pragma unroll
            for (int j = 0; j < kFusedPoints; ++j) {
                if (tid + kFusedWorkers * j < (int)it.aux) {
                    const float4 p = ld_stream_f4_evict_first(tile + tid + kFusedWorkers * j, pol);
                    pre[j] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(p.w));
                }
            }
        } else if (it.role == kRoleBand) {   // the cursor (final: all the frame's tiles have signalled) and, before its value
            // is known, the first records of every thread (a bucket is bucket_cap records of mapped memory)
            const BevRecord* rec = reinterpret_cast<const BevRecord*>(it.ptr);
            n_all_pre = ld_cg_u32(a.cursors + ((size_t)it.slot * nb + it.idx) * kCursorStride);
#pragma unroll
            for (int j = 0; j < kFusedSpecRecords; ++j) {
                const uint32_t i = tid + j * kFusedWorkers;
                pre[j] = (i < a.bucket_cap) ? ld_record(rec + i) : make_uint4(0, 0, 0, 0);
            }
        }
    };
    // the next item's first loads, if the service thread has published it already (it usually has)
    auto prefetch_next = [&](uint32_t i) {
        const uint32_t nb_ = (i + 1u) & 1u;
        if (mbar_test(&item_full[nb_], ((i + 1u) >> 1) & 1u)) {
            first_loads(s_item[nb_]);
            st |= 32u;
        }
    };

    for (uint32_t i = 0;; ++i) {
        const uint32_t buf = i & 1u;
        if (!(st & 32u)) mbar_wait(&item_full[buf], (i >> 1) & 1u);
        const FusedItem it = s_item[buf];
        if (it.role == kRoleDone) break;
        if (!(st & 32u)) first_loads(it);
        st &= ~32u;

        if (it.role == kRoleBin) {
            // =============================== bin tile ===============================
            const int lane = tid & 31, warp = tid >> 5;
            const int n_tile = (int)it.aux;
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
                    if (tid + kFusedWorkers * j >= n_tile) { cell = -1; oob = false; }
                    n_oob += oob ? 1u : 0u;
                    const uint32_t b = band_of((uint32_t)max(cell, 0), a.plan);
                    local[j] = (b << 16) | ((uint32_t)max(cell, 0) - b * (uint32_t)cpb);
                    packed[j] = 0xFFFFFFFFu;
                    if (cell >= 0) packed[j] = (b << 24) | atomicAdd(&hist[b], 1u);
                    pre[j].z = __float_as_uint(z);
                }
                if (!RANGE_SAFE) n_oob_total += n_oob;
            }
            worker_barrier();
            // threads 0..127, one band each: exclusive scan over the bands -> stage slots, and the reservation of the global
            // runs (atomics only ISSUED here; their results are consumed after the staging).  Leaves the histogram zero.
            uint32_t res = 0, slot0 = 0;
            if (tid < kFusedBands) {
                const uint32_t c = hist[tid];
                hist[tid] = 0;
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += v;
                }
                if (lane == 31) wsum[warp] = incl;
                asm volatile("bar.sync 2, %0;" ::"n"(kFusedBands) : "memory");
                uint32_t base = 0;
#pragma unroll
                for (int q = 0; q < kFusedBands / 32 - 1; ++q) base += (q < warp) ? wsum[q] : 0u;
                slot0 = base + incl - c;
                soff[tid] = slot0;
                if (tid == kFusedBands - 1) s_nkept = base + incl;
                if (c) res = atomicAdd(a.cursors + ((size_t)it.slot * nb + tid) * kCursorStride, c);
            }
            worker_barrier();
            {
                const uint32_t i0 = (uint32_t)it.idx * kFusedTile + tid;
#pragma unroll
                for (int j = 0; j < kFusedPoints; ++j) {
                    if (packed[j] != 0xFFFFFFFFu) {
                        const uint32_t s = soff[packed[j] >> 24] + (packed[j] & 0xFFFFFFu);
                        stage[s] = make_uint4(pre[j].z, pre[j].w, i0 + kFusedWorkers * j, local[j]);
                    }
                }
            }
            if (tid < kFusedBands) gpos[tid] = res - slot0;
            prefetch_next(i);   // pre[] is free: the next item's first loads fly during the barrier and the copy-out
            const int n_kept = (int)s_nkept;
            worker_barrier();
            // copy out in sorted order: stage slot s belongs to band (stage[s].w >> 16), record s - soff[band] of its run
            BevRecord* fb = a.buckets + (size_t)it.slot * a.slot_recs;
            for (int s0 = tid; s0 < n_kept; s0 += kFusedWorkers) {
                uint4 r = stage[s0];
                stage[s0] = make_uint4(0, 0, 0, 0);   // the staging area doubles as `inv`, whose idle state is zero
                const uint32_t b = r.w >> 16;
                const uint32_t pos = gpos[b] + (uint32_t)s0;
                if (pos < a.bucket_cap) {
                    r.w &= 0xFFFFu;
                    *reinterpret_cast<uint4*>(fb + (size_t)b * a.bucket_cap + pos) = r;
                } else {   // the band's bucket is full: frame overflow list, record keeps its band tag
                    BevRecord* ovf = fb + (size_t)nb * a.bucket_cap;
                    *reinterpret_cast<uint4*>(ovf + atomicAdd(reinterpret_cast<uint32_t*>(it.out), 1u)) = r;
                }
            }
            mbar_arrive(&tail_bar[buf]);   // this thread's records are stored, its staging reads done
        } else {
            // =============================== band item ===============================
            const BevRecord* rec = reinterpret_cast<const BevRecord*>(it.ptr);
            const uint32_t n_all = n_all_pre;                  // records of this band, in its bucket or overflowed
            const uint32_t n_rec = min(n_all, a.bucket_cap);   // ... of which in the bucket
            const bool overflowed = n_all > a.bucket_cap;
            wait_zero();

            if (!overflowed && n_rec <= (uint32_t)(kFusedRegRecords * kFusedWorkers)) {
                uint4 r[kFusedRegRecords];
#pragma unroll
                for (int j = 0; j < kFusedSpecRecords; ++j) r[j] = pre[j];
#pragma unroll
                for (int j = kFusedSpecRecords; j < kFusedRegRecords; ++j) {
                    const uint32_t k = tid + j * kFusedWorkers;
                    if (k < n_rec) r[j] = ld_record(rec + k);
                }
#pragma unroll
                for (int j = 0; j < kFusedRegRecords; ++j) {
                    const uint32_t k = tid + j * kFusedWorkers;
                    if (k < n_rec) {
                        atomicMax(&zkey[r[j].w], orderable_u32(__uint_as_float(r[j].x), 0u));   // NaN z sorts last (key 0)
                        atomicAdd(&cnt[r[j].w], 1u);
                    }
                }
                worker_barrier();
                // A record alone in its cell is the winner: it writes the cell's final values at once.  Cells with
                // several records vote on the lowest index among their highest-z records.
                uint32_t multi = 0;
#pragma unroll
                for (int j = 0; j < kFusedRegRecords; ++j) {
                    const uint32_t k = tid + j * kFusedWorkers;
                    if (k < n_rec) {
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
                worker_barrier();
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
                const BevRecord* ovf = a.buckets + (size_t)it.slot * a.slot_recs + (size_t)nb * a.bucket_cap;
                const uint32_t n_ovf = overflowed ? ld_cg_u32(reinterpret_cast<const uint32_t*>(it.out)) : 0u;
                worker_barrier();   // nobody is still in the previous item's copy-out / phase 3 (which use `inv`) — cheap on this rare path
                band_stream_reduce<MUL_HEIGHT, kFusedWorkers>(zkey, inv, cnt, inten, lut, rec, n_rec, ovf, n_ovf, (uint32_t)it.idx, a.g.max_h);
            }
            fence_proxy_async_smem();   // this thread's st.shared / atom.shared -> visible to the async proxy (TMA)
            prefetch_next(i);           // lands during the drain
            mbar_arrive(&tail_bar[buf]);
            st |= 2u;                   // the service thread zero-fills the planes after the drain
        }
    }
    if (!RANGE_SAFE && a.status) {
        n_oob_total = __reduce_add_sync(0xFFFFFFFFu, n_oob_total);
        if ((tid & 31) == 0 && n_oob_total) atomicAdd(a.status, n_oob_total);
    }
}

inline int fused_env(const char* name, int dflt, int lo, int hi) { return env_int(name, dflt, lo, hi); }

}  // namespace

// The fused schedule's own band plan (its band count is a tuning constant of this file).
static bool fused_plan(const SfaBevParams* p, BandPlan* plan) {
    return plan_bands(p->height, p->width, plan, kFusedMaxCellsPerBand, kFusedPlanBands) && plan->nb <= kFusedBands &&
           plan->cpb <= kFusedMaxCellsPerBand;
}
// Whether this geometry can run on the fused kernel.
bool fused_supported(const SfaBevParams* p) {
    BandPlan plan;
    return fused_plan(p, &plan);
}
// What SFA_BEV_AUTO picks among the two tiled schedules (SFA_BEV_FUSED=0/1 overrides).
bool fused_is_default() {
    static const int enabled = env_int("SFA_BEV_FUSED", 0, 0, 1);
    return enabled != 0;
}

// Enqueue all B frames as ONE launch.  ws_base: the workspace (header | cursors | ring slots of `slots_bytes`).
int fused_launch(const float* pts, const int64_t* offsets, int B, int64_t max_points, const SfaBevParams* p, const float* lut,
                 float* out, uint32_t* status, unsigned char* ws_base, uint32_t* cursors, unsigned char* slots,
                 size_t slots_bytes, cudaStream_t stream) {
    BandPlan plan;
    if (!fused_plan(p, &plan)) {
        set_error("geometry not supported by the fused kernel");
        return SFA_ERR_UNSUPPORTED;
    }
    const size_t slot_recs = slot_records(max_points, plan.nb);
    const uint32_t bucket_cap = (uint32_t)bucket_records(max_points, plan.nb);
    const int ring_avail = (int)(slots_bytes / (slot_recs * sizeof(BevRecord)));
    if (ring_avail < 1) {
        set_error("workspace too small for the fused schedule");
        return SFA_ERR_WORKSPACE_TOO_SMALL;
    }
    BevRecord* buckets = reinterpret_cast<BevRecord*>(slots);
    static const int ring_want = fused_env("SFA_BEV_FUSED_RING", 16, 2, kFusedMaxRing);
    static const int lag_want = fused_env("SFA_BEV_FUSED_LAG", 10, 1, kFusedMaxRing - 1);
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
    const size_t max_smem = 3 * (size_t)kFusedMaxCellsPerBand * 4 +
                            ((size_t)kFusedMaxCellsPerBand * 4 > kFusedStageBytes ? (size_t)kFusedMaxCellsPerBand * 4 : kFusedStageBytes);
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
