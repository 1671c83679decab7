// libsfa_b200.so: library glue (version, per-thread error text) and the HOST-buffer pipeline.
//
// The reference's call sites hand numpy arrays in and take numpy arrays back
// (data_process/kitti_dataset.py:64-67 `makeBEVMap(...)` -> `torch.from_numpy`; test.py:170-173
// `decode(...)` -> `.cpu().numpy()`), so a drop-in has to accept host memory.  SfaPipeline moves
// chunks of frames host->device, runs the same kernels as the device API, and moves results back,
// on kLanes independent streams so that the H2D copy of chunk i+1, the kernels of chunk i and the
// D2H copy of chunk i-1 overlap (PCIe is full duplex).  All device staging memory is allocated once
// at creation.
#include "sfa_common.cuh"

#include <string.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

namespace sfa {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return SFA_ERR_CUDA;
}

// ---- launch accounting and per-kernel event timing ------------------------------------------
static std::atomic<uint64_t> g_launches{0};
static std::atomic<bool> g_profiling{false};
static std::mutex g_prof_mutex;
struct ProfSpan { std::string name; cudaEvent_t a = nullptr, b = nullptr; };
static std::vector<ProfSpan> g_spans;

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool profiling_on() { return g_profiling.load(std::memory_order_relaxed); }

void profile_mark(const char* name, cudaStream_t stream, bool begin) {
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    if (begin) {
        ProfSpan s;
        s.name = name;
        if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
        cudaEventRecord(s.a, stream);
        g_spans.push_back(s);
    } else if (!g_spans.empty()) {
        cudaEventRecord(g_spans.back().b, stream);
    }
}

}  // namespace sfa

using namespace sfa;

extern "C" int sfa_version(void) { return SFA_B200_VERSION; }
extern "C" const char* sfa_last_error(void) { return g_error; }
extern "C" uint64_t sfa_kernel_launches(void) { return g_launches.load(); }

extern "C" int sfa_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    for (ProfSpan& s : g_spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    g_spans.clear();
    g_profiling.store(true);
    return SFA_OK;
}

extern "C" int sfa_profile_end(SfaKernelStat* stats, int32_t max_stats) {
    g_profiling.store(false);
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    int n = 0;
    int rc = SFA_OK;
    for (ProfSpan& s : g_spans) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(s.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, s.a, s.b);
        if (e != cudaSuccess) { rc = cuda_fail(e, "profile event"); continue; }
        int i = 0;
        for (; i < n; ++i)
            if (s.name == stats[i].name) break;
        if (i == n) {
            if (n >= max_stats || !stats) continue;
            memset(&stats[n], 0, sizeof(SfaKernelStat));
            strncpy(stats[n].name, s.name.c_str(), sizeof(stats[n].name) - 1);
            ++n;
        }
        stats[i].launches += 1;
        stats[i].total_ms += ms;
    }
    for (ProfSpan& s : g_spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    g_spans.clear();
    return rc == SFA_OK ? n : rc;
}

namespace {

constexpr int kLanes = 3;

struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;      // last work enqueued on this lane
    float* d_pts = nullptr;
    int64_t* d_offsets = nullptr;
    int64_t* h_offsets = nullptr;    // pinned
    float* d_out = nullptr;
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    float* d_heads = nullptr;        // hm | off | dir | z | dim for one chunk
    float* d_det = nullptr;
    void* d_dec_ws = nullptr;        // sfa_decode workspace
    size_t dec_ws_bytes = 0;
};

}  // namespace

// Makes pl->device current for the duration of a pipeline call and puts the caller's device back afterwards (a process
// that did cudaSetDevice(rank) must not find itself on another GPU after calling into the library).
struct DeviceScope {
    int prev = -1;
    cudaError_t err;
    explicit DeviceScope(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    }
    ~DeviceScope() {
        int now = -1;
        if (prev >= 0 && cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev);
    }
};

struct SfaPipeline {
    std::mutex mutex;   // calls on one pipeline serialise (its lanes' staging buffers are shared state)
    int device = 0;
    int max_frames = 0, chunk = 0;
    int64_t max_points = 0;
    SfaBevParams params;
    int C = 0, h = 0, w = 0, K = 0;
    float* d_lut = nullptr;
    uint32_t* d_status = nullptr;
    Lane lanes[kLanes];
};

// Waits (blocking, not spinning: the events are created with cudaEventBlockingSync) until every lane has finished what
// was enqueued on it; returns the first CUDA error.  Also the error path's drain: no copy into a caller's buffer is
// left in flight when a pipeline call returns.
static cudaError_t drain_lanes(SfaPipeline* pl) {
    cudaError_t first = cudaSuccess;
    for (Lane& l : pl->lanes) {
        cudaError_t e = cudaEventRecord(l.done, l.stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(l.done);
        if (e != cudaSuccess && first == cudaSuccess) first = e;
    }
    return first;
}

static void pipeline_free(SfaPipeline* pl) {
    if (!pl) return;
    DeviceScope dev(pl->device);
    for (Lane& l : pl->lanes) {
        if (l.stream) cudaStreamSynchronize(l.stream);
        cudaFree(l.d_pts); cudaFree(l.d_offsets); cudaFreeHost(l.h_offsets); cudaFree(l.d_out);
        if (l.d_ws) sfa_bev_workspace_release(l.d_ws);   // the rasteriser's side streams / events that belong to this workspace
        cudaFree(l.d_ws); cudaFree(l.d_heads); cudaFree(l.d_det); cudaFree(l.d_dec_ws);
        if (l.done) cudaEventDestroy(l.done);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    cudaFree(pl->d_lut);
    cudaFree(pl->d_status);
    delete pl;
}

extern "C" SfaPipeline* sfa_pipeline_create(int32_t device, int32_t max_frames, int64_t max_points_per_frame,
                                            const SfaBevParams* p, const float* density_lut_host, int32_t C, int32_t h,
                                            int32_t w, int32_t K) {
    if (!p || !density_lut_host || max_frames <= 0 || max_points_per_frame < 0 || p->height <= 0 || p->width <= 0) {
        set_error("sfa_pipeline_create: invalid argument");
        return nullptr;
    }
    SfaPipeline* pl = new SfaPipeline();
    pl->device = device;
    pl->max_frames = max_frames;
    const char* e = getenv("SFA_PIPELINE_CHUNK");
    int chunk = e ? atoi(e) : 16;
    if (chunk < 1) chunk = 1;
    pl->chunk = max_frames < chunk ? max_frames : chunk;
    pl->max_points = max_points_per_frame;
    pl->params = *p;
    pl->C = C; pl->h = h; pl->w = w; pl->K = K;
    bool ok = true;
    auto chk = [&](cudaError_t err, const char* what) {
        if (ok && err != cudaSuccess) { cuda_fail(err, what); ok = false; }
    };
    DeviceScope dev(device);
    chk(dev.err, "cudaSetDevice");
    chk(cudaMalloc(&pl->d_lut, 64 * sizeof(float)), "cudaMalloc lut");
    chk(cudaMalloc(&pl->d_status, 2 * sizeof(uint32_t)), "cudaMalloc status");
    if (ok) chk(cudaMemcpy(pl->d_lut, density_lut_host, 64 * sizeof(float), cudaMemcpyHostToDevice), "copy lut");
    if (ok) chk(cudaMemset(pl->d_status, 0, 2 * sizeof(uint32_t)), "memset status");
    const size_t cells = (size_t)p->height * p->width;
    const size_t head_ch = (size_t)(C > 0 ? C + 8 : 0);
    for (Lane& l : pl->lanes) {
        if (!ok) break;
        chk(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking), "stream");
        chk(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming | cudaEventBlockingSync), "event");
        size_t npts = (size_t)pl->chunk * (size_t)(max_points_per_frame > 0 ? max_points_per_frame : 1);
        chk(cudaMalloc(&l.d_pts, npts * 16), "cudaMalloc pts");
        chk(cudaMalloc(&l.d_offsets, (pl->chunk + 1) * sizeof(int64_t)), "cudaMalloc offsets");
        chk(cudaMallocHost(&l.h_offsets, (pl->chunk + 1) * sizeof(int64_t)), "cudaMallocHost offsets");
        chk(cudaMalloc(&l.d_out, (size_t)pl->chunk * 3 * cells * sizeof(float)), "cudaMalloc out");
        l.ws_bytes = sfa_bev_workspace_bytes(pl->chunk, max_points_per_frame, p);
        chk(cudaMalloc(&l.d_ws, l.ws_bytes), "cudaMalloc ws");
        if (ok && sfa_bev_workspace_init(l.d_ws, l.ws_bytes, l.stream) != SFA_OK) ok = false;
        if (head_ch) {
            chk(cudaMalloc(&l.d_heads, (size_t)pl->chunk * head_ch * h * w * sizeof(float)), "cudaMalloc heads");
            chk(cudaMalloc(&l.d_det, (size_t)pl->chunk * K * 10 * sizeof(float)), "cudaMalloc det");
            l.dec_ws_bytes = sfa_decode_workspace_bytes(pl->chunk, C, h, w, K);
            chk(cudaMalloc(&l.d_dec_ws, l.dec_ws_bytes), "cudaMalloc decode ws");
            if (ok && sfa_decode_workspace_init(l.d_dec_ws, l.dec_ws_bytes, l.stream) != SFA_OK) ok = false;
        }
        if (ok) chk(cudaStreamSynchronize(l.stream), "sync");
    }
    if (!ok) {
        pipeline_free(pl);
        return nullptr;
    }
    return pl;
}

extern "C" void sfa_pipeline_destroy(SfaPipeline* pl) { pipeline_free(pl); }

extern "C" int sfa_pipeline_bev_host(SfaPipeline* pl, const float* pts_host, const int64_t* offsets_host, int32_t B,
                                     float* out_host, uint32_t* status_host) {
    SFA_REQUIRE(pl && offsets_host && (out_host || B == 0), "NULL pointer argument");
    SFA_REQUIRE(B >= 0, "B must be >= 0");
    std::lock_guard<std::mutex> lock(pl->mutex);
    DeviceScope dev(pl->device);
    SFA_CUDA_TRY(dev.err);
    const size_t cells = (size_t)pl->params.height * pl->params.width;
    for (int f = 0; f < B; ++f) {
        int64_t n = offsets_host[f + 1] - offsets_host[f];
        SFA_REQUIRE(n >= 0 && n <= pl->max_points, "sweep %d has %lld points; pipeline was created for <= %lld", f,
                    (long long)n, (long long)pl->max_points);
    }
    int rc = SFA_OK;
    auto cu = [&](cudaError_t e, const char* what) { if (rc == SFA_OK && e != cudaSuccess) rc = cuda_fail(e, what); return rc == SFA_OK; };
    int lane_i = 0;
    for (int f0 = 0; f0 < B && rc == SFA_OK; f0 += pl->chunk, lane_i = (lane_i + 1) % kLanes) {
        Lane& l = pl->lanes[lane_i];
        const int nf = B - f0 < pl->chunk ? B - f0 : pl->chunk;
        if (!cu(cudaEventSynchronize(l.done), "lane wait")) break;   // pinned offsets of this lane are free again
        int64_t base = offsets_host[f0], mx = 0;
        for (int j = 0; j <= nf; ++j) l.h_offsets[j] = offsets_host[f0 + j] - base;
        for (int j = 0; j < nf; ++j) {
            int64_t n = l.h_offsets[j + 1] - l.h_offsets[j];
            if (n > mx) mx = n;
        }
        const int64_t npts = l.h_offsets[nf];
        if (!cu(cudaMemcpyAsync(l.d_offsets, l.h_offsets, (nf + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, l.stream), "offsets H2D")) break;
        if (npts && !cu(cudaMemcpyAsync(l.d_pts, pts_host + base * 4, (size_t)npts * 16, cudaMemcpyHostToDevice, l.stream), "sweeps H2D")) break;
        rc = sfa_bev_rasterize(l.d_pts, l.d_offsets, nf, mx, &pl->params, pl->d_lut, l.d_out, pl->d_status, l.d_ws, l.ws_bytes, l.stream);
        if (rc != SFA_OK) break;
        if (!cu(cudaMemcpyAsync(out_host + (size_t)f0 * 3 * cells, l.d_out, (size_t)nf * 3 * cells * sizeof(float),
                                cudaMemcpyDeviceToHost, l.stream), "maps D2H")) break;
        cu(cudaEventRecord(l.done, l.stream), "lane record");
    }
    // success or not, nothing stays in flight towards the caller's buffers
    const cudaError_t drained = drain_lanes(pl);
    if (rc == SFA_OK && drained != cudaSuccess) rc = cuda_fail(drained, "drain");
    if (rc != SFA_OK) return rc;
    if (status_host) {
        SFA_CUDA_TRY(cudaMemcpy(status_host, pl->d_status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        SFA_CUDA_TRY(cudaMemset(pl->d_status, 0, 2 * sizeof(uint32_t)));
    }
    return SFA_OK;
}

extern "C" int sfa_pipeline_decode_host(SfaPipeline* pl, const float* hm, const float* cen_offset, const float* direction,
                                        const float* z_coor, const float* dim, int32_t B, float* det_host) {
    SFA_REQUIRE(pl && (B == 0 || (hm && direction && z_coor && dim && det_host)), "NULL pointer argument");
    SFA_REQUIRE(pl->C > 0, "pipeline was created without a decode stage");
    std::lock_guard<std::mutex> lock(pl->mutex);
    DeviceScope dev(pl->device);
    SFA_CUDA_TRY(dev.err);
    const size_t hw = (size_t)pl->h * pl->w;
    const int C = pl->C, K = pl->K;
    int rc = SFA_OK;
    auto cu = [&](cudaError_t e, const char* what) { if (rc == SFA_OK && e != cudaSuccess) rc = cuda_fail(e, what); return rc == SFA_OK; };
    int lane_i = 0;
    for (int f0 = 0; f0 < B && rc == SFA_OK; f0 += pl->chunk, lane_i = (lane_i + 1) % kLanes) {
        Lane& l = pl->lanes[lane_i];
        const int nf = B - f0 < pl->chunk ? B - f0 : pl->chunk;
        float* d_hm = l.d_heads;
        float* d_off = d_hm + (size_t)pl->chunk * C * hw;
        float* d_dir = d_off + (size_t)pl->chunk * 2 * hw;
        float* d_z = d_dir + (size_t)pl->chunk * 2 * hw;
        float* d_dim = d_z + (size_t)pl->chunk * 1 * hw;
        auto up = [&](float* dst, const float* src, int ch) -> bool {
            return cu(cudaMemcpyAsync(dst, src + (size_t)f0 * ch * hw, (size_t)nf * ch * hw * sizeof(float),
                                      cudaMemcpyHostToDevice, l.stream), "heads H2D");
        };
        if (!up(d_hm, hm, C)) break;
        if (cen_offset && !up(d_off, cen_offset, 2)) break;
        if (!up(d_dir, direction, 2) || !up(d_z, z_coor, 1) || !up(d_dim, dim, 3)) break;
        rc = sfa_decode(d_hm, cen_offset ? d_off : nullptr, d_dir, d_z, d_dim, nf, C, pl->h, pl->w, K, l.d_det,
                        nullptr, 0, l.d_dec_ws, l.dec_ws_bytes, l.stream);
        if (rc != SFA_OK) break;
        cu(cudaMemcpyAsync(det_host + (size_t)f0 * K * 10, l.d_det, (size_t)nf * K * 10 * sizeof(float),
                           cudaMemcpyDeviceToHost, l.stream), "detections D2H");
    }
    const cudaError_t drained = drain_lanes(pl);
    if (rc == SFA_OK && drained != cudaSuccess) rc = cuda_fail(drained, "drain");
    return rc;
}
