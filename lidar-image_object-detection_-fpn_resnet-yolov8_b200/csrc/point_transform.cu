// Training-side sweep augmentation in front of the BEV raster:
//   point_transform          data_process/transformation.py:242-285  rows [x y z 1] times a translation
//                            matrix and up to three rotation matrices, one float64 matmul each
//   Random_Rotation (:349)   lidar[:, 0:3] = point_transform(lidar[:, 0:3], 0, 0, 0, rz=angle)  -> float32
//   Random_Scaling  (:367)   lidar[:, 0:3] = lidar[:, 0:3] * factor                              (float32)
// numpy hands the [N,4] x [4,4] products to dgemm, which accumulates the four terms of every output
// in order with fused multiply-adds (probed: acc = x*m0; acc = fma(y, m1, acc); ...); the kernel does
// exactly that in float64, matrix after matrix, so the float32 sweep it writes back is bit-identical.
// One thread per point; HBM-bound: 16 B in, 16 B out (12 + 12 for [N,3]).
#include "sfa_common.cuh"

namespace sfa {
namespace {

constexpr int kXformThreads = 256;
constexpr int kMaxMatrices = 4;   // translation + rx + ry + rz

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kXformThreads)
transform_points_kernel(const TIn* __restrict__ in, int in_stride, const int64_t* __restrict__ offsets, int64_t max_points,
                        const double* __restrict__ mats, int n_mats, const float* __restrict__ scales,
                        TOut* __restrict__ out, int out_stride) {
    const int f = blockIdx.y;
    int64_t base, n;
    if (offsets) {
        base = offsets[f];
        n = offsets[f + 1] - base;
    } else {
        base = (int64_t)f * max_points;
        n = max_points;
    }
    const int64_t i = (int64_t)blockIdx.x * kXformThreads + threadIdx.x;
    if (i >= n) return;
    const TIn* p = in + (base + i) * in_stride;
    double v[4] = {(double)p[0], (double)p[1], (double)p[2], 1.0};
    const double* M = mats + (size_t)f * n_mats * 16;
    for (int m = 0; m < n_mats; ++m, M += 16) {
        double q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = __dmul_rn(v[0], M[j]);
            acc = __fma_rn(v[1], M[4 + j], acc);
            acc = __fma_rn(v[2], M[8 + j], acc);
            acc = __fma_rn(v[3], M[12 + j], acc);
            q[j] = acc;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = q[j];
    }
    TOut* o = out + (base + i) * out_stride;
    if constexpr (sizeof(TOut) == 8) {
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    } else {
        // assignment into the float32 sweep rounds to nearest; the scaling is a float32 product
        float x = n_mats ? __double2float_rn(v[0]) : (float)p[0];
        float y = n_mats ? __double2float_rn(v[1]) : (float)p[1];
        float z = n_mats ? __double2float_rn(v[2]) : (float)p[2];
        if (scales) {
            const float s = scales[f];
            x = __fmul_rn(x, s); y = __fmul_rn(y, s); z = __fmul_rn(z, s);
        }
        o[0] = x; o[1] = y; o[2] = z;
        if constexpr (sizeof(TIn) == 4) {   // the other columns (intensity) travel unchanged
            const int extra = min(in_stride, out_stride);
            if ((const void*)o != (const void*)p)
                for (int c = 3; c < extra; ++c) o[c] = p[c];
        }
    }
}

}  // namespace
}  // namespace sfa

using namespace sfa;

extern "C" int sfa_transform_points(const void* pts, int32_t in_is_f64, int32_t in_stride, const int64_t* offsets,
                                    int32_t B, int64_t max_points, const double* mats, int32_t n_mats,
                                    const float* scales, void* out, int32_t out_is_f64, int32_t out_stride,
                                    sfa_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SFA_REQUIRE(B >= 0 && max_points >= 0, "bad batch B=%d max_points=%lld", B, (long long)max_points);
    SFA_REQUIRE(in_stride >= 3 && out_stride >= 3, "points need at least x, y, z (strides %d, %d)", in_stride, out_stride);
    SFA_REQUIRE(n_mats >= 0 && n_mats <= kMaxMatrices, "n_mats=%d unsupported (0..%d)", n_mats, kMaxMatrices);
    SFA_REQUIRE(!(out_is_f64 && scales), "the float32 scaling of Random_Scaling has no float64 form");
    SFA_REQUIRE(!(in_is_f64 && !out_is_f64 && scales), "scaling expects a float32 sweep");
    if (B == 0 || max_points == 0) return SFA_OK;
    SFA_REQUIRE(pts && out && (mats || n_mats == 0), "NULL pointer argument");
    SFA_REQUIRE(B <= 65535, "B=%d exceeds the grid limit", B);
    dim3 grid((unsigned)((max_points + kXformThreads - 1) / kXformThreads), B);
#define SFA_XFORM(TI, TO)                                                                                      \
    SFA_LAUNCH("transform_points", stream,                                                                     \
               transform_points_kernel<TI, TO><<<grid, kXformThreads, 0, stream>>>(                             \
                   static_cast<const TI*>(pts), in_stride, offsets, max_points, mats, n_mats, scales,           \
                   static_cast<TO*>(out), out_stride))
    if (in_is_f64 && out_is_f64) SFA_XFORM(double, double);
    else if (in_is_f64) SFA_XFORM(double, float);
    else if (out_is_f64) SFA_XFORM(float, double);
    else SFA_XFORM(float, float);
#undef SFA_XFORM
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}
