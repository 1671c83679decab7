// Detections in the lidar frame -> rect camera frame -> axis-aligned image boxes: the step every
// fusion script of the reference runs right after post_processing.
//
// Reference functions replaced:
//   lidar_to_camera / lidar_to_camera_box   data_process/transformation.py:50-60, :99-107
//   convert_sfa3d_to_2d_boxes               test6.py:129-187 (same body: test4.py:128-186, msac.py:130-201,
//                                           slam.py:130-201 with a 0.2 threshold)
//
// One thread per detection, float64 like numpy (the calibration is float32 data widened exactly;
// the box rows are float32 values held in a float64 array).  Every product of the reference's dense
// 3x3 / 3x4 matrix products is kept — also the ones with a structural zero — so that NaN and inf
// propagate the way they do through np.dot.  The work is ~200 flops per box: latency only.
#include "sfa_common.cuh"

namespace sfa {
namespace {

// V2C [3][4], R0 [3][3], P2 [3][4], row-major
constexpr int kCalibDoubles = 33;

// np.min / np.max over the 8 corners: a NaN anywhere gives NaN
__device__ __forceinline__ void nan_minmax(const double (&v)[8], double& lo, double& hi) {
    lo = v[0];
    hi = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        if (!(lo != lo) && (v[i] != v[i] || v[i] < lo)) lo = v[i];
        if (!(hi != hi) && (v[i] != v[i] || v[i] > hi)) hi = v[i];
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
project_boxes_kernel(const T* __restrict__ real, const uint8_t* __restrict__ keep, int n, int K,
                     const double* __restrict__ calib, int calib_per_frame, double img_h, double img_w,
                     double min_confidence, double* __restrict__ cam, double* __restrict__ box_f,
                     int32_t* __restrict__ box, uint8_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* r = real + (size_t)i * 8;
    const double* c = calib + (calib_per_frame ? (size_t)(i / K) * kCalibDoubles : 0);
    const double* V = c;
    const double* R0 = c + 12;
    const double* P = c + 21;
    const double conf = (double)r[0];
    const double x = (double)r[1], y = (double)r[2], z = (double)r[3];
    const double h = (double)r[4], w = (double)r[5], l = (double)r[6], rz = (double)r[7];

    // transformation.py:50-60: p = V2C @ [x, y, z, 1];  p = R0 @ p
    double p[3], q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = ((V[a * 4 + 0] * x + V[a * 4 + 1] * y) + V[a * 4 + 2] * z) + V[a * 4 + 3] * 1.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) q[a] = (R0[a * 3 + 0] * p[0] + R0[a * 3 + 1] * p[1]) + R0[a * 3 + 2] * p[2];
    const double ry = -rz - 1.5707963267948966;   // transformation.py:104: -rz - np.pi / 2
    if (cam) {
        double* o = cam + (size_t)i * 7;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = h; o[4] = w; o[5] = l; o[6] = ry;
    }

    // test6.py:150-169: corners around the origin, rotate about the camera y axis, translate
    const double hl = l / 2, hw = w / 2;
    const double cx[8] = {-hl, -hl, hl, hl, -hl, -hl, hl, hl};
    const double cy[8] = {0, 0, 0, 0, -h, -h, -h, -h};
    const double cz[8] = {-hw, hw, hw, -hw, -hw, hw, hw, -hw};
    const double cs = cos(ry), sn = sin(ry);
    double u[8], v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double X = ((cs * cx[k] + 0.0 * cy[k]) + sn * cz[k]) + q[0];
        const double Y = ((0.0 * cx[k] + 1.0 * cy[k]) + 0.0 * cz[k]) + q[1];
        const double Z = ((-sn * cx[k] + 0.0 * cy[k]) + cs * cz[k]) + q[2];
        // :172-173: P2 @ [X, Y, Z, 1], then divide by the third row
        const double a0 = ((P[0] * X + P[1] * Y) + P[2] * Z) + P[3] * 1.0;
        const double a1 = ((P[4] * X + P[5] * Y) + P[6] * Z) + P[7] * 1.0;
        const double a2 = ((P[8] * X + P[9] * Y) + P[10] * Z) + P[11] * 1.0;
        u[k] = a0 / a2;
        v[k] = a1 / a2;
    }
    double min_x, max_x, min_y, max_y;
    nan_minmax(u, min_x, max_x);
    nan_minmax(v, min_y, max_y);
    // :180-183 with Python's max(0, a) / min(W, a): the first argument wins unless the comparison is true
    min_x = (min_x > 0.0) ? min_x : 0.0;
    min_y = (min_y > 0.0) ? min_y : 0.0;
    max_x = (max_x < img_w) ? max_x : img_w;
    max_y = (max_y < img_h) ? max_y : img_h;
    const bool kept = keep == nullptr || keep[i] != 0;
    const bool ok = kept && !(conf < min_confidence) && (max_x > min_x) && (max_y > min_y);   // :140, :185
    if (box_f) {
        double* o = box_f + (size_t)i * 4;
        o[0] = min_x; o[1] = min_y; o[2] = max_x; o[3] = max_y;
    }
    int32_t* b = box + (size_t)i * 4;
    // :186: int() truncates; everything is inside [0, image size] when ok
    b[0] = ok ? (int32_t)min_x : 0;
    b[1] = ok ? (int32_t)min_y : 0;
    b[2] = ok ? (int32_t)(max_x - min_x) : 0;
    b[3] = ok ? (int32_t)(max_y - min_y) : 0;
    valid[i] = ok ? 1 : 0;
}

}  // namespace
}  // namespace sfa

using namespace sfa;

extern "C" int sfa_project_boxes(const void* real, int32_t real_is_f64, const uint8_t* keep, int32_t B, int32_t K,
                                 const double* calib, int32_t calib_per_frame, int32_t img_h, int32_t img_w,
                                 double min_confidence, double* cam, double* box_f, int32_t* box, uint8_t* valid,
                                 sfa_stream_t stream) {
    SFA_REQUIRE(B >= 0 && K >= 0, "bad shape B=%d K=%d", B, K);
    SFA_REQUIRE(img_h > 0 && img_w > 0, "bad image shape %d x %d", img_h, img_w);
    const long long n = (long long)B * K;
    SFA_REQUIRE(n < 0x7FFFFFFFll, "too many boxes");
    if (n == 0) return SFA_OK;
    SFA_REQUIRE(real && calib && box && valid, "NULL pointer argument");
    const int blocks = (int)((n + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (real_is_f64) {
        SFA_LAUNCH("project_boxes", s,
                   project_boxes_kernel<double><<<blocks, 128, 0, s>>>(
                       static_cast<const double*>(real), keep, (int)n, K, calib, calib_per_frame, (double)img_h,
                       (double)img_w, min_confidence, cam, box_f, box, valid));
    } else {
        SFA_LAUNCH("project_boxes", s,
                   project_boxes_kernel<float><<<blocks, 128, 0, s>>>(
                       static_cast<const float*>(real), keep, (int)n, K, calib, calib_per_frame, (double)img_h,
                       (double)img_w, min_confidence, cam, box_f, box, valid));
    }
    SFA_CUDA_TRY(cudaGetLastError());
    return SFA_OK;
}
