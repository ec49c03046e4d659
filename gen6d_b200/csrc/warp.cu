// Bit-exact device restatement of the two OpenCV warps the estimator applies between the network
// stages: cv2.warpPerspective (look-at crops of refiner.py:285-325 through
// database_utils.py:8-25) and cv2.warpAffine (the detection crop, base_utils.py:646-655), both
// uint8, 3 channels, INTER_LINEAR, BORDER_CONSTANT(0).
//
// OpenCV's 8-bit linear path is fixed point and therefore reproducible exactly:
//   * source coordinates are computed in double, scaled by 32 (INTER_TAB_SIZE) and rounded to the
//     nearest even integer; the low 5 bits select the bilinear weights, the rest the top-left tap;
//   * the four weights are 32*(32-b or b)*(32-a or a) (they sum to 2^15 exactly), the result is
//     (sum + 2^14) >> 15; taps outside the image contribute the border value 0.
// warpPerspective evaluates (M0*x_blk + M1*y + M2) + M0*x1 with x_blk the start of a
// min(1024/min(16,h), w)-wide column block and divides by the same form of the denominator;
// warpAffine pre-rounds M0*x*1024 and (M1*y + M2)*1024 to int separately and adds 16 before the
// shift.  Every double operation below uses an explicit round-to-nearest intrinsic so that nvcc
// cannot contract a*b+c into an FMA (OpenCV's baseline build has none on this path).
#include "common.cuh"

namespace g6d {

__device__ __forceinline__ int sat_round(double v) {
    v = fmax(-2147483648.0, fmin(2147483647.0, v));
    return __double2int_rn(v);
}

__device__ __forceinline__ void fixed_bilinear_u8x3(const g6d_warp_job& jb, int X, int Y, uint8_t* dst) {
    int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const int a = X & 31, b = Y & 31;
    const int w00 = 32 * (32 - b) * (32 - a), w01 = 32 * (32 - b) * a, w10 = 32 * b * (32 - a), w11 = 32 * b * a;
    const bool x0 = (unsigned)sx < (unsigned)jb.cols, x1 = (unsigned)(sx + 1) < (unsigned)jb.cols;
    const bool y0 = (unsigned)sy < (unsigned)jb.rows, y1 = (unsigned)(sy + 1) < (unsigned)jb.rows;
    const uint8_t* p = jb.src + ((long long)sy * jb.cols + sx) * 3;
    const long long rs = (long long)jb.cols * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int v = 0;
        if (y0 && x0) v += p[c] * w00;
        if (y0 && x1) v += p[3 + c] * w01;
        if (y1 && x0) v += p[rs + c] * w10;
        if (y1 && x1) v += p[rs + 3 + c] * w11;
        dst[c] = (uint8_t)((v + 16384) >> 15);
    }
}

__global__ void __launch_bounds__(256)
warp_perspective_u8_kernel(const g6d_warp_job* __restrict__ jobs, uint8_t* __restrict__ out, int h, int w, int bw0) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, j = blockIdx.z;
    if (x >= w) return;
    const g6d_warp_job jb = jobs[j];
    const double* M = jb.M;
    const double xb = (double)((x / bw0) * bw0), x1 = (double)(x % bw0), yy = (double)y;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], xb), __dmul_rn(M[1], yy)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], xb), __dmul_rn(M[4], yy)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], xb), __dmul_rn(M[7], yy)), M[8]);
    double W = __dadd_rn(W0, __dmul_rn(M[6], x1));
    W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
    const int X = sat_round(__dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], x1)), W));
    const int Y = sat_round(__dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], x1)), W));
    fixed_bilinear_u8x3(jb, X, Y, out + (((long long)j * h + y) * w + x) * 3);
}

__global__ void __launch_bounds__(256)
warp_affine_u8_kernel(const g6d_warp_job* __restrict__ jobs, uint8_t* __restrict__ out, int h, int w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, j = blockIdx.z;
    if (x >= w) return;
    const g6d_warp_job jb = jobs[j];
    const double* M = jb.M;
    const double xx = (double)x, yy = (double)y;
    const int adelta = sat_round(__dmul_rn(__dmul_rn(M[0], xx), 1024.0));
    const int bdelta = sat_round(__dmul_rn(__dmul_rn(M[3], xx), 1024.0));
    const int X0 = sat_round(__dmul_rn(__dadd_rn(__dmul_rn(M[1], yy), M[2]), 1024.0)) + 16;
    const int Y0 = sat_round(__dmul_rn(__dadd_rn(__dmul_rn(M[4], yy), M[5]), 1024.0)) + 16;
    fixed_bilinear_u8x3(jb, (X0 + adelta) >> 5, (Y0 + bdelta) >> 5, out + (((long long)j * h + y) * w + x) * 3);
}

}  // namespace g6d

extern "C" int g6d_warp_perspective_u8(const g6d_warp_job* jobs, int n_jobs, uint8_t* out, int h, int w,
                                       g6d_stream_t stream) {
    G6D_REQUIRE(jobs && out && n_jobs > 0 && h > 0 && w > 0 && h <= 65535 && n_jobs <= 65535,
                "g6d_warp_perspective_u8: bad args");
    const int bh0 = h < 16 ? h : 16;
    const int bw0 = (1024 / bh0) < w ? (1024 / bh0) : w;
    dim3 grid(g6d::ceil_div(w, 128), h, n_jobs);
    g6d::warp_perspective_u8_kernel<<<grid, 128, 0, g6d::as_stream(stream)>>>(jobs, out, h, w, bw0);
    G6D_CHECK_LAUNCH("g6d_warp_perspective_u8");
    return G6D_OK;
}

extern "C" int g6d_warp_affine_u8(const g6d_warp_job* jobs, int n_jobs, uint8_t* out, int h, int w,
                                  g6d_stream_t stream) {
    G6D_REQUIRE(jobs && out && n_jobs > 0 && h > 0 && w > 0 && h <= 65535 && n_jobs <= 65535,
                "g6d_warp_affine_u8: bad args");
    dim3 grid(g6d::ceil_div(w, 128), h, n_jobs);
    g6d::warp_affine_u8_kernel<<<grid, 128, 0, g6d::as_stream(stream)>>>(jobs, out, h, w);
    G6D_CHECK_LAUNCH("g6d_warp_affine_u8");
    return G6D_OK;
}
