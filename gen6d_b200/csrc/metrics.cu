// Pose-evaluation errors on the device: per (predicted, ground-truth) pose pair
// the mean reprojection error, the mean 3-D point error (ADD) and, for symmetric objects, the mean
// closest-point error (ADD-S) over the object's points - utils/pose_utils.py:149-158,192-196 with
// utils/base_utils.py:256-265 (project_points, incl. its depth clamp) and :390-394.  The poses can stay
// on the device after the refiner; only q x 3 floats come back for the ADD-0.1d / Prj-5 thresholds.
//
// grid (point chunks of 256, q poses); fp32 point arithmetic like the reference's float32 numpy,
// fp64 partial sums per chunk, a second tiny kernel adds the chunks in order (deterministic).
#include "common.cuh"

namespace g6d {

struct Pose { float r[9]; float t[3]; };

__device__ __forceinline__ Pose load_pose(const float* p) {
    Pose o;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o.r[i * 3 + 0] = p[i * 4 + 0]; o.r[i * 3 + 1] = p[i * 4 + 1]; o.r[i * 3 + 2] = p[i * 4 + 2];
        o.t[i] = p[i * 4 + 3];
    }
    return o;
}
__device__ __forceinline__ float3 apply(const Pose& P, float3 x) {
    return make_float3(P.r[0] * x.x + P.r[1] * x.y + P.r[2] * x.z + P.t[0],
                       P.r[3] * x.x + P.r[4] * x.y + P.r[5] * x.z + P.t[1],
                       P.r[6] * x.x + P.r[7] * x.y + P.r[8] * x.z + P.t[2]);
}
// base_utils.py:258-264: K applied as a full 3x3, then 0 < |d| < 1e-4 -> +1e-4
__device__ __forceinline__ float2 project(const float* K, float3 c) {
    const float u = K[0] * c.x + K[1] * c.y + K[2] * c.z;
    const float v = K[3] * c.x + K[4] * c.y + K[5] * c.z;
    float d = K[6] * c.x + K[7] * c.y + K[8] * c.z;
    if (fabsf(d) < 1e-4f && fabsf(d) > 0.f) d = 1e-4f;
    return make_float2(u / d, v / d);
}

__global__ void __launch_bounds__(256)
pose_errors_kernel(const float* __restrict__ pts, int n, const float* __restrict__ pr, const float* __restrict__ gt,
                   const float* __restrict__ Ks, int symmetric, double* __restrict__ partial) {
    __shared__ float3 tile[256];
    __shared__ double red[3][8];
    const int q = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    const Pose Ppr = load_pose(pr + q * 12), Pgt = load_pose(gt + q * 12);
    const float* K = Ks + q * 9;
    const bool valid = i < n;
    float3 a = make_float3(0.f, 0.f, 0.f), b = a;
    float prj = 0.f, obj = 0.f, sym = 0.f;
    if (valid) {
        const float3 x = make_float3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
        a = apply(Ppr, x); b = apply(Pgt, x);
        const float2 ua = project(K, a), ub = project(K, b);
        prj = sqrtf((ua.x - ub.x) * (ua.x - ub.x) + (ua.y - ub.y) * (ua.y - ub.y));
        obj = sqrtf((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
    }
    if (symmetric) {
        float best = 3.4e38f;
        for (int j0 = 0; j0 < n; j0 += 256) {
            const int j = j0 + threadIdx.x;
            __syncthreads();
            if (j < n) tile[threadIdx.x] = apply(Pgt, make_float3(pts[j * 3], pts[j * 3 + 1], pts[j * 3 + 2]));
            __syncthreads();
            const int m = min(256, n - j0);
            for (int t = 0; t < m; ++t) {
                const float3 g = tile[t];
                const float d2 = (a.x - g.x) * (a.x - g.x) + (a.y - g.y) * (a.y - g.y) + (a.z - g.z) * (a.z - g.z);
                best = fminf(best, d2);
            }
        }
        sym = valid ? sqrtf(best) : 0.f;
    }
    double s0 = warp_sum((double)prj), s1 = warp_sum((double)obj), s2 = warp_sum((double)sym);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = s0; red[1][w] = s1; red[2][w] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
        partial[((long long)q * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
    }
}

__global__ void pose_errors_final_kernel(const double* __restrict__ partial, int chunks, int n, int q, int symmetric,
                                         float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q * 3) return;
    const int p = i / 3, k = i % 3;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += partial[((long long)p * chunks + c) * 3 + k];
    out[i] = (k == 2 && !symmetric) ? __int_as_float(0x7fc00000) : (float)(s / n);
}

}  // namespace g6d

extern "C" long long g6d_pose_errors_workspace_bytes(int n_pts, int n_poses) {
    if (n_pts <= 0 || n_poses <= 0) return -1;
    return (long long)n_poses * g6d::ceil_div(n_pts, 256) * 3 * (long long)sizeof(double);
}

extern "C" int g6d_pose_errors(const float* pts, int n_pts, const float* poses_pr, const float* poses_gt, const float* Ks,
                               int n_poses, int symmetric, float* out, void* ws, g6d_stream_t stream) {
    G6D_REQUIRE(pts && poses_pr && poses_gt && Ks && out && ws && n_pts > 0 && n_poses > 0 && n_poses <= 65535,
                "g6d_pose_errors: bad args");
    const int chunks = g6d::ceil_div(n_pts, 256);
    cudaStream_t st = g6d::as_stream(stream);
    g6d::pose_errors_kernel<<<dim3(chunks, n_poses), 256, 0, st>>>(pts, n_pts, poses_pr, poses_gt, Ks, symmetric, (double*)ws);
    G6D_CHECK_LAUNCH("g6d_pose_errors");
    g6d::pose_errors_final_kernel<<<g6d::ceil_div(n_poses * 3, 128), 128, 0, st>>>((const double*)ws, chunks, n_pts, n_poses,
                                                                                  symmetric, out);
    G6D_CHECK_LAUNCH("g6d_pose_errors(final)");
    return G6D_OK;
}
