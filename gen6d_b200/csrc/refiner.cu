// Refiner-specific kernels: the unproject-and-aggregate volume fill (R2) and the pose heads.
#include "common.cuh"

namespace g6d {

constexpr int kMaxRefViews = 8;

struct VolParams {
    const float* ref_feats; const float* que_feats;
    const float* ref_Ks; const float* ref_poses; const float* que_Ks; const float* que_poses;
    float* mean_in; float* stdv;
    int Q, R, fh, fw, C, sn, img_h, img_w;
};

// grid_sample(bilinear, zeros padding, align_corners=False) tap set for one projected point
struct Taps {
    int idx[4];     // feature-map linear index (y*fw + x) or -1 when the tap is out of bounds
    float w[4];
};

// P = K @ [R|t]  (refiner.py:227,243), then p = v @ P[:, :3]^T + P[:, 3] (refiner.py:195-197)
__device__ __forceinline__ Taps make_taps(const float* __restrict__ K, const float* __restrict__ pose, float vx,
                                          float vy, float vz, int fh, int fw, int img_h, int img_w) {
    float P[12];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            P[r * 4 + c] = fmaf(K[r * 3 + 0], pose[c], fmaf(K[r * 3 + 1], pose[4 + c], K[r * 3 + 2] * pose[8 + c]));
    const float px = fmaf(vx, P[0], fmaf(vy, P[1], fmaf(vz, P[2], P[3])));
    const float py = fmaf(vx, P[4], fmaf(vy, P[5], fmaf(vz, P[6], P[7])));
    float pz = fmaf(vx, P[8], fmaf(vy, P[9], fmaf(vz, P[10], P[11])));
    if (pz < 1e-4f) pz = 1e-4f;                       // refiner.py:199-200
    const float u = px / pz, v = py / pz;
    // normalize_coords (operator.py:4-17) with the IMAGE size, then grid_sample's unnormalise
    // with the FEATURE size (align_corners=False): ix = ((g + 1) * W_f - 1) / 2
    const float gx = ((u + 0.5f) / (float)img_w - 0.5f) * 2.f;
    const float gy = ((v + 0.5f) / (float)img_h - 0.5f) * 2.f;
    const float ix = ((gx + 1.f) * (float)fw - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * (float)fh - 1.f) * 0.5f;
    Taps t;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float lx = ix - fx0, ly = iy - fy0;
    // keep the float->int conversion safe for huge coordinates (points behind the camera)
    const bool near = (fx0 > -2.f) && (fx0 < (float)fw + 1.f) && (fy0 > -2.f) && (fy0 < (float)fh + 1.f);
    const int x0 = near ? (int)fx0 : -10, y0 = near ? (int)fy0 : -10;
    const int xs[2] = {x0, x0 + 1}, ys[2] = {y0, y0 + 1};
    const float wx[2] = {1.f - lx, lx}, wy[2] = {1.f - ly, ly};
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const bool inb = (unsigned)xs[i] < (unsigned)fw && (unsigned)ys[j] < (unsigned)fh;
            t.idx[j * 2 + i] = inb ? ys[j] * fw + xs[i] : -1;
            t.w[j * 2 + i] = wx[i] * wy[j];
        }
    return t;
}

__device__ __forceinline__ float4 sample4(const float* __restrict__ fmap, const Taps& t, int C, int c) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.idx[k] >= 0) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(fmap + (long long)t.idx[k] * C + c));
            acc.x = fmaf(v.x, t.w[k], acc.x); acc.y = fmaf(v.y, t.w[k], acc.y);
            acc.z = fmaf(v.z, t.w[k], acc.z); acc.w = fmaf(v.w, t.w[k], acc.w);
        }
    }
    return acc;
}

// One warp per voxel; lane l owns channels [4l, 4l+4) (+128 per extra pass for C > 128).
// Projection + bilinear weights are recomputed by every lane (a few dozen FMAs, cheaper than a
// shuffle broadcast).  The 7 feature maps (3.7 MB per pose) stay L2-resident; the HBM traffic
// that matters is the 3*C*sn^3*4 B of output, written as full 512 B rows with streaming stores.
__global__ void __launch_bounds__(256) ref_volume_fill_kernel(const VolParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nvox = (long long)p.sn * p.sn * p.sn;
    if (warp_global >= nvox * p.Q) return;
    const int qi = (int)(warp_global / nvox);
    const int vox = (int)(warp_global % nvox);
    const int k = vox % p.sn, j = (vox / p.sn) % p.sn, i = vox / (p.sn * p.sn);
    // torch.linspace(-1, 1, sn): start + step*idx for the first half, end - step*(n-1-idx) after
    const float step = 2.f / (float)(p.sn - 1);
    auto lin = [&](int a) { return a < p.sn / 2 ? -1.f + step * (float)a : 1.f - step * (float)(p.sn - 1 - a); };
    const float ci = lin(i), cj = lin(j), ck = lin(k);
    // row vector @ R_in  (refiner.py:216-220); R_in = poses_in[:, :3, :3]
    const float* qp = p.que_poses + (long long)qi * 12;
    const float vx = fmaf(ci, qp[0], fmaf(cj, qp[4], ck * qp[8]));
    const float vy = fmaf(ci, qp[1], fmaf(cj, qp[5], ck * qp[9]));
    const float vz = fmaf(ci, qp[2], fmaf(cj, qp[6], ck * qp[10]));

    const long long fsz = (long long)p.fh * p.fw * p.C;
    const long long orow = (long long)qi * nvox + vox;
    for (int c = lane * 4; c < p.C; c += 128) {
        float4 s[kMaxRefViews];
        float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxRefViews; ++r) {
            if (r < p.R) {
                const long long v = (long long)qi * p.R + r;
                const Taps t = make_taps(p.ref_Ks + v * 9, p.ref_poses + v * 12, vx, vy, vz, p.fh, p.fw, p.img_h, p.img_w);
                s[r] = sample4(p.ref_feats + v * fsz, t, p.C, c);
                mean.x += s[r].x; mean.y += s[r].y; mean.z += s[r].z; mean.w += s[r].w;
            }
        }
        const float fr = (float)p.R;
        mean.x /= fr; mean.y /= fr; mean.z /= fr; mean.w /= fr;
        float4 var = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxRefViews; ++r) {
            if (r < p.R) {
                float d;
                d = s[r].x - mean.x; var.x = fmaf(d, d, var.x);
                d = s[r].y - mean.y; var.y = fmaf(d, d, var.y);
                d = s[r].z - mean.z; var.z = fmaf(d, d, var.z);
                d = s[r].w - mean.w; var.w = fmaf(d, d, var.w);
            }
        }
        const float fu = (float)(p.R - 1);  // unbiased (torch.std default, refiner.py:237)
        float4 sd;
        sd.x = sqrtf(var.x / fu); sd.y = sqrtf(var.y / fu); sd.z = sqrtf(var.z / fu); sd.w = sqrtf(var.w / fu);
        const Taps qt = make_taps(p.que_Ks + (long long)qi * 9, qp, vx, vy, vz, p.fh, p.fw, p.img_h, p.img_w);
        const float4 qs = sample4(p.que_feats + (long long)qi * fsz, qt, p.C, c);
        float* mrow = p.mean_in + orow * (2 * p.C);
        __stcs(reinterpret_cast<float4*>(mrow + c), mean);
        __stcs(reinterpret_cast<float4*>(mrow + p.C + c), qs);
        __stcs(reinterpret_cast<float4*>(p.stdv + orow * p.C + c), sd);
    }
}

// x [M,K] -> out [M,7]: quaternion (normalised, F.normalize eps 1e-12), 2-D offset, log2 scale
__global__ void ref_pose_heads_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ b, float* __restrict__ out, int K) {
    const int m = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float res[7];
    if (warp < 7) {
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) acc = fmaf(x[(long long)m * K + k], w[(long long)warp * K + k], acc);
        acc = warp_sum(acc);
        if (lane == 0) res[warp] = acc + b[warp];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float n = fmaxf(sqrtf(res[0] * res[0] + res[1] * res[1] + res[2] * res[2] + res[3] * res[3]), 1e-12f);
        for (int i = 0; i < 4; ++i) out[m * 7 + i] = res[i] / n;
        for (int i = 4; i < 7; ++i) out[m * 7 + i] = res[i];
    }
}

}  // namespace g6d

using namespace g6d;

extern "C" int g6d_ref_volume_fill(const float* ref_feats, const float* que_feats, const float* ref_Ks,
                                   const float* ref_poses, const float* que_Ks, const float* que_poses, int Q, int R,
                                   int fh, int fw, int C, int sn, int img_h, int img_w, float* mean_in, float* stdv,
                                   g6d_stream_t stream) {
    G6D_REQUIRE(ref_feats && que_feats && ref_Ks && ref_poses && que_Ks && que_poses && mean_in && stdv,
                "g6d_ref_volume_fill: null pointer");
    G6D_REQUIRE(Q > 0 && R >= 2 && R <= kMaxRefViews && fh > 0 && fw > 0 && C > 0 && (C & 3) == 0 && sn >= 2 &&
                    img_h > 0 && img_w > 0,
                "g6d_ref_volume_fill: bad dims (2 <= R <= %d, C%%4 == 0, sn >= 2)", kMaxRefViews);
    VolParams p{ref_feats, que_feats, ref_Ks, ref_poses, que_Ks, que_poses, mean_in, stdv,
                Q, R, fh, fw, C, sn, img_h, img_w};
    const long long warps = (long long)Q * sn * sn * sn;
    ref_volume_fill_kernel<<<ceil_div(warps, 8), 256, 0, as_stream(stream)>>>(p);
    G6D_CHECK_LAUNCH("g6d_ref_volume_fill");
    return G6D_OK;
}

extern "C" int g6d_ref_pose_heads(const float* x, const float* w, const float* b, float* out, int M, int K,
                                  g6d_stream_t stream) {
    G6D_REQUIRE(x && w && b && out && M > 0 && K > 0, "g6d_ref_pose_heads: bad args");
    ref_pose_heads_kernel<<<M, 256, 0, as_stream(stream)>>>(x, w, b, out, K);
    G6D_CHECK_LAUNCH("g6d_ref_pose_heads");
    return G6D_OK;
}
