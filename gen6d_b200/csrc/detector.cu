// Detector-specific kernels: fused normalise/clip/upsample/resize/stack/score_conv/max-over-refs
// (D2 epilogue + D3 head) and the argmax decode (D4).
#include "common.cuh"

namespace g6d {

struct DetFuseParams {
    g6d_det_maps maps;
    const float* w1; const float* b1; const float* w2; const float* b2;
    float* out;
    int qn;
};

__device__ __forceinline__ void bil_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;   // align_corners=False, clamped at 0 (ATen)
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i0 = i0 > in_size - 1 ? in_size - 1 : i0;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

// One warp per output pixel; lanes stride over the reference views.  For each (pixel, ref) the 3S
// inputs are gathered straight from the raw per-level correlation maps: level l of a scale is
// nearest-upsampled by 2^l (detector.py:225-226: index >> l), normalised and clipped
// (detector.py:207-216) at the 4 bilinear taps of the (Hc,Wc)->(hs,ws) resize (detector.py:243),
// then blended.  score_conv (12->64 ReLU ->64, detector.py:159-163,246) runs in registers with
// the weights in shared memory, and the max over references (detector.py:247) is a warp max.
template <int NIN>
__global__ void __launch_bounds__(128) det_score_fuse_kernel(const DetFuseParams p) {
    constexpr int D = 64;
    __shared__ float w1s[D * NIN], b1s[D], b2s[D];
    __shared__ __align__(16) float w2s[D * D];
    for (int i = threadIdx.x; i < D * NIN; i += blockDim.x) w1s[i] = p.w1[i];
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) w2s[i] = p.w2[i];
    for (int i = threadIdx.x; i < D; i += blockDim.x) { b1s[i] = p.b1[i]; b2s[i] = p.b2[i]; }
    __syncthreads();
    const g6d_det_maps& M = p.maps;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long npix = (long long)p.qn * M.hs * M.ws;
    if (pix >= npix) return;
    const int x = (int)(pix % M.ws), y = (int)((pix / M.ws) % M.hs), qi = (int)(pix / ((long long)M.ws * M.hs));

    auto gather = [&](int r, float* in) {
        const bool valid = r < M.rfn;
#pragma unroll
        for (int s = 0; s < NIN / 3; ++s) {
            const int Hc = M.H[s][0], Wc = M.W[s][0];
            int y0, y1, x0, x1; float ly, lx;
            bil_src(y, (float)Hc / (float)M.hs, Hc, y0, y1, ly);
            bil_src(x, (float)Wc / (float)M.ws, Wc, x0, x1, lx);
            const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                float v = 0.f;
                if (valid) {
                    const int Hl = M.H[s][l], Wl = M.W[s][l];
                    const float* mp = M.map[s][l] + (long long)qi * Hl * Wl * M.rfn + r;
                    auto tap = [&](int yy, int xx) {
                        const float raw = __ldg(mp + ((long long)(yy >> l) * Wl + (xx >> l)) * M.rfn);
                        const float n = (raw - M.mu[l]) * M.inv_sigma[l];
                        return fminf(fmaxf(n, -M.clip), M.clip);
                    };
                    v = hy * (hx * tap(y0, x0) + lx * tap(y0, x1)) + ly * (hx * tap(y1, x0) + lx * tap(y1, x1));
                }
                in[s * 3 + l] = v;
            }
        }
    };
    auto hidden = [&](const float* in, float* hid) {
#pragma unroll
        for (int h = 0; h < D; ++h) {
            float a = b1s[h];
#pragma unroll
            for (int i = 0; i < NIN; ++i) a = fmaf(w1s[h * NIN + i], in[i], a);
            hid[h] = fmaxf(a, 0.f);
        }
    };

    // references are processed 64 at a time (two per lane) so the hidden vectors stay in registers
    for (int r0 = 0; r0 < M.rfn; r0 += 64) {
        float in[NIN], hidA[D], hidB[D];
        const int ra = r0 + lane, rb = r0 + 32 + lane;
        gather(ra, in); hidden(in, hidA);
        gather(rb, in); hidden(in, hidB);
        const bool va = ra < M.rfn, vb = rb < M.rfn;
#pragma unroll 2
        for (int o = 0; o < D; ++o) {
            float a = b2s[o], b = a;
            const float4* wr = reinterpret_cast<const float4*>(&w2s[o * D]);
#pragma unroll
            for (int h4 = 0; h4 < D / 4; ++h4) {
                const float4 w = wr[h4];
                a = fmaf(w.x, hidA[h4 * 4 + 0], fmaf(w.y, hidA[h4 * 4 + 1], fmaf(w.z, hidA[h4 * 4 + 2], fmaf(w.w, hidA[h4 * 4 + 3], a))));
                b = fmaf(w.x, hidB[h4 * 4 + 0], fmaf(w.y, hidB[h4 * 4 + 1], fmaf(w.z, hidB[h4 * 4 + 2], fmaf(w.w, hidB[h4 * 4 + 3], b))));
            }
            float m = fmaxf(va ? a : -INFINITY, vb ? b : -INFINITY);
            m = warp_max(m);
            if (lane == 0) {
                float* dst = p.out + pix * D + o;
                *dst = r0 == 0 ? m : fmaxf(*dst, m);
            }
        }
    }
}

__global__ void det_parse_kernel(const float* __restrict__ scores, const float* __restrict__ scales,
                                 const float* __restrict__ offsets, int hs, int ws, int pool, float* __restrict__ out,
                                 long long* __restrict__ out_idx) {
    const int qi = blockIdx.x;
    const int n = hs * ws;
    const float* sc = scores + (long long)qi * n;
    // first-max argmax (torch.argmax returns the lowest index among ties and treats NaN as the
    // maximum, detector.py:91): candidate (v, i) beats (bv, bi) under that order
    auto beats = [](float v, int i, float bv, int bi) {
        const bool vn = v != v, bn = bv != bv;
        if (vn || bn) return vn && (!bn || i < bi);
        return v > bv || (v == bv && i < bi);
    };
    float bv = -INFINITY; int bi = 0x7fffffff;      // sentinel: loses to every real element (even -inf) on the index
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = sc[i];
        if (beats(v, i, bv, bi)) { bv = v; bi = i; }
    }
    __shared__ float sv[256]; __shared__ int si[256];
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const float v = sv[threadIdx.x + o]; const int i = si[threadIdx.x + o];
            if (beats(v, i, sv[threadIdx.x], si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = i; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int idx = min(max(si[0], 0), n - 1);
        const int y = idx / ws, x = idx % ws;
        const float ox = offsets[((long long)qi * n + idx) * 2 + 0], oy = offsets[((long long)qi * n + idx) * 2 + 1];
        out[qi * 4 + 0] = ((float)x + ox + 0.5f) * (float)pool - 0.5f;
        out[qi * 4 + 1] = ((float)y + oy + 0.5f) * (float)pool - 0.5f;
        out[qi * 4 + 2] = exp2f(scales[(long long)qi * n + idx]);
        out[qi * 4 + 3] = sv[0];
        out_idx[qi] = idx;
    }
}

// Second half of the row-decomposed sliding inner product (detector.py:222-224 F.conv2d(que, ref) with
// the reference features as k x k kernels):  out[y,x,r] = sum_ky partial[y + ky, x, ky*rfn + r], where
// partial = (1 x k convolution with k*rfn output channels (ky-major), zero padding k/2 in BOTH axes) holds,
// for every input row y' = y + ky - k/2, the contribution of kernel row ky.  The 1 x k form turns the
// N = rfn (32) GEMM of the direct formulation into an N = k*rfn (480) one: full-width tensor-core tiles.
__global__ void det_corr_rowsum_kernel(const float4* __restrict__ partial, float4* __restrict__ out, long long total,
                                       int H, int W, int k, int rfn4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)(i % rfn4);
    long long t = i / rfn4;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const long long q = t / H;
    const long long row_stride = (long long)W * k * rfn4;            // float4 per partial row y'
    const float4* p = partial + (q * (H + k - 1) + y) * row_stride + (long long)x * k * rfn4 + r;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = 0; ky < k; ++ky) {
        const float4 v = __ldg(p + ky * row_stride + ky * rfn4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[i] = acc;
}

}  // namespace g6d

using namespace g6d;

extern "C" int g6d_det_corr_rowsum(const float* partial, float* out, int qn, int H, int W, int k, int rfn,
                                   g6d_stream_t stream) {
    G6D_REQUIRE(partial && out && qn > 0 && H > 0 && W > 0 && k > 0 && rfn > 0 && (rfn & 3) == 0,
                "g6d_det_corr_rowsum: bad args (rfn must be a multiple of 4)");
    const long long total = (long long)qn * H * W * (rfn / 4);
    det_corr_rowsum_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(partial), reinterpret_cast<float4*>(out), total, H, W, k, rfn / 4);
    G6D_CHECK_LAUNCH("g6d_det_corr_rowsum");
    return G6D_OK;
}

extern "C" int g6d_det_score_fuse(const g6d_det_maps* host_maps, int qn, const float* w1, const float* b1,
                                  const float* w2, const float* b2, float* out, g6d_stream_t stream) {
    G6D_REQUIRE(host_maps && w1 && b1 && w2 && b2 && out && qn > 0, "g6d_det_score_fuse: bad args");
    G6D_REQUIRE(host_maps->n_scales >= 1 && host_maps->n_scales <= G6D_DET_MAX_SCALES && host_maps->rfn > 0,
                "g6d_det_score_fuse: bad map table");
    DetFuseParams p;
    p.maps = *host_maps; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.out = out; p.qn = qn;
    const long long npix = (long long)qn * host_maps->hs * host_maps->ws;
    const int grid = ceil_div(npix, 4);
    cudaStream_t st = as_stream(stream);
    switch (host_maps->n_scales) {
        case 1: det_score_fuse_kernel<3><<<grid, 128, 0, st>>>(p); break;
        case 2: det_score_fuse_kernel<6><<<grid, 128, 0, st>>>(p); break;
        case 3: det_score_fuse_kernel<9><<<grid, 128, 0, st>>>(p); break;
        case 4: det_score_fuse_kernel<12><<<grid, 128, 0, st>>>(p); break;
        case 5: det_score_fuse_kernel<15><<<grid, 128, 0, st>>>(p); break;
        case 6: det_score_fuse_kernel<18><<<grid, 128, 0, st>>>(p); break;
        default:
            set_error("g6d_det_score_fuse: %d scales not instantiated (max 6)", host_maps->n_scales);
            return G6D_EINVAL;
    }
    G6D_CHECK_LAUNCH("g6d_det_score_fuse");
    return G6D_OK;
}

extern "C" int g6d_det_parse(const float* scores, const float* scales, const float* offsets, int qn, int hs, int ws,
                             int pool_ratio, float* out, long long* out_idx, g6d_stream_t stream) {
    G6D_REQUIRE(scores && scales && offsets && out && out_idx && qn > 0 && hs > 0 && ws > 0, "g6d_det_parse: bad args");
    det_parse_kernel<<<qn, 256, 0, as_stream(stream)>>>(scores, scales, offsets, hs, ws, pool_ratio, out, out_idx);
    G6D_CHECK_LAUNCH("g6d_det_parse");
    return G6D_OK;
}
