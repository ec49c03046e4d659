// Layout, resize, pooling, normalisation kernels (all HBM/L2-bound streaming passes over
// channels-last fp32 data; float4-vectorised where the channel count allows).
#include "common.cuh"

namespace g6d {

__constant__ float kMean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float kStd[3] = {0.229f, 0.224f, 0.225f};

// u8 [n,3] -> f32 [n,out_c] (out_c = 3 or 4; channel 3 is zero padding for the conv loader)
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, long long n, int out_c,
                                     int norm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = (float)img[i * 3 + c] / 255.f;
        if (norm) x = (x - kMean[c]) / kStd[c];
        v[c] = x;
    }
    if (out_c == 4) {
        *reinterpret_cast<float4*>(out + i * 4) = make_float4(v[0], v[1], v[2], 0.f);
    } else {
        out[i * 3 + 0] = v[0]; out[i * 3 + 1] = v[1]; out[i * 3 + 2] = v[2];
    }
}

// f32 [n,in_c] (first 3 channels) -> normalised [n,out_c]
__global__ void imagenet_norm_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int in_c,
                                     int out_c) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[i * out_c + c] = (in[i * in_c + c] - kMean[c]) / kStd[c];
    if (out_c == 4) out[i * 4 + 3] = 0.f;
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int out_c) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = in + (long long)n * C * HW;
    float* dst = out + (long long)n * HW * out_c;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? src[(long long)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + threadIdx.x;
        if (p < HW && c < out_c) dst[(long long)p * out_c + c] = tile[threadIdx.x][j];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int in_c) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = in + (long long)n * HW * in_c;
    float* dst = out + (long long)n * C * HW;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? src[(long long)p * in_c + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + threadIdx.x;
        if (p < HW && c < C) dst[(long long)c * HW + p] = tile[threadIdx.x][j];
    }
}

// PyTorch upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0,
// i1 = min(i0+1, in-1), scale = in/out computed in fp32.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i0 = i0 > in_size - 1 ? in_size - 1 : i0;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

__global__ void resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int Hi, int Wi,
                                       int Ho, int Wo, int C, int ocs, int oco) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Ho * Wo * C;
    if (i >= total) return;
    const int c = (int)(i % C);
    long long r = i / C;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    int y0, y1, x0, x1; float ly, lx;
    bilinear_src(yo, sh, Hi, y0, y1, ly);
    bilinear_src(xo, sw, Wi, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* b = in + (long long)n * Hi * Wi * C + c;
    const float v00 = b[((long long)y0 * Wi + x0) * C], v01 = b[((long long)y0 * Wi + x1) * C];
    const float v10 = b[((long long)y1 * Wi + x0) * C], v11 = b[((long long)y1 * Wi + x1) * C];
    // same association as ATen: hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11)
    out[(((long long)n * Ho + yo) * Wo + xo) * ocs + oco + c] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

__global__ void resize_nearest_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int Hi, int Wi,
                                      int Ho, int Wo, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Ho * Wo * C;
    if (i >= total) return;
    const int c = (int)(i % C);
    long long r = i / C;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    int yi = (int)floorf((float)yo * sh); yi = yi > Hi - 1 ? Hi - 1 : yi;
    int xi = (int)floorf((float)xo * sw); xi = xi > Wi - 1 ? Wi - 1 : xi;
    out[i] = in[(((long long)n * Hi + yi) * Wi + xi) * C + c];
}

__global__ void maxpool2x2_kernel(const float4* __restrict__ in, float4* __restrict__ out, int N, int H, int W, int C4) {
    const int Ho = H / 2, Wo = W / 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Ho * Wo * C4;
    if (i >= total) return;
    const int c = (int)(i % C4);
    long long r = i / C4;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const float4* b = in + (((long long)n * H + 2 * yo) * W + 2 * xo) * C4 + c;
    const float4 a0 = b[0], a1 = b[C4], a2 = b[(long long)W * C4], a3 = b[(long long)W * C4 + C4];
    float4 m;
    m.x = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
    m.y = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
    m.z = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
    m.w = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
    out[i] = m;
}

// one warp per row of C channels: y = x / max(||x||, eps)
__global__ void l2norm_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* src = reinterpret_cast<const float4*>(in + row * C);
    float4* dst = reinterpret_cast<float4*>(out + row * C);
    float ss = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
        const float4 v = src[i];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float d = fmaxf(sqrtf(ss), eps);
    for (int i = lane; i < C / 4; i += 32) {
        float4 v = src[i];
        v.x /= d; v.y /= d; v.z /= d; v.w /= d;
        dst[i] = v;
    }
}

__global__ void affine_act_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C4,
                                  long long rows_per_group, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int act, int ics, int ico, int ocs, int oco) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C4) return;
    const int c = (int)(i % C4) * 4;
    const long long r = i / C4;
    const long long g = r / rows_per_group;
    float4 v = *reinterpret_cast<const float4*>(in + r * ics + ico + c);
    const float4 s = *reinterpret_cast<const float4*>(scale + g * C4 * 4 + c);
    const float4 b = *reinterpret_cast<const float4*>(shift + g * C4 * 4 + c);
    v.x = fmaf(v.x, s.x, b.x); v.y = fmaf(v.y, s.y, b.y); v.z = fmaf(v.z, s.z, b.z); v.w = fmaf(v.w, s.w, b.w);
    if (act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(out + r * ocs + oco + c) = v;
}

__global__ void avgpool_affine_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_out,
                                      int spatial, int C, long long rows_per_group, const float* __restrict__ scale,
                                      const float* __restrict__ shift, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out * C) return;
    const int c = (int)(i % C);
    const long long o = i / C;
    float acc = 0.f;
    for (int s = 0; s < spatial; ++s) {
        const long long r = o * spatial + s;
        float v = in[r * C + c];
        if (scale) {
            const long long g = r / rows_per_group;
            v = fmaf(v, scale[g * C + c], shift[g * C + c]);
        }
        if (act == 1) v = fmaxf(v, 0.f);
        acc += v;
    }
    out[i] = acc / (float)spatial;
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                           long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

// ---- instance-norm statistics: per (group, channel) sum and sum of squares in double
__global__ void in_stats_partial_kernel(const float* __restrict__ x, long long rows, int C, int cstride, int coff,
                                        long long rows_per_group, int rows_per_block, double* __restrict__ ws) {
    // block handles rows [r0, r0+rows_per_block) of ONE group (grid.y = group), all channels strided by threads
    const long long g = blockIdx.y;
    const long long r0 = g * rows_per_group + (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, (g + 1) * rows_per_group);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        float f1 = 0.f, f2 = 0.f;
        int cnt = 0;
        for (long long r = r0; r < r1; ++r) {
            const float v = x[r * cstride + coff + c];
            f1 += v; f2 = fmaf(v, v, f2);
            if (++cnt == 32) { s1 += f1; s2 += f2; f1 = f2 = 0.f; cnt = 0; }
        }
        s1 += f1; s2 += f2;
        atomicAdd(&ws[(g * C + c) * 2 + 0], s1);
        atomicAdd(&ws[(g * C + c) * 2 + 1], s2);
    }
}

__global__ void in_stats_final_kernel(const double* __restrict__ ws, long long n, long long rows_per_group, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cnt = (double)rows_per_group;
    const double mean = ws[i * 2] / cnt;
    double var = ws[i * 2 + 1] / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    scale[i] = (float)rstd;
    shift[i] = (float)(-mean * rstd);
}

}  // namespace g6d

using namespace g6d;

extern "C" int g6d_preprocess_u8(const uint8_t* img, float* out, long long n_pixels, int out_c, int imagenet_norm,
                                 g6d_stream_t stream) {
    G6D_REQUIRE(img && out && n_pixels > 0 && (out_c == 3 || out_c == 4), "g6d_preprocess_u8: bad args");
    preprocess_u8_kernel<<<ceil_div(n_pixels, 256), 256, 0, as_stream(stream)>>>(img, out, n_pixels, out_c,
                                                                                 imagenet_norm);
    G6D_CHECK_LAUNCH("g6d_preprocess_u8");
    return G6D_OK;
}

extern "C" int g6d_imagenet_norm(const float* in, float* out, long long n_pixels, int in_c, int out_c,
                                 g6d_stream_t stream) {
    G6D_REQUIRE(in && out && n_pixels > 0 && (in_c == 3 || in_c == 4) && (out_c == 3 || out_c == 4),
                "g6d_imagenet_norm: bad args");
    imagenet_norm_kernel<<<ceil_div(n_pixels, 256), 256, 0, as_stream(stream)>>>(in, out, n_pixels, in_c, out_c);
    G6D_CHECK_LAUNCH("g6d_imagenet_norm");
    return G6D_OK;
}

extern "C" int g6d_nchw_to_nhwc(const float* in, float* out, int N, int C, int H, int W, int out_c,
                                g6d_stream_t stream) {
    G6D_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0 && out_c >= C, "g6d_nchw_to_nhwc: bad args");
    dim3 grid(ceil_div((long long)H * W, 32), ceil_div(out_c, 32), N);
    nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, out, C, H * W, out_c);
    G6D_CHECK_LAUNCH("g6d_nchw_to_nhwc");
    return G6D_OK;
}

extern "C" int g6d_nhwc_to_nchw(const float* in, float* out, int N, int C, int H, int W, int in_c,
                                g6d_stream_t stream) {
    G6D_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0 && in_c >= C, "g6d_nhwc_to_nchw: bad args");
    dim3 grid(ceil_div((long long)H * W, 32), ceil_div(C, 32), N);
    nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, out, C, H * W, in_c);
    G6D_CHECK_LAUNCH("g6d_nhwc_to_nchw");
    return G6D_OK;
}

extern "C" int g6d_resize_bilinear(const float* in, float* out, int N, int Hi, int Wi, int Ho, int Wo, int C,
                                   int out_cstride, int out_coff, g6d_stream_t stream) {
    G6D_REQUIRE(in && out && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0 && out_coff + C <= out_cstride,
                "g6d_resize_bilinear: bad args");
    const long long total = (long long)N * Ho * Wo * C;
    resize_bilinear_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(in, out, N, Hi, Wi, Ho, Wo, C,
                                                                                out_cstride, out_coff);
    G6D_CHECK_LAUNCH("g6d_resize_bilinear");
    return G6D_OK;
}

extern "C" int g6d_resize_nearest(const float* in, float* out, int N, int Hi, int Wi, int Ho, int Wo, int C,
                                  g6d_stream_t stream) {
    G6D_REQUIRE(in && out && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0, "g6d_resize_nearest: bad args");
    const long long total = (long long)N * Ho * Wo * C;
    resize_nearest_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(in, out, N, Hi, Wi, Ho, Wo, C);
    G6D_CHECK_LAUNCH("g6d_resize_nearest");
    return G6D_OK;
}

extern "C" int g6d_maxpool2x2(const float* in, float* out, int N, int H, int W, int C, g6d_stream_t stream) {
    G6D_REQUIRE(in && out && N > 0 && H >= 2 && W >= 2 && (C & 3) == 0, "g6d_maxpool2x2: need H, W >= 2 and C%%4==0");
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    maxpool2x2_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), N, H, W, C / 4);
    G6D_CHECK_LAUNCH("g6d_maxpool2x2");
    return G6D_OK;
}

extern "C" int g6d_l2norm_channels(const float* in, float* out, long long rows, int C, float eps,
                                   g6d_stream_t stream) {
    G6D_REQUIRE(in && out && rows > 0 && C > 0 && (C & 3) == 0, "g6d_l2norm_channels: bad args");
    l2norm_kernel<<<ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(in, out, rows, C, eps);
    G6D_CHECK_LAUNCH("g6d_l2norm_channels");
    return G6D_OK;
}

extern "C" int g6d_affine_act(const float* in, float* out, long long rows, int C, long long rows_per_group,
                              const float* scale, const float* shift, int act, int in_cstride, int in_coff,
                              int out_cstride, int out_coff, g6d_stream_t stream) {
    G6D_REQUIRE(in && out && scale && shift && rows > 0 && rows_per_group > 0 && (C & 3) == 0 &&
                    (in_cstride & 3) == 0 && (in_coff & 3) == 0 && (out_cstride & 3) == 0 && (out_coff & 3) == 0,
                "g6d_affine_act: bad args (all channel counts/offsets must be multiples of 4)");
    affine_act_kernel<<<ceil_div(rows * (C / 4), 256), 256, 0, as_stream(stream)>>>(
        in, out, rows, C / 4, rows_per_group, scale, shift, act, in_cstride, in_coff, out_cstride, out_coff);
    G6D_CHECK_LAUNCH("g6d_affine_act");
    return G6D_OK;
}

extern "C" int g6d_avgpool_affine(const float* in, float* out, long long n_out, int spatial, int C,
                                  long long rows_per_group, const float* scale, const float* shift, int act,
                                  g6d_stream_t stream) {
    G6D_REQUIRE(in && out && n_out > 0 && spatial > 0 && C > 0 && (!scale || (shift && rows_per_group > 0)),
                "g6d_avgpool_affine: bad args");
    avgpool_affine_kernel<<<ceil_div(n_out * C, 256), 256, 0, as_stream(stream)>>>(in, out, n_out, spatial, C,
                                                                                   rows_per_group, scale, shift, act);
    G6D_CHECK_LAUNCH("g6d_avgpool_affine");
    return G6D_OK;
}

extern "C" int g6d_add(const float* a, const float* b, float* out, long long n, g6d_stream_t stream) {
    G6D_REQUIRE(a && b && out && n > 0, "g6d_add: bad args");
    add_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(a, b, out, n);
    G6D_CHECK_LAUNCH("g6d_add");
    return G6D_OK;
}

static int instnorm_partial_impl(const float* x, long long rows, int C, int cstride, int coff, long long rows_per_group,
                                 double* ws, cudaStream_t st) {
    const long long groups = rows / rows_per_group;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * groups * C, st);
    if (e != cudaSuccess) { set_error("g6d_instnorm: memset failed: %s", cudaGetErrorString(e)); return G6D_ECUDA; }
    // enough blocks per group to fill the machine, at least 32 rows each
    long long want = (4ll * kNumSMs + groups - 1) / groups;
    long long rpb = (rows_per_group + want - 1) / want;
    rpb = rpb < 32 ? 32 : rpb;
    const int bx = ceil_div(rows_per_group, rpb);
    const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
    in_stats_partial_kernel<<<dim3(bx, (unsigned)groups), threads, 0, st>>>(x, rows, C, cstride, coff, rows_per_group,
                                                                           (int)rpb, ws);
    G6D_CHECK_LAUNCH("g6d_instnorm_stats(partial)");
    return G6D_OK;
}

extern "C" int g6d_instnorm_partial(const float* x, long long rows, int C, int cstride, int coff,
                                    long long rows_per_group, double* ws, g6d_stream_t stream) {
    G6D_REQUIRE(x && ws && rows > 0 && C > 0 && rows_per_group > 0 && rows % rows_per_group == 0 && coff + C <= cstride,
                "g6d_instnorm_partial: bad args");
    G6D_REQUIRE(rows / rows_per_group <= 65535, "g6d_instnorm_partial: too many groups");
    return instnorm_partial_impl(x, rows, C, cstride, coff, rows_per_group, ws, as_stream(stream));
}

extern "C" int g6d_instnorm_finalize(const double* ws, long long groups, int C, long long count, float eps,
                                     float* scale, float* shift, g6d_stream_t stream) {
    G6D_REQUIRE(ws && scale && shift && groups > 0 && C > 0 && count > 0, "g6d_instnorm_finalize: bad args");
    in_stats_final_kernel<<<ceil_div(groups * C, 256), 256, 0, as_stream(stream)>>>(ws, groups * C, count, eps, scale, shift);
    G6D_CHECK_LAUNCH("g6d_instnorm_stats(final)");
    return G6D_OK;
}

extern "C" int g6d_instnorm_stats(const float* x, long long rows, int C, int cstride, int coff,
                                  long long rows_per_group, float eps, float* scale, float* shift, double* ws,
                                  g6d_stream_t stream) {
    G6D_REQUIRE(x && scale && shift && ws && rows > 0 && C > 0 && rows_per_group > 0 && rows % rows_per_group == 0 &&
                    coff + C <= cstride,
                "g6d_instnorm_stats: bad args");
    const long long groups = rows / rows_per_group;
    G6D_REQUIRE(groups <= 65535, "g6d_instnorm_stats: too many groups");
    int rc = instnorm_partial_impl(x, rows, C, cstride, coff, rows_per_group, ws, as_stream(stream));
    if (rc != G6D_OK) return rc;
    return g6d_instnorm_finalize(ws, groups, C, rows_per_group, eps, scale, shift, stream);
}
