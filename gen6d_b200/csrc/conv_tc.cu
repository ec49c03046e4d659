// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05), fp32-faithful via a
// 3xTF32 operand split:  A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  with x_hi = tf32(x),
// x_lo = tf32(x - x_hi); the dropped A_lo*B_lo term is O(2^-22) relative, so index selections
// (detection cell, viewpoint) stay bit-exact against the fp32 reference while the contraction
// runs on the tensor pipe instead of the FFMA pipe.
//
// GEMM view (same as conv_ffma.cu): M = B*Do*Ho*Wo, N = Cout, K = taps*Cin, channels-last.
// One CTA computes a 128 x BLOCK_N output tile (UMMA M=128, cta_group::1, accumulator in TMEM).
// Warp roles (320 threads):
//   warps 0-7  A producers: gather the im2col rows of the K-block from global memory, apply the
//              folded InstanceNorm(+ReLU) / selector q(.)ref prologue to in-bounds elements, split
//              into tf32 hi/lo and write both tiles into shared memory in the canonical K-major
//              SWIZZLE_128B layout the UMMA descriptor expects; then the epilogue (TMEM -> regs
//              -> bias/activation -> global).
//   warp 8     B producer: one elected lane issues TMA (cp.async.bulk.tensor.2d) loads of the
//              pre-split weight tiles W_hi / W_lo [Cout, K] (K-major, 128B swizzle) signalling an
//              mbarrier with complete_tx.
//   warp 9     MMA issuer: one elected lane waits on the full barriers and issues 12
//              tcgen05.mma.kind::tf32 per K-block (4 K-steps x 3 split terms); tcgen05.commit
//              releases the stage back to the producers; the last commit signals the epilogue.
//              This warp also owns the TMEM allocation.
// Small-M / huge-K problems are split along K over blockIdx.z into a workspace (same
// deterministic reduce kernel as the FFMA path).
#include <stdlib.h>

#include "tc_common.cuh"

namespace g6d {

constexpr int TC_BM = 128;       // rows per tile (UMMA M)
constexpr int TC_BK = 32;        // fp32 elements per K-block = one 128-byte swizzle row
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_MAX_KB_PER_SPLIT = 64;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 2) * 32;

struct ConvTcP {
    const float* x; const float* bias; const float* ps; const float* pb;
    float* y; float* ws;
    int B, D, H, W, Cin, ics, ico, Cout, kd, kh, kw, stride, pd, ph, pw, Do, Ho, Wo, ocs, oco, pro, act;
    long long group_rows;
    int M, K, kblocks, splits, kb_per_split;
};


template <int BN> struct TcCfg {
    static constexpr int A_BYTES = TC_BM * 128;            // one A tile (hi or lo)
    static constexpr int B_BYTES = BN * 128;               // one B tile (hi or lo)
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    // TMEM accumulators: NMAIN round-robin chains for the main (hi*hi) term + one for the cross terms.
    // Each tensor-core accumulate truncates to fp32; spreading the K-blocks over several shorter,
    // smaller-magnitude chains (summed in fp32 round-to-nearest by the epilogue) divides the
    // resulting bias on same-sign data by ~NMAIN at no cost.
    static constexpr int NMAIN = BN == 32 ? 7 : (BN == 256 ? 1 : 3);
    static constexpr int TMEM_COLS = (NMAIN + 1) * BN;             // 256 / 256 / 512 / 512 columns
};

// ------------------------------------------------------------------------------------------ kernel
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const ConvTcP p, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // carve: [stages x (A_hi | A_lo | B_hi | B_lo)] 1024-aligned, then barriers
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_base = base + STAGES * Cfg::STAGE_BYTES;
    auto a_hi = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
    auto a_lo = [&](int s) { return base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
    auto b_hi = [&](int s) { return base + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; };
    auto b_lo = [&](int s) { return base + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };
    auto full_a = [&](int s) { return bar_base + 8 * s; };
    auto full_b = [&](int s) { return bar_base + 8 * (STAGES + s); };
    auto empty = [&](int s) { return bar_base + 8 * (2 * STAGES + s); };
    const uint32_t tmem_full = bar_base + 8 * (3 * STAGES);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 1));
    __shared__ int4 row_info[TC_BM];     // (b, z0, y0, x0) of each tile row; b < 0 -> row beyond M

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_base = blockIdx.x * TC_BM;
    const int n_base = blockIdx.y * BN;
    const int split = blockIdx.z;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.kblocks, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;

    // ---- one-time setup
    if (threadIdx.x < TC_BM) {
        int m = m_base + threadIdx.x;
        int4 ri = make_int4(-1, 0, 0, 0);
        if (m < p.M) {
            int xo = m % p.Wo; m /= p.Wo;
            int yo = m % p.Ho; m /= p.Ho;
            int zo = m % p.Do; m /= p.Do;
            ri = make_int4(m, zo * p.stride - p.pd, yo * p.stride - p.ph, xo * p.stride - p.pw);
        }
        row_info[threadIdx.x] = ri;
    }
    if (warp == TC_PRODUCER_WARPS && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_a(s), TC_PRODUCER_WARPS);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    }
    if (warp == TC_PRODUCER_WARPS + 1) {   // TMEM allocation (warp-collective)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        // =============================== A producers ===============================
        const int chunk = threadIdx.x & 7;          // 16-byte chunk (4 floats) within the 128-byte K row
        const int rg = threadIdx.x >> 3;            // 0..31; rows rg + 32*i
        // Per-row constants: pointer to the tap-(0,0,0) element of this thread's chunk (may point
        // before the tensor for padded rows; only dereferenced when the tap is in bounds), the
        // base input coordinates, and the prologue operand rows.
        const float* rowp[4];
        const float* scp[4];
        const float* shp[4];
        int rz[4], ry[4], rx[4];
        bool rv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int4 ri = row_info[rg + 32 * i];
            rv[i] = ri.x >= 0;
            rz[i] = ri.y; ry[i] = ri.z; rx[i] = ri.w;
            const long long b = rv[i] ? ri.x : 0;
            const long long sp0 = ((long long)ri.y * p.H + ri.z) * p.W + ri.w;
            rowp[i] = p.x + (b * p.D * p.H * p.W + sp0) * p.ics + p.ico + chunk * 4;
            if (p.pro == G6D_PRO_CORR) {
                scp[i] = p.ps + sp0 * p.Cin + chunk * 4;
                shp[i] = p.pb + chunk * 4;
            } else {
                const long long g = b / p.group_rows;
                scp[i] = p.ps + g * p.Cin + chunk * 4;
                shp[i] = p.pb + g * p.Cin + chunk * 4;
            }
        }
        // K-block cursor (tap, channel), advanced incrementally: no divisions inside the loop
        int c0, kx, ky, kz;
        {
            const int k = kb_begin * TC_BK;
            int tap = 0;
            c0 = k;
            if (p.K != p.Cin) { tap = k / p.Cin; c0 = k - tap * p.Cin; }
            kx = tap % p.kw;
            const int tq = tap / p.kw;
            ky = tq % p.kh;
            kz = tq / p.kh;
        }
        auto advance = [&]() {
            c0 += TC_BK;
            if (c0 == p.Cin) {
                c0 = 0;
                if (++kx == p.kw) { kx = 0; if (++ky == p.kh) { ky = 0; ++kz; } }
            }
        };

        float4 cur[4], nxt[4];
        unsigned cur_ok = 0, nxt_ok = 0;            // bit i: row i's tap is in bounds
        int cur_c = 0, nxt_c = 0;
        long long cur_sp = 0, nxt_sp = 0;           // spatial offset of the tap (for the CORR prologue)

        auto issue_loads = [&]() {
            const long long tap_sp = ((long long)kz * p.H + ky) * p.W + kx;
            const long long off = tap_sp * p.ics + c0;
            nxt_ok = 0; nxt_c = c0; nxt_sp = tap_sp;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool inb = rv[i] && (unsigned)(rz[i] + kz) < (unsigned)p.D && (unsigned)(ry[i] + ky) < (unsigned)p.H &&
                                 (unsigned)(rx[i] + kx) < (unsigned)p.W;
                nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb) {
                    nxt[i] = __ldg(reinterpret_cast<const float4*>(rowp[i] + off));
                    nxt_ok |= 1u << i;
                }
            }
            advance();
        };

        if (nkb > 0) issue_loads();
        for (int it = 0; it < nkb; ++it) {
            const int s = it % STAGES;
            const uint32_t n_use = it / STAGES;
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
            cur_ok = nxt_ok; cur_c = nxt_c; cur_sp = nxt_sp;
            if (it + 1 < nkb) issue_loads();
            // prologue on in-bounds elements (zero padding stays zero)
            if (p.pro != G6D_PRO_NONE) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (cur_ok & (1u << i)) {
                        const long long so = p.pro == G6D_PRO_CORR ? cur_sp * p.Cin + cur_c : (long long)cur_c;
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(scp[i] + so));
                        const float4 sh = __ldg(reinterpret_cast<const float4*>(shp[i] + cur_c));
                        float4 v = cur[i];
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                        if (p.pro == G6D_PRO_AFFINE_RELU) {
                            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                        }
                        cur[i] = v;
                    }
                }
            }
            mbar_wait(empty(s), (n_use & 1) ^ 1, 1, it);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // 3xTF32 split with full-rate integer ops: hi = round-to-nearest (ties away) of the
                // fp32 mantissa to 10 bits (same as cvt.rna.tf32.f32 for finite values), lo = v - hi
                // (exact in fp32; the tensor core ignores its low 13 mantissa bits).
                const float4 v = cur[i];
                float4 hi, lo;
                hi.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u);
                hi.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u);
                hi.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u);
                hi.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u);
                lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                const int r = rg + 32 * i;
                const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);     // Swizzle<3,4,3>
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_hi(s) + off), "f"(hi.x), "f"(hi.y),
                             "f"(hi.z), "f"(hi.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_lo(s) + off), "f"(lo.x), "f"(lo.y),
                             "f"(lo.z), "f"(lo.w) : "memory");
            }
            fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
        }

        // =============================== epilogue ===============================
        mbar_wait(tmem_full, 0, 2, nkb);
        tc_fence_after();
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        const int m = m_base + row;
        constexpr int HALF = BN / 2;                    // warps 0-3: columns [0,HALF), warps 4-7: [HALF,BN)
        const int col0 = (warp >> 2) * HALF;
        const bool partial = p.splits > 1;
#pragma unroll
        for (int cc = 0; cc < HALF; cc += 16) {
            float accv[16];
            const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(col0 + cc);
            const int n_acc = nkb < Cfg::NMAIN ? nkb : Cfg::NMAIN;    // main chains that were written
#pragma unroll
            for (int a = 0; a <= Cfg::NMAIN; ++a) {
                const bool used = a == Cfg::NMAIN || a < n_acc;      // last = cross-term accumulator
                uint32_t r[16];
                if (used) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr + (uint32_t)(a * BN)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) accv[j] = a == 0 ? __uint_as_float(r[j]) : accv[j] + __uint_as_float(r[j]);
                }
            }
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n_base + col0 + cc + j;
                    if (n < p.Cout) {
                        float v = accv[j];                                           // main chains + cross terms
                        if (partial) {
                            p.ws[((long long)split * p.M + m) * p.Cout + n] = v;
                        } else {
                            if (p.bias) v += __ldg(p.bias + n);
                            p.y[(long long)m * p.ocs + p.oco + n] = tc_act(v, p.act);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp == TC_PRODUCER_WARPS) {
        // =============================== B producer (TMA) ===============================
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t n_use = it / STAGES;
                mbar_wait(empty(s), (n_use & 1) ^ 1, 3, it);
                mbar_expect_tx(full_b(s), 2 * Cfg::B_BYTES);
                const int k = (kb_begin + it) * TC_BK;
                tma_load_2d(b_hi(s), &map_hi, full_b(s), k, n_base);
                tma_load_2d(b_lo(s), &map_lo, full_b(s), k, n_base);
            }
        }
    } else {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t n_use = it / STAGES;
                mbar_wait(full_a(s), n_use & 1, 4, it);
                mbar_wait(full_b(s), n_use & 1, 5, it);
                tc_fence_after();
                const uint64_t dah = umma_desc_sw128(a_hi(s)), dal = umma_desc_sw128(a_lo(s));
                const uint64_t dbh = umma_desc_sw128(b_hi(s)), dbl = umma_desc_sw128(b_lo(s));
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 32) >> 4);   // +32 bytes of K per step, in 16-byte units
                    // The small cross terms get their own TMEM accumulator, the main term rotates over
                    // NMAIN accumulators (see TcCfg).
                    const uint32_t main_acc = tmem_acc + (uint32_t)((it % Cfg::NMAIN) * BN);
                    const uint32_t cross_acc = tmem_acc + (uint32_t)(Cfg::NMAIN * BN);
                    umma_tf32(cross_acc, dal + adv, dbh + adv, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(cross_acc, dah + adv, dbl + adv, idesc, 1u);
                    umma_tf32(main_acc, dah + adv, dbh + adv, idesc, (it >= Cfg::NMAIN || ks > 0) ? 1u : 0u);
                }
                umma_commit(empty(s));            // frees the stage once these MMAs have read it
            }
            umma_commit(tmem_full);               // accumulator complete -> epilogue
        }
    }
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ==========================================================================================
// conv_tc2_kernel: persistent version of conv_tc_kernel.
//   * one CTA per SM loops over (M tile, N tile, K split) work items, so there is no wave tail and
//     the per-CTA set-up (TMEM allocation, barrier init, descriptor prefetch) is paid once;
//   * two TMEM accumulator buffers: the MMA warp starts the next tile while four dedicated
//     epilogue warps drain the previous one (tmem_full / tmem_empty barriers);
//   * the eight producer warps keep a two-deep register prefetch (loads of K-blocks it+1, it+2 in
//     flight while K-block it is transformed and stored);
//   * the smem ring never drains between tiles (one global K-block counter).
// Warps: 0-7 A producers, 8 B producer (TMA), 9 MMA issuer + TMEM owner, 10-13 epilogue.
constexpr int TC2_THREADS = 14 * 32;
constexpr int TC2_PF = 12;          // weight-tile L2 prefetch distance (K-blocks)

template <int BN> struct Tc2Cfg {
    static constexpr int A_BYTES = TC_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (208 * 1024) / STAGE_BYTES > 6 ? 6 : (208 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int NMAIN = BN == 32 ? 7 : (BN == 64 ? 3 : 1);      // per accumulator buffer
    static constexpr int BUF_COLS = (NMAIN + 1) * BN;                     // 256
    static constexpr int TMEM_COLS = 2 * BUF_COLS;                        // 512: two buffers
};

struct Tc2Work { int m_tiles, n_tiles, total; };

template <int BN>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const ConvTcP p, const Tc2Work wk, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
    using Cfg = Tc2Cfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NMAIN = Cfg::NMAIN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_base = base + STAGES * Cfg::STAGE_BYTES;
    auto a_hi = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
    auto a_lo = [&](int s) { return base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
    auto b_hi = [&](int s) { return base + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; };
    auto b_lo = [&](int s) { return base + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };
    auto full_a = [&](int s) { return bar_base + 8 * s; };
    auto full_b = [&](int s) { return bar_base + 8 * (STAGES + s); };
    auto empty = [&](int s) { return bar_base + 8 * (2 * STAGES + s); };
    auto tmem_full = [&](int b) { return bar_base + 8 * (3 * STAGES + b); };
    auto tmem_empty = [&](int b) { return bar_base + 8 * (3 * STAGES + 2 + b); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_a(s), 8);          // the 8 producer warps
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    // work item -> (m tile, n tile, split); n fastest so neighbouring CTAs share the gathered A rows in L2
    auto decode = [&](int w, int& mt, int& nt, int& sp) {
        nt = w % wk.n_tiles; w /= wk.n_tiles;
        mt = w % wk.m_tiles;
        sp = w / wk.m_tiles;
    };

    if (warp < 8) {
        // =============================== A producers ===============================
        // All 8 warps fill every K-block (4 rows x one 16-byte chunk per thread) with a two-deep
        // register prefetch: the global loads of K-blocks it+1 and it+2 are in flight while
        // K-block it is transformed and stored.
        const int chunk = threadIdx.x & 7;
        const int r0 = threadIdx.x >> 3;                   // rows r0 + 32*j, j = 0..3
        int git = 0;                                       // global K-block counter of this CTA
        for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
            int mt, nt, sp;
            decode(w, mt, nt, sp);
            const int m_base = mt * TC_BM;
            const int kb_begin = sp * p.kb_per_split;
            const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
            int rb[4], rsp[4], rc[4];
            unsigned rvmask = 0;
            const long long plane_sz = (long long)p.D * p.H * p.W;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int m = m_base + r0 + 32 * j;
                const bool v = m < p.M;
                if (!v) m = 0;
                const int xo = m % p.Wo; m /= p.Wo;
                const int yo = m % p.Ho; m /= p.Ho;
                const int zo = m % p.Do; m /= p.Do;
                rb[j] = m;
                const int z = zo * p.stride - p.pd, y = yo * p.stride - p.ph, x = xo * p.stride - p.pw;
                rsp[j] = (z * p.H + y) * p.W + x;
                rc[j] = ((z + 8) << 24) | ((y + 8) << 12) | (x + 8);
                if (v) rvmask |= 1u << j;
            }
            int c0, kx, ky, kz;
            {
                const int k = kb_begin * TC_BK;
                int tap = 0; c0 = k;
                if (p.K != p.Cin) { tap = k / p.Cin; c0 = k - tap * p.Cin; }
                kx = tap % p.kw; const int tq = tap / p.kw; ky = tq % p.kh; kz = tq / p.kh;
            }
            // prefetch ring of 3 K-blocks: slot q holds K-block (it % 3 == q)
            float4 v[3][4];
            unsigned okm[3]; int kc[3]; int ksp[3];
            auto issue_loads = [&](int q) {
                const int tap_sp = (kz * p.H + ky) * p.W + kx;
                okm[q] = 0; kc[q] = c0; ksp[q] = tap_sp;
                const float* xb = p.x + p.ico + chunk * 4 + c0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int z = ((rc[j] >> 24) & 0xff) - 8 + kz, y = ((rc[j] >> 12) & 0xfff) - 8 + ky, x = (rc[j] & 0xfff) - 8 + kx;
                    const bool inb = ((rvmask >> j) & 1u) && (unsigned)z < (unsigned)p.D && (unsigned)y < (unsigned)p.H &&
                                     (unsigned)x < (unsigned)p.W;
                    v[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (inb) {
                        v[q][j] = __ldg(reinterpret_cast<const float4*>(xb + ((long long)rb[j] * plane_sz + rsp[j] + tap_sp) * p.ics));
                        okm[q] |= 1u << j;
                    }
                }
                c0 += TC_BK;
                if (c0 == p.Cin) { c0 = 0; if (++kx == p.kw) { kx = 0; if (++ky == p.kh) { ky = 0; ++kz; } } }
            };
            auto process = [&](int q, int it) {
                const int g_it = git + it;
                const int s = g_it % STAGES;
                const uint32_t n_use = g_it / STAGES;
                if (p.pro != G6D_PRO_NONE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (okm[q] & (1u << j)) {
                            float4 sc, sh;
                            if (p.pro == G6D_PRO_CORR) {
                                sc = __ldg(reinterpret_cast<const float4*>(p.ps + (long long)(rsp[j] + ksp[q]) * p.Cin + kc[q] + chunk * 4));
                                sh = __ldg(reinterpret_cast<const float4*>(p.pb + kc[q] + chunk * 4));
                            } else {
                                const long long g = rb[j] / (int)p.group_rows;       // 32-bit divide (host checks the range)
                                sc = __ldg(reinterpret_cast<const float4*>(p.ps + g * p.Cin + kc[q] + chunk * 4));
                                sh = __ldg(reinterpret_cast<const float4*>(p.pb + g * p.Cin + kc[q] + chunk * 4));
                            }
                            float4 x4 = v[q][j];
                            x4.x = fmaf(x4.x, sc.x, sh.x); x4.y = fmaf(x4.y, sc.y, sh.y);
                            x4.z = fmaf(x4.z, sc.z, sh.z); x4.w = fmaf(x4.w, sc.w, sh.w);
                            if (p.pro == G6D_PRO_AFFINE_RELU) {
                                x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
                            }
                            v[q][j] = x4;
                        }
                    }
                }
                mbar_wait(empty(s), (n_use & 1) ^ 1, 1, g_it);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 x4 = v[q][j];
                    float4 hi, lo;
                    hi.x = __uint_as_float((__float_as_uint(x4.x) + 0x1000u) & 0xFFFFE000u);
                    hi.y = __uint_as_float((__float_as_uint(x4.y) + 0x1000u) & 0xFFFFE000u);
                    hi.z = __uint_as_float((__float_as_uint(x4.z) + 0x1000u) & 0xFFFFE000u);
                    hi.w = __uint_as_float((__float_as_uint(x4.w) + 0x1000u) & 0xFFFFE000u);
                    lo.x = x4.x - hi.x; lo.y = x4.y - hi.y; lo.z = x4.z - hi.z; lo.w = x4.w - hi.w;
                    const int r = r0 + 32 * j;
                    const uint32_t so = r * 128 + ((chunk ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_hi(s) + so), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_lo(s) + so), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a(s));
            };
            // software pipeline, unrolled by 3 so the ring slots are compile-time register names
            if (nkb > 0) issue_loads(0);
            if (nkb > 1) issue_loads(1);
            for (int it = 0; it < nkb; it += 3) {
                if (it + 2 < nkb) issue_loads(2);
                process(0, it);
                if (it + 1 < nkb) {
                    if (it + 3 < nkb) issue_loads(0);
                    process(1, it + 1);
                }
                if (it + 2 < nkb) {
                    if (it + 4 < nkb) issue_loads(1);
                    process(2, it + 2);
                }
            }
            git += nkb;
        }
    } else if (warp == 8) {
        // =============================== B producer (TMA) ===============================
        if (lane == 0) {
            int git = 0;
            for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_begin = sp * p.kb_per_split;
                const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
                // The weight tiles of a small-M layer are touched once and come from HBM: with only
                // STAGES tiles in flight the ring is latency-bound (measured 5000 cycles per K-block
                // at M = 660).  An L2 prefetch running TC2_PF K-blocks ahead costs no shared memory.
                for (int it = 0; it < min(nkb, TC2_PF); ++it) {
                    tma_prefetch_2d(&map_hi, (kb_begin + it) * TC_BK, nt * BN);
                    tma_prefetch_2d(&map_lo, (kb_begin + it) * TC_BK, nt * BN);
                }
                for (int it = 0; it < nkb; ++it, ++git) {
                    const int s = git % STAGES;
                    if (it + TC2_PF < nkb) {
                        tma_prefetch_2d(&map_hi, (kb_begin + it + TC2_PF) * TC_BK, nt * BN);
                        tma_prefetch_2d(&map_lo, (kb_begin + it + TC2_PF) * TC_BK, nt * BN);
                    }
                    mbar_wait(empty(s), ((git / STAGES) & 1) ^ 1, 3, git);
                    mbar_expect_tx(full_b(s), 2 * Cfg::B_BYTES);
                    const int k = (kb_begin + it) * TC_BK;
                    tma_load_2d(b_hi(s), &map_hi, full_b(s), k, nt * BN);
                    tma_load_2d(b_lo(s), &map_lo, full_b(s), k, nt * BN);
                }
            }
        }
    } else if (warp == 9) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
            int git = 0, tile = 0;
            for (int w = blockIdx.x; w < wk.total; w += gridDim.x, ++tile) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_begin = sp * p.kb_per_split;
                const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
                const int buf = tile & 1;
                mbar_wait(tmem_empty(buf), ((tile >> 1) & 1) ^ 1, 6, tile);     // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t acc0 = tmem_acc + (uint32_t)(buf * Cfg::BUF_COLS);
                for (int it = 0; it < nkb; ++it, ++git) {
                    const int s = git % STAGES;
                    mbar_wait(full_a(s), (git / STAGES) & 1, 4, git);
                    mbar_wait(full_b(s), (git / STAGES) & 1, 5, git);
                    tc_fence_after();
                    const uint64_t dah = umma_desc_sw128(a_hi(s)), dal = umma_desc_sw128(a_lo(s));
                    const uint64_t dbh = umma_desc_sw128(b_hi(s)), dbl = umma_desc_sw128(b_lo(s));
                    const uint32_t main_acc = acc0 + (uint32_t)((it % NMAIN) * BN);
                    const uint32_t cross_acc = acc0 + (uint32_t)(NMAIN * BN);
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        umma_tf32(cross_acc, dal + adv, dbh + adv, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(cross_acc, dah + adv, dbl + adv, idesc, 1u);
                        umma_tf32(main_acc, dah + adv, dbh + adv, idesc, (it >= NMAIN || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(empty(s));
                }
                umma_commit(tmem_full(buf));
            }
        }
    } else {
        // =============================== epilogue (warps 10-13) ===============================
        const int quad = warp & 3;                     // TMEM lane quadrant = warp id % 4
        int tile = 0;
        for (int w = blockIdx.x; w < wk.total; w += gridDim.x, ++tile) {
            int mt, nt, sp;
            decode(w, mt, nt, sp);
            const int kb_begin = sp * p.kb_per_split;
            const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
            const int buf = tile & 1;
            mbar_wait(tmem_full(buf), (tile >> 1) & 1, 2, tile);
            tc_fence_after();
            const int m = mt * TC_BM + quad * 32 + lane;
            const int n_base = nt * BN;
            const bool partial = p.splits > 1;
            // 128-bit stores need 16-byte aligned rows: channel strides / offsets multiples of 4 floats
            const bool vec_ok = partial ? (p.Cout & 3) == 0
                                        : ((p.ocs & 3) == 0 && (p.oco & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
                                           (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
            const int n_acc = nkb < NMAIN ? nkb : NMAIN;
            const uint32_t tbase = tmem_acc + (uint32_t)(buf * Cfg::BUF_COLS) + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 16) {
                float accv[16];
#pragma unroll
                for (int a = 0; a <= NMAIN; ++a) {
                    const bool used = a == NMAIN || a < n_acc;
                    uint32_t r[16];
                    if (used) {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(tbase + (uint32_t)(a * BN + cc)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) accv[j] = a == 0 ? __uint_as_float(r[j]) : accv[j] + __uint_as_float(r[j]);
                    }
                }
                if (m < p.M) {
                    const int n0 = n_base + cc;
                    float* dst = partial ? p.ws + ((long long)sp * p.M + m) * p.Cout + n0
                                         : p.y + (long long)m * p.ocs + p.oco + n0;
                    const bool vec = vec_ok && n0 + 16 <= p.Cout;
                    if (vec) {      // 4 x 128-bit stores per thread instead of 16 scalar ones
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            float4 v = make_float4(accv[j4 * 4], accv[j4 * 4 + 1], accv[j4 * 4 + 2], accv[j4 * 4 + 3]);
                            if (!partial) {
                                if (p.bias) {
                                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j4);
                                    v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                                }
                                v.x = tc_act(v.x, p.act); v.y = tc_act(v.y, p.act); v.z = tc_act(v.z, p.act); v.w = tc_act(v.w, p.act);
                            }
                            reinterpret_cast<float4*>(dst)[j4] = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (n0 + j < p.Cout) {
                                float v = accv[j];
                                if (!partial) {
                                    if (p.bias) v += __ldg(p.bias + n0 + j);
                                    v = tc_act(v, p.act);
                                }
                                dst[j] = v;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(buf));
        }
    }
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ==========================================================================================
// conv_tc3_kernel: conv_tc2_kernel with the A operand in TENSOR MEMORY.
//
// Measured with tools/probes/mma_rate.cu on B200: a kind::tf32 M=128 MMA takes max(N/2, 32 + N/4)
// cycles when A comes from shared memory (the 4 KB A slice is re-read at 128 B/clk for every one of
// the three split terms) but N/2 cycles - the tensor floor - when A is in TMEM.  In the 3xTF32 scheme
// the shared-memory traffic of v2 (A stores + A re-reads x3 + B) is 160 KB per K-block at N = 128,
// i.e. ~1250 cycles against 768 cycles of tensor time; here the producers write their tf32 hi/lo
// halves straight from registers into TMEM with tcgen05.st and the MMAs are issued in the
// [D], [A], B-descriptor form, so shared memory only carries the TMA-fed weight tiles.
//   * producers: thread = one tile row (TMEM lane = 32*(warp%4) + lane), warp/4 = which 16 of the
//     K-block's 32 channels; 64 contiguous bytes per thread per K-block, two-deep register prefetch;
//   * TMEM: accumulators in columns [0, 256) (N=128: main + cross, single buffer; N=64/32: two
//     buffers), A ring of 4 stages x (32 hi + 32 lo) columns in [256, 512);
//   * B ring in shared memory (TMA, 128B swizzle) with its own, deeper set of stages.
// Warps: 0-7 A producers, 8 B producer (TMA), 9 MMA issuer + TMEM owner, 10-13 epilogue.
template <int BN, bool STAGED = false> struct Tc3Cfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int RAW_BYTES = TC_BM * 128;                         // one unsplit A tile (STAGED)
    static constexpr int BSTAGES = STAGED ? (BN == 128 ? 4 : (BN == 64 ? 6 : 8))
                                          : ((192 * 1024) / (2 * B_BYTES) > 8 ? 8 : (192 * 1024) / (2 * B_BYTES));
    static constexpr int RSTAGES = STAGED ? (BN == 128 ? 5 : (BN == 64 ? 6 : 8)) : 0;
    static constexpr int ASTAGES = 4;
    static constexpr int SMEM_BYTES = BSTAGES * 2 * B_BYTES + RSTAGES * RAW_BYTES + 1024 + 512;
    static constexpr int NMAIN = BN == 32 ? 3 : 1;
    static constexpr int NBUF = BN == 128 ? 1 : 2;
    static constexpr int BUF_COLS = (NMAIN + 1) * BN;                     // 256 / 128 / 128
    static constexpr int A_COL0 = 256;                                    // A ring: columns [256, 512)
    static constexpr int TMEM_COLS = 512;
    static_assert(NBUF * BUF_COLS <= A_COL0, "accumulators overlap the A ring");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// RAWB (G6D_CONV_TC_V=4): the weight tiles are the dominant L2 -> SM stream of this kernel (hi + lo =
// 2 x BN x 128 B per K-block, against 16 KB of gathered A); with RAWB only the UNSPLIT fp32 tile is
// loaded (map_hi = raw weights, map_lo unused).  The tensor core reads it as the hi operand (it
// ignores the 13 low mantissa bits, i.e. hi = trunc(w)) and the producer warps derive
// lo = rn_tf32(w - trunc(w)) from it in shared memory, halving the weight traffic.
//
// STAGED (G6D_CONV_TC_V=5): ncu showed that v3's thread-per-row gather (32 distinct 128-byte lines per
// LDG.128) costs 8x the L1 data-pipe wavefronts of a coalesced one and saturates that pipe (93 % of
// active cycles), just as the tensor core's operand reads + the producers' st.shared do in v2 (LSU 51 % +
// TC 50 %).  STAGED keeps the coalesced access of v2 but never holds the tile in registers: the
// producers issue cp.async (16 B, zero-fill for padding / rows >= M) into a ring of unsplit A tiles in
// shared memory, RSTAGES - 1 K-blocks ahead (cp.async.mbarrier.arrive tracks completion), and later
// read their own row back with bank-conflict-free swizzled ld.shared (4 wavefronts per LDS.128), split
// it and tcgen05.st it to TMEM.  Data-pipe budget per K-block at N = 128: A 128 (cp.async fill) + 128
// (LDS) wavefronts, B 256 (TMA fill) + 384 (MMA operand reads) = ~900, against ~1150-1280 in v2.
template <int BN, bool RAWB, int PF = 2, bool STAGED = false>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc3_kernel(const ConvTcP p, const Tc2Work wk, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
    using Cfg = Tc3Cfg<BN, STAGED>;
    constexpr int BS = Cfg::BSTAGES, AS = Cfg::ASTAGES, NMAIN = Cfg::NMAIN, NBUF = Cfg::NBUF, RS = Cfg::RSTAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t raw_base = base + BS * 2 * Cfg::B_BYTES;              // STAGED: ring of unsplit A tiles
    const uint32_t bar_base = raw_base + RS * Cfg::RAW_BYTES;
    auto b_hi = [&](int s) { return base + s * 2 * Cfg::B_BYTES; };
    auto b_lo = [&](int s) { return base + s * 2 * Cfg::B_BYTES + Cfg::B_BYTES; };
    auto full_a = [&](int s) { return bar_base + 8 * s; };
    auto empty_a = [&](int s) { return bar_base + 8 * (AS + s); };
    auto full_b = [&](int s) { return bar_base + 8 * (2 * AS + s); };
    auto empty_b = [&](int s) { return bar_base + 8 * (2 * AS + BS + s); };
    auto tmem_full = [&](int b) { return bar_base + 8 * (2 * AS + 2 * BS + b); };
    auto tmem_empty = [&](int b) { return bar_base + 8 * (2 * AS + 2 * BS + 2 + b); };
    auto raw_full = [&](int s) { return bar_base + 8 * (2 * AS + 2 * BS + 4 + s); };
    auto araw_full = [&](int s) { return bar_base + 8 * (2 * AS + 3 * BS + 4 + s); };
    auto araw_empty = [&](int s) { return bar_base + 8 * (2 * AS + 3 * BS + 4 + RS + s); };
    auto araw = [&](int s) { return raw_base + s * Cfg::RAW_BYTES; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + (bar_base - base) + 8 * (2 * AS + 3 * BS + 4 + 2 * RS));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8 && lane == 0) {
        for (int s = 0; s < AS; ++s) { mbar_init(full_a(s), 8); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < BS; ++s) { mbar_init(full_b(s), RAWB ? 8 : 1); mbar_init(empty_b(s), 1); mbar_init(raw_full(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
        for (int s = 0; s < RS; ++s) { mbar_init(araw_full(s), 256); mbar_init(araw_empty(s), 8); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        if (!RAWB) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_slot;

    auto decode = [&](int w, int& mt, int& nt, int& sp) {
        nt = w % wk.n_tiles; w /= wk.n_tiles;
        mt = w % wk.m_tiles;
        sp = w / wk.m_tiles;
    };

    if (warp < 8) {
        // =============================== A producers ===============================
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const int cofs = half * 16;                                   // this thread's 16 channels of the K-block
        const uint32_t lane_addr = tmem0 + ((uint32_t)(quad * 32) << 16) + (uint32_t)(Cfg::A_COL0 + cofs);
        int git = 0;
        if constexpr (STAGED) {
            // copy mapping (coalesced, as v2): 16-byte chunk `cchunk` of rows cr0 + 32*j; process mapping: own row
            const int cchunk = threadIdx.x & 7, cr0 = threadIdx.x >> 3;
            constexpr int DEPTH = RS - 1;                                 // K-blocks of cp.async in flight
            for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_begin = sp * p.kb_per_split;
                const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
                const long long plane_sz = (long long)p.D * p.H * p.W;
                // ---- the 4 rows this thread copies
                int rb[4], rsp4[4], rc[4];
                unsigned rvmask = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int m = mt * TC_BM + cr0 + 32 * j;
                    const bool v = m < p.M;
                    if (!v) m = 0;
                    const int xo = m % p.Wo; m /= p.Wo;
                    const int yo = m % p.Ho; m /= p.Ho;
                    const int zo = m % p.Do; m /= p.Do;
                    rb[j] = m;
                    const int z = zo * p.stride - p.pd, y = yo * p.stride - p.ph, x = xo * p.stride - p.pw;
                    rsp4[j] = (z * p.H + y) * p.W + x;
                    rc[j] = ((z + 8) << 24) | ((y + 8) << 12) | (x + 8);
                    if (v) rvmask |= 1u << j;
                }
                // ---- the row this thread transforms
                int m = mt * TC_BM + row;
                const bool rvalid = m < p.M;
                if (!rvalid) m = 0;
                const int xo = m % p.Wo; m /= p.Wo;
                const int yo = m % p.Ho; m /= p.Ho;
                const int zo = m % p.Do; m /= p.Do;
                const int z0 = zo * p.stride - p.pd, y0 = yo * p.stride - p.ph, x0 = xo * p.stride - p.pw;
                const int rsp = (z0 * p.H + y0) * p.W + x0;
                const long long grp = m / (int)p.group_rows;
                // two cursors over (tap, channel block): one for the copies (ahead), one for the transforms
                int ic0, ikx, iky, ikz, pc0, pkx, pky, pkz;
                {
                    const int k = kb_begin * TC_BK;
                    int tap = 0; ic0 = k;
                    if (p.K != p.Cin) { tap = k / p.Cin; ic0 = k - tap * p.Cin; }
                    ikx = tap % p.kw; const int tq = tap / p.kw; iky = tq % p.kh; ikz = tq / p.kh;
                    pc0 = ic0; pkx = ikx; pky = iky; pkz = ikz;
                }
                auto issue = [&](int it) {
                    const int g_it = git + it;
                    const int s = g_it % RS;
                    mbar_wait(araw_empty(s), ((g_it / RS) & 1) ^ 1, 8, g_it);       // all 8 warps have read the previous tile
                    const int tap_sp = (ikz * p.H + iky) * p.W + ikx;
                    const float* xb = p.x + p.ico + cchunk * 4 + ic0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int z = ((rc[j] >> 24) & 0xff) - 8 + ikz, y = ((rc[j] >> 12) & 0xfff) - 8 + iky, x = (rc[j] & 0xfff) - 8 + ikx;
                        const bool inb = ((rvmask >> j) & 1u) && (unsigned)z < (unsigned)p.D && (unsigned)y < (unsigned)p.H &&
                                         (unsigned)x < (unsigned)p.W;
                        const float* src = inb ? xb + ((long long)rb[j] * plane_sz + rsp4[j] + tap_sp) * p.ics : p.x;
                        const int r = cr0 + 32 * j;
                        const uint32_t dst = araw(s) + r * 128 + ((cchunk ^ (r & 7)) << 4);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(inb ? 16 : 0) : "memory");
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(araw_full(s)) : "memory");
                    ic0 += TC_BK;
                    if (ic0 == p.Cin) { ic0 = 0; if (++ikx == p.kw) { ikx = 0; if (++iky == p.kh) { iky = 0; ++ikz; } } }
                };
                auto process = [&](int it) {
                    const int g_it = git + it;
                    const int s = g_it % RS, sa = g_it % AS;
                    const bool ok = rvalid && (unsigned)(z0 + pkz) < (unsigned)p.D && (unsigned)(y0 + pky) < (unsigned)p.H &&
                                    (unsigned)(x0 + pkx) < (unsigned)p.W;
                    const int tap_sp = (pkz * p.H + pky) * p.W + pkx;
                    const int kc = pc0 + cofs;
                    mbar_wait(araw_full(s), (g_it / RS) & 1, 9, g_it);
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t src = araw(s) + row * 128 + (((half * 4 + j) ^ (row & 7)) << 4);
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w) : "r"(src) : "memory");
                    }
                    if (p.pro != G6D_PRO_NONE && ok) {                  // padding / rows >= M stay zero (zero-filled copies)
                        const float4* scp; const float4* shp;
                        if (p.pro == G6D_PRO_CORR) {
                            scp = reinterpret_cast<const float4*>(p.ps + (long long)(rsp + tap_sp) * p.Cin + kc);
                            shp = reinterpret_cast<const float4*>(p.pb + kc);
                        } else {
                            scp = reinterpret_cast<const float4*>(p.ps + grp * p.Cin + kc);
                            shp = reinterpret_cast<const float4*>(p.pb + grp * p.Cin + kc);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 sc = __ldg(scp + j), sh = __ldg(shp + j);
                            float4 x4 = v[j];
                            x4.x = fmaf(x4.x, sc.x, sh.x); x4.y = fmaf(x4.y, sc.y, sh.y);
                            x4.z = fmaf(x4.z, sc.z, sh.z); x4.w = fmaf(x4.w, sc.w, sh.w);
                            if (p.pro == G6D_PRO_AFFINE_RELU) {
                                x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
                            }
                            v[j] = x4;
                        }
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float xs[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t h = (__float_as_uint(xs[e]) + 0x1000u) & 0xFFFFE000u;
                            hi[j * 4 + e] = h;
                            lo[j * 4 + e] = __float_as_uint(xs[e] - __uint_as_float(h));
                        }
                    }
                    __syncwarp();                                       // every lane's ld.shared has returned (values consumed above)
                    if (lane == 0) mbar_arrive(araw_empty(s));
                    mbar_wait(empty_a(sa), ((g_it / AS) & 1) ^ 1, 1, g_it);
                    tc_fence_after();
                    const uint32_t ta = lane_addr + (uint32_t)(sa * 64);
                    tmem_st16(ta, hi);
                    tmem_st16(ta + 32, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full_a(sa));
                    pc0 += TC_BK;
                    if (pc0 == p.Cin) { pc0 = 0; if (++pkx == p.kw) { pkx = 0; if (++pky == p.kh) { pky = 0; ++pkz; } } }
                };
                for (int it = 0; it < min(nkb, DEPTH); ++it) issue(it);
                for (int it = 0; it < nkb; ++it) {
                    if (it + DEPTH < nkb) issue(it + DEPTH);
                    process(it);
                }
                git += nkb;
            }
        } else
        for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
            int mt, nt, sp;
            decode(w, mt, nt, sp);
            const int kb_begin = sp * p.kb_per_split;
            const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
            int m = mt * TC_BM + row;
            const bool rvalid = m < p.M;
            if (!rvalid) m = 0;
            const int xo = m % p.Wo; m /= p.Wo;
            const int yo = m % p.Ho; m /= p.Ho;
            const int zo = m % p.Do; m /= p.Do;
            const int rb = m;
            const int z0 = zo * p.stride - p.pd, y0 = yo * p.stride - p.ph, x0 = xo * p.stride - p.pw;
            const int rsp = (z0 * p.H + y0) * p.W + x0;              // spatial index of tap (0,0,0); may be negative
            const long long rbase = (long long)rb * p.D * p.H * p.W + rsp;
            const long long grp = rb / (int)p.group_rows;
            int c0, kx, ky, kz;
            {
                const int k = kb_begin * TC_BK;
                int tap = 0; c0 = k;
                if (p.K != p.Cin) { tap = k / p.Cin; c0 = k - tap * p.Cin; }
                kx = tap % p.kw; const int tq = tap / p.kw; ky = tq % p.kh; kz = tq / p.kh;
            }
            // ring of R = PF + 1 register slots: PF K-blocks of loads in flight while one is processed
            constexpr int R = PF + 1;
            float4 v[R][4];
            bool ok[R]; int kc[R]; int ksp[R];
            auto issue_loads = [&](int q) {
                const int tap_sp = (kz * p.H + ky) * p.W + kx;
                kc[q] = c0 + cofs; ksp[q] = tap_sp;
                ok[q] = rvalid && (unsigned)(z0 + kz) < (unsigned)p.D && (unsigned)(y0 + ky) < (unsigned)p.H &&
                        (unsigned)(x0 + kx) < (unsigned)p.W;
#pragma unroll
                for (int j = 0; j < 4; ++j) v[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[q]) {
                    const float4* src = reinterpret_cast<const float4*>(p.x + p.ico + (rbase + tap_sp) * p.ics + c0 + cofs);
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[q][j] = __ldg(src + j);
                }
                c0 += TC_BK;
                if (c0 == p.Cin) { c0 = 0; if (++kx == p.kw) { kx = 0; if (++ky == p.kh) { ky = 0; ++kz; } } }
            };
            auto process = [&](int q, int it) {
                const int g_it = git + it;
                const int s = g_it % AS;
                const uint32_t n_use = g_it / AS;
                if (p.pro != G6D_PRO_NONE && ok[q]) {
                    const float4* scp; const float4* shp;
                    if (p.pro == G6D_PRO_CORR) {
                        scp = reinterpret_cast<const float4*>(p.ps + (long long)(rsp + ksp[q]) * p.Cin + kc[q]);
                        shp = reinterpret_cast<const float4*>(p.pb + kc[q]);
                    } else {
                        scp = reinterpret_cast<const float4*>(p.ps + grp * p.Cin + kc[q]);
                        shp = reinterpret_cast<const float4*>(p.pb + grp * p.Cin + kc[q]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 sc = __ldg(scp + j), sh = __ldg(shp + j);
                        float4 x4 = v[q][j];
                        x4.x = fmaf(x4.x, sc.x, sh.x); x4.y = fmaf(x4.y, sc.y, sh.y);
                        x4.z = fmaf(x4.z, sc.z, sh.z); x4.w = fmaf(x4.w, sc.w, sh.w);
                        if (p.pro == G6D_PRO_AFFINE_RELU) {
                            x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
                        }
                        v[q][j] = x4;
                    }
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xs[4] = {v[q][j].x, v[q][j].y, v[q][j].z, v[q][j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t h = (__float_as_uint(xs[e]) + 0x1000u) & 0xFFFFE000u;
                        hi[j * 4 + e] = h;
                        lo[j * 4 + e] = __float_as_uint(xs[e] - __uint_as_float(h));
                    }
                }
                mbar_wait(empty_a(s), (n_use & 1) ^ 1, 1, g_it);
                tc_fence_after();
                const uint32_t ta = lane_addr + (uint32_t)(s * 64);
                tmem_st16(ta, hi);
                tmem_st16(ta + 32, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a(s));
                if (RAWB) {
                    // weight tile of the same K-block: lo = rn_tf32(w - trunc(w)), elementwise on the swizzled tile
                    const int sb = g_it % BS;
                    mbar_wait(raw_full(sb), (g_it / BS) & 1, 7, g_it);
#pragma unroll
                    for (int ch = 0; ch < BN * 8 / 256; ++ch) {
                        const uint32_t off = (uint32_t)((ch * 256 + (int)threadIdx.x) * 16);
                        float4 wv;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(wv.x), "=f"(wv.y), "=f"(wv.z), "=f"(wv.w) : "r"(b_hi(sb) + off) : "memory");
                        auto lo_of = [](float w) {
                            const float d = w - __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
                            return __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
                        };
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(b_lo(sb) + off), "f"(lo_of(wv.x)), "f"(lo_of(wv.y)), "f"(lo_of(wv.z)), "f"(lo_of(wv.w)) : "memory");
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full_b(sb));
                }
            };
#pragma unroll
            for (int q = 0; q < PF; ++q)
                if (q < nkb) issue_loads(q);
            for (int it = 0; it < nkb; it += R) {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (it + q < nkb) {
                        if (it + q + PF < nkb) issue_loads((q + PF) % R);
                        process(q, it + q);
                    }
                }
            }
            git += nkb;
        }
    } else if (warp == 8) {
        // =============================== B producer (TMA) ===============================
        if (lane == 0) {
            int git = 0;
            for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_begin = sp * p.kb_per_split;
                const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
                for (int it = 0; it < min(nkb, TC2_PF); ++it) {
                    tma_prefetch_2d(&map_hi, (kb_begin + it) * TC_BK, nt * BN);
                    tma_prefetch_2d(&map_lo, (kb_begin + it) * TC_BK, nt * BN);
                }
                for (int it = 0; it < nkb; ++it, ++git) {
                    const int s = git % BS;
                    if (it + TC2_PF < nkb) {
                        tma_prefetch_2d(&map_hi, (kb_begin + it + TC2_PF) * TC_BK, nt * BN);
                        tma_prefetch_2d(&map_lo, (kb_begin + it + TC2_PF) * TC_BK, nt * BN);
                    }
                    mbar_wait(empty_b(s), ((git / BS) & 1) ^ 1, 3, git);
                    const int k = (kb_begin + it) * TC_BK;
                    if (RAWB) {
                        mbar_expect_tx(raw_full(s), Cfg::B_BYTES);
                        tma_load_2d(b_hi(s), &map_hi, raw_full(s), k, nt * BN);
                    } else {
                        mbar_expect_tx(full_b(s), 2 * Cfg::B_BYTES);
                        tma_load_2d(b_hi(s), &map_hi, full_b(s), k, nt * BN);
                        tma_load_2d(b_lo(s), &map_lo, full_b(s), k, nt * BN);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // =============================== MMA issuer ===============================
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
            int git = 0, tile = 0;
            for (int w = blockIdx.x; w < wk.total; w += gridDim.x, ++tile) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_begin = sp * p.kb_per_split;
                const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
                const int buf = tile % NBUF;
                mbar_wait(tmem_empty(buf), ((tile / NBUF) & 1) ^ 1, 6, tile);     // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t acc0 = tmem0 + (uint32_t)(buf * Cfg::BUF_COLS);
                const uint32_t cross_acc = acc0 + (uint32_t)(NMAIN * BN);
                for (int it = 0; it < nkb; ++it, ++git) {
                    const int sa = git % AS, sb = git % BS;
                    mbar_wait(full_a(sa), (git / AS) & 1, 4, git);
                    mbar_wait(full_b(sb), (git / BS) & 1, 5, git);
                    tc_fence_after();
                    const uint32_t ah = tmem0 + (uint32_t)(Cfg::A_COL0 + sa * 64), al = ah + 32;
                    const uint64_t dbh = umma_desc_sw128(b_hi(sb)), dbl = umma_desc_sw128(b_lo(sb));
                    const uint32_t main_acc = acc0 + (uint32_t)((it % NMAIN) * BN);
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        umma_tf32_ts(cross_acc, al + ks * 8, dbh + adv, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32_ts(cross_acc, ah + ks * 8, dbl + adv, idesc, 1u);
                        umma_tf32_ts(main_acc, ah + ks * 8, dbh + adv, idesc, (it >= NMAIN || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_a(sa));
                    umma_commit(empty_b(sb));
                }
                umma_commit(tmem_full(buf));
            }
        }
    } else {
        // =============================== epilogue (warps 10-13) ===============================
        const int quad = warp & 3;                     // TMEM lane quadrant = warp id % 4
        int tile = 0;
        for (int w = blockIdx.x; w < wk.total; w += gridDim.x, ++tile) {
            int mt, nt, sp;
            decode(w, mt, nt, sp);
            const int kb_begin = sp * p.kb_per_split;
            const int nkb = min(p.kblocks, kb_begin + p.kb_per_split) - kb_begin;
            const int buf = tile % NBUF;
            mbar_wait(tmem_full(buf), (tile / NBUF) & 1, 2, tile);
            tc_fence_after();
            const int m = mt * TC_BM + quad * 32 + lane;
            const int n_base = nt * BN;
            const bool partial = p.splits > 1;
            const bool vec_ok = partial ? (p.Cout & 3) == 0
                                        : ((p.ocs & 3) == 0 && (p.oco & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
                                           (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
            const int n_acc = nkb < NMAIN ? nkb : NMAIN;
            const uint32_t tbase = tmem0 + (uint32_t)(buf * Cfg::BUF_COLS) + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 16) {
                float accv[16];
#pragma unroll
                for (int a = 0; a <= NMAIN; ++a) {
                    const bool used = a == NMAIN || a < n_acc;
                    uint32_t r[16];
                    if (used) {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(tbase + (uint32_t)(a * BN + cc)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) accv[j] = a == 0 ? __uint_as_float(r[j]) : accv[j] + __uint_as_float(r[j]);
                    }
                }
                if (m < p.M) {
                    const int n0 = n_base + cc;
                    float* dst = partial ? p.ws + ((long long)sp * p.M + m) * p.Cout + n0
                                         : p.y + (long long)m * p.ocs + p.oco + n0;
                    const bool vec = vec_ok && n0 + 16 <= p.Cout;
                    if (vec) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            float4 v = make_float4(accv[j4 * 4], accv[j4 * 4 + 1], accv[j4 * 4 + 2], accv[j4 * 4 + 3]);
                            if (!partial) {
                                if (p.bias) {
                                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j4);
                                    v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                                }
                                v.x = tc_act(v.x, p.act); v.y = tc_act(v.y, p.act); v.z = tc_act(v.z, p.act); v.w = tc_act(v.w, p.act);
                            }
                            reinterpret_cast<float4*>(dst)[j4] = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (n0 + j < p.Cout) {
                                float v = accv[j];
                                if (!partial) {
                                    if (p.bias) v += __ldg(p.bias + n0 + j);
                                    v = tc_act(v, p.act);
                                }
                                dst[j] = v;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(buf));
        }
    }
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

__global__ void conv_tc_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias,
                                      float* __restrict__ y, int M, int Cout, int splits, int ocs, int oco, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * Cout) return;
    const int n = (int)(i % Cout);
    const long long m = i / Cout;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += ws[(long long)s * M * Cout + i];
    if (bias) v += bias[n];
    y[m * ocs + oco + n] = tc_act(v, act);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D tensor map over W [rows = Cout_pad, cols = K] fp32, box = [32 cols, bn rows], 128B swizzle
static int make_weight_map(CUtensorMap* map, const float* w, int rows, int K, int bn) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("g6d_conv_tc: cuTensorMapEncodeTiled unavailable"); return G6D_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("g6d_conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return G6D_ECUDA; }
    return G6D_OK;
}

// N = 256 tiles (instantiated, selectable with G6D_CONV_N256=1) halve the A re-reads per flop but leave
// room for only 2 pipeline stages and halve the CTA count: measured 18.4 vs 13.2 ms of convolution
// time per step on B200, so N <= 128 stays the default.
static int tc_block_n(int Cout) {
    static int n256 = -1;
    if (n256 < 0) { const char* e = getenv("G6D_CONV_N256"); n256 = (e && e[0] == '1') ? 1 : 0; }
    if (n256 && Cout > 128) return 256;
    return Cout > 64 ? 128 : (Cout > 32 ? 64 : 32);
}
static int flat_block_n(int Cout) { return Cout > 64 ? 128 : (Cout > 32 ? 64 : 32); }   // A-reuse kernel: smem goes to the halo

static int fill_tc_params(const g6d_conv_desc* d, ConvTcP& p) {
    G6D_REQUIRE(d != nullptr, "g6d_conv_tc: null desc");
    G6D_REQUIRE(d->B > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "g6d_conv_tc: bad dims");
    G6D_REQUIRE(d->kd > 0 && d->kh > 0 && d->kw > 0 && d->stride > 0, "g6d_conv_tc: bad kernel/stride");
    G6D_REQUIRE((d->Cin % TC_BK) == 0, "g6d_conv_tc: Cin (%d) must be a multiple of %d", d->Cin, TC_BK);
    G6D_REQUIRE((d->in_cstride & 3) == 0 && (d->in_coff & 3) == 0, "g6d_conv_tc: in_cstride/in_coff must be multiples of 4");
    G6D_REQUIRE(d->in_coff + d->Cin <= d->in_cstride, "g6d_conv_tc: input channel slice out of row");
    G6D_REQUIRE(d->out_coff + d->Cout <= d->out_cstride, "g6d_conv_tc: output channel slice out of row");
    const int Do = (d->D + 2 * d->pd - d->kd) / d->stride + 1;
    const int Ho = (d->H + 2 * d->ph - d->kh) / d->stride + 1;
    const int Wo = (d->W + 2 * d->pw - d->kw) / d->stride + 1;
    G6D_REQUIRE(Do == d->Do && Ho == d->Ho && Wo == d->Wo, "g6d_conv_tc: output dims mismatch");
    G6D_REQUIRE(d->prologue >= 0 && d->prologue <= 3 && d->act >= 0 && d->act <= 2, "g6d_conv_tc: bad prologue/act");
    const long long M = (long long)d->B * Do * Ho * Wo;
    const long long K = (long long)d->kd * d->kh * d->kw * d->Cin;
    G6D_REQUIRE(M < (1ll << 31) && K < (1ll << 31), "g6d_conv_tc: problem too large");
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.ics = d->in_cstride; p.ico = d->in_coff;
    p.Cout = d->Cout; p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pd = d->pd; p.ph = d->ph;
    p.pw = d->pw; p.Do = Do; p.Ho = Ho; p.Wo = Wo; p.ocs = d->out_cstride; p.oco = d->out_coff; p.pro = d->prologue;
    p.act = d->act; p.group_rows = d->group_rows > 0 ? d->group_rows : 1;
    p.M = (int)M; p.K = (int)K; p.kblocks = (int)(K / TC_BK);
    const int bn = tc_block_n(d->Cout);
    const long long ctas = (long long)ceil_div(M, TC_BM) * ceil_div(d->Cout, bn);
    int splits = 1;
    if (ctas < kNumSMs && p.kblocks >= 16) {
        // as many K splits as still fit in ONE wave of the 148 persistent CTAs (a second, partial wave
        // of long items costs more than the parallelism it adds)
        splits = (int)(kNumSMs / ctas);
        splits = splits > p.kblocks / 8 ? p.kblocks / 8 : splits;
        splits = splits < 1 ? 1 : splits;
    }
    // The tensor core adds each K-step into the fp32 accumulator with truncation; over very long
    // K chains of same-sign products (detector correlation: K = 115200 of post-ReLU features)
    // that is a systematic bias of ~4e-5 relative.  For long-K problems (K > 8192) the chain per
    // CTA is bounded to 64 K-blocks (2048 terms) and the partials are summed in fp32 round-to-nearest.
    const int min_splits = p.kblocks > 256 ? (p.kblocks + TC_MAX_KB_PER_SPLIT - 1) / TC_MAX_KB_PER_SPLIT : 1;
    splits = splits < min_splits ? min_splits : splits;
    splits = splits > 64 ? 64 : splits;
    p.kb_per_split = (p.kblocks + splits - 1) / splits;
    p.splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
    return G6D_OK;
}

template <int BN>
static int launch_tc(const ConvTcP& p, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("g6d_conv_tc: cannot opt in to %d B of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return G6D_ECUDA; }
        configured = true;
    }
    dim3 grid(ceil_div(p.M, TC_BM), ceil_div(p.Cout, BN), p.splits);
    conv_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(p, mh, ml);
    G6D_CHECK_LAUNCH("g6d_conv_tc");
    return G6D_OK;
}

static int tc_version() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("G6D_CONV_TC_V"); v = (e && e[0] >= '1' && e[0] <= '5') ? e[0] - '0' : 2; }
    return v;
}

static int tc3_prefetch() {      // K-blocks of activation loads in flight per producer thread (G6D_CONV_PF = 2 | 3)
    static int v = -1;
    if (v < 0) { const char* e = getenv("G6D_CONV_PF"); v = (e && e[0] == '3') ? 3 : 2; }
    return v;
}

template <int BN, bool RAWB, int PF = 2, bool STAGED = false>
static int launch_tc3(const ConvTcP& p, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
    using Cfg = Tc3Cfg<BN, STAGED>;
    if (!RAWB && !STAGED && PF == 2 && tc3_prefetch() == 3) return launch_tc3<BN, false, 3>(p, mh, ml, st);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<BN, RAWB, PF, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("g6d_conv_tc: cannot opt in to %d B of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return G6D_ECUDA; }
        configured = true;
    }
    Tc2Work wk;
    wk.m_tiles = ceil_div(p.M, TC_BM); wk.n_tiles = ceil_div(p.Cout, BN);
    const long long total = (long long)wk.m_tiles * wk.n_tiles * p.splits;
    wk.total = (int)total;
    const int grid = total < kNumSMs ? (int)total : kNumSMs;
    conv_tc3_kernel<BN, RAWB, PF, STAGED><<<grid, TC2_THREADS, Cfg::SMEM_BYTES, st>>>(p, wk, mh, ml);
    G6D_CHECK_LAUNCH("g6d_conv_tc(v3)");
    return G6D_OK;
}

template <int BN>
static int launch_tc2(const ConvTcP& p, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
    using Cfg = Tc2Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("g6d_conv_tc: cannot opt in to %d B of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return G6D_ECUDA; }
        configured = true;
    }
    Tc2Work wk;
    wk.m_tiles = ceil_div(p.M, TC_BM); wk.n_tiles = ceil_div(p.Cout, BN);
    const long long total = (long long)wk.m_tiles * wk.n_tiles * p.splits;
    wk.total = (int)total;
    const int grid = total < kNumSMs ? (int)total : kNumSMs;
    conv_tc2_kernel<BN><<<grid, TC2_THREADS, Cfg::SMEM_BYTES, st>>>(p, wk, mh, ml);
    G6D_CHECK_LAUNCH("g6d_conv_tc(v2)");
    return G6D_OK;
}

// elementwise tf32 split of an fp32 array (detector reference features used as kernels)
__global__ void split_tf32_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = in[i];
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
}

// [Cout, Cin, taps] (reference layout) -> hi/lo [rows_pad, taps*Cin_pad], K index = tap*Cin_pad + c
__global__ void pack_conv_weight_tc_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                           float* __restrict__ raw, int Cout, int Cin, int Cin_pad, int taps, int rows_pad,
                                           const float* __restrict__ scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long K = (long long)taps * Cin_pad;
    if (i >= K * rows_pad) return;
    const int o = (int)(i / K);
    const long long k = i % K;
    const int tap = (int)(k / Cin_pad), c = (int)(k % Cin_pad);
    float v = 0.f;
    if (o < Cout && c < Cin) {
        v = w[((long long)o * Cin + c) * taps + tap];
        if (scale) v *= scale[o];
    }
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
    if (raw) raw[i] = v;
}

// ==========================================================================================
// conv_tcflat_kernel: stride-1 convolutions with A-operand reuse across taps.
//
// The output positions of one image plane are enumerated over the PADDED width Wp = W + 2*pw:
// f = y*Wp + x.  Tap (ky,kx) of output f reads padded-input position f + ky*Wp + kx, so for a
// tile of 128 consecutive f the A operand of EVERY tap is a window of 128 consecutive rows of
// one shared-memory buffer holding padded-input positions [f0, f0 + 127 + (kh-1)*Wp + kw-1]:
// the tap is selected by the UMMA descriptor's start address (+shift*128 B; the 128B swizzle is a
// function of the absolute smem address, verified by g6d_debug_umma_shift).  The producers
// therefore gather (and prologue-transform, and hi/lo split) each input element ONCE per channel
// block instead of once per tap: 9x less producer work / L2 traffic for 3x3 ("FLAT" mode).  When
// the halo (kh-1)*Wp does not fit in shared memory (wide images, 15x15 correlation kernels) the
// buffer holds one kernel row at a time ("ROW" mode: kw-fold reuse).  Columns x >= Wo of the
// padded enumeration are computed and dropped.  B tiles stream by TMA per (channel block, tap).
struct ConvFlatP {
    const float* x; const float* bias; const float* ps; const float* pb; float* y; float* ws;
    int B, D, H, W, Cin, ics, ico, Cout, kd, kh, kw, pd, ph, pw, Do, Ho, Wo, ocs, oco, pro, act;
    long long group_rows;
    int Wp, tiles_per_plane, mode, nseg, taps_per_seg, seg_rows, rows_pad, ntab, cblocks;
    int a_stages, b_stages, splits, cb_per_split, M;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tcflat_kernel(const ConvFlatP p, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
    constexpr int NMAIN = TcCfg<BN>::NMAIN;
    constexpr int TMEM_COLS = TcCfg<BN>::TMEM_COLS;
    constexpr int B_BYTES = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int A_TILE = p.rows_pad * 128;                 // one of hi / lo
    const uint32_t a_base = base;
    const uint32_t b_base = base + p.a_stages * 2 * A_TILE;
    const uint32_t bar_base = b_base + p.b_stages * 2 * B_BYTES;
    auto a_hi = [&](int s) { return a_base + s * 2 * A_TILE; };
    auto a_lo = [&](int s) { return a_base + s * 2 * A_TILE + A_TILE; };
    auto b_hi = [&](int s) { return b_base + s * 2 * B_BYTES; };
    auto b_lo = [&](int s) { return b_base + s * 2 * B_BYTES + B_BYTES; };
    auto a_full = [&](int s) { return bar_base + 8 * s; };
    auto a_empty = [&](int s) { return bar_base + 8 * (4 + s); };
    auto b_full = [&](int s) { return bar_base + 8 * (8 + s); };
    auto b_empty = [&](int s) { return bar_base + 8 * (12 + s); };
    const uint32_t tmem_full = bar_base + 8 * 16;
    const uint32_t bar_off = (bar_base - base);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8 * 17);
    int* rowtab = reinterpret_cast<int*>(base_ptr + bar_off + 256);   // [ntab][seg_rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tile = blockIdx.x;
    const int t_in_plane = tile % p.tiles_per_plane; tile /= p.tiles_per_plane;
    const int zo = tile % p.Do;
    const int b = tile / p.Do;
    const int f0 = t_in_plane * TC_BM;
    const int n_base = blockIdx.y * BN;
    const int split = blockIdx.z;
    const int cb_begin = split * p.cb_per_split;
    const int cb_end = min(p.cblocks, cb_begin + p.cb_per_split);
    const int nunits = (cb_end - cb_begin) * p.nseg;

    // ---- setup: row tables (element offset of each gathered row inside its image plane, -1 = zero)
    for (int e = threadIdx.x; e < p.ntab * p.seg_rows; e += blockDim.x) {
        const int tb = e / p.seg_rows, i = e % p.seg_rows;
        const int g = f0 + (p.mode == 1 ? tb * p.Wp : 0) + i;     // padded-input flat position
        const int yy = g / p.Wp - p.ph, xx = g % p.Wp - p.pw;
        rowtab[e] = ((unsigned)yy < (unsigned)p.H && (unsigned)xx < (unsigned)p.W) ? yy * p.W + xx : -1;
    }
    if (warp == TC_PRODUCER_WARPS && lane == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_init(a_full(s), TC_PRODUCER_WARPS);
            mbar_init(a_empty(s), 1);
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    }
    if (warp == TC_PRODUCER_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        // =============================== A producers ===============================
        const int chunk = threadIdx.x & 7;
        const int r0 = threadIdx.x >> 3;                        // rows r0 + 32*j
        const long long plane = (long long)p.H * p.W;
        const long long gi = (long long)b / p.group_rows;
        for (int u = 0; u < nunits; ++u) {
            const int cb = cb_begin + u / p.nseg, seg = u % p.nseg;
            const int s = u % p.a_stages;
            const uint32_t n_use = u / p.a_stages;
            // FLAT: seg = kz, table 0.  ROW: seg = kz*kh + ky, table ky.
            const int kz = p.mode == 1 ? seg / p.kh : seg;
            const int tb = p.mode == 1 ? seg % p.kh : 0;
            const int zz = zo + kz - p.pd;
            const bool zok = (unsigned)zz < (unsigned)p.D;
            const float* xplane = p.x + ((long long)b * p.D + (zok ? zz : 0)) * plane * p.ics + p.ico + cb * TC_BK + chunk * 4;
            const int* tab = rowtab + tb * p.seg_rows;
            const int c = cb * TC_BK + chunk * 4;
            mbar_wait(a_empty(s), (n_use & 1) ^ 1, 1, u);
            for (int rbase = 0; rbase < p.seg_rows; rbase += 128) {        // 4 rows per thread per trip
                float4 v[4]; int off[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = rbase + r0 + 32 * j;
                    off[j] = (r < p.seg_rows && zok) ? tab[r] : -1;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (off[j] >= 0) v[j] = __ldg(reinterpret_cast<const float4*>(xplane + (long long)off[j] * p.ics));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = rbase + r0 + 32 * j;
                    if (r >= p.rows_pad) continue;
                    float4 x4 = v[j];
                    if (p.pro != G6D_PRO_NONE && off[j] >= 0) {
                        float4 sc, sh;
                        if (p.pro == G6D_PRO_CORR) {
                            const long long sp = (long long)zz * plane + off[j];
                            sc = __ldg(reinterpret_cast<const float4*>(p.ps + sp * p.Cin + c));
                            sh = __ldg(reinterpret_cast<const float4*>(p.pb + c));
                        } else {
                            sc = __ldg(reinterpret_cast<const float4*>(p.ps + gi * p.Cin + c));
                            sh = __ldg(reinterpret_cast<const float4*>(p.pb + gi * p.Cin + c));
                        }
                        x4.x = fmaf(x4.x, sc.x, sh.x); x4.y = fmaf(x4.y, sc.y, sh.y);
                        x4.z = fmaf(x4.z, sc.z, sh.z); x4.w = fmaf(x4.w, sc.w, sh.w);
                        if (p.pro == G6D_PRO_AFFINE_RELU) {
                            x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
                        }
                    }
                    float4 hi, lo;
                    hi.x = __uint_as_float((__float_as_uint(x4.x) + 0x1000u) & 0xFFFFE000u);
                    hi.y = __uint_as_float((__float_as_uint(x4.y) + 0x1000u) & 0xFFFFE000u);
                    hi.z = __uint_as_float((__float_as_uint(x4.z) + 0x1000u) & 0xFFFFE000u);
                    hi.w = __uint_as_float((__float_as_uint(x4.w) + 0x1000u) & 0xFFFFE000u);
                    lo.x = x4.x - hi.x; lo.y = x4.y - hi.y; lo.z = x4.z - hi.z; lo.w = x4.w - hi.w;
                    const uint32_t so = r * 128 + ((chunk ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_hi(s) + so), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_lo(s) + so), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(s));
        }

        // =============================== epilogue ===============================
        mbar_wait(tmem_full, 0, 2, nunits);
        tc_fence_after();
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int f = f0 + row;
        const int yo = f / p.Wp, xo = f % p.Wp;
        const bool valid = yo < p.Ho && xo < p.Wo;
        const long long m = (((long long)b * p.Do + zo) * p.Ho + yo) * p.Wo + xo;
        constexpr int HALF = BN / 2;
        const int col0 = (warp >> 2) * HALF;
        const bool partial = p.splits > 1;
        const bool vec_ok = partial ? (p.Cout & 3) == 0
                                    : ((p.ocs & 3) == 0 && (p.oco & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
                                       (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
        const int total_mm = nunits * p.taps_per_seg;
        const int n_acc = total_mm < NMAIN ? total_mm : NMAIN;
#pragma unroll
        for (int cc = 0; cc < HALF; cc += 16) {
            float accv[16];
            const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(col0 + cc);
#pragma unroll
            for (int a = 0; a <= NMAIN; ++a) {
                const bool used = a == NMAIN || a < n_acc;
                uint32_t r[16];
                if (used) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr + (uint32_t)(a * BN)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) accv[j] = a == 0 ? __uint_as_float(r[j]) : accv[j] + __uint_as_float(r[j]);
                }
            }
            if (valid) {
                const int n0 = n_base + col0 + cc;
                float* dst = partial ? p.ws + ((long long)split * p.M + m) * p.Cout + n0 : p.y + m * p.ocs + p.oco + n0;
                if (vec_ok && n0 + 16 <= p.Cout) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 v = make_float4(accv[j4 * 4], accv[j4 * 4 + 1], accv[j4 * 4 + 2], accv[j4 * 4 + 3]);
                        if (!partial) {
                            if (p.bias) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j4);
                                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                            }
                            v.x = tc_act(v.x, p.act); v.y = tc_act(v.y, p.act); v.z = tc_act(v.z, p.act); v.w = tc_act(v.w, p.act);
                        }
                        reinterpret_cast<float4*>(dst)[j4] = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n0 + j < p.Cout) {
                            float v = accv[j];
                            if (!partial) {
                                if (p.bias) v += __ldg(p.bias + n0 + j);
                                v = tc_act(v, p.act);
                            }
                            dst[j] = v;
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp == TC_PRODUCER_WARPS) {
        // =============================== B producer (TMA) ===============================
        if (lane == 0) {
            int bi = 0;
            for (int u = 0; u < nunits; ++u) {
                const int cb = cb_begin + u / p.nseg, seg = u % p.nseg;
                for (int t = 0; t < p.taps_per_seg; ++t, ++bi) {
                    const int s = bi % p.b_stages;
                    const uint32_t n_use = bi / p.b_stages;
                    mbar_wait(b_empty(s), (n_use & 1) ^ 1, 3, bi);
                    mbar_expect_tx(b_full(s), 2 * B_BYTES);
                    const int k = (seg * p.taps_per_seg + t) * p.Cin + cb * TC_BK;
                    tma_load_2d(b_hi(s), &map_hi, b_full(s), k, n_base);
                    tma_load_2d(b_lo(s), &map_lo, b_full(s), k, n_base);
                }
            }
        }
    } else {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
            int bi = 0;
            for (int u = 0; u < nunits; ++u) {
                const int sa = u % p.a_stages;
                mbar_wait(a_full(sa), (u / p.a_stages) & 1, 4, u);
                tc_fence_after();
                for (int t = 0; t < p.taps_per_seg; ++t, ++bi) {
                    const int sb = bi % p.b_stages;
                    mbar_wait(b_full(sb), (bi / p.b_stages) & 1, 5, bi);
                    tc_fence_after();
                    const int shift = p.mode == 1 ? t : (t / p.kw) * p.Wp + (t % p.kw);     // rows
                    const uint64_t dah = umma_desc_sw128(a_hi(sa) + shift * 128), dal = umma_desc_sw128(a_lo(sa) + shift * 128);
                    const uint64_t dbh = umma_desc_sw128(b_hi(sb)), dbl = umma_desc_sw128(b_lo(sb));
                    const uint32_t main_acc = tmem_acc + (uint32_t)((bi % NMAIN) * BN);
                    const uint32_t cross_acc = tmem_acc + (uint32_t)(NMAIN * BN);
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        umma_tf32(cross_acc, dal + adv, dbh + adv, idesc, (bi > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(cross_acc, dah + adv, dbl + adv, idesc, 1u);
                        umma_tf32(main_acc, dah + adv, dbh + adv, idesc, (bi >= NMAIN || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(b_empty(sb));
                }
                umma_commit(a_empty(sa));
            }
            umma_commit(tmem_full);
        }
    }
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(TMEM_COLS) : "memory");
    }
}

// G6D_CONV_FLAT: 0 = never use the A-reuse kernel, 1 = FLAT mode only (default), 2 = FLAT and ROW.
// Measured on B200 (tools/conv_breakdown.py): the 3xTF32 MMAs are shared-memory-bandwidth bound
// (every MMA re-reads 4 KB of A and N*32 B of B; at N = 128 that alone is 128 B/clk/SM), so the
// 3x-reuse ROW mode does not pay for its junk columns, while FLAT (9x reuse, and the prologue
// applied once per element instead of once per tap) gains 26 % on the selector's first tower conv.
static int flat_level() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("G6D_CONV_FLAT"); v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }
    return v;
}
static bool flat_disabled() { return flat_level() == 0; }

static int fill_flat_params(const g6d_conv_desc* d, ConvFlatP& p, int* smem_bytes) {
    if (!d || d->stride != 1 || (d->Cin % TC_BK) != 0 || d->Cout < 16 || (d->in_cstride & 3) || (d->in_coff & 3)) return -1;
    const int Do = d->D + 2 * d->pd - d->kd + 1, Ho = d->H + 2 * d->ph - d->kh + 1, Wo = d->W + 2 * d->pw - d->kw + 1;
    if (Do != d->Do || Ho != d->Ho || Wo != d->Wo || Do < 1 || Ho < 1 || Wo < 1) return -1;
    if (d->kd * d->kh * d->kw == 1) return -1;                        // 1x1: nothing to reuse, old kernel
    const int bn = flat_block_n(d->Cout);
    const int Wp = d->W + 2 * d->pw;
    const int flat_rows = TC_BM + (d->kh - 1) * Wp + d->kw - 1;
    const int row_rows = TC_BM + d->kw - 1;
    const int budget = 220 * 1024;
    const int b_stage = 2 * bn * 128;
    // tiles never span image planes: small planes (selector 4x4 / 8x8 maps) would leave most of a
    // 128-row tile empty -> keep those on the batch-flattened kernel
    {
        const long long tiles = ((long long)Ho * Wp + TC_BM - 1) / TC_BM;
        if ((long long)Ho * Wo * 100 < tiles * TC_BM * 60) return -1;
    }
    auto a_stage = [](int rows) { return 2 * ((rows + 7) / 8 * 8) * 128; };
    int mode, rows, ntab;
    // FLAT when two A stages of the full halo + >= 3 B stages fit
    if (2 * a_stage(flat_rows) + 3 * b_stage + 4 * flat_rows + 2048 <= budget) { mode = 0; rows = flat_rows; ntab = 1; }
    else if (d->kw > 1 && flat_level() >= 2) { mode = 1; rows = row_rows; ntab = d->kh; }
    else return -1;
    p.B = d->B; p.D = d->D; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.ics = d->in_cstride; p.ico = d->in_coff;
    p.Cout = d->Cout; p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
    p.Do = Do; p.Ho = Ho; p.Wo = Wo; p.ocs = d->out_cstride; p.oco = d->out_coff; p.pro = d->prologue; p.act = d->act;
    p.group_rows = d->group_rows > 0 ? d->group_rows : 1;
    p.Wp = Wp; p.tiles_per_plane = (Ho * Wp + TC_BM - 1) / TC_BM; p.mode = mode;
    p.nseg = mode == 0 ? d->kd : d->kd * d->kh; p.taps_per_seg = mode == 0 ? d->kh * d->kw : d->kw;
    p.seg_rows = rows; p.rows_pad = (rows + 7) / 8 * 8; p.ntab = ntab; p.cblocks = d->Cin / TC_BK;
    const long long M = (long long)d->B * Do * Ho * Wo;
    if (M >= (1ll << 31)) return -1;
    p.M = (int)M;
    const int tab_bytes = 4 * ntab * rows;
    int a_st = 2, b_st = 3;
    int used = a_st * a_stage(rows) + b_st * b_stage + tab_bytes + 2048;
    while (b_st < 4 && used + b_stage <= budget) { ++b_st; used += b_stage; }
    while (a_st < 4 && used + a_stage(rows) <= budget) { ++a_st; used += a_stage(rows); }
    if (used > budget) return -1;
    p.a_stages = a_st; p.b_stages = b_st;
    *smem_bytes = a_st * a_stage(rows) + b_st * b_stage + 256 + tab_bytes + 1024 + 64;
    // split over channel blocks when the tile grid cannot fill the machine, or to bound accumulate chains
    const long long ctas = (long long)d->B * Do * p.tiles_per_plane * ((d->Cout + bn - 1) / bn);
    const long long kblocks = (long long)p.cblocks * d->kd * d->kh * d->kw;
    int splits = 1;
    if (ctas < kNumSMs && p.cblocks >= 2) splits = (int)((kNumSMs + ctas - 1) / ctas);
    if (kblocks > 256) { const int ms = (int)((kblocks + TC_MAX_KB_PER_SPLIT - 1) / TC_MAX_KB_PER_SPLIT); splits = splits < ms ? ms : splits; }
    splits = splits > p.cblocks ? p.cblocks : splits;
    splits = splits < 1 ? 1 : splits;
    p.cb_per_split = (p.cblocks + splits - 1) / splits;
    p.splits = (p.cblocks + p.cb_per_split - 1) / p.cb_per_split;
    return 0;
}

template <int BN>
static int launch_flat(const ConvFlatP& p, int smem, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
    static int configured = 0;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcflat_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("g6d_conv_tc(flat): cannot opt in to shared memory: %s", cudaGetErrorString(e)); return G6D_ECUDA; }
        configured = 227 * 1024;
    }
    dim3 grid((unsigned)((long long)p.B * p.Do * p.tiles_per_plane), ceil_div(p.Cout, BN), p.splits);
    conv_tcflat_kernel<BN><<<grid, TC_THREADS, smem, st>>>(p, mh, ml);
    G6D_CHECK_LAUNCH("g6d_conv_tc(flat)");
    return G6D_OK;
}

}  // namespace g6d

using namespace g6d;

// Debug aid: copies the 8-int timeout record (0 = no timeout; else [1]=waiter role 1 A-producer/empty,
// 2 epilogue/tmem_full, 3 B-producer/empty, 4 MMA/full_a, 5 MMA/full_b; [2]=iteration; [3]=parity;
// [4..6]=block; [7]=thread) and clears it.  Synchronises the device.
extern "C" int g6d_conv_tc_debug(int* host_out8) {
    G6D_REQUIRE(host_out8 != nullptr, "g6d_conv_tc_debug: null");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("g6d_conv_tc_debug: sync: %s", cudaGetErrorString(e)); return G6D_ECUDA; }
    e = cudaMemcpyFromSymbol(host_out8, g_tc_timeout, sizeof(int) * 8);
    if (e != cudaSuccess) { set_error("g6d_conv_tc_debug: %s", cudaGetErrorString(e)); return G6D_ECUDA; }
    int zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_tc_timeout, zeros, sizeof(zeros));
    return G6D_OK;
}

extern "C" int g6d_conv_tc_supported(const g6d_conv_desc* d) {
    if (!d) return 0;
    return (d->Cin % TC_BK) == 0 && d->Cout >= 16 && (d->in_cstride & 3) == 0 && (d->in_coff & 3) == 0 ? 1 : 0;
}



extern "C" long long g6d_conv_tc_workspace_bytes(const g6d_conv_desc* desc) {
    {
        ConvFlatP fp{}; int smem = 0;
        if (!flat_disabled() && fill_flat_params(desc, fp, &smem) == 0)
            return fp.splits > 1 ? (long long)fp.splits * fp.M * fp.Cout * (long long)sizeof(float) : 0;
    }
    ConvTcP p{};
    if (fill_tc_params(desc, p) != G6D_OK) return -1;
    return p.splits > 1 ? (long long)p.splits * p.M * p.Cout * (long long)sizeof(float) : 0;
}

extern "C" int g6d_conv_tc(const g6d_conv_desc* desc, const float* x, const float* w_hi, const float* w_lo,
                           const float* w_raw, int w_rows, const float* bias, const float* pro_scale, const float* pro_shift, float* y,
                           void* ws, g6d_stream_t stream) {
    {   // stride-1 multi-tap convolutions: A-reuse kernel
        ConvFlatP fp{}; int smem = 0;
        if (!flat_disabled() && fill_flat_params(desc, fp, &smem) == 0) {
            G6D_REQUIRE(x && w_hi && w_lo && y, "g6d_conv_tc: null tensor pointer");
            G6D_REQUIRE(w_rows >= fp.Cout, "g6d_conv_tc: weight rows (%d) < Cout (%d)", w_rows, fp.Cout);
            if (fp.pro != G6D_PRO_NONE) G6D_REQUIRE(pro_scale && pro_shift, "g6d_conv_tc: prologue operands missing");
            if (fp.splits > 1) G6D_REQUIRE(ws != nullptr, "g6d_conv_tc: split workspace required (%d splits)", fp.splits);
            fp.x = x; fp.bias = bias; fp.ps = pro_scale; fp.pb = pro_shift; fp.y = y; fp.ws = (float*)ws;
            const int bn = flat_block_n(fp.Cout);
            const int K = fp.kd * fp.kh * fp.kw * fp.Cin;
            CUtensorMap mh, ml;
            int rc2;
            if ((rc2 = make_weight_map(&mh, w_hi, w_rows, K, bn)) != G6D_OK) return rc2;
            if ((rc2 = make_weight_map(&ml, w_lo, w_rows, K, bn)) != G6D_OK) return rc2;
            cudaStream_t st = as_stream(stream);
            if (bn == 128) rc2 = launch_flat<128>(fp, smem, mh, ml, st);
            else if (bn == 64) rc2 = launch_flat<64>(fp, smem, mh, ml, st);
            else rc2 = launch_flat<32>(fp, smem, mh, ml, st);
            if (rc2 != G6D_OK) return rc2;
            if (fp.splits > 1) {
                const long long n = (long long)fp.M * fp.Cout;
                conv_tc_reduce_kernel<<<ceil_div(n, 256), 256, 0, st>>>(fp.ws, bias, y, fp.M, fp.Cout, fp.splits, fp.ocs, fp.oco, fp.act);
                G6D_CHECK_LAUNCH("g6d_conv_tc(flat reduce)");
            }
            return G6D_OK;
        }
    }
    ConvTcP p{};
    int rc = fill_tc_params(desc, p);
    if (rc != G6D_OK) return rc;
    G6D_REQUIRE(x && w_hi && w_lo && y, "g6d_conv_tc: null tensor pointer");
    G6D_REQUIRE(w_rows >= p.Cout, "g6d_conv_tc: weight rows (%d) < Cout (%d)", w_rows, p.Cout);
    if (p.pro != G6D_PRO_NONE) G6D_REQUIRE(pro_scale && pro_shift, "g6d_conv_tc: prologue operands missing");
    if (p.splits > 1) G6D_REQUIRE(ws != nullptr, "g6d_conv_tc: split-K workspace required (%d splits)", p.splits);
    p.x = x; p.bias = bias; p.ps = pro_scale; p.pb = pro_shift; p.y = y; p.ws = (float*)ws;
    const int bn = tc_block_n(p.Cout);
    CUtensorMap mh, ml;
    if ((rc = make_weight_map(&mh, w_hi, w_rows, p.K, bn)) != G6D_OK) return rc;
    if ((rc = make_weight_map(&ml, w_lo, w_rows, p.K, bn)) != G6D_OK) return rc;
    cudaStream_t st = as_stream(stream);
    // v2 packs (z, y, x) + 8 into 8/12/12 bits and uses 32-bit spatial offsets and group indices
    const bool v2_ok = p.D + p.pd + 8 < 256 && p.H + p.ph + 8 < 4096 && p.W + p.pw + 8 < 4096 &&
                       (long long)p.D * p.H * p.W < (1ll << 30) && p.group_rows < (1ll << 31);
    if (bn == 256) rc = launch_tc<256>(p, mh, ml, st);
    else if (tc_version() == 4 && v2_ok && w_raw) { // A operand in tensor memory, weights split in shared memory
        CUtensorMap mr;
        if ((rc = make_weight_map(&mr, w_raw, w_rows, p.K, bn)) != G6D_OK) return rc;
        if (bn == 128) rc = launch_tc3<128, true>(p, mr, mr, st);
        else if (bn == 64) rc = launch_tc3<64, true>(p, mr, mr, st);
        else rc = launch_tc3<32, true>(p, mr, mr, st);
    } else if (tc_version() == 5 && v2_ok) {        // A staged by cp.async in shared memory, operand in tensor memory
        if (bn == 128) rc = launch_tc3<128, false, 2, true>(p, mh, ml, st);
        else if (bn == 64) rc = launch_tc3<64, false, 2, true>(p, mh, ml, st);
        else rc = launch_tc3<32, false, 2, true>(p, mh, ml, st);
    } else if (tc_version() >= 3 && v2_ok) {        // A operand in tensor memory
        if (bn == 128) rc = launch_tc3<128, false>(p, mh, ml, st);
        else if (bn == 64) rc = launch_tc3<64, false>(p, mh, ml, st);
        else rc = launch_tc3<32, false>(p, mh, ml, st);
    } else if (tc_version() == 2 && v2_ok) {
        if (bn == 128) rc = launch_tc2<128>(p, mh, ml, st);
        else if (bn == 64) rc = launch_tc2<64>(p, mh, ml, st);
        else rc = launch_tc2<32>(p, mh, ml, st);
    } else {
        if (bn == 128) rc = launch_tc<128>(p, mh, ml, st);
        else if (bn == 64) rc = launch_tc<64>(p, mh, ml, st);
        else rc = launch_tc<32>(p, mh, ml, st);
    }
    if (rc != G6D_OK) return rc;
    if (p.splits > 1) {
        const long long n = (long long)p.M * p.Cout;
        conv_tc_reduce_kernel<<<ceil_div(n, 256), 256, 0, st>>>(p.ws, bias, y, p.M, p.Cout, p.splits, p.ocs, p.oco, p.act);
        G6D_CHECK_LAUNCH("g6d_conv_tc(splitk reduce)");
    }
    return G6D_OK;
}

extern "C" int g6d_split_tf32(const float* in, float* hi, float* lo, long long n, g6d_stream_t stream) {
    G6D_REQUIRE(in && hi && lo && n > 0, "g6d_split_tf32: bad args");
    split_tf32_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(in, hi, lo, n);
    G6D_CHECK_LAUNCH("g6d_split_tf32");
    return G6D_OK;
}

extern "C" int g6d_pack_conv_weight_tc(const float* w, float* out_hi, float* out_lo, float* out_raw, int Cout, int Cin, int Cin_pad,
                                       int taps, int rows_pad, const float* cout_scale, g6d_stream_t stream) {
    G6D_REQUIRE(w && out_hi && out_lo && Cout > 0 && Cin > 0 && Cin_pad >= Cin && taps > 0 && rows_pad >= Cout,
                "g6d_pack_conv_weight_tc: bad args");
    const long long total = (long long)taps * Cin_pad * rows_pad;
    pack_conv_weight_tc_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, out_hi, out_lo, out_raw, Cout, Cin,
                                                                                    Cin_pad, taps, rows_pad, cout_scale);
    G6D_CHECK_LAUNCH("g6d_pack_conv_weight_tc");
    return G6D_OK;
}

// ------------------------------------------------------------------------------------------
// Probe (debug/test only): does a K-major SWIZZLE_128B A operand tolerate a start address that is
// shifted by `shift` rows (shift*128 B, not 1024-aligned) when the data was written with the
// swizzle phase of its ABSOLUTE shared-memory row?  D[128 x 32] = A[shift .. shift+128) x B^T with
// K = 32, A[r][k] = r + k/64 (exactly representable), B = 32x32 identity.  `mode` selects how the
// descriptor's base_offset field is set: 0 -> 0, 1 -> (start_address >> 7) & 7.
namespace g6d {
__global__ void __launch_bounds__(128) umma_shift_probe_kernel(float* out, int shift, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a_base = base;                 // 160 rows x 128 B
    const uint32_t b_base = base + 160 * 128;     // 32 rows x 128 B (20480 is 1024-aligned)
    const uint32_t bar = b_base + 32 * 128;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + 160 * 128 + 32 * 128 + 16);
    const int t = threadIdx.x;
    for (int r = t; r < 160; r += 128)
        for (int c = 0; c < 8; ++c) {
            float4 v = make_float4(r + (c * 4 + 0) / 64.f, r + (c * 4 + 1) / 64.f, r + (c * 4 + 2) / 64.f, r + (c * 4 + 3) / 64.f);
            *reinterpret_cast<float4*>(bp + r * 128 + ((c ^ (r & 7)) << 4)) = v;
        }
    for (int r = t; r < 32; r += 128)
        for (int c = 0; c < 8; ++c) {
            float4 v = make_float4(r == c * 4 ? 1.f : 0.f, r == c * 4 + 1 ? 1.f : 0.f, r == c * 4 + 2 ? 1.f : 0.f, r == c * 4 + 3 ? 1.f : 0.f);
            *reinterpret_cast<float4*>(bp + 160 * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
        }
    fence_proxy_async();
    if (t == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32((const void*)slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (t == 0) {
        const uint32_t start = a_base + shift * 128;
        uint64_t da = umma_desc_sw128(start);
        if (mode == 1) da |= (uint64_t)((start >> 7) & 7) << 49;
        const uint64_t db = umma_desc_sw128(b_base);
        const uint32_t idesc = umma_idesc_tf32(128, 32);
        for (int ks = 0; ks < 4; ++ks) umma_tf32(tmem, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, ks > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0, 9, 0);
    tc_fence_after();
    const int warp = t >> 5, lane = t & 31;
    for (int cc = 0; cc < 32; cc += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 32 + cc + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}
}  // namespace g6d

extern "C" int g6d_debug_umma_shift(float* out, int shift, int mode, g6d_stream_t stream) {
    G6D_REQUIRE(out && shift >= 0 && shift <= 31, "g6d_debug_umma_shift: bad args");
    g6d::umma_shift_probe_kernel<<<1, 128, 160 * 128 + 32 * 128 + 1024 + 64, g6d::as_stream(stream)>>>(out, shift, mode);
    G6D_CHECK_LAUNCH("g6d_debug_umma_shift");
    return G6D_OK;
}
