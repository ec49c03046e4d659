// tcgen05 / TMA / mbarrier PTX helpers shared by the tensor-core convolution kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace g6d {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps inside the instruction until the phase completes
// (or the hint expires) instead of re-polling.  ncu on the round-2 kernel: the un-hinted polls of the
// waiting roles (MMA issuer, TMA and epilogue warps) were 5.7 M SYNCS + 4.5 M bail-out flag reads per
// launch, ~15 % of the L1 / shared-memory data-pipe wavefronts the operand feed competes for.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU.  On timeout (~0.25 s) the waiter records
// who was waiting on what in g_tc_timeout (read back with g6d_conv_tc_debug) and every wait in the
// grid falls through, so the kernel terminates (with garbage output) instead of spinning.
static __device__ int g_tc_timeout[8] = {0, 0, 0, 0, 0, 0, 0, 0};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who = 0, int iter = 0) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 63u) != 0) continue;                 // the bail-out flag is global memory: look at it rarely
        if (*(volatile int*)&g_tc_timeout[0] != 0) return;
        if (clock64() - t0 > 500000000ll) {
            if (atomicCAS(&g_tc_timeout[0], 0, 1) == 0) {
                g_tc_timeout[1] = who; g_tc_timeout[2] = iter; g_tc_timeout[3] = (int)parity;
                g_tc_timeout[4] = (int)blockIdx.x; g_tc_timeout[5] = (int)blockIdx.y; g_tc_timeout[6] = (int)blockIdx.z;
                g_tc_timeout[7] = (int)threadIdx.x;
            }
            return;
        }
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// L2 prefetch of a TMA box (no shared memory, no barrier): turns the later tma_load_2d into an L2 hit
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start
// address >> 4 in bits [0,14); LBO (ignored for swizzled K-major) = 1 in [16,30); SBO = 1024 B
// (8 rows x 128 B) >> 4 in [32,46); descriptor version 1 in [46,48); layout SWIZZLE_128B (=2)
// in [61,64).  The tile base must be 1024-byte aligned (base_offset = 0).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major:
// c_format F32 (1) at [4,6); a_format/b_format TF32 (2) at [7,10)/[10,13); n_dim = N>>3 at
// [17,23); m_dim = M>>4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__device__ __forceinline__ float tc_act(float v, int act) {
    if (act == G6D_ACT_RELU) return fmaxf(v, 0.f);
    if (act == G6D_ACT_LEAKY01) return v > 0.f ? v : 0.1f * v;
    return v;
}


}  // namespace g6d
