// tcgen05.mma kind::tf32 issue-rate probe (B200): cycles per MMA for
//   N in {32,64,128,256}; A from shared memory (SS) or from tensor memory (TS);
//   all MMAs accumulating into ONE accumulator vs rotating over R independent accumulators.
// One CTA per SM, one issuing thread; operands are zero-filled (timing only).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gen6d_b200/csrc/tc_common.cuh"
using namespace g6d;

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS, mode 1: TS (A in TMEM).  rot = number of accumulators rotated over (1 = dependent chain).
template <int N, int mode, int rot, int ELECT>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
    constexpr int nb = 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    // A: 2 tiles of 16 KB; B: nb tiles of N*128 B; barrier after
    const uint32_t a0 = base, b0 = base + 2 * 16384, bar = b0 + nb * 256 * 128;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(bp + (bar - base) + 16);
    for (uint32_t i = threadIdx.x; i < (bar - base) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x < 32 && (ELECT ? elect_one_sync() : threadIdx.x == 0)) {
        const uint32_t idesc = umma_idesc_tf32(128, N);
        const uint32_t a_tmem = tm + 448;            // 64 columns of A (garbage) at the top of TMEM
        long long t0 = clock64();
        int n = 0;
        for (int it = 0; it < iters; ++it) {
            const uint64_t da = umma_desc_sw128(a0 + (it & 1) * 16384);
            const uint64_t db = umma_desc_sw128(b0 + (it & (nb - 1)) * N * 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks, ++n) {
                const uint32_t d = tm + (uint32_t)((n % rot) * N);
                const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                if (mode == 0) umma_tf32(d, da + adv, db + adv, idesc, 1u);
                else umma_tf32_ts(d, a_tmem + ks * 8, db + adv, idesc, 1u);
            }
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
    }
}

template <int N, int mode, int rot, int EL>
void run(long long* out) {
    const int iters = 2000;
    cudaFuncSetAttribute(probe<N, mode, rot, EL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int ctas : {1, 148}) {
        probe<N, mode, rot, EL><<<ctas, 128, 190 * 1024>>>(iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s (N=%d mode=%d rot=%d)\n", cudaGetErrorString(e), N, mode, rot); exit(1); }
        long long mx = 0; for (int i = 0; i < ctas; ++i) mx = out[i] > mx ? out[i] : mx;
        printf("%s %s N=%3d rot=%d ctas=%3d : %7.1f clk/MMA  (tensor floor %d)\n", EL ? "elect " : "lane==0", mode ? "TS" : "SS", N, rot, ctas,
               (double)mx / (iters * 4), 128 * N / 256 * 2);
    }
}
int main() {
    long long* out; cudaMallocManaged(&out, 148 * sizeof(long long));
    run<32, 0, 1, 0>(out); run<32, 0, 1, 1>(out); run<32, 0, 4, 1>(out); run<32, 1, 1, 1>(out); run<32, 1, 4, 1>(out);
    run<64, 0, 1, 0>(out); run<64, 0, 1, 1>(out); run<64, 0, 4, 1>(out); run<64, 1, 1, 1>(out); run<64, 1, 4, 1>(out);
    run<128, 0, 1, 0>(out); run<128, 0, 1, 1>(out); run<128, 0, 2, 1>(out); run<128, 1, 1, 1>(out); run<128, 1, 2, 1>(out);
    run<256, 0, 1, 0>(out); run<256, 0, 1, 1>(out); run<256, 1, 1, 1>(out);
    return 0;
}
