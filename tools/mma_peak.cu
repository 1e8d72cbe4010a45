// Bare tcgen05 main-loop probe: what the tensor pipe sustains on THIS chip when its operands already sit in shared memory.
// No TMA, no epilogue: one thread per CTA (pair) issues `iters` k-blocks of 4 tcgen05.mma (M = 128 per CTA, N = 256, one
// 128-byte swizzle row of K each) cycling over `stages` operand buffers, commits, waits.  Optional background warps copy
// 16 KB tiles through shared memory like the operand-split warps of conv_tc.cu do (bg = 1) to show the shared-memory
// contention.  Prints TFLOP/s per variant: kind::tf32 vs kind::f16 (bf16 operands), cta_group::1 vs ::2.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_peak tools/mma_peak.cu && tools/mma_peak
//
// These are the denominators bench.py's tensor-side fractions use; measured, not assumed.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// BF16: kind::f16 with bf16 operands (K = 16 per instruction), else kind::tf32 (K = 8).  CTA2: cta_group::2 pairs.
template <bool BF16, bool CTA2, int BN>
__global__ void __launch_bounds__(256, 1) mma_peak_kernel(int iters, int stages, int bg, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    constexpr uint32_t A_BYTES = 128 * 128;                       // 128 rows x one swizzle row
    constexpr uint32_t B_BYTES = (CTA2 ? BN / 2 : BN) * 128;      // a pair stages half of the weight tile per CTA
    const uint32_t ring = stages * (A_BYTES + B_BYTES);
    const uint32_t bar = base + ring;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + ring + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
    // operand data: small finite numbers (not zeros: zero operands draw less power and flatter the clocks)
    for (uint32_t i = threadIdx.x; i < ring / 4; i += blockDim.x) {
        uint32_t h = (i * 2654435761u) >> 9;
        if (BF16) reinterpret_cast<uint32_t*>(gen)[i] = 0x3C003C00u | (h & 0x007F007Fu);          // two bf16 near 0.0078
        else reinterpret_cast<uint32_t*>(gen)[i] = 0x3C000000u | (h & 0x007FE000u);                // fp32 near 0.0078
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        *reinterpret_cast<volatile uint32_t*>(gen + ring + 128) = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CTA2) cluster_sync_all();
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t MM = CTA2 ? 256 : 128;
    constexpr uint32_t fmt = BF16 ? 1u : 2u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
    unsigned long long t0 = 0, t1 = 0;
    if (warp == 1 && lane == 0 && rank == 0) {
        t0 = clock64();
        int s = 0;
        for (int it = 0; it < iters; ++it) {
            const uint64_t da = make_desc(base + s * (A_BYTES + B_BYTES)), db = make_desc(base + s * (A_BYTES + B_BYTES) + A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t koff = (uint64_t)((k * 32) >> 4);
                const uint32_t acc = (it | k) != 0 ? 1u : 0u;
                if (CTA2) {
                    if (BF16)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da + koff), "l"(db + koff), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da + koff), "l"(db + koff), "r"(idesc), "r"(acc) : "memory");
                } else {
                    if (BF16)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da + koff), "l"(db + koff), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da + koff), "l"(db + koff), "r"(idesc), "r"(acc) : "memory");
                }
            }
            if (++s == stages) s = 0;
        }
        if (CTA2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        mbar_wait(bar, 0);
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
        *reinterpret_cast<volatile uint32_t*>(gen + ring + 128) = 1u;      // tells the background warps to stop
    } else if (warp >= 4 && bg) {
        // background shared-memory traffic: 128 threads copy a 16 KB tile (read 16 KB + write 16 KB per round) from the
        // LAST stage's A buffer onto a scratch tile behind the ring, until the issuer is done
        const int t = threadIdx.x - 128;
        const float4* src = reinterpret_cast<const float4*>(gen + (stages - 1) * (A_BYTES + B_BYTES));
        float4* dst = reinterpret_cast<float4*>(gen + ring + 1024);
        volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(gen + ring + 128);
        const bool watch = !(CTA2 && rank != 0);                      // the peer has no issuer: it runs a fixed number of rounds
        for (int round = 0; watch ? (*stop == 0u) : (round < iters * bg); ++round) {
#pragma unroll 4
            for (int i = t; i < 1024; i += 128) {
                float4 v = src[i];
                v.x += 1.f;
                dst[i] = v;
            }
        }
    }
    if (CTA2 && rank != 0 && warp == 1 && lane == 0) mbar_wait(bar, 0);       // the multicast commit lands here too
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CTA2) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
    }
}

template <bool BF16, bool CTA2, int BN>
static void run(const char* name, int iters, int stages, int bg, int sms, unsigned long long* d_cycles) {
    auto kern = mma_peak_kernel<BF16, CTA2, BN>;
    const int smem = stages * (128 * 128 + (CTA2 ? BN / 2 : BN) * 128) + 1024 + 1024 + 16384 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CTA2 ? 2 : 1;
    at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(CTA2 ? (sms & ~1) : sms));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, stages, bg, d_cycles);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaEventSynchronize(e1);
        if (e != cudaSuccess || e2 != cudaSuccess) {
            printf("%-28s FAILED: %s / %s\n", name, cudaGetErrorString(e), cudaGetErrorString(e2));
            return;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double k_per_block = BF16 ? 64.0 : 32.0;
    const double macs_per_cta = (double)iters * 128.0 * BN * k_per_block;
    const double flops = 2.0 * macs_per_cta * (CTA2 ? (sms & ~1) : sms);
    printf("{\"variant\": \"%s\", \"bg\": %d, \"stages\": %d, \"ms\": %.4f, \"tflops\": %.1f, \"cycles_per_mma\": %.1f}\n", name, bg, stages, best,
           flops / (best * 1e-3) / 1e12, (double)cyc / (4.0 * iters));
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 20000;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d_cycles;
    cudaMalloc(&d_cycles, 4096 * sizeof(unsigned long long));
    for (int bg = 0; bg <= 1; ++bg) {
        run<false, false, 256>("tf32 cta1 N256", iters, 4, bg, sms, d_cycles);
        run<false, true, 256>("tf32 cta2 N256", iters, 4, bg, sms, d_cycles);
        run<true, false, 256>("bf16 cta1 N256", iters, 4, bg, sms, d_cycles);
        run<true, true, 256>("bf16 cta2 N256", iters, 4, bg, sms, d_cycles);
        run<true, false, 128>("bf16 cta1 N128", iters, 4, bg, sms, d_cycles);
        run<true, true, 128>("bf16 cta2 N128", iters, 4, bg, sms, d_cycles);
    }
    return 0;
}
