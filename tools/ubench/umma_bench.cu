// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, M = 128 or 64, K = 16) as a function of N, of the
// shared-memory swizzle of the K-major operands, of a row-shifted A start address and of A-in-TMEM.  One CTA, one issuing
// thread, operands all zero (timing only).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

struct Result { int m, n, sw, mode, shift; float cyc; };

// mode 0: A and B from shared memory; 1: A from TMEM; 2: SS with two alternating accumulators (no D dependency)
__global__ void __launch_bounds__(256, 1) bench(Result* out, int* nout, int m_count, int mode_mask, int nissue, int same_d) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), nissue); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    uint32_t parity = 0;
    int cnt = 0;
    {
        const int ITERS = 128;
        for (int sw = 0; sw < 2; ++sw) {                       // 0: 128-byte swizzle (64-channel rows), 1: 64-byte swizzle
            const uint32_t rowb = sw == 0 ? 128 : 64;
            const uint32_t desc_hi = (uint32_t)((sw == 0 ? 1024u : 512u) >> 4) | (1u << 14) | ((sw == 0 ? 2u : 4u) << 29);
            const int ksteps = sw == 0 ? 4 : 2;
            for (int mi = 0; mi < m_count; ++mi) {
                const int m = mi == 0 ? 128 : 64;
                for (int mode = 0; mode < 3; ++mode)
                    for (int shift = 0; shift < 2; ++shift) {
                        if ((mode == 1 && shift) || !((mode_mask >> mode) & 1)) continue;
                        for (int n = 32; n <= 256; n *= 2) {
                            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
                            const uint32_t a_addr = smem_u32(smem) + (shift ? 3 * rowb : 0);
                            const uint32_t b_addr = smem_u32(smem) + 64 * 1024;
                            const uint64_t ad = ((uint64_t)desc_hi << 32) | ((a_addr >> 4) | 0x10000u);
                            const uint64_t bd = ((uint64_t)desc_hi << 32) | ((b_addr >> 4) | 0x10000u);
                            long long t0 = 0, t1 = 0;
                            // warp-uniform operands (uniform registers feed UTCHMMA directly, no R2UR in the loop)
                            const uint32_t ad_lo = __shfl_sync(0xffffffffu, (uint32_t)ad, 0), bd_lo = __shfl_sync(0xffffffffu, (uint32_t)bd, 0);
                            const uint32_t d_u = __shfl_sync(0xffffffffu, tmem + (uint32_t)(same_d ? 0 : (warp & 7) * 32), 0);
                            const bool active = warp < nissue;
                            for (int rep = 0; rep < 2; ++rep) {             // first repetition warms up
                                asm volatile("bar.sync 1, 256;" ::: "memory");
                                t0 = clock64();
                                if (active) {
#pragma unroll 1
                                    for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
                                        for (int u = 0; u < 4; ++u)
#pragma unroll
                                            for (int k = 0; k < 4; ++k) {
                                                if (k >= ksteps) continue;
                                                const uint32_t d = d_u + ((mode == 2 && (u & 1)) ? 256u : 0u);
                                                const uint64_t a64 = ((uint64_t)desc_hi << 32) | (ad_lo + 2 * k), b64 = ((uint64_t)desc_hi << 32) | (bd_lo + 2 * k);
                                                if (lane == 0) {
                                                    if (mode == 1) umma_ts(d, tmem + 256u + 8u * k, b64, idesc, 1u);
                                                    else umma_ss(d, a64, b64, idesc, 1u);
                                                }
                                            }
                                    }
                                    if (lane == 0) umma_commit(smem_u32(&bar));
                                    __syncwarp();
                                }
                                while (!mbar_try_wait(smem_u32(&bar), parity)) {}
                                parity ^= 1;
                                t1 = clock64();
                            }
                            if (threadIdx.x == 0) out[cnt] = Result{m, n, sw, mode, shift, (float)(t1 - t0) / (float)(ITERS * ksteps)};
                            ++cnt;
                        }
                    }
            }
        }
        if (threadIdx.x == 0) *nout = cnt;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

int main(int argc, char** argv) {
    const int m_count = argc > 1 ? atoi(argv[1]) : 1, mode_mask = argc > 2 ? atoi(argv[2]) : 5, nissue = argc > 3 ? atoi(argv[3]) : 1, same_d = argc > 4 ? atoi(argv[4]) : 0;
    Result* d_out; int* d_n;
    cudaMalloc(&d_out, sizeof(Result) * 256); cudaMalloc(&d_n, sizeof(int));
    cudaMemset(d_n, 0, sizeof(int));
    const int smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<<<1, 256, smem>>>(d_out, d_n, m_count, mode_mask, nissue, same_d);
    printf("issuing warps: %d, %s accumulator columns (cycles are per MMA of ONE warp; all warps issue concurrently)\n", nissue, same_d ? "the SAME" : "private");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    Result h[256]; int n = 0;
    cudaMemcpy(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(h, d_out, sizeof(Result) * 256, cudaMemcpyDeviceToHost);
    const char* modes[3] = {"A smem", "A tmem", "A smem, 2 accumulators"};
    printf("cycles per tcgen05.mma kind::f16 (bf16, K=16), one issuing thread, back-to-back\n");
    for (int i = 0; i < n; ++i)
        printf("M=%3d N=%3d swizzle=%3dB %-24s shift=%d : %7.1f cycles/MMA\n", h[i].m, h[i].n, h[i].sw == 0 ? 128 : 64, modes[h[i].mode], h[i].shift, h[i].cyc);
    return 0;
}
