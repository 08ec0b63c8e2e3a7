// Probe for the round-2 design of the up-sampling layers (data flow specified and checked on the
// CPU by tools/ubench/upconv_dataflow.py).  NOT part of the library - a standalone program to answer, on a B200, whether
// the "tap-free GEMM + combining epilogue" formulation of conv_transpose2d(stride 2) + 4x4 blur keeps the tensor pipe fed:
//
//   per 8x16 patch of input pixels (one-pixel halo, 6x14 interior) and per N tile of 16 output channels:
//     GEMM   Z[m, (k, o)] = sum_i xs[m, i] * W[k][o][i]        M = 128 pixels, N = 9 taps x 16 channels = 144, K = Cin
//            operands pre-split to bf16 hi/lo planes (x * style on the activation side), x_hi w_hi + x_lo w_hi + x_hi w_lo,
//            both loaded by TMA (4-D box for the patch, hardware zero fill outside the image) - no transform warps;
//     epilogue: TMEM -> horizontal combination with the x-neighbours (warp shuffles) -> shared-memory exchange ->
//            vertical combination with the rows above / below -> demodulation, bias, leaky ReLU -> 64-byte stores.
//
// Single region, ONE MMA-issuing warp (the product kernel would use three).  Prints the error against a
// double-precision conv_transpose2d + blur on a small case, then the time of a production shape and where CTA 0 waited.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o upconv_probe upconv_probe.cu -lcuda
// Run:   ./upconv_probe [B H W Cin Cout]          (default 16 32 32 512 512 = layer c6 of the 1024x1024 generator)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

constexpr int TH = 8, TW = 16, IH = 6, IW = 14;       // patch and interior (tools/ubench/upconv_dataflow.py)
constexpr int NTC = 16, N = 9 * NTC;                  // output channels per N tile; GEMM columns
constexpr int KC = 32, ROWB = KC * 2;                 // K chunk (64-byte swizzle rows)
constexpr int A_PLANE = 128 * ROWB, B_PLANE = N * ROWB;
constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;      // a_hi, a_lo, b_hi, b_lo = 34816 B
constexpr int NSTAGE = 5;
constexpr int EXCH = 6 * NTC * 128 * 4;               // g[(ky, px)][channel][pixel] fp32 = 49152 B
constexpr int NUM_THREADS = 32 * 6;                   // warp 0 TMA, warp 1 MMA issue, warps 2-5 epilogue
constexpr int TMEM_COLS = 512;                        // two accumulators of 144 columns
constexpr float SQRT2 = 1.41421356237309515f;

struct Params {
    const float* demod;   // [B, Cout]
    const float* bias;    // [Cout]
    float* y;             // [B, 2H, 2W, Cout]
    float bf[4];          // flipped 1-D blur taps of one axis (outer(bf, bf) = the flipped 4x4 FIR)
    int batch, h, w, cin, cout;
    int tiles_x, tiles_y, n_tiles, items;
    long long* prof;      // [3 roles][2]: role cycles, cycles in barrier waits (CTA 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {       // bounded: a protocol bug traps instead of hanging
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major operand, 64-byte swizzle: 8-row atoms 512 B apart, layout code 4 (csrc/modconv_tcr.cu:smem_desc<32>)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// four consecutive accumulator columns of this thread's TMEM lane; the registers are valid only after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Item {
    int b, ty, tx, nt;
};
__device__ __forceinline__ Item decode_item(const Params& p, int it) {
    Item r;
    const int ptiles = p.tiles_x * p.tiles_y * p.batch;
    r.nt = it / ptiles;
    int pt = it - r.nt * ptiles;
    r.tx = pt % p.tiles_x;
    pt /= p.tiles_x;
    r.ty = pt % p.tiles_y;
    r.b = pt / p.tiles_y;
    return r;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
upconv_probe_kernel(const __grid_constant__ CUtensorMap xh_map, const __grid_constant__ CUtensorMap xl_map,
                    const __grid_constant__ CUtensorMap wh_map, const __grid_constant__ CUtensorMap wl_map, Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stages = smem;
    float* exch = reinterpret_cast<float*>(smem + NSTAGE * STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE + EXCH);
    const int FULL = 0, EMPTY = NSTAGE, ACC_FULL = 2 * NSTAGE, ACC_EMPTY = ACC_FULL + 2, NBARS = ACC_EMPTY + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.cin / KC;
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && lane == 0;
    long long waited = 0;
    const long long t_role = clock64();

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) mbar_init(smem_u32(&bars[FULL + i]), 1), mbar_init(smem_u32(&bars[EMPTY + i]), 1);
        for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&bars[ACC_FULL + i]), 1), mbar_init(smem_u32(&bars[ACC_EMPTY + i]), 128);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================ producer: four TMA loads per K chunk
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
                const Item item = decode_item(p, it);
                const int x0 = item.tx * IW - 1, y0 = item.ty * IH - 1;
                for (int kc = 0; kc < nchunks; ++kc) {
                    const long long t0 = clock64();
                    mbar_wait(smem_u32(&bars[EMPTY + st]), ph ^ 1);
                    waited += clock64() - t0;
                    const uint32_t full = smem_u32(&bars[FULL + st]);
                    const uint32_t dst = smem_u32(stages + (size_t)st * STAGE);
                    mbar_expect_tx(full, STAGE);
                    tma_load_4d(dst, &xh_map, kc * KC, x0, y0, item.b, full);
                    tma_load_4d(dst + A_PLANE, &xl_map, kc * KC, x0, y0, item.b, full);
                    tma_load_2d(dst + 2 * A_PLANE, &wh_map, kc * KC, item.nt * N, full);
                    tma_load_2d(dst + 2 * A_PLANE + B_PLANE, &wl_map, kc * KC, item.nt * N, full);
                    if (++st == NSTAGE) st = 0, ph ^= 1;
                }
            }
        }
        if (prof_on) p.prof[0] = clock64() - t_role, p.prof[1] = waited;
    } else if (warp == 1) {
        // ================================================================ MMA issue (one thread)
        constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (lane == 0) {
            int st = 0, acc = 0;
            uint32_t ph = 0, pacc[2] = {0, 0};
            for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
                long long t0 = clock64();
                mbar_wait(smem_u32(&bars[ACC_EMPTY + acc]), pacc[acc] ^ 1);
                waited += clock64() - t0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                for (int kc = 0; kc < nchunks; ++kc) {
                    t0 = clock64();
                    mbar_wait(smem_u32(&bars[FULL + st]), ph);
                    waited += clock64() - t0;
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(stages + (size_t)st * STAGE), a_lo = a_hi + A_PLANE;
                    const uint32_t b_hi = a_hi + 2 * A_PLANE, b_lo = b_hi + B_PLANE;
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k) {
                        const uint32_t ko = k * 32;
                        umma_bf16(d_tmem, smem_desc(a_hi + ko), smem_desc(b_hi + ko), IDESC, (kc | k) ? 1u : 0u);
                        umma_bf16(d_tmem, smem_desc(a_lo + ko), smem_desc(b_hi + ko), IDESC, 1u);
                        umma_bf16(d_tmem, smem_desc(a_hi + ko), smem_desc(b_lo + ko), IDESC, 1u);
                    }
                    umma_commit(smem_u32(&bars[EMPTY + st]));
                    if (++st == NSTAGE) st = 0, ph ^= 1;
                }
                umma_commit(smem_u32(&bars[ACC_FULL + acc]));
                pacc[acc] ^= 1;
                acc ^= 1;
            }
        }
        if (prof_on) p.prof[2] = clock64() - t_role, p.prof[3] = waited;
    } else {
        // ================================================================ epilogue: 128 threads, TMEM lane = patch pixel
        const uint32_t quarter = (uint32_t)(warp & 3);
        const int m = (int)quarter * 32 + lane;                 // patch pixel: row m >> 4, column m & 15
        const int py_ = m >> 4, px_ = m & 15;
        const bool interior = py_ >= 1 && py_ <= IH && px_ >= 1 && px_ <= IW;
        const float bf0 = p.bf[0], bf1 = p.bf[1], bf2 = p.bf[2], bf3 = p.bf[3];
        const int ho = 2 * p.h, wo = 2 * p.w;
        int acc = 0;
        uint32_t pacc[2] = {0, 0};
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const long long t0 = clock64();
            mbar_wait(smem_u32(&bars[ACC_FULL + acc]), pacc[acc]);
            waited += clock64() - t0;
            pacc[acc] ^= 1;
            tc_fence_after();
            const uint32_t lanes = tmem_base + ((quarter * 32u) << 16) + (uint32_t)(acc * N);
            // ---- horizontal combination, four channels at a time: z[k][c] = column k * 16 + c of this pixel's lane
#pragma unroll 1
            for (int cg = 0; cg < NTC / 4; ++cg) {
                uint32_t zr[9][4];
#pragma unroll
                for (int k = 0; k < 9; ++k) tmem_ld4(lanes + (uint32_t)(k * NTC + cg * 4), zr[k]);
                tmem_wait_ld();
                float z[9][4];
#pragma unroll
                for (int k = 0; k < 9; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) z[k][c] = __uint_as_float(zr[k][c]);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float c0 = z[ky * 3 + 0][c], c1 = z[ky * 3 + 1][c], c2 = z[ky * 3 + 2][c];
                        const float l1 = __shfl_up_sync(0xffffffffu, c1, 1), l2 = __shfl_up_sync(0xffffffffu, c2, 1);       // pixel x - 1
                        const float r0 = __shfl_down_sync(0xffffffffu, c0, 1), r1 = __shfl_down_sync(0xffffffffu, c1, 1);   // pixel x + 1
                        // upconv_dataflow.py:axis_coefficients - a = 2 (d - 1) + k + 1 - p
                        const float g0 = bf0 * l1 + bf1 * l2 + bf1 * c0 + bf2 * c1 + bf3 * c2 + bf3 * r0;                   // output column 2 x
                        const float g1 = bf0 * l2 + bf0 * c0 + bf1 * c1 + bf2 * c2 + bf2 * r0 + bf3 * r1;                   // output column 2 x + 1
                        const int ch = cg * 4 + c;
                        exch[((ky * 2 + 0) * NTC + ch) * 128 + m] = g0;
                        exch[((ky * 2 + 1) * NTC + ch) * 128 + m] = g1;
                    }
                }
            }
            // the accumulator is free as soon as every epilogue thread has read it
            tc_fence_before();
            mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
            acc ^= 1;
            epi_barrier();
            // ---- vertical combination + per-output epilogue, interior pixels only
            const int iy = item.ty * IH + py_ - 1, ix = item.tx * IW + px_ - 1;
            if (interior && iy < p.h && ix < p.w) {
                const int n0 = item.nt * NTC;
                const float* dm = p.demod + (size_t)item.b * p.cout + n0;
#pragma unroll
                for (int opx = 0; opx < 2; ++opx) {
                    float o0[NTC], o1[NTC];
#pragma unroll
                    for (int ch = 0; ch < NTC; ++ch) {
                        const float* gp = exch + (size_t)ch * 128 + m;                    // + (ky * 2 + opx) * NTC * 128, rows at -16 / 0 / +16
                        const float u1 = gp[((1 * 2 + opx) * NTC) * 128 - 16], u2 = gp[((2 * 2 + opx) * NTC) * 128 - 16];
                        const float m0 = gp[((0 * 2 + opx) * NTC) * 128], m1 = gp[((1 * 2 + opx) * NTC) * 128], m2 = gp[((2 * 2 + opx) * NTC) * 128];
                        const float d0 = gp[((0 * 2 + opx) * NTC) * 128 + 16], d1 = gp[((1 * 2 + opx) * NTC) * 128 + 16];
                        o0[ch] = bf0 * u1 + bf1 * u2 + bf1 * m0 + bf2 * m1 + bf3 * m2 + bf3 * d0;      // output row 2 y
                        o1[ch] = bf0 * u2 + bf0 * m0 + bf1 * m1 + bf2 * m2 + bf2 * d0 + bf3 * d1;      // output row 2 y + 1
                    }
#pragma unroll
                    for (int opy = 0; opy < 2; ++opy) {
                        float* dst = p.y + (((size_t)item.b * ho + (2 * iy + opy)) * wo + (2 * ix + opx)) * p.cout + n0;
#pragma unroll
                        for (int q = 0; q < NTC / 4; ++q) {
                            float4 v;
                            float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int ch = 4 * q + e;
                                float t = (opy ? o1[ch] : o0[ch]) * __ldg(dm + ch) + __ldg(p.bias + n0 + ch);
                                vv[e] = (t > 0.f ? t : 0.2f * t) * SQRT2;
                            }
                            *reinterpret_cast<float4*>(dst + 4 * q) = v;
                        }
                    }
                }
            }
            epi_barrier();                                   // the exchange buffer is rewritten by the next item
        }
        if (prof_on && warp == 2) p.prof[4] = clock64() - t_role, p.prof[5] = waited;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// x * style -> bf16 hi / lo planes (what the producing layer's epilogue would emit in the product)
__global__ void split_planes_kernel(const float* __restrict__ x, const float* __restrict__ s, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, size_t per_sample, int cin, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i] * s[(i / per_sample) * cin + (i % cin)];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline void ck(cudaError_t e, const char* file, int line) {
    if (e != cudaSuccess) {
        printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line);
        exit(1);
    }
}
#define CK(call) ck((call), __FILE__, __LINE__)

struct Problem {
    int b, h, w, cin, cout;
};

float frand(uint32_t& st) {
    st = st * 1664525u + 1013904223u;
    return ((st >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}

// returns max-rel error against the double-precision reference (check = true, small problems) and the kernel time
void run(const Problem& pr, bool check, int iters) {
    const int B = pr.b, H = pr.h, W = pr.w, CI = pr.cin, CO = pr.cout;
    if (CI % KC || CO % NTC) {
        printf("Cin must be a multiple of %d, Cout of %d\n", KC, NTC);
        exit(1);
    }
    uint32_t seed = 12345u + (uint32_t)(B * 131 + H * 17 + CI);
    const size_t nx = (size_t)B * H * W * CI, ny = (size_t)B * 4 * H * W * CO;
    std::vector<float> x(nx), s((size_t)B * CI), wt((size_t)9 * CO * CI), dm((size_t)B * CO), bias(CO);
    for (auto& v : x) v = frand(seed);
    for (auto& v : s) v = 1.0f + 0.3f * frand(seed);
    const float wscale = 1.0f / std::sqrt(9.0f * CI);
    for (auto& v : wt) v = frand(seed) * wscale;           // [k][o][i]
    for (auto& v : dm) v = 1.0f + 0.2f * frand(seed);
    for (auto& v : bias) v = 0.1f * frand(seed);
    const float taps[4] = {0.25f, 0.75f, 0.75f, 0.25f};     // [1,3,3,1] / 8 * sqrt(4): outer(t, t) = the up-sampling blur

    // weights as GEMM rows [nt][k][o16][Cin], split to bf16 hi / lo on the host
    const int n_tiles = CO / NTC;
    std::vector<__nv_bfloat16> wh((size_t)n_tiles * N * CI), wl(wh.size());
    for (int nt = 0; nt < n_tiles; ++nt)
        for (int k = 0; k < 9; ++k)
            for (int o = 0; o < NTC; ++o)
                for (int i = 0; i < CI; ++i) {
                    const float v = wt[((size_t)k * CO + nt * NTC + o) * CI + i];
                    const __nv_bfloat16 hb = __float2bfloat16_rn(v);
                    const size_t dst = ((size_t)nt * N + k * NTC + o) * CI + i;
                    wh[dst] = hb;
                    wl[dst] = __float2bfloat16_rn(v - __bfloat162float(hb));
                }

    float *d_x, *d_s, *d_dm, *d_bias, *d_y;
    __nv_bfloat16 *d_xh, *d_xl, *d_wh, *d_wl;
    long long* d_prof;
    CK(cudaMalloc(&d_x, nx * 4)), CK(cudaMalloc(&d_s, s.size() * 4)), CK(cudaMalloc(&d_dm, dm.size() * 4));
    CK(cudaMalloc(&d_bias, bias.size() * 4)), CK(cudaMalloc(&d_y, ny * 4)), CK(cudaMalloc(&d_xh, nx * 2)), CK(cudaMalloc(&d_xl, nx * 2));
    CK(cudaMalloc(&d_wh, wh.size() * 2)), CK(cudaMalloc(&d_wl, wl.size() * 2)), CK(cudaMalloc(&d_prof, 6 * sizeof(long long)));
    CK(cudaMemcpy(d_x, x.data(), nx * 4, cudaMemcpyHostToDevice)), CK(cudaMemcpy(d_s, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_dm, dm.data(), dm.size() * 4, cudaMemcpyHostToDevice)), CK(cudaMemcpy(d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_wh, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice)), CK(cudaMemcpy(d_wl, wl.data(), wl.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_y, 0xff, ny * 4)), CK(cudaMemset(d_prof, 0, 6 * sizeof(long long)));
    split_planes_kernel<<<1024, 256>>>(d_x, d_s, d_xh, d_xl, (size_t)H * W * CI, CI, nx);
    CK(cudaGetLastError());

    EncodeTiledFn enc = nullptr;
    {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
        enc = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    CUtensorMap xh_map, xl_map, wh_map, wl_map;
    {
        cuuint64_t xd[4] = {(cuuint64_t)CI, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t xs[3] = {(cuuint64_t)CI * 2, (cuuint64_t)W * CI * 2, (cuuint64_t)H * W * CI * 2};
        cuuint32_t xb[4] = {(cuuint32_t)KC, (cuuint32_t)TW, (cuuint32_t)TH, 1}, one4[4] = {1, 1, 1, 1};
        cuuint64_t wd[2] = {(cuuint64_t)CI, (cuuint64_t)n_tiles * N};
        cuuint64_t ws[1] = {(cuuint64_t)CI * 2};
        cuuint32_t wb[2] = {(cuuint32_t)KC, (cuuint32_t)N}, one2[2] = {1, 1};
        CUresult r1 = enc(&xh_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_xh, xd, xs, xb, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&xl_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_xl, xd, xs, xb, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r3 = enc(&wh_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_wh, wd, ws, wb, one2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r4 = enc(&wl_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_wl, wd, ws, wb, one2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 || r2 || r3 || r4) {
            printf("cuTensorMapEncodeTiled failed: %d %d %d %d\n", (int)r1, (int)r2, (int)r3, (int)r4);
            exit(1);
        }
    }

    Params p{};
    p.demod = d_dm, p.bias = d_bias, p.y = d_y;
    for (int i = 0; i < 4; ++i) p.bf[i] = taps[3 - i];
    p.batch = B, p.h = H, p.w = W, p.cin = CI, p.cout = CO;
    p.tiles_x = (W + IW - 1) / IW, p.tiles_y = (H + IH - 1) / IH, p.n_tiles = n_tiles;
    p.items = p.tiles_x * p.tiles_y * B * n_tiles;
    p.prof = d_prof;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = p.items < sms ? p.items : sms;
    const size_t smem = 1024 + (size_t)NSTAGE * STAGE + EXCH + 14 * 8 + 64;
    CK(cudaFuncSetAttribute(upconv_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upconv_probe_kernel<<<grid, NUM_THREADS, smem>>>(xh_map, xl_map, wh_map, wl_map, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    if (check) {
        // double-precision conv_transpose2d(stride 2) + blur (true convolution with outer(taps, taps), pad (1, 1))
        std::vector<float> y(ny);
        CK(cudaMemcpy(y.data(), d_y, ny * 4, cudaMemcpyDeviceToHost));
        const int UH = 2 * H + 1, UW = 2 * W + 1;
        double worst = 0.0, big = 0.0;
        std::vector<double> u((size_t)UH * UW);
        for (int b = 0; b < B; ++b)
            for (int o = 0; o < CO; ++o) {
                std::fill(u.begin(), u.end(), 0.0);
                for (int qy = 0; qy < H; ++qy)
                    for (int qx = 0; qx < W; ++qx)
                        for (int k = 0; k < 9; ++k) {
                            double acc = 0.0;
                            const float* xr = &x[(((size_t)b * H + qy) * W + qx) * CI];
                            const float* wr = &wt[((size_t)k * CO + o) * CI];
                            for (int i = 0; i < CI; ++i) acc += (double)xr[i] * s[(size_t)b * CI + i] * wr[i];
                            u[(size_t)(2 * qy + k / 3) * UW + 2 * qx + k % 3] += acc;
                        }
                for (int oy = 0; oy < 2 * H; ++oy)
                    for (int ox = 0; ox < 2 * W; ++ox) {
                        double v = 0.0;
                        for (int a = 0; a < 4; ++a)
                            for (int c = 0; c < 4; ++c) {
                                const int uy = oy + a - 1, ux = ox + c - 1;
                                if (uy < 0 || uy >= UH || ux < 0 || ux >= UW) continue;
                                v += (double)taps[3 - a] * taps[3 - c] * u[(size_t)uy * UW + ux];
                            }
                        double t = v * dm[(size_t)b * CO + o] + bias[o];
                        t = (t > 0 ? t : 0.2 * t) * 1.4142135623730951;
                        const double got = y[(((size_t)b * 2 * H + oy) * 2 * W + ox) * CO + o];
                        worst = std::fmax(worst, std::fabs(got - t));
                        big = std::fmax(big, std::fabs(t));
                    }
            }
        printf("check  B=%d %dx%d %d->%d: max|err| / max|ref| = %.3e (%s)\n", B, H, W, CI, CO, worst / big, worst / big < 1e-4 ? "ok" : "WRONG");
    }
    if (iters > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)), CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) upconv_probe_kernel<<<grid, NUM_THREADS, smem>>>(xh_map, xl_map, wh_map, wl_map, p);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
        long long prof[6];
        CK(cudaMemcpy(prof, d_prof, sizeof(prof), cudaMemcpyDeviceToHost));
        const double flops = 2.0 * 9 * CI * CO * (double)B * H * W;
        printf("time   B=%d %dx%d %d->%d: %.3f ms, %.1f algorithmic TFLOP/s (x3 bf16 issued, x%.2f for the patch halo); items %d on %d CTAs\n", B, H, W,
               CI, CO, ms, flops / (ms * 1e-3) / 1e12, 128.0 / (IH * IW), p.items, grid);
        const char* role[3] = {"TMA producer", "MMA issue", "epilogue (warp 2)"};
        for (int r = 0; r < 3; ++r)
            printf("  CTA 0 %-18s %12lld cycles, %5.1f %% in barrier waits\n", role[r], prof[2 * r], prof[2 * r] ? 100.0 * prof[2 * r + 1] / prof[2 * r] : 0.0);
    }
    cudaFree(d_x), cudaFree(d_s), cudaFree(d_dm), cudaFree(d_bias), cudaFree(d_y), cudaFree(d_xh), cudaFree(d_xl), cudaFree(d_wh), cudaFree(d_wl), cudaFree(d_prof);
}

}  // namespace

int main(int argc, char** argv) {
    run({2, 13, 31, 64, 32}, true, 0);          // ragged tiles, two N tiles, two K chunks
    run({1, 6, 14, 32, 16}, true, 0);           // exactly one tile, one chunk
    Problem pr{16, 32, 32, 512, 512};           // c6: 512 -> 512, 32x32 -> 64x64, 16 faces
    if (argc >= 6) pr = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
    run(pr, false, 20);
    return 0;
}
