// Region-selected modulated 3x3 convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same contract as modconv_simt.cu (one launch = one StyledConv.forward of the reference for all regions,
// src/models/stylegan2/model.py:382-406), executed as an implicit GEMM per CTA tile:
//
//     D[128 pixels x NT out-channels] += A_tap[128 x 64] * B_tap[64 x NT]      for 9 taps x Cin/64 chunks
//
//   * M = 128 output pixels = an 8 x 16 patch of the (input-grid) image, 14 of the 16 columns valid; the two
//     spare columns let every tap be a pure ROW SHIFT of one staged halo tile (10 x 16 pixels): tap (dy,dx)
//     reads halo row r + 16*dy + dx + 1, so the activation is staged ONCE per 64-channel chunk and the nine
//     taps are nine shared-memory descriptors over the same bytes (start address moved by whole 128-B rows,
//     `base_offset` telling the MMA unit the swizzle phase of the first row).
//   * precision: the reference computes in fp32 and the parity bar is 1e-3, which single-pass bf16 (2^-9 per
//     product) misses and single-pass tf32 only grazes.  Both operands are split x = hi + lo (bf16 each) and
//     three MMAs hi*hi + hi*lo + lo*hi accumulate in fp32 TMEM: ~2^-17 relative error per product at 3 bf16
//     MMAs per tap - cheaper than one tf32 pass plus corrections, and the operands cost the same 4 B/element.
//   * B (weights): constant, pre-split to bf16 hi/lo planes [hilo][phase][tap][Cout][Cin] once on the host
//     side; streamed by TMA (128-B swizzle, K-major) through a 4-stage mbarrier ring.
//   * A (activations): fp32 pixel-major in HBM.  Four transform warps load the halo tile with 128-bit loads,
//     multiply by the style of the region being computed (model.py:276-277 folded onto the activation, the
//     reference's own non-fused form model.py:254), split to bf16 hi/lo and write the swizzled K-major
//     operand planes (2-stage ring), then fence the async proxy and signal the MMA warp.
//   * one elected thread issues tcgen05.mma; tcgen05.commit releases smem stages / publishes the accumulator.
//   * regions: a tile whose 112 valid pixels carry k distinct classes runs the main loop k times (k = 1 for
//     the vast majority of tiles of a face mask); the epilogue of pass c stores only rows of class c.
//   * epilogue (same four warps): tcgen05.ld 32 columns at a time, demodulate, add noise and bias, leaky-ReLU,
//     128-bit stores; each thread owns one pixel, so the pixel-major output row is contiguous.
//   * up-sampling layers use the folded parity kernels (see modconv_simt.cu / DESIGN.md): parity is one more
//     tile coordinate.
//
// Every mbarrier wait is bounded: a protocol bug traps the kernel (CUDA error at the next sync) instead of
// hanging the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>

#include "common.cuh"

namespace tc {

constexpr int TH = 8;            // tile rows
constexpr int TWP = 16;          // tile columns computed (M = TH*TWP = 128)
constexpr int TW = 14;           // ... of which valid
constexpr int KC = 64;           // input channels per chunk = one 128-byte swizzle row of bf16
constexpr int A_ROWS = 168;      // (TH+2)*16 = 160 halo rows, +1 leading, +slack; 168*128 B keeps 1024-B alignment
constexpr int A_PLANE = A_ROWS * 128;
constexpr int A_STAGE = 2 * A_PLANE;          // hi + lo
constexpr int NSTAGE_A = 2;
constexpr int NSTAGE_B = 4;
constexpr int NUM_WORKERS = 128;              // warps 2..5: transform + epilogue
constexpr int NUM_THREADS = 64 + NUM_WORKERS; // warp 0: TMA producer, warp 1: MMA issuer / TMEM owner

struct Params {
    const float* x;
    const float* s;
    const float* demod;
    const uint8_t* label;
    const float* noise;
    const float* noise_w;
    const float* bias;
    float* y;
    int batch, h, w, cin, cout, ncls, up, noise_b, act;
    int tiles_x, tiles_y, nt, shift_mode;
};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: ~seconds of spinning, then trap (a hang would cost a GPU strike; a trap is a clean CUDA error).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000ll) __trap();      // ~2 s at 2 GHz
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, 128-byte swizzle, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | base_offset [49,52) | layout [61,64) (2 = SW128)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t base_offset) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                       // LBO = 16 B (canonical K-major value; unused by swizzled layouts)
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));   // first source -> upper half
    return r;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---------------------------------------------------------------------------------------- kernel
template <int NT>
__global__ void __launch_bounds__(NUM_THREADS, 1) modconv3x3_tc_kernel(const __grid_constant__ CUtensorMap wmap, Params p) {
    constexpr int B_PLANE = NT * 128;        // NT rows x 64 bf16
    constexpr int B_STAGE = 2 * B_PLANE;     // hi + lo
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int TMEM_COLS = NT < 32 ? 32 : NT;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_buf = smem;                                   // [NSTAGE_A][hi|lo][A_ROWS][128]
    uint8_t* b_buf = a_buf + NSTAGE_A * A_STAGE;             // [NSTAGE_B][hi|lo][NT][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + NSTAGE_B * B_STAGE);
    // barrier indices
    constexpr int A_FULL = 0, A_EMPTY = A_FULL + NSTAGE_A, B_FULL = A_EMPTY + NSTAGE_A, B_EMPTY = B_FULL + NSTAGE_B,
                  ACC_FULL = B_EMPTY + NSTAGE_B, ACC_EMPTY = ACC_FULL + 1, NBARS = ACC_EMPTY + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    uint32_t* cls_mask = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates
    int bid = blockIdx.x;
    const int tile_x = bid % p.tiles_x;
    bid /= p.tiles_x;
    const int tile_y = bid % p.tiles_y;
    bid /= p.tiles_y;
    const int nphase = p.up ? 4 : 1;
    const int phase = bid % nphase;
    const int b = bid / nphase;
    const int py = phase >> 1, px = phase & 1;
    const int n0 = blockIdx.y * NT;
    const int mul = p.up ? 2 : 1;
    const int ho = p.h * mul, wo = p.w * mul;
    const int y0 = tile_y * TH, x0 = tile_x * TW;
    const int nchunks = p.cin / KC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE_A; ++i) mbar_init(smem_u32(&bars[A_FULL + i]), NUM_WORKERS), mbar_init(smem_u32(&bars[A_EMPTY + i]), 1);
        for (int i = 0; i < NSTAGE_B; ++i) mbar_init(smem_u32(&bars[B_FULL + i]), 1), mbar_init(smem_u32(&bars[B_EMPTY + i]), 1);
        mbar_init(smem_u32(&bars[ACC_FULL]), 1);
        mbar_init(smem_u32(&bars[ACC_EMPTY]), NUM_WORKERS);
        *cls_mask = 0u;
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    if (warp == 1) {   // TMEM allocation (one warp owns alloc + dealloc)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();

    // ---- which classes occur among this tile's valid output pixels (worker thread t owns M-row t)
    int my_cls = -1, my_oy = 0, my_ox = 0;
    if (warp >= 2) {
        const int r = threadIdx.x - 64;
        const int ty = r >> 4, tx = r & 15;
        const int iy = y0 + ty, ix = x0 + tx;
        if (tx < TW && iy < p.h && ix < p.w) {
            my_oy = iy * mul + py, my_ox = ix * mul + px;
            my_cls = p.label ? min((int)p.label[((int64_t)b * ho + my_oy) * wo + my_ox], p.ncls - 1) : 0;
            atomicOr(cls_mask, 1u << my_cls);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t classes = *cls_mask;

    if (warp == 0) {
        // =========================================================== TMA producer: weight tiles
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            const int rows_per_plane = nphase * 9 * p.cout;       // rows of the hi plane; lo plane follows
            for (uint32_t cm = classes; cm; cm &= cm - 1) {
                for (int kc = 0; kc < nchunks; ++kc) {
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(smem_u32(&bars[B_EMPTY + stage]), ph ^ 1);
                        const uint32_t full = smem_u32(&bars[B_FULL + stage]);
                        mbar_expect_tx(full, B_STAGE);
                        const uint32_t dst = smem_u32(b_buf + stage * B_STAGE);
                        const int row = (phase * 9 + tap) * p.cout + n0;
                        tma_load_2d(dst, &wmap, kc * KC, row, full);
                        tma_load_2d(dst + B_PLANE, &wmap, kc * KC, rows_per_plane + row, full);
                        if (++stage == NSTAGE_B) stage = 0, ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();      // lanes 1..31 must not run ahead to the block-wide barrier (aligned barriers need the whole warp)
    } else if (warp == 1) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0, pacc = 0;
            for (uint32_t cm = classes; cm; cm &= cm - 1) {
                mbar_wait(smem_u32(&bars[ACC_EMPTY]), pacc ^ 1);      // epilogue of the previous pass has drained TMEM
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kc = 0; kc < nchunks; ++kc) {
                    mbar_wait(smem_u32(&bars[A_FULL + sa]), pa);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(a_buf + sa * A_STAGE), a_lo = a_hi + A_PLANE;
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(smem_u32(&bars[B_FULL + sb]), pb);
                        tc_fence_after();
                        const uint32_t b_hi = smem_u32(b_buf + sb * B_STAGE), b_lo = b_hi + B_PLANE;
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        const uint32_t row_off = (uint32_t)(dy * TWP + dx + 1) * 128u;     // tap = row shift of the halo tile
                        const uint32_t boff = p.shift_mode == 0 ? ((a_hi + row_off) >> 7) & 7u : 0u;
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k) {
                            const uint64_t dah = smem_desc_sw128(a_hi + row_off + k * 32, boff);
                            const uint64_t dal = smem_desc_sw128(a_lo + row_off + k * 32, boff);
                            const uint64_t dbh = smem_desc_sw128(b_hi + k * 32, 0);
                            const uint64_t dbl = smem_desc_sw128(b_lo + k * 32, 0);
                            umma_bf16(tmem_base, dah, dbh, IDESC, accumulate);
                            umma_bf16(tmem_base, dah, dbl, IDESC, 1u);
                            umma_bf16(tmem_base, dal, dbh, IDESC, 1u);
                            accumulate = 1u;
                        }
                        umma_commit(smem_u32(&bars[B_EMPTY + sb]));
                        if (++sb == NSTAGE_B) sb = 0, pb ^= 1;
                    }
                    umma_commit(smem_u32(&bars[A_EMPTY + sa]));
                    if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                }
                umma_commit(smem_u32(&bars[ACC_FULL]));
                pacc ^= 1;
            }
        }
        __syncwarp();
    } else {
        // =========================================================== workers: transform (A operand) + epilogue
        const int t = threadIdx.x - 64;                 // 0..127
        const int c8 = t & 7;                           // which 8-channel (16-byte) chunk of the 64-channel row
        const int pbase = t >> 3;                       // halo pixel = pbase + 16*i, i = 0..9
        const float* xb = p.x + (int64_t)b * p.h * p.w * p.cin;
        const uint32_t quarter = (uint32_t)(warp & 3);  // TMEM lanes this warp may read
        const int m_row = quarter * 32 + lane;          // accumulator row (pixel) owned in the epilogue
        // epilogue pixel of this thread (may differ from the class-detection row `t`)
        int e_cls = -1, e_oy = 0, e_ox = 0;
        {
            const int ty = m_row >> 4, tx = m_row & 15;
            const int iy = y0 + ty, ix = x0 + tx;
            if (tx < TW && iy < p.h && ix < p.w) {
                e_oy = iy * mul + py, e_ox = ix * mul + px;
                e_cls = p.label ? min((int)p.label[((int64_t)b * ho + e_oy) * wo + e_ox], p.ncls - 1) : 0;
            }
        }
        (void)my_cls; (void)my_oy; (void)my_ox;
        int sa = 0;
        uint32_t pa = 0, pacc = 0;
        for (uint32_t cm = classes; cm; cm &= cm - 1) {
            const int cls = __ffs(cm) - 1;
            const float* sc = p.s + ((int64_t)b * p.ncls + cls) * p.cin;
            for (int kc = 0; kc < nchunks; ++kc) {
                const int ch = kc * KC + 8 * c8;
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + ch));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(sc + ch + 4));
                // issue all global loads of this chunk before touching shared memory
                float4 v0[10], v1[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const int hp = pbase + 16 * i;              // halo pixel index 0..159
                    const int gy = y0 - 1 + (hp >> 4), gx = x0 - 1 + (hp & 15);
                    v0[i] = make_float4(0.f, 0.f, 0.f, 0.f), v1[i] = v0[i];
                    if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
                        const float* src = xb + ((int64_t)gy * p.w + gx) * p.cin + ch;
                        v0[i] = __ldg(reinterpret_cast<const float4*>(src));
                        v1[i] = __ldg(reinterpret_cast<const float4*>(src + 4));
                    }
                }
                mbar_wait(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1);
                uint8_t* hi_plane = a_buf + sa * A_STAGE;
                uint8_t* lo_plane = hi_plane + A_PLANE;
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const int row = pbase + 16 * i + 1;          // +1: room for the dx = -1 shift
                    float f[8] = {v0[i].x * s0.x, v0[i].y * s0.y, v0[i].z * s0.z, v0[i].w * s0.w,
                                  v1[i].x * s1.x, v1[i].y * s1.y, v1[i].z * s1.z, v1[i].w * s1.w};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float h0 = bf16_round(f[2 * j]), h1 = bf16_round(f[2 * j + 1]);
                        hi[j] = pack_bf16x2(h0, h1);
                        lo[j] = pack_bf16x2(f[2 * j] - h0, f[2 * j + 1] - h1);
                    }
                    const uint32_t off = (uint32_t)row * 128u + (uint32_t)((c8 ^ (row & 7)) << 4);   // 128-B swizzle
                    *reinterpret_cast<uint4*>(hi_plane + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(lo_plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async();                              // generic-proxy stores -> visible to tcgen05 (async proxy)
                mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
            }
            // ---- epilogue of this class pass
            mbar_wait(smem_u32(&bars[ACC_FULL]), pacc);
            pacc ^= 1;
            tc_fence_after();
            const bool mine = (e_cls == cls);
            const float nw = (p.noise && p.noise_w) ? __ldg(p.noise_w) : 0.f;
            float nz = 0.f;
            if (mine && p.noise) nz = nw * __ldg(p.noise + ((int64_t)(p.noise_b == 1 ? 0 : b) * ho + e_oy) * wo + e_ox);
            const float* dm = p.demod ? p.demod + ((int64_t)b * p.ncls + cls) * p.cout + n0 : nullptr;
            float* dst = p.y + (((int64_t)b * ho + e_oy) * wo + e_ox) * p.cout + n0;
#pragma unroll 1
            for (int j = 0; j < NT / 32; ++j) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((quarter * 32u) << 16) + (uint32_t)(j * 32), r);
                if (mine) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int co = j * 32 + 4 * q;
                        float4 d = dm ? __ldg(reinterpret_cast<const float4*>(dm + co)) : make_float4(1.f, 1.f, 1.f, 1.f);
                        float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 o;
                        o.x = __uint_as_float(r[4 * q + 0]) * d.x + nz + bv.x;
                        o.y = __uint_as_float(r[4 * q + 1]) * d.y + nz + bv.y;
                        o.z = __uint_as_float(r[4 * q + 2]) * d.z + nz + bv.z;
                        o.w = __uint_as_float(r[4 * q + 3]) * d.w + nz + bv.w;
                        if (p.act) {
                            const float k = 1.41421356237309515f;
                            o.x = lrelu_scaled(o.x, 0.2f, k), o.y = lrelu_scaled(o.y, 0.2f, k);
                            o.z = lrelu_scaled(o.z, 0.2f, k), o.w = lrelu_scaled(o.w, 0.2f, k);
                        }
                        *reinterpret_cast<float4*>(dst + co) = o;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars[ACC_EMPTY]));
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            cudaGetLastError();
    });
    return fn;
}

template <int NT>
static int launch(const CUtensorMap& map, Params p, cudaStream_t st) {
    p.nt = NT;
    constexpr size_t smem = 1024 + NSTAGE_A * A_STAGE + NSTAGE_B * 2 * NT * 128 + 256;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(modconv3x3_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)cudaGetLastError();
        attr = true;
    }
    int64_t nblk = (int64_t)p.tiles_x * p.tiles_y * (p.up ? 4 : 1) * p.batch;
    if (nblk >= (1ll << 31)) return E4S_ERR_SHAPE;
    dim3 grid((unsigned)nblk, (unsigned)(p.cout / NT));
    modconv3x3_tc_kernel<NT><<<grid, NUM_THREADS, smem, st>>>(map, p);
    return e4s_launch_status();
}

}  // namespace tc

extern "C" int e4s_modconv3x3_tc_fwd(const float* x, const void* w_hilo_bf16, const float* s, const float* demod,
                                     const uint8_t* label, const float* noise, const float* noise_w, const float* bias,
                                     float* y, int batch, int h, int w, int cin, int cout, int ncls, int up, int noise_b,
                                     int act, int shift_mode, void* stream) {
    E4S_REQUIRE(x && w_hilo_bf16 && s && y, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && ncls > 0 && ncls <= 32, E4S_ERR_ARG);
    E4S_REQUIRE((cin % tc::KC) == 0 && (cout % 32) == 0 && (cout <= 128 || cout % 128 == 0), E4S_ERR_SHAPE);
    E4S_REQUIRE(label || ncls == 1, E4S_ERR_ARG);
    E4S_REQUIRE(!noise || (noise_w && (noise_b == 1 || noise_b == batch)), E4S_ERR_ARG);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w_hilo_bf16) && e4s_aligned16(s) && e4s_aligned16(y) &&
                    (!demod || e4s_aligned16(demod)) && (!bias || e4s_aligned16(bias)),
                E4S_ERR_ALIGN);
    tc::EncodeTiledFn enc = tc::encode_fn();
    E4S_REQUIRE(enc != nullptr, E4S_ERR_ARCH);
    const int nt = cout >= 128 ? 128 : cout;       // 128, 64 or 32
    const int nphase = up ? 4 : 1;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)2 * nphase * 9 * cout};
    cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::KC, (cuuint32_t)nt};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_hilo_bf16), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return 700 + (int)cr;
    tc::Params p{x, s, demod, label, noise, noise_w, bias, y, batch, h, w, cin, cout, ncls, up ? 1 : 0, noise_b, act,
                 (int)e4s_ceil_div(w, tc::TW), (int)e4s_ceil_div(h, tc::TH), nt, shift_mode};
    cudaStream_t st = (cudaStream_t)stream;
    if (nt == 128) return tc::launch<128>(map, p, st);
    if (nt == 64) return tc::launch<64>(map, p, st);
    return tc::launch<32>(map, p, st);
}
