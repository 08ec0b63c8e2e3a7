// Input- and style-gradient of the region-selected modulated 3x3 convolution on tcgen05 tensor cores (sm_100a).
//
// Math (csrc/modconv_bwd.cu header): with gv = act'(y) * gy,
//     G_c[q,i] = sum_{o,k} Wd[k][i][o] * ( gv[q+k-1,o] * d[c,o] * [region(q+k-1) == c] )       one dgrad conv per region
//     gx[q,i]  = sum_c s[c,i] * G_c[q,i]                gs_conv[c,i] = sum_q x[q,i] * G_c[q,i]
// Execution = the persistent pipeline of modconv_tcp.cu (TMA weight ring, single-thread tcgen05.mma issue, TMEM
// double buffering, split-bf16 x3) with three changes:
//   * the staged operand is the OUTPUT gradient: the transform warps read gy (and y for the leaky-ReLU derivative),
//     scale by the region's demodulation row and zero every halo pixel that belongs to another region.  Masking is per
//     staged row, so the nine taps stay row-shifted descriptors over one staged tile;
//   * K runs over Cout (x 4 parity planes for an up-sampling layer, gathered with stride 2 from gy), N over Cin; the
//     weights are the forward planes with taps flipped and channels transposed, pre-split to bf16 hi/lo;
//   * a tile whose halo touches k regions runs k passes; the epilogue of pass c accumulates s[c,:] * acc into gx
//     (first pass writes, later passes add) and, when styles need gradients, reduces x * acc over the tile's pixels
//     with a shuffle tree and adds it to gs[b, c, :].
// This replaces, per region per layer, the cuDNN dgrad AND wgrad the reference's autograd runs (model.py:277-316).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace tcd {

constexpr int TH = 8, TWP = 16, TW = 14;
constexpr int A_ROWS = 168;
constexpr int NSTAGE_A = 2;
constexpr int NUM_MMA_WARPS = 3;              // one issuing warp per split-precision product (see modconv_tcr.cu)
constexpr int NUM_THREADS = 384;              // warps: 0 weights (TMA), 1-3 MMA issue, 4-7 transform, 8-11 epilogue
constexpr int NUM_XFORM = 128, NUM_EPI = 128;
constexpr int SMEM_BUDGET = 227 * 1024 - 2048;
constexpr float SQRT2 = 1.41421356237309515f;

struct Params {
    const float* gy;       // [B, Ho, Wo, Cout]
    const float* y;        // forward output (activation derivative) or NULL
    const float* x;        // [B, H, W, Cin] forward input (style gradient) or NULL
    const float* s;        // [B, ncls, Cin]
    const float* demod;    // [B, ncls, Cout] or NULL
    const uint8_t* label;  // [B, Ho, Wo] or NULL
    float* gx;             // [B, H, W, Cin] or NULL
    float* gs;             // [B, ncls, Cin] (atomic accumulate) or NULL
    int batch, h, w, cin, cout, ncls, act;
    int tiles_x, tiles_y, n_tiles, items, nslot_b;
    int gsplit, hsplit;    // a (pixel tile, N tile) pair is cut into gsplit x hsplit work items: region passes g, g + gsplit, ...
                           // and NPH / hsplit parity planes each; > 1 -> partial sums meet in gx by red.global.add
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
// Bounded wait: a protocol bug traps (clean CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 8000000000ll) __trap();
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
// one lane by elect.sync: single-lane to the compiler, no ELECT / R2UR.BROADCAST waterfall around every MMA (tc_ptx.cuh)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major operand descriptor (cute::UMMA::SmemDescriptor): 128-B swizzle -> 8-row atoms 1024 B apart, layout code 2;
// 64-B swizzle -> 8-row atoms 512 B apart, layout code 4.
template <int KC>
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((KC == 64 ? 1024u : 512u) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(KC == 64 ? 2 : 4) << 61;
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
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void red_add_f4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }


struct Item {
    int b, ty, tx, nt, g, hq;
};
__device__ __forceinline__ Item decode_item(const Params& p, int it) {
    Item r;
    const int nsub = p.gsplit * p.hsplit;
    const int sub = it % nsub;
    it /= nsub;
    r.g = sub % p.gsplit;
    r.hq = sub / p.gsplit;
    const int ptiles = p.tiles_x * p.tiles_y * p.batch;
    r.nt = it / ptiles;
    int pt = it - r.nt * ptiles;
    r.tx = pt % p.tiles_x;
    pt /= p.tiles_x;
    r.ty = pt % p.tiles_y;
    r.b = pt / p.tiles_y;
    return r;
}

// Regions present among the SOURCE pixels of a tile: the 10x16 halo of the input-grid tile, every parity plane.
template <int NPH>
__device__ __forceinline__ uint32_t halo_class_mask(const Params& p, const Item& it, int lane) {
    if (!p.label) return 1u;
    constexpr int MUL = NPH == 4 ? 2 : 1;
    const int ho = p.h * MUL, wo = p.w * MUL;
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int hp = lane + 32 * i;
        const int gy = it.ty * TH - 1 + (hp >> 4), gx = it.tx * TW - 1 + (hp & 15);
        if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
            const uint8_t* lp = p.label + ((int64_t)it.b * ho + gy * MUL) * wo + gx * MUL;
            m |= 1u << min((int)lp[0], p.ncls - 1);
            if (NPH == 4) m |= (1u << min((int)lp[1], p.ncls - 1)) | (1u << min((int)lp[wo], p.ncls - 1)) | (1u << min((int)lp[wo + 1], p.ncls - 1));
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if (p.gsplit > 1) {                       // this work item's share of the region passes: every gsplit-th present region
        uint32_t mine = 0;
        int k = 0;
        for (uint32_t cm = m; cm; cm &= cm - 1, ++k)
            if (k % p.gsplit == it.g) mine |= cm & (0u - cm);
        m = mine;
    }
    return m;
}

// NTI = input channels (N of the MMA) per work item; KC = output channels per K chunk; NPH = parity planes.
template <int NTI, int KC, int NPH>
__global__ void __launch_bounds__(NUM_THREADS, 1) modconv3x3_dgrad_tc_kernel(const __grid_constant__ CUtensorMap wmap, Params p) {
    constexpr int N = NTI;
    constexpr int ROWB = KC * 2;
    constexpr int A_PLANE = A_ROWS * ROWB;
    constexpr int A_STAGE = 2 * A_PLANE;
    constexpr int B_SLOT = N * ROWB;
    constexpr int NACC = (2 * N <= 512) ? 2 : 1;
    constexpr int TMEM_COLS = (NACC * N <= 32) ? 32 : (NACC * N <= 64) ? 64 : (NACC * N <= 128) ? 128 : (NACC * N <= 256) ? 256 : 512;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int KSTEPS = KC / 16;
    constexpr int MUL = NPH == 4 ? 2 : 1;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_buf = smem;
    uint8_t* b_buf = a_buf + ((NSTAGE_A * A_STAGE + 1023) & ~1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)p.nslot_b * B_SLOT);
    const int A_FULL = 0, A_EMPTY = A_FULL + NSTAGE_A, ACC_FULL = A_EMPTY + NSTAGE_A, ACC_EMPTY = ACC_FULL + NACC,
              B_FULL = ACC_EMPTY + NACC, B_EMPTY = B_FULL + p.nslot_b, NBARS = B_EMPTY + p.nslot_b;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ho = p.h * MUL, wo = p.w * MUL;
    const int nchunks = p.cout / KC;                  // K chunks per parity plane

    if (threadIdx.x == 0) {
        // A stages and accumulators are released by all three MMA warps; a w_hi slot (even) is read by two of them
        // (x_hi w_hi, x_lo w_hi), a w_lo slot (odd) by one
        for (int i = 0; i < NSTAGE_A; ++i) mbar_init(smem_u32(&bars[A_FULL + i]), NUM_XFORM), mbar_init(smem_u32(&bars[A_EMPTY + i]), NUM_MMA_WARPS);
        for (int i = 0; i < NACC; ++i) mbar_init(smem_u32(&bars[ACC_FULL + i]), NUM_MMA_WARPS), mbar_init(smem_u32(&bars[ACC_EMPTY + i]), NUM_EPI);
        for (int i = 0; i < p.nslot_b; ++i) mbar_init(smem_u32(&bars[B_FULL + i]), 1), mbar_init(smem_u32(&bars[B_EMPTY + i]), (i & 1) ? 1 : 2);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 8) {                                 // every MMA accumulates (no ordered "first" MMA across three warps)
        const uint32_t lanes = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16);
#pragma unroll 1
        for (int c = 0; c < TMEM_COLS; c += 32) tmem_zero32(lanes + (uint32_t)c);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===================================================================== weight-plane producer (TMA)
        int slot = 0;
        uint32_t ph = 0;
        const int rows_lo = (NPH * 9) * p.cin;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const int npass = __popc(halo_class_mask<NPH>(p, item, lane));
            if (elect_one()) {
                const int q0 = item.hq * (NPH / p.hsplit), q1 = q0 + NPH / p.hsplit;
                for (int pass = 0; pass < npass; ++pass)
                    for (int q = q0; q < q1; ++q)
                        for (int kc = 0; kc < nchunks; ++kc)
                            for (int tap = 0; tap < 9; ++tap)
                                for (int hl = 0; hl < 2; ++hl) {
                                    mbar_wait(smem_u32(&bars[B_EMPTY + slot]), ph ^ 1);
                                    const uint32_t full = smem_u32(&bars[B_FULL + slot]);
                                    mbar_expect_tx(full, B_SLOT);
                                    tma_load_2d(smem_u32(b_buf + (size_t)slot * B_SLOT), &wmap, kc * KC,
                                                hl * rows_lo + (q * 9 + tap) * p.cin + item.nt * NTI, full);
                                    if (++slot == p.nslot_b) slot = 0, ph ^= 1;
                                }
            }
            __syncwarp();
        }
    } else if (warp <= NUM_MMA_WARPS) {
        // ===================================================================== MMA issuers: x_hi w_hi | x_lo w_hi | x_hi w_lo
        const int role = warp - 1;
        const bool lo_w = role == 2;
        int sa = 0, slot = 0, acc = 0;
        uint32_t pa = 0, pb = 0, pacc0 = 0, pacc1 = 0;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t a0 = smem_u32(a_buf) + (role == 1 ? A_PLANE : 0), b0 = smem_u32(b_buf), bars0 = smem_u32(bars);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const int npass = __shfl_sync(0xffffffffu, __popc(halo_class_mask<NPH>(p, item, lane)), 0);
            for (int pass = 0; pass < npass; ++pass) {
                mbar_wait(bars0 + 8 * (ACC_EMPTY + acc), (acc ? pacc1 : pacc0) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + (uint32_t)(acc * N);
#pragma unroll 1
                for (int kk = 0; kk < (NPH / p.hsplit) * nchunks; ++kk) {
                    mbar_wait(bars0 + 8 * (A_FULL + sa), pa);
                    tc_fence_after();
                    const uint32_t ap = a0 + sa * A_STAGE;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        const uint32_t row_off = (uint32_t)(dy * TWP + dx + 1) * ROWB;
                        const int sl = slot + (lo_w ? 1 : 0);          // (hi, lo) pairs never straddle the ring wrap
                        mbar_wait(bars0 + 8 * (B_FULL + sl), pb);
                        tc_fence_after();
                        const uint32_t bb = b0 + sl * B_SLOT;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                umma_bf16(d_tmem, smem_desc<KC>(ap + row_off + k * 32), smem_desc<KC>(bb + k * 32), IDESC, 1u);
                            umma_commit(bars0 + 8 * (B_EMPTY + sl));
                        }
                        slot += 2;
                        if (slot >= p.nslot_b) slot = 0, pb ^= 1;
                    }
                    if (elect_one()) umma_commit(bars0 + 8 * (A_EMPTY + sa));
                    if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                }
                if (elect_one()) umma_commit(bars0 + 8 * (ACC_FULL + acc));
                if (acc) pacc1 ^= 1; else pacc0 ^= 1;
                if (NACC == 2) acc ^= 1;
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ===================================================================== transform: masked, demod-scaled output gradient
        const int t = threadIdx.x - 32 * (1 + NUM_MMA_WARPS);
        constexpr int CPR = KC / 8;
        constexpr int PPI = 128 / CPR;
        constexpr int NSWEEP = 160 / PPI;               // 10 or 5
        const int c8 = t % CPR;
        const int pbase = t / CPR;
        int sa = 0;
        uint32_t pa = 0;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const uint32_t classes = halo_class_mask<NPH>(p, item, lane);
            const float* gyb = p.gy + (int64_t)item.b * ho * wo * p.cout;
            const float* yb = p.y ? p.y + (int64_t)item.b * ho * wo * p.cout : nullptr;
            const int y0 = item.ty * TH, x0 = item.tx * TW;
            for (uint32_t cm = classes; cm; cm &= cm - 1) {
                const int cls = __ffs(cm) - 1;
                const float* dmc = p.demod ? p.demod + ((int64_t)item.b * p.ncls + cls) * p.cout : nullptr;
                for (int q = item.hq * (NPH / p.hsplit); q < (item.hq + 1) * (NPH / p.hsplit); ++q) {
                    const int py = q >> 1, px = q & 1;
                    for (int kc = 0; kc < nchunks; ++kc) {
                        const int ch = kc * KC + 8 * c8;
                        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
                        const float4 d0 = dmc ? __ldg(reinterpret_cast<const float4*>(dmc + ch)) : one4;
                        const float4 d1 = dmc ? __ldg(reinterpret_cast<const float4*>(dmc + ch + 4)) : one4;
                        uint8_t* hi_plane = a_buf + sa * A_STAGE;
                        uint8_t* lo_plane = hi_plane + A_PLANE;
                        bool waited = false;
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {
                            constexpr int HS = (NSWEEP + 1) / 2;
                            float4 g0[HS], g1[HS];
#pragma unroll
                            for (int i = 0; i < HS; ++i) {
                                const int sw = half * HS + i;
                                const int hp = pbase + PPI * sw;
                                const int sy = y0 - 1 + (hp >> 4), sx = x0 - 1 + (hp & 15);
                                g0[i] = make_float4(0.f, 0.f, 0.f, 0.f), g1[i] = g0[i];
                                if (sw < NSWEEP && sy >= 0 && sy < p.h && sx >= 0 && sx < p.w) {
                                    const int oy = sy * MUL + py, ox = sx * MUL + px;
                                    const int cl = p.label ? min((int)p.label[((int64_t)item.b * ho + oy) * wo + ox], p.ncls - 1) : 0;
                                    if (cl == cls) {
                                        const int64_t off = ((int64_t)oy * wo + ox) * p.cout + ch;
                                        g0[i] = __ldg(reinterpret_cast<const float4*>(gyb + off));
                                        g1[i] = __ldg(reinterpret_cast<const float4*>(gyb + off + 4));
                                        if (p.act) {
                                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(yb + off));
                                            const float4 a1 = __ldg(reinterpret_cast<const float4*>(yb + off + 4));
                                            g0[i].x *= a0.x > 0.f ? SQRT2 : 0.2f * SQRT2, g0[i].y *= a0.y > 0.f ? SQRT2 : 0.2f * SQRT2;
                                            g0[i].z *= a0.z > 0.f ? SQRT2 : 0.2f * SQRT2, g0[i].w *= a0.w > 0.f ? SQRT2 : 0.2f * SQRT2;
                                            g1[i].x *= a1.x > 0.f ? SQRT2 : 0.2f * SQRT2, g1[i].y *= a1.y > 0.f ? SQRT2 : 0.2f * SQRT2;
                                            g1[i].z *= a1.z > 0.f ? SQRT2 : 0.2f * SQRT2, g1[i].w *= a1.w > 0.f ? SQRT2 : 0.2f * SQRT2;
                                        }
                                    }
                                }
                            }
                            if (!waited) {
                                mbar_wait(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1);
                                waited = true;
                            }
#pragma unroll
                            for (int i = 0; i < HS; ++i) {
                                const int sw = half * HS + i;
                                if (sw >= NSWEEP) continue;
                                const int row = pbase + PPI * sw + 1;
                                float f[8] = {g0[i].x * d0.x, g0[i].y * d0.y, g0[i].z * d0.z, g0[i].w * d0.w,
                                              g1[i].x * d1.x, g1[i].y * d1.y, g1[i].z * d1.z, g1[i].w * d1.w};
                                uint32_t hi[4], lo[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float h0 = bf16_round(f[2 * j]), h1 = bf16_round(f[2 * j + 1]);
                                    hi[j] = pack_bf16x2(h0, h1);
                                    lo[j] = pack_bf16x2(f[2 * j] - h0, f[2 * j + 1] - h1);
                                }
                                const uint32_t sxz = KC == 64 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                                const uint32_t off = (uint32_t)row * ROWB + (((uint32_t)c8 ^ sxz) << 4);
                                *reinterpret_cast<uint4*>(hi_plane + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(lo_plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                        fence_proxy_async();
                        mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                        if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                    }
                }
            }
        }
    } else {
        // ===================================================================== epilogue: gx (+)= s_c * acc ; gs[c] += sum_q x * acc
        const uint32_t quarter = (uint32_t)(warp & 3);
        const int m_row = quarter * 32 + lane;
        const int ty = m_row >> 4, tx = m_row & 15;
        int acc = 0;
        uint32_t pacc[2] = {0, 0};
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const uint32_t classes = halo_class_mask<NPH>(p, item, lane);
            const int qy = item.ty * TH + ty, qx = item.tx * TW + tx;
            const bool mine = tx < TW && qy < p.h && qx < p.w;
            const int n0 = item.nt * NTI;
            const int64_t pix = ((int64_t)item.b * p.h + qy) * p.w + qx;
            const bool partial = p.gsplit * p.hsplit > 1;       // several work items add into the same gx rows (zeroed by the host side)
            bool first = true;
            for (uint32_t cm = classes; cm; cm &= cm - 1) {
                const int cls = __ffs(cm) - 1;
                mbar_wait(smem_u32(&bars[ACC_FULL + acc]), pacc[acc]);
                pacc[acc] ^= 1;
                tc_fence_after();
                const float* sc = p.s + ((int64_t)item.b * p.ncls + cls) * p.cin + n0;
#pragma unroll 1
                for (int j = 0; j < NTI / 32; ++j) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((quarter * 32u) << 16) + (uint32_t)(acc * N + j * 32), r);
                    tmem_zero32(tmem_base + ((quarter * 32u) << 16) + (uint32_t)(acc * N + j * 32));
                    if (p.gx && mine) {
                        float* dst = p.gx + pix * p.cin + n0 + j * 32;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const float4 sv = __ldg(reinterpret_cast<const float4*>(sc + j * 32 + 4 * g));
                            float4 o = make_float4(__uint_as_float(r[4 * g]) * sv.x, __uint_as_float(r[4 * g + 1]) * sv.y,
                                                   __uint_as_float(r[4 * g + 2]) * sv.z, __uint_as_float(r[4 * g + 3]) * sv.w);
                            if (partial) {
                                red_add_f4(dst + 4 * g, o);
                                continue;
                            }
                            if (!first) {
                                const float4 old = *reinterpret_cast<const float4*>(dst + 4 * g);
                                o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
                            }
                            *reinterpret_cast<float4*>(dst + 4 * g) = o;
                        }
                    }
                    if (p.gs) {
                        // column sums over this warp's 32 pixels of x[q,i] * acc[q,i], then one atomic per column
                        const float* xr = p.x + pix * p.cin + n0 + j * 32;
                        float keep = 0.f;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (mine) xv = *reinterpret_cast<const float4*>(xr + 4 * g);
                            // rows outside the image / spare tile columns hold garbage accumulators (possibly NaN): mask, do not multiply
                            float v[4] = {0.f, 0.f, 0.f, 0.f};
                            if (mine) {
                                v[0] = xv.x * __uint_as_float(r[4 * g]), v[1] = xv.y * __uint_as_float(r[4 * g + 1]);
                                v[2] = xv.z * __uint_as_float(r[4 * g + 2]), v[3] = xv.w * __uint_as_float(r[4 * g + 3]);
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float sum = v[e];
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                                if (lane == 4 * g + e) keep = sum;
                            }
                        }
                        if (keep != 0.f) atomicAdd(p.gs + ((int64_t)item.b * p.ncls + cls) * p.cin + n0 + j * 32 + lane, keep);
                    }
                }
                first = false;
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
                if (NACC == 2) acc ^= 1;
            }
        }
    }

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

static int num_sms() { return e4s_num_sms(); }

// One (pixel tile, N tile) pair is a serial chain of (regions in the tile) x NPH x Cout/KC x 9 weight slots with 4 MMAs
// each, whatever the N width - MMA issue, not bandwidth, sets its length (~150 cycles per MMA and issuing warp).  The
// 4x4 ... 32x32 layers of ONE face have 1-12 pixel tiles with every region in each: 2 CTAs ran 12 x 4 x 8 x 9 slots
// back to back (2.2 ms per layer, profiles/r1_bwd_layers_wide_ntile.log).  When the pairs cannot fill the SMs, cut the
// chain: region passes first (up to ncls ways), then parity planes; the partial sums meet in gx through vector
// red.global.add (gx zeroed first) and in gs through the atomics the kernel uses anyway.
// E4S_B200_DGRAD_SPLIT="G,H" forces a split (tests).
static void choose_split(int64_t pairs, int ncls, int nph, int& gsplit, int& hsplit) {
    gsplit = hsplit = 1;
    if (const char* f = getenv("E4S_B200_DGRAD_SPLIT")) {
        int g = 0, h = 0;
        if (sscanf(f, "%d,%d", &g, &h) == 2 && g >= 1 && (h == 1 || h == 2 || h == 4)) {
            gsplit = g < ncls ? g : ncls, hsplit = h < nph ? h : nph;
            return;
        }
    }
    const int sms = num_sms();
    if (pairs >= 2 * sms) return;
    // aim at ~4 work items per SM: the chains differ in length (regions per tile) and CTAs take items round-robin
    const int64_t g = e4s_ceil_div(4 * sms, pairs);
    gsplit = (int)(g < ncls ? g : ncls);
    while (hsplit < nph && pairs * gsplit * hsplit < sms) hsplit *= 2;
}

template <int NTI, int KC, int NPH>
static int launch(const void* wd_hilo, Params p, cudaStream_t st) {
    constexpr int ROWB = KC * 2;
    constexpr int A_BYTES = ((NSTAGE_A * 2 * A_ROWS * ROWB) + 1023) & ~1023;
    constexpr int B_SLOT = NTI * ROWB;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return E4S_ERR_ARCH;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)p.cout, (cuuint64_t)2 * NPH * 9 * p.cin};
    cuuint64_t strides[1] = {(cuuint64_t)p.cout * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)NTI};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wd_hilo), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return 700 + (int)cr;
    p.tiles_x = (int)e4s_ceil_div(p.w, TW);
    p.tiles_y = (int)e4s_ceil_div(p.h, TH);
    p.n_tiles = p.cin / NTI;
    const int64_t pairs = (int64_t)p.tiles_x * p.tiles_y * p.batch * p.n_tiles;
    choose_split(pairs, p.label ? p.ncls : 1, NPH, p.gsplit, p.hsplit);
    const int64_t items = pairs * p.gsplit * p.hsplit;
    if (items >= (1ll << 31)) return E4S_ERR_SHAPE;
    p.items = (int)items;
    if (p.gsplit * p.hsplit > 1 && p.gx &&
        cudaMemsetAsync(p.gx, 0, (size_t)p.batch * p.h * p.w * p.cin * sizeof(float), st) != cudaSuccess)
        return (int)cudaGetLastError();
    int max_slots = (SMEM_BUDGET - A_BYTES - 1024) / B_SLOT;
    if (max_slots < 2) return E4S_ERR_SHAPE;
    p.nslot_b = max_slots > 8 ? 8 : (max_slots & ~1);      // even: (hi, lo) slot pairs never straddle the ring wrap
    const size_t smem = 1024 + A_BYTES + (size_t)p.nslot_b * B_SLOT + (size_t)(2 * NSTAGE_A + 4 + 2 * p.nslot_b) * 8 + 64;
    static E4sSmemOptIn optin;
    if (const int rc = e4s_smem_optin(optin, modconv3x3_dgrad_tc_kernel<NTI, KC, NPH>, smem)) return rc;
    const int grid = p.items < num_sms() ? p.items : num_sms();
    modconv3x3_dgrad_tc_kernel<NTI, KC, NPH><<<grid, NUM_THREADS, smem, st>>>(map, p);
    return e4s_launch_status();
}

// Width of the N tile (input channels per work item).  Wide tiles re-use a staged gradient tile for more columns, but a
// work item is a whole (pixel tile, N tile) pair with up to ncls x NPH x 9 x Cout/KC weight slots streamed through ONE
// CTA: the 4x4 ... 32x32 layers of a single face have 1-12 pixel tiles, and at 256 columns two CTAs pulled the whole
// 38-MB weight set through their own L2 ports while 146 SMs idled (2.2 ms per layer, profiles/r1_bwd_layers_*.log).
// Take the widest tile that still yields work for half the SMs, else the narrowest.  E4S_B200_NTILE=32|64|128|256
// forces a width (tests).
static int pick_ntile(int channels, int widest, int64_t pixel_tiles, int max_split = 1) {
    static const int cand[4] = {256, 128, 64, 32};
    if (const char* f = getenv("E4S_B200_NTILE")) {
        const int v = atoi(f);
        if ((v == 32 || v == 64 || v == 128 || v == 256) && v <= widest && channels % v == 0) return v;
    }
    int last = 32;
    for (int i = 0; i < 4; ++i) {
        const int c = cand[i];
        if (c > widest || channels % c != 0) continue;
        last = c;
        if (pixel_tiles * (channels / c) * max_split >= num_sms() / 2) return c;
    }
    return last;
}

template <int KC, int NPH>
static int dispatch_n(const void* wd, const Params& p, cudaStream_t st) {
    const int64_t pixel_tiles = e4s_ceil_div(p.w, TW) * e4s_ceil_div(p.h, TH) * p.batch;
    switch (pick_ntile(p.cin, 256, pixel_tiles, (p.label ? p.ncls : 1) * NPH)) {
        case 256: return launch<256, KC, NPH>(wd, p, st);
        case 128: return launch<128, KC, NPH>(wd, p, st);
        case 64: return launch<64, KC, NPH>(wd, p, st);
        default: return launch<32, KC, NPH>(wd, p, st);
    }
}

}  // namespace tcd

extern "C" int e4s_modconv3x3_bwd_tc(const float* gy, const float* y, const float* x, const void* wd_hilo_bf16, const float* s,
                                     const float* demod, const uint8_t* label, float* gx, float* gs, int batch, int h, int w,
                                     int cin, int cout, int ncls, int up, int act, void* stream) {
    E4S_REQUIRE(gy && wd_hilo_bf16 && s && (gx || gs), E4S_ERR_ARG);
    E4S_REQUIRE(!act || y, E4S_ERR_ARG);
    E4S_REQUIRE(!gs || x, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && ncls > 0 && ncls <= 32, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0 && (cout % 32) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(label || ncls == 1, E4S_ERR_ARG);
    tcd::Params p{gy, y, x, s, demod, label, gx, gs, batch, h, w, cin, cout, ncls, act, 0, 0, 0, 0, 0, 1, 1};
    cudaStream_t st = (cudaStream_t)stream;
    const bool k64 = (cout % 64) == 0;
    if (!up) return k64 ? tcd::dispatch_n<64, 1>(wd_hilo_bf16, p, st) : tcd::dispatch_n<32, 1>(wd_hilo_bf16, p, st);
    return k64 ? tcd::dispatch_n<64, 4>(wd_hilo_bf16, p, st) : tcd::dispatch_n<32, 4>(wd_hilo_bf16, p, st);
}

// Host-only: the work list e4s_modconv3x3_bwd_tc would build for this shape (N-tile width, region-pass and parity-plane
// split).  ncls = number of regions the label map can hold (1 without a label map).  No launch, no device access beyond
// the SM count (148 when no device is present) - lets the host-side heuristics be tested without a GPU.
extern "C" int e4s_modconv3x3_bwd_tc_plan(int batch, int h, int w, int cin, int ncls, int up, int* ntile, int* gsplit, int* hsplit) {
    E4S_REQUIRE(ntile && gsplit && hsplit && batch > 0 && h > 0 && w > 0 && cin > 0 && ncls > 0 && ncls <= 32, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0, E4S_ERR_SHAPE);
    const int nph = up ? 4 : 1;
    const int64_t tiles = e4s_ceil_div(w, tcd::TW) * e4s_ceil_div(h, tcd::TH) * batch;
    *ntile = tcd::pick_ntile(cin, 256, tiles, ncls * nph);
    tcd::choose_split(tiles * (cin / *ntile), ncls, nph, *gsplit, *hsplit);
    return E4S_OK;
}

