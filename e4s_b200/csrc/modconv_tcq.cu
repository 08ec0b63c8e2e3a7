// Region-selected modulated 3x3 convolution on tcgen05 tensor cores - third generation (sm_100a).
//
// Contract, implicit-GEMM formulation and pipeline roles as modconv_tcp.cu (read modconv_tc.cu's header for the
// formulation: 8x16 pixel tile with 14 valid columns, taps as row-shifted descriptors over one staged halo tile,
// split-bf16 x3 accumulation in fp32 TMEM).  Two measured weaknesses of the second generation are removed
// (profiles/r1_opbench_tcp_v2*.json, profiles/r1_ncu_tcp_v2_summary.md):
//
//  1. Memory-level parallelism.  The transform warps loaded the activation halo tile straight from global memory,
//     ~20 KB in flight per SM, and every tile paid two exposed DRAM latencies: the 32->32 layer at 1024^2 ran at
//     10 % of DRAM bandwidth.  Now a dedicated producer warp streams raw fp32 halo tiles with 4-D TMA loads
//     (hardware zero fill outside the image) through a 3-stage ring, several tiles ahead of the transform warps,
//     which convert shared -> shared.
//  2. Region passes.  A tile whose pixels belong to k regions ran the whole main loop k times (2-5x on the 32^2-256^2
//     masked layers with a face mask).  Now a mixed tile runs ONE main-loop pass in "row-class" mode: the transform
//     warps materialise each tap's operand rows separately (im2col per tap, 9x the transform work, which hides under
//     the MMAs) and scale every row by the style of that row's own region.  Up-sampling layers, whose four output
//     parities can disagree on the region, add one classic pass per region that occurs only on a disagreeing parity.
//
// K chunks are 32 channels (64-byte-swizzle operand rows) for every layer, which is what lets the three rings
// (raw tiles 3 x 20 KB, operand planes 2 x 21 KB, weight planes up to 16 x N x 64 B) share 227 KB.
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>

#include "common.cuh"

namespace tcq {

constexpr int TH = 8, TWP = 16, TW = 14;
constexpr int KC = 32, ROWB = 64;
constexpr int A_ROWS = 168;
constexpr int A_PLANE = A_ROWS * ROWB;        // 10752
constexpr int A_SLOT = 2 * A_PLANE;           // hi + lo
constexpr int NSLOT_A = 2;
constexpr int XS_STAGE = 160 * KC * 4;        // raw fp32 halo tile of one chunk
constexpr int NXS = 3;
constexpr int NUM_THREADS = 352;              // 11 warps
constexpr int NUM_XFORM = 128, NUM_EPI = 128;
constexpr int SMEM_BUDGET = 227 * 1024 - 2048;

struct Params {
    const float* x;
    const float* s;
    const float* demod;
    const uint8_t* label;
    const float* noise;
    const float* noise_w;
    const float* bias;
    float* y;
    int batch, h, w, cin, cout, ncls, noise_b, act;       // act: 0 none, 1 sqrt(2)*lrelu(0.2), 2 PReLU(slope[c])
    int tiles_x, tiles_y, n_tiles, items, nslot_b, resident;
    const float* shift;
    const float* slope;
    int out_stride;
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major operand descriptor (cute::UMMA::SmemDescriptor): 128-B swizzle -> 8-row atoms 1024 B apart, layout code 2;
// 64-B swizzle -> 8-row atoms 512 B apart, layout code 4.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {      // K-major, 64-byte swizzle: 8-row atoms 512 B apart
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
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
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }


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

// Pass plan of a tile, derived by one whole warp from the label map (every role derives it independently).
//   all : regions present among valid (pixel, parity) outputs
//   fix : regions that occur on a parity whose region differs from the pixel's parity-0 region (NPH == 4 only)
// all has one bit  -> one classic pass (shift mode) for that region;
// otherwise         -> pass 0 in row-class mode (row r uses the region of its parity-0 pixel), then one classic pass
//                      per region in fix.
struct Plan {
    uint32_t all, fix;
};
template <int NPH>
__device__ __forceinline__ Plan tile_plan(const Params& p, const Item& it, int lane) {
    Plan pl{1u, 0u};
    if (!p.label) return pl;
    constexpr int MUL = NPH == 4 ? 2 : 1;
    const int ho = p.h * MUL, wo = p.w * MUL;
    uint32_t all = 0, fix = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = lane + 32 * i;
        const int ty = r >> 4, tx = r & 15;
        const int iy = it.ty * TH + ty, ix = it.tx * TW + tx;
        if (tx < TW && iy < p.h && ix < p.w) {
            const uint8_t* lp = p.label + ((int64_t)it.b * ho + iy * MUL) * wo + ix * MUL;
            const int rc = min((int)lp[0], p.ncls - 1);
            all |= 1u << rc;
            if (NPH == 4) {
                const int c1 = min((int)lp[1], p.ncls - 1), c2 = min((int)lp[wo], p.ncls - 1), c3 = min((int)lp[wo + 1], p.ncls - 1);
                all |= (1u << c1) | (1u << c2) | (1u << c3);
                if (c1 != rc) fix |= 1u << c1;
                if (c2 != rc) fix |= 1u << c2;
                if (c3 != rc) fix |= 1u << c3;
            }
        }
    }
    pl.all = __reduce_or_sync(0xffffffffu, all);
    pl.fix = __reduce_or_sync(0xffffffffu, fix);
    return pl;
}
__device__ __forceinline__ bool plan_mixed(const Plan& pl) { return (pl.all & (pl.all - 1)) != 0; }
__device__ __forceinline__ int plan_passes(const Plan& pl) { return plan_mixed(pl) ? 1 + __popc(pl.fix) : 1; }

// ---------------------------------------------------------------------------------------- kernel
template <int NTC, int NPH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
modconv3x3_tcq_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap xmap, Params p) {
    constexpr int N = NTC * NPH;
    constexpr int B_SLOT = N * ROWB;
    constexpr int NACC = (2 * N <= 512) ? 2 : 1;
    constexpr int TMEM_COLS = (NACC * N <= 32) ? 32 : (NACC * N <= 64) ? 64 : (NACC * N <= 128) ? 128 : (NACC * N <= 256) ? 256 : 512;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int MUL = NPH == 4 ? 2 : 1;
    static_assert(N <= 256 && N % 16 == 0, "UMMA N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xs_buf = smem;                                   // [NXS][160][32] fp32
    uint8_t* a_buf = xs_buf + NXS * XS_STAGE;                 // [NSLOT_A][hi|lo][A_ROWS][64 B]
    uint8_t* b_buf = a_buf + NSLOT_A * A_SLOT;                // [nslot_b][N][64 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)p.nslot_b * B_SLOT);
    const int XS_FULL = 0, XS_EMPTY = XS_FULL + NXS, A_FULL = XS_EMPTY + NXS, A_EMPTY = A_FULL + NSLOT_A,
              ACC_FULL = A_EMPTY + NSLOT_A, ACC_EMPTY = ACC_FULL + NACC, B_FULL = ACC_EMPTY + NACC, B_EMPTY = B_FULL + p.nslot_b,
              NBARS = B_EMPTY + p.nslot_b;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ho = p.h * MUL, wo = p.w * MUL;
    const int nchunks = p.cin / KC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NXS; ++i) mbar_init(smem_u32(&bars[XS_FULL + i]), 1), mbar_init(smem_u32(&bars[XS_EMPTY + i]), NUM_XFORM);
        for (int i = 0; i < NSLOT_A; ++i) mbar_init(smem_u32(&bars[A_FULL + i]), NUM_XFORM), mbar_init(smem_u32(&bars[A_EMPTY + i]), 1);
        for (int i = 0; i < NACC; ++i) mbar_init(smem_u32(&bars[ACC_FULL + i]), 1), mbar_init(smem_u32(&bars[ACC_EMPTY + i]), NUM_EPI);
        for (int i = 0; i < p.nslot_b; ++i) mbar_init(smem_u32(&bars[B_FULL + i]), 1), mbar_init(smem_u32(&bars[B_EMPTY + i]), 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    if (warp == 10 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== weight-plane producer (TMA)
        int slot = 0;
        uint32_t ph = 0;
        bool loaded_resident = false;
        const int rows_lo = (NPH * 9) * p.cout;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const Plan pl = tile_plan<NPH>(p, item, lane);
            if (p.resident && loaded_resident) continue;
            const int npass = p.resident ? 1 : plan_passes(pl);
            if (lane == 0) {
                for (int pass = 0; pass < npass; ++pass)
                    for (int kc = 0; kc < nchunks; ++kc)
                        for (int tap = 0; tap < 9; ++tap)
                            for (int hl = 0; hl < 2; ++hl) {
                                if (!p.resident) mbar_wait(smem_u32(&bars[B_EMPTY + slot]), ph ^ 1);
                                const uint32_t full = smem_u32(&bars[B_FULL + slot]);
                                mbar_expect_tx(full, B_SLOT);
                                const uint32_t dst = smem_u32(b_buf + (size_t)slot * B_SLOT);
#pragma unroll
                                for (int q = 0; q < NPH; ++q)
                                    tma_load_2d(dst + q * NTC * ROWB, &wmap, kc * KC, hl * rows_lo + (q * 9 + tap) * p.cout + item.nt * NTC, full);
                                if (++slot == p.nslot_b) slot = 0, ph ^= 1;
                            }
            }
            loaded_resident = true;
            __syncwarp();
        }
    } else if (warp == 10) {
        // ===================================================================== raw activation tile producer (4-D TMA)
        int st = 0;
        uint32_t ph = 0;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const Plan pl = tile_plan<NPH>(p, item, lane);
            const int npass = plan_passes(pl);
            if (lane == 0) {
                for (int pass = 0; pass < npass; ++pass)
                    for (int kc = 0; kc < nchunks; ++kc) {
                        mbar_wait(smem_u32(&bars[XS_EMPTY + st]), ph ^ 1);
                        const uint32_t full = smem_u32(&bars[XS_FULL + st]);
                        mbar_expect_tx(full, XS_STAGE);
                        tma_load_4d(smem_u32(xs_buf + st * XS_STAGE), &xmap, kc * KC, item.tx * TW - 1, item.ty * TH - 1, item.b, full);
                        if (++st == NXS) st = 0, ph ^= 1;
                    }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        int sa = 0, slot = 0, acc = 0;
        uint32_t pa = 0, pb = 0, pacc[2] = {0, 0};
        bool first_resident_pass = true;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const Plan pl = tile_plan<NPH>(p, item, lane);
            const int npass = plan_passes(pl);
            const bool mixed = plan_mixed(pl);
            if (lane == 0) {
                for (int pass = 0; pass < npass; ++pass) {
                    const bool rowclass = mixed && pass == 0;
                    mbar_wait(smem_u32(&bars[ACC_EMPTY + acc]), pacc[acc] ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                    uint32_t accumulate = 0;
                    if (p.resident) slot = 0;
                    for (int kc = 0; kc < nchunks; ++kc) {
                        if (!rowclass) {
                            mbar_wait(smem_u32(&bars[A_FULL + sa]), pa);
                            tc_fence_after();
                        }
                        for (int tap = 0; tap < 9; ++tap) {
                            uint32_t row_off;
                            if (rowclass) {              // this tap's rows were materialised on their own: wait for them
                                mbar_wait(smem_u32(&bars[A_FULL + sa]), pa);
                                tc_fence_after();
                                row_off = 0;
                            } else {
                                const int dy = tap / 3, dx = tap - 3 * dy;
                                row_off = (uint32_t)(dy * TWP + dx + 1) * ROWB;
                            }
                            const uint32_t a_hi = smem_u32(a_buf + sa * A_SLOT) + row_off, a_lo = a_hi + A_PLANE;
                            if (!p.resident || first_resident_pass) mbar_wait(smem_u32(&bars[B_FULL + slot]), pb);
                            tc_fence_after();
                            uint32_t bb = smem_u32(b_buf + (size_t)slot * B_SLOT);
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint64_t db = smem_desc(bb + k * 32);
                                umma_bf16(d_tmem, smem_desc(a_hi + k * 32), db, IDESC, accumulate);
                                umma_bf16(d_tmem, smem_desc(a_lo + k * 32), db, IDESC, 1u);
                                accumulate = 1u;
                            }
                            if (!p.resident) umma_commit(smem_u32(&bars[B_EMPTY + slot]));
                            if (++slot == p.nslot_b) slot = 0, pb ^= 1;
                            if (!p.resident || first_resident_pass) mbar_wait(smem_u32(&bars[B_FULL + slot]), pb);
                            tc_fence_after();
                            bb = smem_u32(b_buf + (size_t)slot * B_SLOT);
#pragma unroll
                            for (int k = 0; k < 2; ++k) umma_bf16(d_tmem, smem_desc(a_hi + k * 32), smem_desc(bb + k * 32), IDESC, 1u);
                            if (!p.resident) umma_commit(smem_u32(&bars[B_EMPTY + slot]));
                            if (++slot == p.nslot_b) slot = 0, pb ^= 1;
                            if (rowclass) {
                                umma_commit(smem_u32(&bars[A_EMPTY + sa]));
                                if (++sa == NSLOT_A) sa = 0, pa ^= 1;
                            }
                        }
                        if (!rowclass) {
                            umma_commit(smem_u32(&bars[A_EMPTY + sa]));
                            if (++sa == NSLOT_A) sa = 0, pa ^= 1;
                        }
                    }
                    umma_commit(smem_u32(&bars[ACC_FULL + acc]));
                    pacc[acc] ^= 1;
                    if (NACC == 2) acc ^= 1;
                    first_resident_pass = false;
                }
            }
            __syncwarp();
        }
    } else if (warp < 6) {
        // ===================================================================== activation transform (raw tile -> operand planes)
        const int t = threadIdx.x - 64;                  // 0..127
        const int c8 = t & 3;                            // 16-byte chunk (8 channels) of the 64-byte operand row
        const int pb4 = t >> 2;                          // 0..31
        int sa = 0, sx = 0;
        uint32_t pa = 0, px = 0;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const Plan pl = tile_plan<NPH>(p, item, lane);
            const int npass = plan_passes(pl);
            const bool mixed = plan_mixed(pl);
            const int y0 = item.ty * TH, x0 = item.tx * TW;
            // parity-0 region of this thread's four im2col rows (row = pb4 + 32*i)
            int rcls[4] = {0, 0, 0, 0};
            if (mixed) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = pb4 + 32 * i;
                    const int iy = y0 + (r >> 4), ix = x0 + (r & 15);
                    if ((r & 15) < TW && iy < p.h && ix < p.w)
                        rcls[i] = min((int)p.label[((int64_t)item.b * ho + iy * MUL) * wo + ix * MUL], p.ncls - 1);
                }
            }
            uint32_t rem = pl.fix;
            for (int pass = 0; pass < npass; ++pass) {
                const bool rowclass = mixed && pass == 0;
                int cls = 0;
                if (!mixed) cls = __ffs(pl.all) - 1;
                else if (pass > 0) cls = __ffs(rem) - 1, rem &= rem - 1;
                const float* sbase = p.s ? p.s + (int64_t)item.b * p.ncls * p.cin : nullptr;
                const float* shbase = p.shift ? p.shift + (int64_t)item.b * p.ncls * p.cin : nullptr;
                for (int kc = 0; kc < nchunks; ++kc) {
                    const int ch = kc * KC + 8 * c8;
                    mbar_wait(smem_u32(&bars[XS_FULL + sx]), px);
                    const float* xs = reinterpret_cast<const float*>(xs_buf + sx * XS_STAGE);
                    if (!rowclass) {
                        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 s0 = sbase ? __ldg(reinterpret_cast<const float4*>(sbase + (int64_t)cls * p.cin + ch)) : one4;
                        const float4 s1 = sbase ? __ldg(reinterpret_cast<const float4*>(sbase + (int64_t)cls * p.cin + ch + 4)) : one4;
                        const float4 t0 = shbase ? __ldg(reinterpret_cast<const float4*>(shbase + (int64_t)cls * p.cin + ch)) : zero4;
                        const float4 t1 = shbase ? __ldg(reinterpret_cast<const float4*>(shbase + (int64_t)cls * p.cin + ch + 4)) : zero4;
                        mbar_wait(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1);
                        uint8_t* hi_plane = a_buf + sa * A_SLOT;
                        uint8_t* lo_plane = hi_plane + A_PLANE;
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const int hp = pb4 + 32 * i;                         // halo pixel 0..159
                            const float4 v0 = *reinterpret_cast<const float4*>(xs + hp * KC + 8 * c8);
                            const float4 v1 = *reinterpret_cast<const float4*>(xs + hp * KC + 8 * c8 + 4);
                            float f[8] = {v0.x * s0.x, v0.y * s0.y, v0.z * s0.z, v0.w * s0.w, v1.x * s1.x, v1.y * s1.y, v1.z * s1.z, v1.w * s1.w};
                            if (shbase) {       // zero padding applies to the NORMALISED tensor: shift in-image pixels only
                                const int gy = y0 - 1 + (hp >> 4), gx = x0 - 1 + (hp & 15);
                                if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
                                    f[0] += t0.x, f[1] += t0.y, f[2] += t0.z, f[3] += t0.w;
                                    f[4] += t1.x, f[5] += t1.y, f[6] += t1.z, f[7] += t1.w;
                                }
                            }
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float h0 = bf16_round(f[2 * j]), h1 = bf16_round(f[2 * j + 1]);
                                hi[j] = pack_bf16x2(h0, h1);
                                lo[j] = pack_bf16x2(f[2 * j] - h0, f[2 * j + 1] - h1);
                            }
                            const int row = hp + 1;
                            const uint32_t off = (uint32_t)row * ROWB + (((uint32_t)c8 ^ (uint32_t)((row >> 1) & 3)) << 4);
                            *reinterpret_cast<uint4*>(hi_plane + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(lo_plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                        fence_proxy_async();
                        mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                        if (++sa == NSLOT_A) sa = 0, pa ^= 1;
                    } else {
                        // row-class mode: one operand slot per tap, row r scaled by the style of its own region
                        float4 sr0[4], sr1[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            sr0[i] = __ldg(reinterpret_cast<const float4*>(sbase + (int64_t)rcls[i] * p.cin + ch));
                            sr1[i] = __ldg(reinterpret_cast<const float4*>(sbase + (int64_t)rcls[i] * p.cin + ch + 4));
                        }
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const int dy = tap / 3, dx = tap - 3 * dy;
                            mbar_wait(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1);
                            uint8_t* hi_plane = a_buf + sa * A_SLOT;
                            uint8_t* lo_plane = hi_plane + A_PLANE;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = pb4 + 32 * i;                      // operand row = output pixel of the tile
                                const int hp = min(((r >> 4) + dy) * 16 + (r & 15) + dx, 159);
                                const float4 v0 = *reinterpret_cast<const float4*>(xs + hp * KC + 8 * c8);
                                const float4 v1 = *reinterpret_cast<const float4*>(xs + hp * KC + 8 * c8 + 4);
                                float f[8] = {v0.x * sr0[i].x, v0.y * sr0[i].y, v0.z * sr0[i].z, v0.w * sr0[i].w,
                                              v1.x * sr1[i].x, v1.y * sr1[i].y, v1.z * sr1[i].z, v1.w * sr1[i].w};
                                uint32_t hi[4], lo[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float h0 = bf16_round(f[2 * j]), h1 = bf16_round(f[2 * j + 1]);
                                    hi[j] = pack_bf16x2(h0, h1);
                                    lo[j] = pack_bf16x2(f[2 * j] - h0, f[2 * j + 1] - h1);
                                }
                                const uint32_t off = (uint32_t)r * ROWB + (((uint32_t)c8 ^ (uint32_t)((r >> 1) & 3)) << 4);
                                *reinterpret_cast<uint4*>(hi_plane + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<uint4*>(lo_plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            }
                            fence_proxy_async();
                            mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                            if (++sa == NSLOT_A) sa = 0, pa ^= 1;
                        }
                    }
                    mbar_arrive(smem_u32(&bars[XS_EMPTY + sx]));          // raw tile fully consumed by this thread
                    if (++sx == NXS) sx = 0, px ^= 1;
                }
            }
        }
    } else if (warp < 10) {
        // ===================================================================== epilogue
        const uint32_t quarter = (uint32_t)(warp & 3);
        const int m_row = quarter * 32 + lane;
        const int ty = m_row >> 4, tx = m_row & 15;
        int acc = 0;
        uint32_t pacc[2] = {0, 0};
        const float nw = (p.noise && p.noise_w) ? __ldg(p.noise_w) : 0.f;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = decode_item(p, it);
            const Plan pl = tile_plan<NPH>(p, item, lane);
            const int npass = plan_passes(pl);
            const bool mixed = plan_mixed(pl);
            const int iy = item.ty * TH + ty, ix = item.tx * TW + tx;
            const bool strided = (NPH == 1 && p.out_stride == 2);
            const bool in_img = tx < TW && iy < p.h && ix < p.w && (!strided || ((iy | ix) & 1) == 0);
            const int n0 = item.nt * NTC;
            int pcls[NPH];
#pragma unroll
            for (int q = 0; q < NPH; ++q) {
                pcls[q] = -1;
                if (in_img) {
                    const int oy = iy * MUL + (q >> 1), ox = ix * MUL + (q & 1);
                    pcls[q] = p.label ? min((int)p.label[((int64_t)item.b * ho + oy) * wo + ox], p.ncls - 1) : 0;
                }
            }
            const int rc = pcls[0];
            uint32_t rem = pl.fix;
            for (int pass = 0; pass < npass; ++pass) {
                const bool rowclass = mixed && pass == 0;
                int cls = 0;
                if (!mixed) cls = __ffs(pl.all) - 1;
                else if (pass > 0) cls = __ffs(rem) - 1, rem &= rem - 1;
                mbar_wait(smem_u32(&bars[ACC_FULL + acc]), pacc[acc]);
                pacc[acc] ^= 1;
                tc_fence_after();
                const int dcls = rowclass ? max(rc, 0) : cls;            // region whose demodulation applies to my rows
                const float* dm = p.demod ? p.demod + ((int64_t)item.b * p.ncls + dcls) * p.cout + n0 : nullptr;
#pragma unroll
                for (int q = 0; q < NPH; ++q) {
                    bool mine;
                    if (rowclass) mine = in_img && pcls[q] == rc;
                    else if (mixed) mine = in_img && pcls[q] == cls && pcls[q] != rc;
                    else mine = in_img;
                    if (!__any_sync(0xffffffffu, mine)) continue;
                    const int oy = strided ? (iy >> 1) : iy * MUL + (q >> 1), ox = strided ? (ix >> 1) : ix * MUL + (q & 1);
                    const int oh = strided ? (p.h >> 1) : ho, ow = strided ? (p.w >> 1) : wo;
                    float nz = 0.f;
                    if (mine && p.noise) nz = nw * __ldg(p.noise + ((int64_t)(p.noise_b == 1 ? 0 : item.b) * oh + oy) * ow + ox);
                    float* dst = p.y + (((int64_t)item.b * oh + oy) * ow + ox) * p.cout + n0;
#pragma unroll 1
                    for (int j = 0; j < NTC / 32; ++j) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + ((quarter * 32u) << 16) + (uint32_t)(acc * N + q * NTC + j * 32), r);
                        if (mine) {
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                const int co = j * 32 + 4 * g;
                                const float4 d = dm ? __ldg(reinterpret_cast<const float4*>(dm + co)) : make_float4(1.f, 1.f, 1.f, 1.f);
                                const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                float4 o;
                                o.x = __uint_as_float(r[4 * g + 0]) * d.x + nz + bv.x;
                                o.y = __uint_as_float(r[4 * g + 1]) * d.y + nz + bv.y;
                                o.z = __uint_as_float(r[4 * g + 2]) * d.z + nz + bv.z;
                                o.w = __uint_as_float(r[4 * g + 3]) * d.w + nz + bv.w;
                                if (p.act == 1) {
                                    const float k = 1.41421356237309515f;
                                    o.x = lrelu_scaled(o.x, 0.2f, k), o.y = lrelu_scaled(o.y, 0.2f, k);
                                    o.z = lrelu_scaled(o.z, 0.2f, k), o.w = lrelu_scaled(o.w, 0.2f, k);
                                } else if (p.act == 2) {
                                    const float4 sl = __ldg(reinterpret_cast<const float4*>(p.slope + n0 + co));
                                    o.x = o.x > 0.f ? o.x : o.x * sl.x, o.y = o.y > 0.f ? o.y : o.y * sl.y;
                                    o.z = o.z > 0.f ? o.z : o.z * sl.z, o.w = o.w > 0.f ? o.w : o.w * sl.w;
                                }
                                *reinterpret_cast<float4*>(dst + co) = o;
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
                if (NACC == 2) acc ^= 1;
            }
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

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = E4S_NUM_SMS;
    }
    return n;
}

template <int NTC, int NPH>
static int launch(const void* w_hilo, Params p, cudaStream_t st) {
    constexpr int N = NTC * NPH;
    constexpr int B_SLOT = N * ROWB;
    constexpr int FIXED = NXS * XS_STAGE + NSLOT_A * A_SLOT;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return E4S_ERR_ARCH;
    CUtensorMap wmap, xmap;
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.cin, (cuuint64_t)2 * NPH * 9 * p.cout};
        cuuint64_t strides[1] = {(cuuint64_t)p.cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)NTC};
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_hilo), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return 700 + (int)cr;
    }
    {
        // activation [B, H, W, C] fp32 as a 4-D tensor (C fastest); box = 32 channels x 16 columns x 10 rows of one sample
        cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.batch};
        cuuint64_t strides[3] = {(cuuint64_t)p.cin * 4, (cuuint64_t)p.w * p.cin * 4, (cuuint64_t)p.h * p.w * p.cin * 4};
        cuuint32_t box[4] = {(cuuint32_t)KC, 16, 10, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return 800 + (int)cr;
    }
    p.tiles_x = (int)e4s_ceil_div(p.w, TW);
    p.tiles_y = (int)e4s_ceil_div(p.h, TH);
    p.n_tiles = p.cout / NTC;
    const int64_t items = (int64_t)p.tiles_x * p.tiles_y * p.batch * p.n_tiles;
    if (items >= (1ll << 31)) return E4S_ERR_SHAPE;
    p.items = (int)items;
    const int planes = (p.cin / KC) * 18;
    int max_slots = (SMEM_BUDGET - FIXED - 1024) / B_SLOT;
    if (max_slots > 40) max_slots = 40;
    if (max_slots < 2) return E4S_ERR_SHAPE;
    p.resident = (p.n_tiles == 1 && planes <= max_slots) ? 1 : 0;
    p.nslot_b = p.resident ? planes : (max_slots > 16 ? 16 : max_slots);
    const size_t smem = 1024 + FIXED + (size_t)p.nslot_b * B_SLOT + (size_t)(2 * NXS + 2 * NSLOT_A + 4 + 2 * p.nslot_b) * 8 + 64;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(modconv3x3_tcq_kernel<NTC, NPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)cudaGetLastError();
        smem_set = smem;
    }
    const int grid = p.items < num_sms() ? p.items : num_sms();
    modconv3x3_tcq_kernel<NTC, NPH><<<grid, NUM_THREADS, smem, st>>>(wmap, xmap, p);
    return e4s_launch_status();
}

int dispatch(const void* w_hilo_bf16, Params p, int up, cudaStream_t st) {
    const int cout = p.cout;
    if (!up) {
        if (cout % 256 == 0) return launch<256, 1>(w_hilo_bf16, p, st);
        if (cout % 128 == 0) return launch<128, 1>(w_hilo_bf16, p, st);
        if (cout % 64 == 0) return launch<64, 1>(w_hilo_bf16, p, st);
        return launch<32, 1>(w_hilo_bf16, p, st);
    }
    if (cout % 64 == 0) return launch<64, 4>(w_hilo_bf16, p, st);
    return launch<32, 4>(w_hilo_bf16, p, st);
}

}  // namespace tcq

extern "C" int e4s_modconv3x3_tcq_fwd(const float* x, const void* w_hilo_bf16, const float* s, const float* demod,
                                      const uint8_t* label, const float* noise, const float* noise_w, const float* bias,
                                      float* y, int batch, int h, int w, int cin, int cout, int ncls, int up, int noise_b,
                                      int act, void* stream) {
    E4S_REQUIRE(x && w_hilo_bf16 && s && y, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && ncls > 0 && ncls <= 32, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0 && (cout % 32) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(label || ncls == 1, E4S_ERR_ARG);
    E4S_REQUIRE(!noise || (noise_w && (noise_b == 1 || noise_b == batch)), E4S_ERR_ARG);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w_hilo_bf16) && e4s_aligned16(s) && e4s_aligned16(y) &&
                    (!demod || e4s_aligned16(demod)) && (!bias || e4s_aligned16(bias)),
                E4S_ERR_ALIGN);
    tcq::Params p{x, s, demod, label, noise, noise_w, bias, y, batch, h, w, cin, cout, ncls, noise_b, act ? 1 : 0,
                  0, 0, 0, 0, 0, 0, nullptr, nullptr, 1};
    return tcq::dispatch(w_hilo_bf16, p, up, (cudaStream_t)stream);
}

extern "C" int e4s_conv3x3_tcq_f32(const float* x, const void* w_hilo_bf16, const float* scale, const float* shift,
                                   const float* prelu_slope, float* y, int batch, int h, int w, int cin, int cout,
                                   int out_stride, void* stream) {
    E4S_REQUIRE(x && w_hilo_bf16 && y, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0 && (cout % 32) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(out_stride == 1 || (out_stride == 2 && (h % 2) == 0 && (w % 2) == 0), E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w_hilo_bf16) && e4s_aligned16(y) && (!scale || e4s_aligned16(scale)) &&
                    (!shift || e4s_aligned16(shift)) && (!prelu_slope || e4s_aligned16(prelu_slope)),
                E4S_ERR_ALIGN);
    tcq::Params p{x, scale, nullptr, nullptr, nullptr, nullptr, nullptr, y, batch, h, w, cin, cout, 1, 1, prelu_slope ? 2 : 0,
                  0, 0, 0, 0, 0, 0, shift, prelu_slope, out_stride};
    return tcq::dispatch(w_hilo_bf16, p, 0, (cudaStream_t)stream);
}
