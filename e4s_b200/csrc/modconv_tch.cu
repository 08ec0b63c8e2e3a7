// Up-sampling StyledConv in the H-FORM on tcgen05 tensor cores (sm_100a): conv_transpose2d(stride 2) + 4x4 blur
// (reference src/models/stylegan2/model.py:287-300) + region selection + noise / bias / activation in one kernel, at HALF
// the multiply-accumulates of the polyphase form of modconv_tcr.cu.
//
// Formulation (executable specification: tools/ubench/hform_dataflow.py, checked against conv_transpose2d + upfirdn2d).
// The blur is separable.  Its VERTICAL half is folded into the weights on the host, its HORIZONTAL half runs in the
// epilogue:
//
//     T[py, kx][m, n'] = sum_dy V[py, kx][dy] . (s_c * x)[m + dy - 1, n']        V[py, kx][dy] = sum_ky Ay[py][dy, ky] W[ky, kx]
//     out[2m + py, 2n + px] = sum_{dx, kx} Ax[px][dx, kx] T[py, kx][m, n + dx - 1]
//
// i.e. an implicit GEMM with M = 128 input pixels (the 8x16 patch of modconv_tcr.cu: patch column 0 is image column
// x0 - 1), N = 6 x NTC accumulator columns (the six (py, kx) groups of NTC output channels) and K = 3 row taps x Cin -
// 18 tap-MACs per input pixel against 36 in the polyphase form - followed by a six-term combination per output parity of
// the pixel's own accumulators with those of its left / right neighbours (lane -1 / +1 of the same warp: two warp
// shuffles per term that crosses a pixel).  Row taps are pure row shifts of one staged halo tile by 16 operand rows, the
// same trick (and the same operand staging code) as the 3x3 kernel's taps with dx = 0.
//
// Regions: the style is that of the OUTPUT pixel's region (the reference masks after the blur, model.py:389-397), and an
// accumulator row feeds output pixels of three columns, so rows cannot be scaled by "their" region.  A tile is processed
// in PASSES of up to two regions, one accumulator buffer per region (x * s_A and x * s_B staged side by side, the same
// weight slots multiplied into both); the epilogue combines within one region's buffer and each output pixel takes the
// result of its own region.  One region: one pass, double-buffered across tiles.  k regions: ceil(k / 2) passes.
//
// Roles (persistent CTAs, 20 warps: 1 TMA, 3 MMA issue, 8 transform, 8 epilogue), barriers, split-bf16 x3 accumulation, zeroing of read accumulators by the epilogue:
// as in modconv_tcr.cu.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tch {
using namespace tcx;

constexpr int TH = 8, TW = 14;
constexpr int A_ROWS = 168;                     // 160 halo pixels + the leading row of the 3x3 kernel's layout
constexpr int NSTAGE_A = 2;
constexpr int NUM_MMA_WARPS = 3;
constexpr int W_XFORM0 = 1 + NUM_MMA_WARPS, W_EPI0 = W_XFORM0 + 8, W_END = W_EPI0 + 8;
constexpr int NUM_THREADS = 32 * W_END;         // 640: registers are capped at 96 per thread
constexpr int NUM_XFORM = 256, NUM_EPI = 256;
constexpr int SMEM_BUDGET = 227 * 1024 - 2048;
constexpr int NG = 6;                           // (py, kx) accumulator column groups
constexpr int ACC_COLS = 256, NACC = 2, TMEM_COLS = 512;

struct Params {
    const float* x;
    const float* s;
    const float* demod;
    const uint8_t* label;
    const float* noise;
    const float* noise_w;
    const float* bias;
    float* y;
    float f0, f1, f2, f3;                       // flipped horizontal FIR taps
    int batch, h, w, cin, cout, ncls, noise_b, act;
    int tiles_x, tiles_y, n_tiles, items, nslot_b;
    int det;              // 1: one warp issues the three split-precision products in a fixed order (bit-reproducible)
    long long* prof;      // optional [4 roles][4] cycle counters of CTA 0 (e4s_tch_set_profile; diagnostic build only)
};

// Stall attribution (tools/opbench.py --prof): only in the diagnostic build (-DE4S_TC_PROFILE, libe4s_b200_prof.so).
#ifdef E4S_TC_PROFILE
#define TCH_WAIT(bar, parity, k) do { if (prof_on) { const long long t_ = clock64(); mbar_wait(bar, parity); pw[k] += clock64() - t_; } else mbar_wait(bar, parity); } while (0)
#else
#define TCH_WAIT(bar, parity, k) mbar_wait(bar, parity)
#endif

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
struct Walk {                                   // work items blockIdx.x, + gridDim.x, ...: decoded once, advanced with carries
    Item cur, step;
    __device__ __forceinline__ void init(const Params& p, int first, int stride) {
        cur = decode_item(p, first);
        step = decode_item(p, stride);
    }
    __device__ __forceinline__ void advance(const Params& p) {
        cur.tx += step.tx;
        int carry = 0;
        if (cur.tx >= p.tiles_x) cur.tx -= p.tiles_x, carry = 1;
        cur.ty += step.ty + carry, carry = 0;
        if (cur.ty >= p.tiles_y) cur.ty -= p.tiles_y, carry = 1;
        cur.b += step.b + carry, carry = 0;
        if (cur.b >= p.batch) cur.b -= p.batch, carry = 1;
        cur.nt += step.nt + carry;
    }
};

// Regions among the outputs of a tile (all four parities of its valid pixels: patch columns 1..14); one whole warp.
__device__ __forceinline__ uint32_t tile_classes(const Params& p, const Item& it, int lane) {
    if (!p.label) return 1u;
    const int wo = 2 * p.w;
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = lane + 32 * i;
        const int tx = r & 15;
        const int iy = it.ty * TH + (r >> 4), ix = it.tx * TW + tx - 1;
        if (tx >= 1 && tx <= TW && iy < p.h && ix < p.w) {
            const uint8_t* lp = p.label + ((int64_t)it.b * 2 * p.h + 2 * iy) * wo + 2 * ix;
            m |= (1u << min((int)lp[0], p.ncls - 1)) | (1u << min((int)lp[1], p.ncls - 1)) | (1u << min((int)lp[wo], p.ncls - 1)) |
                 (1u << min((int)lp[wo + 1], p.ncls - 1));
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    return m ? m : 1u;
}
// next pass of a tile: its one or two regions (cb < 0: one), taken off the remaining-regions mask
__device__ __forceinline__ void next_pass(uint32_t& m, int& ca, int& cb) {
    ca = __ffs(m) - 1;
    m &= m - 1;
    cb = -1;
    if (m) cb = __ffs(m) - 1, m &= m - 1;
}

template <int NTC, int KC>
__global__ void __launch_bounds__(NUM_THREADS, 1) modconv3x3_up_tch_kernel(const __grid_constant__ CUtensorMap wmap, Params p) {
    constexpr int N = NG * NTC;
    constexpr int ROWB = KC * 2;
    constexpr int A_PLANE = A_ROWS * ROWB;
    constexpr int A_STAGE = 2 * A_PLANE;
    constexpr int B_SLOT = N * ROWB;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
    constexpr uint32_t DESC_HI = (uint32_t)((KC == 64 ? 1024u : 512u) >> 4) | (1u << 14) | ((KC == 64 ? 2u : 4u) << 29);
    constexpr int KSTEPS = KC / 16;
    constexpr int CPR = KC / 8;                       // 16-byte chunks per operand row
    constexpr int PPS = NUM_XFORM / CPR;              // halo pixels covered per sweep of the transform threads
    constexpr int NSW = (160 + PPS - 1) / PPS;
    static_assert(N <= ACC_COLS && N % 16 == 0 && NTC % 32 == 0, "UMMA N; the epilogue zeroes 32-column blocks per row parity");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_buf = smem;                                                // [NSTAGE_A][hi|lo][A_ROWS][ROWB]
    uint8_t* b_buf = a_buf + ((NSTAGE_A * A_STAGE + 1023) & ~1023);       // [nslot_b][N][ROWB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)p.nslot_b * B_SLOT);
    const int A_FULL = 0, A_EMPTY = A_FULL + NSTAGE_A, ACC_FULL = A_EMPTY + NSTAGE_A, ACC_EMPTY = ACC_FULL + NACC,
              B_FULL = ACC_EMPTY + NACC, B_EMPTY = B_FULL + p.nslot_b, NBARS = B_EMPTY + p.nslot_b;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = p.cin / KC;
#ifdef E4S_TC_PROFILE
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0;
    long long pw[4] = {0, 0, 0, 0};                      // [0] role time, [1..3] cycles in its barrier waits
    const long long t_start = prof_on ? clock64() : 0;
#endif

    if (threadIdx.x == 0) {
        const uint32_t nmma = p.det ? 1u : (uint32_t)NUM_MMA_WARPS;      // issuing warps that commit to each barrier
        for (int i = 0; i < NSTAGE_A; ++i) mbar_init(smem_u32(&bars[A_FULL + i]), NUM_XFORM), mbar_init(smem_u32(&bars[A_EMPTY + i]), nmma);
        for (int i = 0; i < NACC; ++i) mbar_init(smem_u32(&bars[ACC_FULL + i]), nmma), mbar_init(smem_u32(&bars[ACC_EMPTY + i]), NUM_EPI);
        for (int i = 0; i < p.nslot_b; ++i) mbar_init(smem_u32(&bars[B_FULL + i]), 1), mbar_init(smem_u32(&bars[B_EMPTY + i]), nmma);
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
    if (warp >= W_EPI0 && warp < W_EPI0 + 4) {          // every MMA accumulates: both accumulator buffers start at zero
        const uint32_t lanes = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16);
#pragma unroll 1
        for (int c = 0; c < TMEM_COLS; c += 32) tmem_zero32(lanes + (uint32_t)c);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===================================================================== TMA producer: per (pass, chunk, row tap) ONE 4-D box =
        // the (hi, lo) planes of all six column groups = one slot pair
        int slot = 0;
        uint32_t ph = 0;
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        Walk wk2 = wk;                                     // two items ahead: pulls the (streaming) noise rows into L2 for the epilogue
        wk2.advance(p), wk2.advance(p);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x, wk.advance(p), wk2.advance(p)) {
            const Item item = wk.cur;
            const int npass = (__popc(tile_classes(p, item, lane)) + 1) >> 1;
            if (p.noise && it + 2 * (int)gridDim.x < p.items && lane < 2 * TH) {
                const Item f = wk2.cur;
                const int oy = 2 * f.ty * TH + lane;
                if (oy < 2 * p.h)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.noise + ((int64_t)(p.noise_b == 1 ? 0 : f.b) * 2 * p.h + oy) * 2 * p.w + 2 * f.tx * TW));
            }
            if (elect_one()) {
                for (int ps = 0; ps < npass; ++ps)
                    for (int kc = 0; kc < nchunks; ++kc)
                        for (int tap = 0; tap < 3; ++tap) {
                            TCH_WAIT(smem_u32(&bars[B_EMPTY + slot]), ph ^ 1, 1);
                            const uint32_t full = smem_u32(&bars[B_FULL + slot]);
                            mbar_expect_tx(full, 2 * B_SLOT);
                            tma_load_4d(smem_u32(b_buf + (size_t)slot * B_SLOT), &wmap, kc * KC, item.nt * NTC, tap, 0, full);
                            slot += 2;
                            if (slot >= p.nslot_b) slot = 0, ph ^= 1;
                        }
            }
            __syncwarp();
        }
    } else if (warp <= NUM_MMA_WARPS) {
      if (!p.det || warp == 1) {
        // ===================================================================== MMA issuers: product 0 x_hi w_hi, 1 x_lo w_hi, 2 x_hi w_lo,
        // one warp each; deterministic mode (p.det): warp 1 issues all three in this order, the other two idle
        const int role = warp - 1;
        constexpr uint32_t A_LO = (uint32_t)A_PLANE >> 4, W_LO = (uint32_t)B_SLOT >> 4;
        const uint32_t a_role = role == 1 ? A_LO : 0u, w_role = role == 2 ? W_LO : 0u;
        int sa = 0, slot = 0, acc = 0;
        uint32_t pa = 0, pb = 0, pacc = 0;                // bit b = phase of accumulator buffer b
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t bars0 = smem_u32(bars);
        const uint32_t a0 = smem_u32(a_buf), b0 = smem_u32(b_buf);
        auto desc = [](uint32_t lo) -> uint64_t { return ((uint64_t)DESC_HI << 32) | lo; };
        auto lo_of = [](uint32_t addr) -> uint32_t { return (addr >> 4) | 0x10000u; };
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x, wk.advance(p)) {
            const Item item = wk.cur;
            uint32_t rem = __shfl_sync(0xffffffffu, tile_classes(p, item, lane), 0);
            while (rem) {
                int ca, cb;
                next_pass(rem, ca, cb);
                const bool two = cb >= 0;
                TCH_WAIT(bars0 + 8 * (ACC_EMPTY + acc), ((pacc >> acc) & 1u) ^ 1u, 1);
                if (two) TCH_WAIT(bars0 + 8 * (ACC_EMPTY + (acc ^ 1)), ((pacc >> (acc ^ 1)) & 1u) ^ 1u, 1);
                tc_fence_after();
                const uint32_t d_a = tmem_u + (uint32_t)(acc * ACC_COLS), d_b = tmem_u + (uint32_t)((acc ^ 1) * ACC_COLS);
#pragma unroll 1
                for (int kc = 0; kc < nchunks; ++kc) {
                    // region A's operand in stage sa; region B's (two-region pass) in the next stage of the ring
                    const int sb = (sa + 1 == NSTAGE_A) ? 0 : sa + 1;
                    const uint32_t pbs = (sa + 1 == NSTAGE_A) ? pa ^ 1 : pa;
                    TCH_WAIT(bars0 + 8 * (A_FULL + sa), pa, 2);
                    if (two) TCH_WAIT(bars0 + 8 * (A_FULL + sb), pbs, 2);
                    tc_fence_after();
                    const uint32_t apA = lo_of(a0 + sa * A_STAGE), apB = lo_of(a0 + sb * A_STAGE);
#pragma unroll 1
                    for (int tap = 0; tap < 3; ++tap) {
                        TCH_WAIT(bars0 + 8 * (B_FULL + slot), pb, 3);
                        tc_fence_after();
                        const uint32_t bp = lo_of(b0 + slot * B_SLOT);
                        const uint32_t roff = (uint32_t)((1 + 16 * tap) * ROWB) >> 4;       // halo pixel hp is operand row hp + 1
                        if (elect_one()) {
                            if (!p.det) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_a, desc(apA + a_role + roff + 2 * k), desc(bp + w_role + 2 * k), IDESC, 1u);
                                if (two) {
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_b, desc(apB + a_role + roff + 2 * k), desc(bp + w_role + 2 * k), IDESC, 1u);
                                }
                            } else {
#pragma unroll 1
                                for (int r = 0; r < 3; ++r) {
                                    const uint32_t ao = r == 1 ? A_LO : 0u, wo2 = r == 2 ? W_LO : 0u;
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_a, desc(apA + ao + roff + 2 * k), desc(bp + wo2 + 2 * k), IDESC, 1u);
                                    if (two) {
#pragma unroll
                                        for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_b, desc(apB + ao + roff + 2 * k), desc(bp + wo2 + 2 * k), IDESC, 1u);
                                    }
                                }
                            }
                            umma_commit(bars0 + 8 * (B_EMPTY + slot));
                        }
                        slot += 2;
                        if (slot >= p.nslot_b) slot = 0, pb ^= 1;
                    }
                    if (elect_one()) {
                        umma_commit(bars0 + 8 * (A_EMPTY + sa));
                        if (two) umma_commit(bars0 + 8 * (A_EMPTY + sb));
                    }
                    if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                    if (two && ++sa == NSTAGE_A) sa = 0, pa ^= 1;
                }
                if (elect_one()) umma_commit(bars0 + 8 * (ACC_FULL + acc));
                pacc ^= 1u << acc;
                if (two) {                                   // both buffers used: hand both over, buffer order unchanged
                    if (elect_one()) umma_commit(bars0 + 8 * (ACC_FULL + (acc ^ 1)));
                    pacc ^= 1u << (acc ^ 1);
                } else {
                    acc ^= 1;
                }
                __syncwarp();
            }
        }
      }
    } else if (warp < W_EPI0) {
        // ===================================================================== activation transform: fp32 x region style -> bf16 hi/lo stage
        const int t = threadIdx.x - 32 * W_XFORM0;       // 0..255
        const int c8 = t % CPR;
        const int pix0 = t / CPR;
        int sa = 0;
        uint32_t pa = 0;
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x, wk.advance(p)) {
            const Item item = wk.cur;
            uint32_t rem = tile_classes(p, item, lane);
            const float* xb = p.x + (int64_t)item.b * p.h * p.w * p.cin;
            const int y0 = item.ty * TH, x0 = item.tx * TW;
            while (rem) {
                int cls2[2];
                next_pass(rem, cls2[0], cls2[1]);
                const int nclass = cls2[1] >= 0 ? 2 : 1;
                for (int kc = 0; kc < nchunks; ++kc) {
                    const int ch = kc * KC + 8 * c8;
                    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v0[NSW], v1[NSW];
#pragma unroll
                    for (int i = 0; i < NSW; ++i) {
                        const int hp = pix0 + PPS * i;
                        const int gy = y0 - 1 + (hp >> 4), gx = x0 - 1 + (hp & 15);
                        v0[i] = zero4, v1[i] = zero4;
                        if (hp < 160 && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
                            const float* src = xb + ((int64_t)gy * p.w + gx) * p.cin + ch;
                            v0[i] = __ldg(reinterpret_cast<const float4*>(src));
                            v1[i] = __ldg(reinterpret_cast<const float4*>(src + 4));
                        }
                    }
#pragma unroll 1
                    for (int ci = 0; ci < nclass; ++ci) {
                        const float* sc = p.s + ((int64_t)item.b * p.ncls + cls2[ci]) * p.cin;
                        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + ch));
                        const float4 s1 = __ldg(reinterpret_cast<const float4*>(sc + ch + 4));
                        TCH_WAIT(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1, 1);
                        const uint32_t hi_plane = smem_u32(a_buf) + (uint32_t)sa * A_STAGE, lo_plane = hi_plane + A_PLANE;
#pragma unroll
                        for (int i = 0; i < NSW; ++i) {
                            const int hp = pix0 + PPS * i;
                            if (hp >= 160) continue;
                            const int row = hp + 1;
                            const uint32_t sx = KC == 64 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                            const uint32_t off = (uint32_t)row * ROWB + (((uint32_t)c8 ^ sx) << 4);
                            scale_split_store8(v0[i], v1[i], s0, s1, s0, s1, false, hi_plane + off, lo_plane + off);
                        }
                        fence_proxy_async();
                        mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                        if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                    }
                }
            }
        }
    } else {
        // ===================================================================== epilogue: horizontal half of the blur, region selection,
        // demodulation, noise, bias, activation.  One thread = one patch pixel (m, n') and ONE output row parity py (two output
        // columns): warps W_EPI0..+3 take py = 0, the next four py = 1 - eight warps, two per TMEM lane quarter (the first
        // version's four warps, one per scheduler with nothing to switch to, ran this role at 0.27 instructions per cycle and
        // bound the kernel: profiles/r2_stall_attribution_tch_v1.log).  Arithmetic on packed fp32 pairs (FFMA2 / FMUL2 / FADD2).
        const uint32_t quarter = (uint32_t)(warp & 3);
        const int py = (warp - W_EPI0) >> 2;
        const int m_row = quarter * 32 + lane;
        const int ty = m_row >> 4, tx = m_row & 15;
        int acc = 0;
        uint32_t pacc = 0;                                // bit b = phase of accumulator buffer b
        const float nw = (p.noise && p.noise_w) ? __ldg(p.noise_w) : 0.f;
        const uint64_t F0 = pk2(p.f0, p.f0), F1 = pk2(p.f1, p.f1), F2 = pk2(p.f2, p.f2), F3 = pk2(p.f3, p.f3);
        const float k2 = 1.41421356237309515f;
        const uint64_t S2 = pk2(k2, k2), S02 = pk2(0.2f * k2, 0.2f * k2);
        const int ho = 2 * p.h, wo = 2 * p.w;
        const uint32_t lanes = tmem_base + ((quarter * 32u) << 16);
        auto fetch = [&](bool valid, const Item& i2, int (&c)[2], float (&z)[2]) {
            c[0] = c[1] = 0, z[0] = z[1] = 0.f;
            if (!valid) return;
            const int iy = i2.ty * TH + ty, ix = i2.tx * TW + tx - 1;
            if (!(tx >= 1 && tx <= TW && iy < p.h && ix < p.w)) return;
            const int64_t o = ((int64_t)i2.b * ho + 2 * iy + py) * wo + 2 * ix;
            if (p.label) c[0] = min((int)p.label[o], p.ncls - 1), c[1] = min((int)p.label[o + 1], p.ncls - 1);
            if (p.noise) {
                const float* nzp = p.noise + ((int64_t)(p.noise_b == 1 ? 0 : i2.b) * ho + 2 * iy + py) * wo + 2 * ix;
                z[0] = __ldg(nzp), z[1] = __ldg(nzp + 1);
            }
        };
        // six-term horizontal combination of 8 channels (4 packed pairs) from one region's accumulators
        auto combine = [&](uint32_t col, uint64_t (&o0)[4], uint64_t (&o1)[4]) {
            uint32_t t0[8], t1[8], t2[8];
            tmem_ld8x3(lanes + col, lanes + col + NTC, lanes + col + 2 * NTC, t0, t1, t2);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint64_t a0 = pk2u(t0[2 * e], t0[2 * e + 1]), a1 = pk2u(t1[2 * e], t1[2 * e + 1]), a2 = pk2u(t2[2 * e], t2[2 * e + 1]);
                const uint64_t l1 = pk2u(__shfl_up_sync(0xffffffffu, t1[2 * e], 1), __shfl_up_sync(0xffffffffu, t1[2 * e + 1], 1));
                const uint64_t l2 = pk2u(__shfl_up_sync(0xffffffffu, t2[2 * e], 1), __shfl_up_sync(0xffffffffu, t2[2 * e + 1], 1));
                const uint64_t r0 = pk2u(__shfl_down_sync(0xffffffffu, t0[2 * e], 1), __shfl_down_sync(0xffffffffu, t0[2 * e + 1], 1));
                const uint64_t r1 = pk2u(__shfl_down_sync(0xffffffffu, t1[2 * e], 1), __shfl_down_sync(0xffffffffu, t1[2 * e + 1], 1));
                // px = 0: f0 T1[n-1] + f1 T2[n-1] + f1 T0[n] + f2 T1[n] + f3 T2[n] + f3 T0[n+1]
                o0[e] = fma2(F3, r0, fma2(F3, a2, fma2(F2, a1, fma2(F1, a0, fma2(F1, l2, mul2(F0, l1))))));
                // px = 1: f0 T2[n-1] + f0 T0[n] + f1 T1[n] + f2 T2[n] + f2 T0[n+1] + f3 T1[n+1]
                o1[e] = fma2(F3, r1, fma2(F2, r0, fma2(F2, a2, fma2(F1, a1, fma2(F0, a0, mul2(F0, l2))))));
            }
        };
        int cls_next[2];
        float nz_next[2];
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        fetch(blockIdx.x < p.items, wk.cur, cls_next, nz_next);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = wk.cur;
            wk.advance(p);
            const int iy = item.ty * TH + ty, ix = item.tx * TW + tx - 1;
            const bool mine = tx >= 1 && tx <= TW && iy < p.h && ix < p.w;
            const int n0 = item.nt * NTC;
            const int c0 = cls_next[0], c1 = cls_next[1];
            const uint64_t Z0 = pk2(nw * nz_next[0], nw * nz_next[0]), Z1 = pk2(nw * nz_next[1], nw * nz_next[1]);
            uint32_t rem = tile_classes(p, item, lane);
            fetch(it + (int)gridDim.x < p.items, wk.cur, cls_next, nz_next);
            const float* dm0 = p.demod ? p.demod + ((int64_t)item.b * p.ncls + c0) * p.cout + n0 : nullptr;
            const float* dm1 = p.demod ? p.demod + ((int64_t)item.b * p.ncls + c1) * p.cout + n0 : nullptr;
            float* dst0 = p.y + (((int64_t)item.b * ho + 2 * iy + py) * wo + 2 * ix) * p.cout + n0;
            float* dst1 = dst0 + p.cout;
            while (rem) {
                int ca, cb;
                next_pass(rem, ca, cb);
                const bool two = cb >= 0;
                TCH_WAIT(smem_u32(&bars[ACC_FULL + acc]), (pacc >> acc) & 1u, 1);
                pacc ^= 1u << acc;
                if (two) {
                    TCH_WAIT(smem_u32(&bars[ACC_FULL + (acc ^ 1)]), (pacc >> (acc ^ 1)) & 1u, 1);
                    pacc ^= 1u << (acc ^ 1);
                }
                tc_fence_after();
                const uint32_t cbase = (uint32_t)(acc * ACC_COLS + 3 * py * NTC), cother = (uint32_t)((acc ^ 1) * ACC_COLS + 3 * py * NTC);
                const bool w0 = mine && (c0 == ca || c0 == cb), w1 = mine && (c1 == ca || c1 == cb);
                const bool sel0 = two && c0 == cb, sel1 = two && c1 == cb;      // this output takes region B's combination
#pragma unroll 1
                for (int jb = 0; jb < NTC / 8; ++jb) {
                    const int co = 8 * jb;
                    // epilogue operands first: their (L1-hit) latency overlaps the TMEM loads and shuffles below
                    float4 bv[2], d0v[2], d1v[2];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        bv[h2] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + co + 4 * h2)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        d0v[h2] = dm0 ? __ldg(reinterpret_cast<const float4*>(dm0 + co + 4 * h2)) : make_float4(1.f, 1.f, 1.f, 1.f);
                        d1v[h2] = dm1 ? __ldg(reinterpret_cast<const float4*>(dm1 + co + 4 * h2)) : make_float4(1.f, 1.f, 1.f, 1.f);
                    }
                    uint64_t o0[4], o1[4];
                    combine(cbase + (uint32_t)co, o0, o1);
                    if (two) {
                        uint64_t q0[4], q1[4];
                        combine(cother + (uint32_t)co, q0, q1);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (sel0) o0[e] = q0[e];
                            if (sel1) o1[e] = q1[e];
                        }
                    }
                    auto finish = [&](const uint64_t (&o)[4], const float4 (&d)[2], uint64_t Z, float* dst) {
                        const uint64_t dp[4] = {pk2(d[0].x, d[0].y), pk2(d[0].z, d[0].w), pk2(d[1].x, d[1].y), pk2(d[1].z, d[1].w)};
                        const uint64_t bp[4] = {pk2(bv[0].x, bv[0].y), pk2(bv[0].z, bv[0].w), pk2(bv[1].x, bv[1].y), pk2(bv[1].z, bv[1].w)};
                        float r[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint64_t t = fma2(o[e], dp[e], add2(bp[e], Z));
                            if (p.act) {                                  // sqrt(2) * lrelu_0.2(t) = max(sqrt(2) t, 0.2 sqrt(2) t)
                                float a_lo, a_hi, b_lo, b_hi;
                                upk2(mul2(t, S2), a_lo, a_hi);
                                upk2(mul2(t, S02), b_lo, b_hi);
                                r[2 * e] = fmaxf(a_lo, b_lo), r[2 * e + 1] = fmaxf(a_hi, b_hi);
                            } else {
                                upk2(t, r[2 * e], r[2 * e + 1]);
                            }
                        }
                        st_global_v8(dst + co, make_float4(r[0], r[1], r[2], r[3]), make_float4(r[4], r[5], r[6], r[7]));
                    };
                    if (w0) finish(o0, d0v, Z0, dst0);
                    if (w1) finish(o1, d1v, Z1, dst1);
                }
                // every MMA accumulates: hand this row parity's column groups back zeroed
#pragma unroll
                for (int c = 0; c < 3 * NTC; c += 32) tmem_zero32(lanes + cbase + (uint32_t)c);
                if (two) {
#pragma unroll
                    for (int c = 0; c < 3 * NTC; c += 32) tmem_zero32(lanes + cother + (uint32_t)c);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
                if (two) mbar_arrive(smem_u32(&bars[ACC_EMPTY + (acc ^ 1)]));
                else acc ^= 1;
            }
        }
    }

#ifdef E4S_TC_PROFILE
    if (prof_on && lane == 0) {
        const int role = warp == 0 ? 0 : warp == 1 ? 1 : warp == W_XFORM0 ? 2 : warp == W_EPI0 ? 3 : -1;
        if (role >= 0) {
            pw[0] = clock64() - t_start;
#pragma unroll
            for (int k = 0; k < 4; ++k) p.prof[role * 4 + k] = pw[k];
        }
    }
#endif
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

static long long* g_prof = nullptr;

template <int NTC, int KC>
static int launch(const void* v_hilo, Params p, cudaStream_t st) {
    p.prof = g_prof;
    p.det = e4s_get_deterministic();
    constexpr int N = NG * NTC, ROWB = KC * 2;
    constexpr int A_BYTES = ((NSTAGE_A * 2 * A_ROWS * ROWB) + 1023) & ~1023;
    constexpr int B_SLOT = N * ROWB;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return E4S_ERR_ARCH;
    CUtensorMap map;
    // weights [2][6][3][Cout][Cin] bf16 as a 4-D tensor (Cin, Cout, row tap, hl * 6 + column group)
    cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.cout, 3, (cuuint64_t)2 * NG};
    cuuint64_t strides[3] = {(cuuint64_t)p.cin * 2, (cuuint64_t)p.cout * p.cin * 2, (cuuint64_t)3 * p.cout * p.cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)NTC, 1, (cuuint32_t)2 * NG};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v_hilo), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return 700 + (int)cr;

    p.tiles_x = (int)e4s_ceil_div(p.w, TW);
    p.tiles_y = (int)e4s_ceil_div(p.h, TH);
    p.n_tiles = p.cout / NTC;
    const int64_t items = (int64_t)p.tiles_x * p.tiles_y * p.batch * p.n_tiles;
    if (items >= (1ll << 31)) return E4S_ERR_SHAPE;
    p.items = (int)items;
    int max_slots = (SMEM_BUDGET - A_BYTES - 1024) / B_SLOT;
    if (max_slots < 2) return E4S_ERR_SHAPE;
    p.nslot_b = max_slots > 12 ? 12 : (max_slots & ~1);      // even: (hi, lo) slot pairs never straddle the ring wrap
    const size_t smem = 1024 + A_BYTES + (size_t)p.nslot_b * B_SLOT + (size_t)(2 * NSTAGE_A + 2 * NACC + 2 * p.nslot_b) * 8 + 64;
    static E4sSmemOptIn optin;
    if (const int rc = e4s_smem_optin(optin, modconv3x3_up_tch_kernel<NTC, KC>, smem)) return rc;
    const int sms = e4s_num_sms();
    const int grid = p.items < sms ? p.items : sms;
    modconv3x3_up_tch_kernel<NTC, KC><<<grid, NUM_THREADS, smem, st>>>(map, p);
    return e4s_launch_status();
}

}  // namespace tch

extern "C" int e4s_modconv3x3_up_tch_fwd(const float* x, const void* v_hilo_bf16, const float* s, const float* demod,
                                         const uint8_t* label, const float* noise, const float* noise_w, const float* bias,
                                         float* y, float fx0, float fx1, float fx2, float fx3, int batch, int h, int w, int cin,
                                         int cout, int ncls, int noise_b, int act, void* stream) {
    E4S_REQUIRE(x && v_hilo_bf16 && s && y, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && ncls > 0 && ncls <= 32, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0 && (cout % 32) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(label || ncls == 1, E4S_ERR_ARG);
    E4S_REQUIRE(!noise || (noise_w && (noise_b == 1 || noise_b == batch)), E4S_ERR_ARG);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(v_hilo_bf16) && e4s_aligned16(s) && (!demod || e4s_aligned16(demod)) &&
                    (!bias || e4s_aligned16(bias)),
                E4S_ERR_ALIGN);
    E4S_REQUIRE((reinterpret_cast<uintptr_t>(y) & 31) == 0, E4S_ERR_ALIGN);          // 256-bit stores
    tch::Params p{x, s, demod, label, noise, noise_w, bias, y, fx0, fx1, fx2, fx3, batch, h, w, cin, cout, ncls, noise_b, act ? 1 : 0,
                  0, 0, 0, 0, 0, 0, nullptr};
    if ((cin % 64) == 0) return tch::launch<32, 64>(v_hilo_bf16, p, (cudaStream_t)stream);
    return tch::launch<32, 32>(v_hilo_bf16, p, (cudaStream_t)stream);
}

// Diagnostic: per-role stall attribution of CTA 0 of every following H-form launch ([4 roles][4] int64 cycle counters in
// device memory: role time, then the cycles it spent in its barrier waits).  nullptr switches it off (the default).
extern "C" int e4s_tch_set_profile(long long* device_counters) {
#ifdef E4S_TC_PROFILE
    tch::g_prof = device_counters;
    return E4S_OK;
#else
    return device_counters ? E4S_ERR_ARG : E4S_OK;      // production build carries no counters
#endif
}
