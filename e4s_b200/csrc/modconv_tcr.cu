// Region-selected modulated 3x3 convolution on tcgen05 tensor cores - fourth generation (sm_100a).
//
// Contract and implicit-GEMM formulation: modconv_tc.cu's header (8x16 pixel tile with 14 valid columns, taps as
// row-shifted descriptors over one staged halo tile, split-bf16 x3 accumulation in fp32 TMEM, weights streamed by TMA
// or resident).  Persistent CTAs (one per SM), 16 warps in four roles connected by mbarrier rings:
//
//   warp 0       TMA producer: per tap ONE 4-D box = the (w_hi, w_lo) slot pair of all parities; in XS mode (Cin <= 64)
//                also the raw fp32 halo tile of every chunk, issued one chunk ahead of the weights it is used with
//   warps 1-3    MMA issue, one warp per split-precision product (x_hi w_hi, x_lo w_hi, x_hi w_lo)
//   warps 4-11   transform: fp32 activations x region style -> bf16 hi/lo operand stage (128-/64-byte swizzle)
//   warps 12-15  epilogue: TMEM -> demodulate, noise, bias, activation -> HBM; zero the accumulator buffer
//
// What the measurements under profiles/ built into it:
//
//  * ONE main-loop pass per tile whatever the mask.  Region-pure tiles stage the operand once per K chunk (scaled by the
//    region's style) and use the row-shift trick.  Two-region tiles of an up-sampling layer (most mixed tiles of a face
//    mask) stage it twice per chunk, once per region, accumulate the two in the two TMEM buffers with the same pure-tile
//    MMA sequence (N = 4 x NTC, tensor-pipe bound) and let the epilogue pick per (pixel, parity).  Other mixed tiles use
//    ROW-CLASS staging: the operand of every (tap, parity) is materialised separately, each row scaled by the style of
//    that row's own output pixel, MMAs per (tap, parity) with N = NTC.
//  * MMA issue is a per-warp resource.  tools/ubench/umma_bench.cu: one warp gets a tcgen05.mma out every ~120-190
//    cycles whatever its shape, while the pipe needs 16 (N = 32) ... 128 (N = 256) cycles; two / four issuing warps
//    overlap (80-94 / 47-56 cycles per MMA at N <= 64).  Hence three issuing warps.  MMAs of different warps have no
//    defined order, so none may be the "first" (accumulate = 0): the epilogue zeroes a buffer after reading it
//    (tcgen05.st) and every MMA accumulates.
//  * A consumer that only LOADS a TMA-filled buffer and then arrives on its "empty" barrier does not wait for the loads:
//    the raw-tile ring is released after the stores that consume the values (a parity bug found at 1024x1024 shapes).
//  * The weight producer was issue-bound with one 2-D box per (tap, hi/lo, parity) (144 per tile on the 64->32
//    up-sampling layer): one 4-D box per tap.  The noise map is a streaming tensor: the epilogue fetches its operands one
//    work item ahead.
//
// K chunk: 64 channels (128-byte swizzle) when Cin % 64 == 0 and Cin > 64, else 32 channels (64-byte swizzle).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tcr {
using namespace tcx;

constexpr int TH = 8, TWP = 16, TW = 14;
constexpr int A_ROWS = 168;
constexpr int NSTAGE_A = 2;
// Warp roles: 0 TMA producer (weights; raw activation tiles in XS mode), 1-3 MMA issue (one per split-precision
// product), 4-11 transform, 12-15 epilogue.  512 threads: 128 registers each.
constexpr int NUM_MMA_WARPS = 3;
constexpr int W_XFORM0 = 1 + NUM_MMA_WARPS, W_EPI0 = W_XFORM0 + 8, W_XS = W_EPI0 + 4;
constexpr int NUM_THREADS = 32 * W_XS;        // 512
constexpr int NUM_THREADS_XS = NUM_THREADS;  // XS mode: the raw-tile producer is the weights thread
constexpr int NXS = 3;                        // raw-tile ring depth (XS mode, 32-channel chunks: 20 KB per stage)
constexpr int XS_STAGE = 160 * 32 * 4;
constexpr int NUM_XFORM = 256, NUM_EPI = 128;
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
    int det;              // 1: ONE warp issues the three split-precision products in a fixed order (bit-reproducible accumulation)
    long long* prof;      // optional [5 roles][4] cycle counters of CTA 0 (e4s_tcr_set_profile); nullptr in production
    int tap_mask;         // bit t set: tap t (row-major 3x3) is multiplied; 0 = all nine.  Taps whose weights are zero by construction
                          // (a stride-2 convolution on a space-to-depth tensor uses 4 of 9) are neither loaded nor issued.
};

// Stall attribution (tools/opbench.py --prof): only in the diagnostic build (-DE4S_TCR_PROFILE, libe4s_b200_prof.so) -
// the eight counter registers cost the mixed-tile transform path ~10 % through spills.
#if defined(E4S_TC_PROFILE) && !defined(E4S_TCR_PROFILE)
#define E4S_TCR_PROFILE 1
#endif
#ifdef E4S_TCR_PROFILE
__device__ __forceinline__ void mbar_wait_p(uint32_t bar, uint32_t parity, long long& ctr, bool on) {
    if (!on) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    ctr += clock64() - t0;
}
#define MBAR_WAIT_P(bar, parity, k) mbar_wait_p(bar, parity, pw[k], prof_on)
#define PROF_BEGIN() const long long prof_t0 = prof_on ? clock64() : 0
#define PROF_END(k) do { if (prof_on) pw[k] += clock64() - prof_t0; } while (0)
#else
#define MBAR_WAIT_P(bar, parity, k) mbar_wait(bar, parity)
#define PROF_BEGIN() do {} while (0)
#define PROF_END(k) do {} while (0)
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

// Work items of one CTA: it = blockIdx.x, + gridDim.x, ...  The divisions of decode_item cost every role ~300 cycles per
// item (9 % of all warp samples on the 32->32 layer): decoded once, then advanced with carries.
struct Walk {
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

// Regions present among the valid (pixel, parity) outputs of a tile; one whole warp, every role recomputes it.
// UP2 (parity work items of an up-sampling layer): only the outputs (2 iy + py, 2 ix + px) of the item's own parity count.
template <int NPH, bool UP2>
__device__ __forceinline__ uint32_t tile_class_mask(const Params& p, const Item& it, int lane) {
    if (!p.label) return 1u;
    constexpr int MUL = (NPH == 4 || UP2) ? 2 : 1;
    const int ho = p.h * MUL, wo = p.w * MUL;
    const int py = UP2 ? ((it.nt >> 1) & 1) : 0, px = UP2 ? (it.nt & 1) : 0;
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = lane + 32 * i;
        const int iy = it.ty * TH + (r >> 4), ix = it.tx * TW + (r & 15);
        if ((r & 15) < TW && iy < p.h && ix < p.w) {
            const uint8_t* lp = p.label + ((int64_t)it.b * ho + iy * MUL + py) * wo + ix * MUL + px;
            m |= 1u << min((int)lp[0], p.ncls - 1);
            if (NPH == 4) m |= (1u << min((int)lp[1], p.ncls - 1)) | (1u << min((int)lp[wo], p.ncls - 1)) | (1u << min((int)lp[wo + 1], p.ncls - 1));
        }
    }
    return __reduce_or_sync(0xffffffffu, m);
}

__device__ __forceinline__ void xform_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------- kernel
// XS = true (32-channel chunks only): the raw fp32 halo tile of every chunk is streamed by a 4-D TMA load (hardware zero
// fill outside the image) through a 4-stage shared-memory ring by a dedicated producer warp, and the transform warps
// convert shared -> shared.  Used for the small-K layers (Cin <= 64), which are HBM-bound: without it each tile
// exposes a full DRAM latency in the transform warps (ncu: 11 % tensor pipe, 15 % DRAM on the 32->32 layer).
//
// UP2 = true: an up-sampling layer run as PARITY WORK ITEMS.  A work item is (pixel tile, N tile, output parity): it is a
// plain 3x3 convolution with that parity's folded kernel (NPH = 1 inside the MMA, N = NTC up to 256) whose outputs go to
// (2 iy + py, 2 ix + px).  Against the NPH = 4 form (four parities along N, NTC <= 64) a region-pure tile costs the same
// (N = 256 either way, one operand stage per chunk and 256 accumulator columns), but a tile that mixes regions stages
// its row-class operand 9 times per chunk for MMAs of N = 256 instead of 36 times for MMAs of N = 64 - the masked
// low-resolution up-sampling layers (>= 3 regions in most tiles at <= 64x64) were transform- and issue-bound at 27-33 %
// tensor pipe (profiles/r1_ncu_tcr_17_layers.md).  Item::nt carries (N tile) * 4 + parity.
template <int NTC, int KC, int NPH, bool XS, bool UP2, bool STK>
__global__ void __launch_bounds__(XS ? NUM_THREADS_XS : NUM_THREADS, 1)
modconv3x3_tcr_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap xmap, Params p) {
    static_assert(!XS || KC == 32, "XS mode stages 32-channel chunks");
    static_assert(!UP2 || NPH == 1, "parity work items are plain-convolution items");
    // STK (small N tiles of plain layers): the w_hi and w_lo slots of a tap are contiguous in shared memory, so ONE MMA with
    // N = 2 NTC multiplies x_hi by both (columns [0, NTC): x_hi w_hi, [NTC, 2 NTC): x_hi w_lo) and a second one adds x_lo w_hi to
    // the first half: two MMAs and two operand fetches per (tap, K step) instead of three.  At N <= 64 an MMA costs its A-operand
    // fetch from shared memory (~80 cycles for 16 of arithmetic at N = 32), not its arithmetic: the 32->32 layer at 1024x1024
    // spent 54 such MMAs per tile.  The epilogue adds the two halves.
    static_assert(!STK || (NPH == 1 && !UP2 && 2 * NTC <= 256), "stacked hi/lo weights: plain layers, N tile <= 128");
    constexpr int N = NTC * NPH;
    constexpr int ROWB = KC * 2;
    constexpr int A_PLANE = A_ROWS * ROWB;
    constexpr int A_STAGE = 2 * A_PLANE;
    constexpr int B_SLOT = N * ROWB;
    // MMA ISSUE: one warp issues a tcgen05.mma every ~120-190 cycles whatever its shape (tools/ubench/umma_bench.cu,
    // profiles/r1_umma_bench_*.log), while the tensor pipe needs 16 (N = 32) ... 128 (N = 256) cycles for it; several
    // warps issuing concurrently do overlap.  The three split-precision products x_hi*w_hi, x_lo*w_hi, x_hi*w_lo are
    // therefore issued by three warps.  MMAs of different warps have no defined order, so none of them may be the
    // "first" one (accumulate = 0): the epilogue warps zero an accumulator buffer after reading it (tcgen05.st) and
    // every MMA accumulates.
    constexpr int ACC_COLS = 256;
    // Two-region tiles of an up-sampling layer (most mixed tiles of a face mask): see the transform role.
    constexpr bool TWO_CLASS = (NPH == 4) && (NSTAGE_A == 2);
    constexpr int NACC = 2;
    constexpr int TMEM_COLS = 512;
    constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC_N = IDESC_BASE | ((uint32_t)(N >> 3) << 17);         // all parities in one MMA
    constexpr uint32_t IDESC_Q = IDESC_BASE | ((uint32_t)(NTC >> 3) << 17);       // one parity
    constexpr uint32_t IDESC_2N = IDESC_BASE | ((uint32_t)((2 * N) >> 3) << 17);  // STK: w_hi and w_lo stacked along N
    constexpr int NMMA = STK ? 2 : NUM_MMA_WARPS;                                 // issuing warps in use
    constexpr uint32_t DESC_HI = (uint32_t)((KC == 64 ? 1024u : 512u) >> 4) | (1u << 14) | ((KC == 64 ? 2u : 4u) << 29);
    constexpr int KSTEPS = KC / 16;
    constexpr int MUL = (NPH == 4 || UP2) ? 2 : 1;
    constexpr int CPR = KC / 8;                       // 16-byte chunks per operand row
    constexpr int PPS = NUM_XFORM / CPR;              // pixels (rows) covered per sweep of the transform threads
    static_assert(N <= 256 && N % 16 == 0 && NTC % 16 == 0, "UMMA N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_buf = smem;                                                // [NSTAGE_A][hi|lo][A_ROWS][ROWB]
    uint8_t* xs_buf = a_buf + ((NSTAGE_A * A_STAGE + 1023) & ~1023);      // [NXS][160][32] fp32 (XS mode only)
    uint8_t* b_buf = xs_buf + (XS ? NXS * XS_STAGE : 0);                  // [nslot_b][N][ROWB]
    float* s_tab = reinterpret_cast<float*>(b_buf + (size_t)p.nslot_b * B_SLOT);   // [2][ncls][KC] styles of the current chunk
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_tab + 2 * p.ncls * KC);
    const int A_FULL = 0, A_EMPTY = A_FULL + NSTAGE_A, ACC_FULL = A_EMPTY + NSTAGE_A, ACC_EMPTY = ACC_FULL + NACC,
              B_FULL = ACC_EMPTY + NACC, B_EMPTY = B_FULL + p.nslot_b, XS_FULL = B_EMPTY + p.nslot_b, XS_EMPTY = XS_FULL + NXS,
              NBARS = XS_EMPTY + NXS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ho = p.h * MUL, wo = p.w * MUL;
    const int nchunks = p.cin / KC;
    const uint32_t tmask = p.tap_mask ? (uint32_t)p.tap_mask : 0x1FFu;
#ifdef E4S_TCR_PROFILE
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0;
    long long pw[4] = {0, 0, 0, 0};                      // [0] role time, [1..3] cycles in its barrier waits
    const long long t_start = prof_on ? clock64() : 0;
#endif

    if (threadIdx.x == 0) {
        // every MMA warp commits to the barriers of what it read: both operand planes are read by two of them
        // (x_hi: warps 0 and 2, w_hi: warps 0 and 1) - A stages, accumulators and (hi, lo) weight slot PAIRS (the ring unit:
        // one TMA box, the barriers of the even slot) are released by all three
        const uint32_t nmma = p.det ? 1u : (uint32_t)NMMA;               // issuing warps that commit to each barrier
        for (int i = 0; i < NSTAGE_A; ++i) mbar_init(smem_u32(&bars[A_FULL + i]), NUM_XFORM), mbar_init(smem_u32(&bars[A_EMPTY + i]), nmma);
        for (int i = 0; i < NACC; ++i) mbar_init(smem_u32(&bars[ACC_FULL + i]), nmma), mbar_init(smem_u32(&bars[ACC_EMPTY + i]), NUM_EPI);
        for (int i = 0; i < p.nslot_b; ++i) mbar_init(smem_u32(&bars[B_FULL + i]), 1), mbar_init(smem_u32(&bars[B_EMPTY + i]), nmma);
        for (int i = 0; i < NXS; ++i) mbar_init(smem_u32(&bars[XS_FULL + i]), 1), mbar_init(smem_u32(&bars[XS_EMPTY + i]), NUM_XFORM);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    if (XS && warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= W_EPI0 && warp < W_XS) {                 // every MMA accumulates: both accumulator buffers start at zero
        const uint32_t lanes = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16);
#pragma unroll 1
        for (int c = 0; c < TMEM_COLS; c += 32) tmem_zero32(lanes + (uint32_t)c);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===================================================================== TMA producer: weight planes (one pass per tile, or
        // once when resident) and, in XS mode, the raw activation tile of every chunk.  One thread issues both in consumption
        // order: x(chunk) precedes the weight taps of that chunk, so a full weight ring never holds back an activation tile
        // the MMA warps are (indirectly) waiting for.
        int slot = 0, st = 0;
        uint32_t ph = 0, phx = 0;
        bool loaded_resident = false;
        if (elect_one()) {
            // the activation tile runs one chunk AHEAD of the weights it is consumed with (the transform warps need a
            // few thousand cycles to turn it into an operand stage)
            auto issue_x = [&](const Item& i2, int kc2) {
                MBAR_WAIT_P(smem_u32(&bars[XS_EMPTY + st]), phx ^ 1, 2);
                const uint32_t full = smem_u32(&bars[XS_FULL + st]);
                mbar_expect_tx(full, XS_STAGE);
                tma_load_4d(smem_u32(xs_buf + st * XS_STAGE), &xmap, kc2 * KC, i2.tx * TW - 1, i2.ty * TH - 1, i2.b, full);
                if (++st == NXS) st = 0, phx ^= 1;
            };
            Walk wk;
            wk.init(p, blockIdx.x, gridDim.x);
            // The noise map is a streaming tensor: the epilogue fetches its value one work item ahead, which on the HBM-bound
            // top layers is less than a DRAM round trip under load (the epilogue warps stalled on it: ncu source view of round
            // 2).  This otherwise idle thread pulls the noise rows of the tile TWO items ahead into L2.
            Walk wk2 = wk;
            wk2.advance(p), wk2.advance(p);
            constexpr int OM = (NPH == 4 || UP2) ? 2 : 1;
            if (XS && blockIdx.x < p.items) issue_x(wk.cur, 0);
            for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
                const bool load_w = !(p.resident && loaded_resident);
                if (!XS && !load_w) break;
                const Item item = wk.cur;
                wk.advance(p);
                if (p.noise && it + 2 * (int)gridDim.x < p.items) {
                    const Item f = wk2.cur;
                    const int oh2 = p.h * OM, ow2 = p.w * OM;
                    const float* nrow = p.noise + ((int64_t)(p.noise_b == 1 ? 0 : f.b) * oh2 + f.ty * TH * OM) * ow2 + f.tx * TW * OM;
                    for (int r = 0; r < TH * OM && f.ty * TH * OM + r < oh2; ++r)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (int64_t)r * ow2));
                }
                wk2.advance(p);
                for (int kc = 0; kc < nchunks; ++kc) {
                    if (XS) {
                        if (kc + 1 < nchunks) issue_x(item, kc + 1);
                        else if (it + (int)gridDim.x < p.items) issue_x(wk.cur, 0);
                    }
                    if (load_w) {
                        // one 4-D box per tap: [hi | lo] x parities x NTC rows x KC channels = the (hi, lo) slot pair, contiguous
                        for (int tap = 0; tap < 9; ++tap) {
                            if (!((tmask >> tap) & 1u)) continue;
                            if (!p.resident) MBAR_WAIT_P(smem_u32(&bars[B_EMPTY + slot]), ph ^ 1, 1);
                            const uint32_t full = smem_u32(&bars[B_FULL + slot]);
                            mbar_expect_tx(full, 2 * B_SLOT);
                            if (UP2) tma_load_5d(smem_u32(b_buf + (size_t)slot * B_SLOT), &wmap, kc * KC, (item.nt >> 2) * NTC, tap, item.nt & 3, 0, full);
                            else tma_load_4d(smem_u32(b_buf + (size_t)slot * B_SLOT), &wmap, kc * KC, item.nt * NTC, tap, 0, full);
                            slot += 2;
                            if (slot >= p.nslot_b) slot = 0, ph ^= 1;
                        }
                    }
                }
                loaded_resident = true;
            }
        }
        __syncwarp();
    } else if (warp <= NUM_MMA_WARPS) {
      if ((!p.det || warp == 1) && warp - 1 < NMMA) {
        // ===================================================================== MMA issuers (one split-precision product each)
        // product 0: x_hi * w_hi, 1: x_lo * w_hi, 2: x_hi * w_lo.  Each warp walks the loops with warp-uniform state (barrier
        // waits included) and one lane, chosen by elect.sync, issues.  Deterministic mode (p.det): warp 1 alone issues the
        // three products of every (tap, K step) in this order, the other two warps idle - MMAs of one thread execute in issue
        // order, so the accumulation order, hence every output bit, is the same in every run (at ~1/3 of the issue rate).
        const int role = warp - 1;
        constexpr uint32_t A_LO = (uint32_t)A_PLANE >> 4, W_LO = (uint32_t)B_SLOT >> 4;      // descriptor offsets of the x_lo plane / w_lo slot
        const uint32_t a_role = role == 1 ? A_LO : 0u, w_role = role == 2 ? W_LO : 0u;     // this warp's product (default mode)
        // STK: product 0 = x_hi [w_hi | w_lo] (N doubled), product 1 = x_lo w_hi; deterministic mode walks r = 0 .. NPROD-1
        constexpr int NPROD = STK ? 2 : 3;
        const uint32_t idn_role = (STK && role == 0) ? IDESC_2N : IDESC_N, idq_role = (STK && role == 0) ? IDESC_2N : IDESC_Q;
        auto idn_of = [&](int r) -> uint32_t { return (STK && r == 0) ? IDESC_2N : IDESC_N; };
        auto idq_of = [&](int r) -> uint32_t { return (STK && r == 0) ? IDESC_2N : IDESC_Q; };
        int sa = 0, slot = 0, acc = 0;
        uint32_t pa = 0, pb = 0, pacc0 = 0, pacc1 = 0;
        bool b_ready = false;                            // resident weights: waited for once
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t bars0 = smem_u32(bars);
        const uint32_t a0 = smem_u32(a_buf), b0 = smem_u32(b_buf);
        auto desc = [](uint32_t lo) -> uint64_t { return ((uint64_t)DESC_HI << 32) | lo; };
        auto lo_of = [](uint32_t addr) -> uint32_t { return (addr >> 4) | 0x10000u; };
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x, wk.advance(p)) {
            const Item item = wk.cur;
            const uint32_t classes = __shfl_sync(0xffffffffu, tile_class_mask<NPH, UP2>(p, item, lane), 0);
            const bool mixed = (classes & (classes - 1)) != 0;
            const bool two = TWO_CLASS && __popc(classes) == 2;      // two regions: one accumulator buffer per region
            MBAR_WAIT_P(bars0 + 8 * (ACC_EMPTY + acc), (acc ? pacc1 : pacc0) ^ 1, 1);
            if (two) MBAR_WAIT_P(bars0 + 8 * (ACC_EMPTY + (acc ^ 1)), (acc ? pacc0 : pacc1) ^ 1, 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + (uint32_t)(acc * ACC_COLS);
            if (p.resident) slot = 0;
            const bool wait_b = !p.resident || !b_ready;
            if (two) {
                // ---------------- two-region tile: per chunk the two stages hold x * s_A and x * s_B; taps are row shifts
                const uint32_t d_other = tmem_u + (uint32_t)((acc ^ 1) * ACC_COLS);
#pragma unroll 1
                for (int kc = 0; kc < nchunks; ++kc) {
                    const int sb = sa ^ 1;
                    MBAR_WAIT_P(bars0 + 8 * (A_FULL + sa), pa, 2);
                    MBAR_WAIT_P(bars0 + 8 * (A_FULL + sb), sa == NSTAGE_A - 1 ? pa ^ 1 : pa, 2);
                    tc_fence_after();
                    const uint32_t apA = lo_of(a0 + sa * A_STAGE), apB = lo_of(a0 + sb * A_STAGE);
                    uint32_t roff = (uint32_t)ROWB >> 4;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        if (wait_b) {
                            MBAR_WAIT_P(bars0 + 8 * (B_FULL + slot), pb, 3);
                            tc_fence_after();
                        }
                        const uint32_t bp = lo_of(b0 + slot * B_SLOT);
                        if (elect_one()) {
                            if (!p.det) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, desc(apA + a_role + roff + 2 * k), desc(bp + w_role + 2 * k), idn_role, 1u);
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_other, desc(apB + a_role + roff + 2 * k), desc(bp + w_role + 2 * k), idn_role, 1u);
                            } else {
#pragma unroll 1
                                for (int r = 0; r < NPROD; ++r) {
                                    const uint32_t ao = r == 1 ? A_LO : 0u, wo2 = r == 2 ? W_LO : 0u;
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, desc(apA + ao + roff + 2 * k), desc(bp + wo2 + 2 * k), idn_of(r), 1u);
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_other, desc(apB + ao + roff + 2 * k), desc(bp + wo2 + 2 * k), idn_of(r), 1u);
                                }
                            }
                            if (!p.resident) umma_commit(bars0 + 8 * (B_EMPTY + slot));
                        }
                        slot += 2;
                        if (slot >= p.nslot_b) slot = 0, pb ^= 1;
                        roff += (uint32_t)(((tap % 3) == 2 ? (TWP - 2) : 1) * ROWB) >> 4;
                    }
                    if (elect_one()) {
                        umma_commit(bars0 + 8 * (A_EMPTY + sa));
                        umma_commit(bars0 + 8 * (A_EMPTY + sb));
                    }
                    pa ^= 1;                             // both stages consumed: same stage index, next phase
                }
#ifndef E4S_TCR_NO_FAST_ISSUE
            } else if (!mixed && !p.det) {
                // ---------------- region-pure tile, fast issue path.  ONE elected thread runs the whole tile: waits, then per chunk
                // 9 taps x KSTEPS MMAs fully unrolled, operand descriptors = two bases + compile-time offsets.  (The loop below,
                // walked by all lanes with a barrier check, an elect and a descriptor rebuild per tap, cost ~35 dependent
                // instructions per tap: the three issuing warps were 85 % busy on the 32->32 layer while the tensor pipe idled,
                // profiles/r2_stall_attribution_tcr_elect.log.)  State (stage / slot rings) is advanced by every lane afterwards.
                if (elect_one()) {
                    const uint32_t a_mine = a0 + (role == 1 ? A_PLANE : 0), b_mine = b0 + ((!STK && role == 2) ? B_SLOT : 0);
                    int sa2 = sa, slot2 = slot;
                    uint32_t pa2 = pa, pb2 = pb;
#pragma unroll 1
                    for (int kc = 0; kc < nchunks; ++kc) {
                        MBAR_WAIT_P(bars0 + 8 * (A_FULL + sa2), pa2, 2);
                        tc_fence_after();
                        const uint32_t ap = lo_of(a_mine + sa2 * A_STAGE);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (!((tmask >> tap) & 1u)) continue;
                            const uint32_t roff = (uint32_t)((1 + (tap / 3) * TWP + (tap % 3)) * ROWB) >> 4;   // halo pixel hp is operand row hp + 1
                            if (wait_b) {
                                MBAR_WAIT_P(bars0 + 8 * (B_FULL + slot2), pb2, 3);
                                tc_fence_after();
                            }
                            const uint32_t bp = lo_of(b_mine + slot2 * B_SLOT);
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, desc(ap + roff + 2 * k), desc(bp + 2 * k), idn_role, 1u);
                            if (!p.resident) umma_commit(bars0 + 8 * (B_EMPTY + slot2));
                            slot2 += 2;
                            if (slot2 >= p.nslot_b) slot2 = 0, pb2 ^= 1;
                        }
                        umma_commit(bars0 + 8 * (A_EMPTY + sa2));
                        if (++sa2 == NSTAGE_A) sa2 = 0, pa2 ^= 1;
                    }
                }
                __syncwarp();
                for (int kc = 0; kc < nchunks; ++kc) {               // the same ring arithmetic on every lane
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!((tmask >> tap) & 1u)) continue;
                        slot += 2;
                        if (slot >= p.nslot_b) slot = 0, pb ^= 1;
                    }
                    if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                }
#endif
            } else if (!mixed) {
                // ---------------- region-pure tile, generic loop (deterministic mode): operand staged once per chunk, taps are row shifts
#pragma unroll 1
                for (int kc = 0; kc < nchunks; ++kc) {
                    MBAR_WAIT_P(bars0 + 8 * (A_FULL + sa), pa, 2);
                    tc_fence_after();
                    const uint32_t ap = lo_of(a0 + sa * A_STAGE);
                    uint32_t roff = (uint32_t)ROWB >> 4;                   // tap (0,0): row shift 1
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!((tmask >> tap) & 1u)) {
                            roff += (uint32_t)(((tap % 3) == 2 ? (TWP - 2) : 1) * ROWB) >> 4;
                            continue;
                        }
                        if (wait_b) {
                            MBAR_WAIT_P(bars0 + 8 * (B_FULL + slot), pb, 3);
                            tc_fence_after();
                        }
                        const uint32_t bp = lo_of(b0 + slot * B_SLOT);
                        if (elect_one()) {
                            if (!p.det) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, desc(ap + a_role + roff + 2 * k), desc(bp + w_role + 2 * k), idn_role, 1u);
                            } else {
#pragma unroll 1
                                for (int r = 0; r < NPROD; ++r) {
                                    const uint32_t ao = r == 1 ? A_LO : 0u, wo2 = r == 2 ? W_LO : 0u;
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, desc(ap + ao + roff + 2 * k), desc(bp + wo2 + 2 * k), idn_of(r), 1u);
                                }
                            }
                            if (!p.resident) umma_commit(bars0 + 8 * (B_EMPTY + slot));
                        }
                        slot += 2;
                        if (slot >= p.nslot_b) slot = 0, pb ^= 1;
                        // next tap: +1 row, or to the start of the next halo row (+16 - 2) after dx = 2
                        roff += (uint32_t)(((tap % 3) == 2 ? (TWP - 2) : 1) * ROWB) >> 4;
                    }
                    if (elect_one()) umma_commit(bars0 + 8 * (A_EMPTY + sa));
                    if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                }
            } else {
                // ---------------- mixed tile: one operand stage per (tap, parity), MMAs of N = NTC
                for (int kc = 0; kc < nchunks; ++kc) {
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        if (wait_b) MBAR_WAIT_P(bars0 + 8 * (B_FULL + slot), pb, 3);
                        const uint32_t bp = lo_of(b0 + slot * B_SLOT);
#pragma unroll
                        for (int q = 0; q < NPH; ++q) {
                            MBAR_WAIT_P(bars0 + 8 * (A_FULL + sa), pa, 2);
                            tc_fence_after();
                            const uint32_t ap = lo_of(a0 + sa * A_STAGE);
                            const uint32_t boff = (uint32_t)(q * NTC * ROWB) >> 4;
                            const uint32_t dq = d_tmem + (uint32_t)(q * NTC);
                            if (elect_one()) {
                                if (!p.det) {
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) umma_bf16(dq, desc(ap + a_role + 2 * k), desc(bp + w_role + boff + 2 * k), idq_role, 1u);
                                } else {
#pragma unroll 1
                                    for (int r = 0; r < NPROD; ++r) {
                                        const uint32_t ao = r == 1 ? A_LO : 0u, wo2 = r == 2 ? W_LO : 0u;
#pragma unroll
                                        for (int k = 0; k < KSTEPS; ++k) umma_bf16(dq, desc(ap + ao + 2 * k), desc(bp + wo2 + boff + 2 * k), idq_of(r), 1u);
                                    }
                                }
                                umma_commit(bars0 + 8 * (A_EMPTY + sa));
                            }
                            if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                        }
                        if (!p.resident && elect_one()) umma_commit(bars0 + 8 * (B_EMPTY + slot));
                        slot += 2;
                        if (slot >= p.nslot_b) slot = 0, pb ^= 1;
                    }
                }
            }
            if (elect_one()) umma_commit(bars0 + 8 * (ACC_FULL + acc));
            if (acc) pacc1 ^= 1; else pacc0 ^= 1;
            acc ^= 1;
            if (two) {                                   // the tile used both buffers: hand both over, buffer order unchanged
                if (elect_one()) umma_commit(bars0 + 8 * (ACC_FULL + acc));
                if (acc) pacc1 ^= 1; else pacc0 ^= 1;
                acc ^= 1;
            }
            b_ready = true;
            __syncwarp();
        }
      }
    } else if (warp < W_EPI0) {
        // ===================================================================== activation transform (A producers), 256 threads
        const int t = threadIdx.x - 32 * W_XFORM0;       // 0..255
        const int c8 = t % CPR;
        const int pix0 = t / CPR;                        // 0..PPS-1
        constexpr int NSW_SHIFT = (160 + PPS - 1) / PPS; // sweeps over the 160 halo pixels (5 or 3)
        constexpr int NSW_ROWS = 128 / PPS;              // sweeps over the 128 operand rows (4 or 2)
        int sa = 0, xstage = 0;
        uint32_t pa = 0, px = 0;
        uint32_t chunk_ctr = 0;                          // selects the s_tab buffer
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x, wk.advance(p)) {
            const Item item = wk.cur;
            const uint32_t classes = tile_class_mask<NPH, UP2>(p, item, lane);
            const bool mixed = (classes & (classes - 1)) != 0;
            const float* xb = p.x + (int64_t)item.b * p.h * p.w * p.cin;
            const int y0 = item.ty * TH, x0 = item.tx * TW;
            const int nclass = TWO_CLASS && __popc(classes) == 2 ? 2 : 1;      // operand stages per chunk (uniform-style staging)
            if (!mixed || nclass == 2) {
                // region-pure tile: one stage per chunk scaled by the region's style.  Two-region tile of an up-sampling
                // layer: TWO such stages per chunk (one per region, same loaded activations); the MMA warps accumulate
                // them into the two accumulator buffers and the epilogue picks per (pixel, parity).
                const int cls2[2] = {__ffs(classes) - 1, 31 - __clz(classes)};
                for (int kc = 0; kc < nchunks; ++kc) {
                    const int ch = kc * KC + 8 * c8;
                    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v0[NSW_SHIFT], v1[NSW_SHIFT];
                    bool inb[NSW_SHIFT];
                    if (XS) MBAR_WAIT_P(smem_u32(&bars[XS_FULL + xstage]), px, 1);
                    const uint32_t xs = smem_u32(xs_buf + xstage * XS_STAGE);
#pragma unroll
                    for (int i = 0; i < NSW_SHIFT; ++i) {
                        const int hp = pix0 + PPS * i;
                        const int gy = y0 - 1 + (hp >> 4), gx = x0 - 1 + (hp & 15);
                        v0[i] = zero4, v1[i] = zero4;
                        inb[i] = hp < 160 && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
                        if (XS) {
                            if (hp < 160) {                       // the TMA box is already zero outside the image
                                v0[i] = lds_f4(xs + (uint32_t)(hp * 128 + 32 * c8));
                                v1[i] = lds_f4(xs + (uint32_t)(hp * 128 + 32 * c8 + 16));
                            }
                        } else if (inb[i]) {
                            const float* src = xb + ((int64_t)gy * p.w + gx) * p.cin + ch;
                            v0[i] = __ldg(reinterpret_cast<const float4*>(src));
                            v1[i] = __ldg(reinterpret_cast<const float4*>(src + 4));
                        }
                    }
#pragma unroll 1
                    for (int ci = 0; ci < nclass; ++ci) {
                    const float* sc = p.s ? p.s + ((int64_t)item.b * p.ncls + cls2[ci]) * p.cin : nullptr;
                    const float* sh = p.shift ? p.shift + ((int64_t)item.b * p.ncls + cls2[ci]) * p.cin : nullptr;
                    const float4 s0 = sc ? __ldg(reinterpret_cast<const float4*>(sc + ch)) : one4;
                    const float4 s1 = sc ? __ldg(reinterpret_cast<const float4*>(sc + ch + 4)) : one4;
                    const float4 t0 = sh ? __ldg(reinterpret_cast<const float4*>(sh + ch)) : zero4;
                    const float4 t1 = sh ? __ldg(reinterpret_cast<const float4*>(sh + ch + 4)) : zero4;
                    MBAR_WAIT_P(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1, 2);
                    const uint32_t hi_plane = smem_u32(a_buf) + (uint32_t)sa * A_STAGE, lo_plane = hi_plane + A_PLANE;
#pragma unroll
                    for (int i = 0; i < NSW_SHIFT; ++i) {
                        const int hp = pix0 + PPS * i;
                        if (hp >= 160) continue;
                        const int row = hp + 1;
                        // zero padding applies to the NORMALISED tensor: shift in-image pixels only
                        const uint32_t sx = KC == 64 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
                        const uint32_t off = (uint32_t)row * ROWB + (((uint32_t)c8 ^ sx) << 4);
                        scale_split_store8(v0[i], v1[i], s0, s1, t0, t1, sh && inb[i], hi_plane + off, lo_plane + off);
                    }
                    fence_proxy_async();
                    mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                    if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                    }
                    if (XS) {
                        // Release the raw tile only AFTER the stores that consume the loaded values: an arrive issued right
                        // behind the loads does not wait for them (no register dependence), and a load still in flight when
                        // the next TMA write lands returns the NEXT tile's first bytes (seen on hardware: ~25 % of the work
                        // items had the first 8 halo pixels of a chunk replaced).
                        mbar_arrive(smem_u32(&bars[XS_EMPTY + xstage]));
                        if (++xstage == NXS) xstage = 0, px ^= 1;
                    }
                }
            } else {
                // ---- mixed tile: per (tap, parity) operand tiles, every row scaled by the style of its own output pixel's region
                uint32_t rcls[NSW_ROWS];                 // 4 x 8-bit regions (one per parity) of each of my rows
#pragma unroll
                for (int i = 0; i < NSW_ROWS; ++i) {
                    const int r = pix0 + PPS * i;
                    const int iy = y0 + (r >> 4), ix = x0 + (r & 15);
                    rcls[i] = 0;
                    if ((r & 15) < TW && iy < p.h && ix < p.w) {
                        const uint8_t* lp = p.label + ((int64_t)item.b * ho + iy * MUL + (UP2 ? ((item.nt >> 1) & 1) : 0)) * wo + ix * MUL + (UP2 ? (item.nt & 1) : 0);
                        uint32_t c0 = min((int)lp[0], p.ncls - 1);
                        rcls[i] = c0;
                        if (NPH == 4)
                            rcls[i] = c0 | ((uint32_t)min((int)lp[1], p.ncls - 1) << 8) | ((uint32_t)min((int)lp[wo], p.ncls - 1) << 16) |
                                      ((uint32_t)min((int)lp[wo + 1], p.ncls - 1) << 24);
                    }
                }
                // my rows' halo origin (element offset inside the image; h * w * cin < 2^31 is checked on the host) and which of
                // the nine taps stay inside the image: computed once per tile - recomputing the 64-bit index and the bounds per
                // (tap, row) was ~110 of the 262 instructions a tap costs a thread, and the transform bounds mixed tiles
                int roff[NSW_ROWS];
                uint32_t vmask[NSW_ROWS];
#pragma unroll
                for (int i = 0; i < NSW_ROWS; ++i) {
                    const int r = pix0 + PPS * i;
                    const int gy0 = y0 - 1 + (r >> 4), gx0 = x0 - 1 + (r & 15);
                    roff[i] = (gy0 * p.w + gx0) * p.cin + 8 * c8;
                    uint32_t m = 0;
#pragma unroll
                    for (int tp = 0; tp < 9; ++tp) {
                        const int gy = gy0 + tp / 3, gx = gx0 + tp % 3;
                        m |= (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) ? (1u << tp) : 0u;
                    }
                    vmask[i] = m;
                }
                const float* sbase = p.s + (int64_t)item.b * p.ncls * p.cin;
                for (int kc = 0; kc < nchunks; ++kc) {
                    // styles of every region for this chunk -> shared table (double-buffered across chunks)
                    float* tab = s_tab + (chunk_ctr & 1) * p.ncls * KC;
                    ++chunk_ctr;
                    for (int e = t; e < p.ncls * (KC / 4); e += NUM_XFORM) {
                        const int c = e / (KC / 4), j = e - c * (KC / 4);
                        *reinterpret_cast<float4*>(tab + c * KC + 4 * j) =
                            __ldg(reinterpret_cast<const float4*>(sbase + (int64_t)c * p.cin + kc * KC + 4 * j));
                    }
                    xform_barrier();
                    const uint32_t tab_s = smem_u32(tab);
                    const int ch = kc * KC + 8 * c8;
                    if (XS) MBAR_WAIT_P(smem_u32(&bars[XS_FULL + xstage]), px, 1);
                    const uint32_t xs = smem_u32(xs_buf + xstage * XS_STAGE);
                    const float* xc = xb + kc * KC;
                    int toff = 0, thp = 0;               // (dy * w + dx) * cin and dy * 16 + dx of the running tap
#pragma unroll 1
                    for (int tap = 0, dx = 0; tap < 9; ++tap) {
                        float4 v0[NSW_ROWS], v1[NSW_ROWS];
#pragma unroll
                        for (int i = 0; i < NSW_ROWS; ++i) {
                            v0[i] = make_float4(0.f, 0.f, 0.f, 0.f), v1[i] = v0[i];
                            if (XS) {
                                const int hp = min(pix0 + PPS * i + thp, 159);
                                v0[i] = lds_f4(xs + (uint32_t)(hp * 128 + 32 * c8));
                                v1[i] = lds_f4(xs + (uint32_t)(hp * 128 + 32 * c8 + 16));
                            } else if ((vmask[i] >> tap) & 1u) {
                                const float* src = xc + (roff[i] + toff);
                                v0[i] = __ldg(reinterpret_cast<const float4*>(src));
                                v1[i] = __ldg(reinterpret_cast<const float4*>(src + 4));
                            }
                        }
                        if (++dx == 3) dx = 0, toff += (p.w - 2) * p.cin, thp += 14;
                        else toff += p.cin, thp += 1;
#pragma unroll
                        for (int q = 0; q < NPH; ++q) {
                            MBAR_WAIT_P(smem_u32(&bars[A_EMPTY + sa]), pa ^ 1, 2);
                            const uint32_t hi_plane = smem_u32(a_buf) + (uint32_t)sa * A_STAGE, lo_plane = hi_plane + A_PLANE;
#pragma unroll
                            for (int i = 0; i < NSW_ROWS; ++i) {
                                const int r = pix0 + PPS * i;
                                const uint32_t cl = (rcls[i] >> (8 * q)) & 0xffu;
                                const float4 s0 = lds_f4(tab_s + (cl * KC + 8 * c8) * 4), s1 = lds_f4(tab_s + (cl * KC + 8 * c8) * 4 + 16);
                                const uint32_t sx = KC == 64 ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
                                const uint32_t off = (uint32_t)r * ROWB + (((uint32_t)c8 ^ sx) << 4);
                                scale_split_store8(v0[i], v1[i], s0, s1, s0, s1, false, hi_plane + off, lo_plane + off);
                            }
                            fence_proxy_async();
                            mbar_arrive(smem_u32(&bars[A_FULL + sa]));
                            if (++sa == NSTAGE_A) sa = 0, pa ^= 1;
                        }
                    }
                    if (XS) {
                        mbar_arrive(smem_u32(&bars[XS_EMPTY + xstage]));
                        if (++xstage == NXS) xstage = 0, px ^= 1;
                    }
                }
            }
        }
    } else if (warp < W_XS) {
        // ===================================================================== epilogue: one pass per tile, region per (pixel, parity)
        const uint32_t quarter = (uint32_t)(warp & 3);
        const int m_row = quarter * 32 + lane;
        const int ty = m_row >> 4, tx = m_row & 15;
        int acc = 0;
        uint32_t pacc[2] = {0, 0};
        const float nw = (p.noise && p.noise_w) ? __ldg(p.noise_w) : 0.f;
        const bool strided = (NPH == 1 && p.out_stride == 2);
        const bool s2d = (NPH == 1 && p.out_stride == 4);        // store [B, H/2, W/2, (y & 1, x & 1, Cout)]: the next layer's stride 2 becomes 4 taps
        const int oh = strided ? (p.h >> 1) : ho, ow = strided ? (p.w >> 1) : wo;
        // Region and noise of my pixel do not depend on the accumulator, and the noise map is a streaming tensor (every
        // read is a DRAM miss): they are fetched one work item AHEAD, so that an epilogue-bound layer does not pay a DRAM
        // round trip per tile.
        auto fetch = [&](bool valid, const Item& i2, int (&c)[NPH], float (&z)[NPH]) {
#pragma unroll
            for (int q = 0; q < NPH; ++q) c[q] = 0, z[q] = 0.f;
            if (!valid) return;
            const int iy = i2.ty * TH + ty, ix = i2.tx * TW + tx;
            if (!(tx < TW && iy < p.h && ix < p.w && (!strided || ((iy | ix) & 1) == 0))) return;
#pragma unroll
            for (int q = 0; q < NPH; ++q) {
                const int qq = UP2 ? (i2.nt & 3) : q;
                const int oy = strided ? (iy >> 1) : iy * MUL + (qq >> 1), ox = strided ? (ix >> 1) : ix * MUL + (qq & 1);
                if (p.label) c[q] = min((int)p.label[((int64_t)i2.b * oh + oy) * ow + ox], p.ncls - 1);
                if (p.noise) z[q] = __ldg(p.noise + ((int64_t)(p.noise_b == 1 ? 0 : i2.b) * oh + oy) * ow + ox);
            }
        };
        int cls_next[NPH];
        float nz_next[NPH];
        Walk wk;
        wk.init(p, blockIdx.x, gridDim.x);
        fetch(blockIdx.x < p.items, wk.cur, cls_next, nz_next);
        for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
            const Item item = wk.cur;
            wk.advance(p);
            const int iy = item.ty * TH + ty, ix = item.tx * TW + tx;
            const bool mine = tx < TW && iy < p.h && ix < p.w && (!strided || ((iy | ix) & 1) == 0);
            const int n0 = (UP2 ? (item.nt >> 2) : item.nt) * NTC;
            int cls[NPH];
            float nz[NPH];
#pragma unroll
            for (int q = 0; q < NPH; ++q) cls[q] = cls_next[q], nz[q] = nw * nz_next[q];
            fetch(it + (int)gridDim.x < p.items, wk.cur, cls_next, nz_next);
            // two-region tile: region A (lowest index) accumulated in buffer `acc`, region B in the other one
            uint32_t classes = 1u;
            if (TWO_CLASS) classes = tile_class_mask<NPH, UP2>(p, item, lane);
            const bool two = TWO_CLASS && __popc(classes) == 2;
            const int cls_b = 31 - __clz(classes);
            MBAR_WAIT_P(smem_u32(&bars[ACC_FULL + acc]), pacc[acc], 1);
            pacc[acc] ^= 1;
            if (two) {
                MBAR_WAIT_P(smem_u32(&bars[ACC_FULL + (acc ^ 1)]), pacc[acc ^ 1], 1);
                pacc[acc ^ 1] ^= 1;
            }
            tc_fence_after();
            const uint32_t cbase = (uint32_t)(acc * ACC_COLS), cother = (uint32_t)((acc ^ 1) * ACC_COLS);
#pragma unroll
            for (int q = 0; q < NPH; ++q) {
                const int qq = UP2 ? (item.nt & 3) : q;
                const int oy = strided ? (iy >> 1) : iy * MUL + (qq >> 1), ox = strided ? (ix >> 1) : ix * MUL + (qq & 1);
                const float* dm = p.demod ? p.demod + ((int64_t)item.b * p.ncls + cls[q]) * p.cout + n0 : nullptr;
                float* dst = s2d ? p.y + (((((int64_t)item.b * (oh >> 1) + (oy >> 1)) * (ow >> 1) + (ox >> 1)) * 4 + ((oy & 1) * 2 + (ox & 1))) * p.cout + n0)
                                 : p.y + (((int64_t)item.b * oh + oy) * ow + ox) * p.cout + n0;
#pragma unroll 1
                for (int j = 0; j < NTC / 32; ++j) {
                    // demodulation and bias are fetched two channel groups ahead of their use (inside a plain load -> use
                    // -> store loop every L1 hit was exposed behind the previous store: ncu source view, 40 % of the
                    // epilogue warps' samples on the first FFMA of a group)
                    auto ld_dm = [&](int g) { return dm ? __ldg(reinterpret_cast<const float4*>(dm + j * 32 + 4 * g)) : make_float4(1.f, 1.f, 1.f, 1.f); };
                    auto ld_bv = [&](int g) { return p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j * 32 + 4 * g)) : make_float4(0.f, 0.f, 0.f, 0.f); };
                    float4 dmr[2] = {ld_dm(0), ld_dm(1)}, bvr[2] = {ld_bv(0), ld_bv(1)};
                    uint32_t r[32];
                    const uint32_t lanes = tmem_base + ((quarter * 32u) << 16) + (uint32_t)(q * NTC + j * 32);
                    {
                        PROF_BEGIN();
                        tmem_ld32(lanes + cbase, r);
                        tmem_zero32(lanes + cbase);      // the next tile's MMAs only accumulate
                        PROF_END(2);
                    }
                    if (STK) {                           // second half of the stacked accumulator: x_hi w_lo
                        uint32_t rl[32];
                        tmem_ld32(lanes + cbase + (uint32_t)N, rl);
                        tmem_zero32(lanes + cbase + (uint32_t)N);
#pragma unroll
                        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(rl[e]));
                    }
                    if (two) {
                        uint32_t r2[32];
                        tmem_ld32(lanes + cother, r2);
                        tmem_zero32(lanes + cother);
                        if (cls[q] == cls_b) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) r[e] = r2[e];
                        }
                    }
                    if (mine) {
                        // packed fp32 pairs (FFMA2 / FMUL2 / FADD2): t = acc * demod + (noise + bias), and for the fused leaky
                        // ReLU sqrt(2) * lrelu_0.2(t) = max(sqrt(2) t, 0.2 sqrt(2) t) with sqrt(2) folded into both operands -
                        // ~3 instructions per output instead of ~7 (the epilogue warps, one per scheduler, bound the
                        // small-K layers: profiles/r2_stall_attribution_tcr_elect.log)
                        const float k2 = 1.41421356237309515f;
                        const uint64_t S2 = pk2(k2, k2), P2 = pk2(0.2f, 0.2f);
                        const uint64_t ZQ = p.act == 1 ? pk2(k2 * nz[q], k2 * nz[q]) : pk2(nz[q], nz[q]);
                        float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const int co = j * 32 + 4 * g;
                            const float4 d = dmr[g & 1], bv = bvr[g & 1];
                            if (g + 2 < 8) dmr[g & 1] = ld_dm(g + 2), bvr[g & 1] = ld_bv(g + 2);
#ifdef E4S_TCR_SCALAR_EPI      // A/B switch of the diagnostic twin builds: the scalar arithmetic of round 1
                            float4 o;
                            {
                                const float zq = p.act == 1 ? nz[q] : nz[q];
                                o.x = __uint_as_float(r[4 * g + 0]) * d.x + zq + bv.x;
                                o.y = __uint_as_float(r[4 * g + 1]) * d.y + zq + bv.y;
                                o.z = __uint_as_float(r[4 * g + 2]) * d.z + zq + bv.z;
                                o.w = __uint_as_float(r[4 * g + 3]) * d.w + zq + bv.w;
                                if (p.act == 1) {
                                    o.x = lrelu_scaled(o.x, 0.2f, k2), o.y = lrelu_scaled(o.y, 0.2f, k2);
                                    o.z = lrelu_scaled(o.z, 0.2f, k2), o.w = lrelu_scaled(o.w, 0.2f, k2);
                                }
                            }
#else
                            uint64_t dd[2] = {pk2(d.x, d.y), pk2(d.z, d.w)}, zz[2] = {pk2(bv.x, bv.y), pk2(bv.z, bv.w)};
                            const uint64_t rr[2] = {pk2u(r[4 * g + 0], r[4 * g + 1]), pk2u(r[4 * g + 2], r[4 * g + 3])};
                            float o4[4];
#pragma unroll
                            for (int h2 = 0; h2 < 2; ++h2) {
                                if (p.act == 1) dd[h2] = mul2(dd[h2], S2), zz[h2] = fma2(zz[h2], S2, ZQ);
                                else zz[h2] = add2(zz[h2], ZQ);
                                const uint64_t t = fma2(rr[h2], dd[h2], zz[h2]);
                                upk2(t, o4[2 * h2], o4[2 * h2 + 1]);
                                if (p.act == 1) {
                                    float u0, u1;
                                    upk2(mul2(t, P2), u0, u1);
                                    o4[2 * h2] = fmaxf(o4[2 * h2], u0), o4[2 * h2 + 1] = fmaxf(o4[2 * h2 + 1], u1);
                                }
                            }
                            float4 o = make_float4(o4[0], o4[1], o4[2], o4[3]);
#endif
                            if (p.act == 2) {
                                const float4 sl = __ldg(reinterpret_cast<const float4*>(p.slope + n0 + co));
                                o.x = o.x > 0.f ? o.x : o.x * sl.x, o.y = o.y > 0.f ? o.y : o.y * sl.y;
                                o.z = o.z > 0.f ? o.z : o.z * sl.z, o.w = o.w > 0.f ? o.w : o.w * sl.w;
                            }
                            // 256-bit stores (one full 32-byte sector per lane): the odd group waits for its even neighbour
                            if (g & 1) st_global_v8(dst + co - 4, keep, o);
                            else keep = o;
                        }
                    }
                }
            }
            {
                PROF_BEGIN();
                tmem_wait_st();
                PROF_END(3);
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
            acc ^= 1;
            if (two) {
                mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
                acc ^= 1;
            }
        }
    }

#ifdef E4S_TCR_PROFILE
    if (prof_on) {
        int role = -1;
        if (lane == 0) role = warp == 0 ? 0 : warp == 1 ? 1 : warp == W_XFORM0 ? 2 : warp == W_EPI0 ? 3 : warp == W_XS ? 4 : -1;
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

static int num_sms() { return e4s_num_sms(); }

static long long* g_prof = nullptr;

template <int NTC, int KC, int NPH, bool XS = false, bool UP2 = false, bool STK = false>
static int launch(const void* w_hilo, Params p, cudaStream_t st) {
    p.prof = g_prof;
    p.det = e4s_get_deterministic();
    constexpr int N = NTC * NPH, ROWB = KC * 2;
    constexpr int A_BYTES = (((NSTAGE_A * 2 * A_ROWS * ROWB) + 1023) & ~1023) + (XS ? NXS * XS_STAGE : 0);
    constexpr int B_SLOT = N * ROWB;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return E4S_ERR_ARCH;
    CUtensorMap map;
    CUresult cr;
    if (UP2) {
        // weights [2][4][9][Cout][Cin] bf16 as a 5-D tensor (Cin, Cout, tap, parity, hl); one box = the (hi, lo) planes of ONE parity
        cuuint64_t dims[5] = {(cuuint64_t)p.cin, (cuuint64_t)p.cout, 9, 4, 2};
        cuuint64_t strides[4] = {(cuuint64_t)p.cin * 2, (cuuint64_t)p.cout * p.cin * 2, (cuuint64_t)9 * p.cout * p.cin * 2,
                                 (cuuint64_t)4 * 9 * p.cout * p.cin * 2};
        cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)NTC, 1, 1, 2};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(w_hilo), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // weights [2][NPH][9][Cout][Cin] bf16 as a 4-D tensor (Cin, Cout, tap, hl * NPH + parity)
        cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.cout, 9, (cuuint64_t)2 * NPH};
        cuuint64_t strides[3] = {(cuuint64_t)p.cin * 2, (cuuint64_t)p.cout * p.cin * 2, (cuuint64_t)9 * p.cout * p.cin * 2};
        cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)NTC, 1, (cuuint32_t)2 * NPH};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(w_hilo), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) return 700 + (int)cr;
    CUtensorMap xmap = map;          // placeholder when the activation is not TMA-staged
    if (XS) {
        // activation [B, H, W, C] fp32 as a 4-D tensor (C fastest); box = 32 channels x 16 columns x 10 rows of one sample
        cuuint64_t xd[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.batch};
        cuuint64_t xs[3] = {(cuuint64_t)p.cin * 4, (cuuint64_t)p.w * p.cin * 4, (cuuint64_t)p.h * p.w * p.cin * 4};
        cuuint32_t xb[4] = {32, 16, 10, 1};
        cuuint32_t xe[4] = {1, 1, 1, 1};
        cr = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.x), xd, xs, xb, xe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return 800 + (int)cr;
    }

    p.tiles_x = (int)e4s_ceil_div(p.w, TW);
    p.tiles_y = (int)e4s_ceil_div(p.h, TH);
    p.n_tiles = p.cout / NTC;
    const int64_t items = (int64_t)p.tiles_x * p.tiles_y * p.batch * p.n_tiles * (UP2 ? 4 : 1);
    if (items >= (1ll << 31)) return E4S_ERR_SHAPE;
    p.items = (int)items;
    const int planes = (p.cin / KC) * 18;
    const int tab_bytes = 2 * p.ncls * KC * 4;
    int max_slots = (SMEM_BUDGET - A_BYTES - tab_bytes - 1024) / B_SLOT;
    if (max_slots > 36) max_slots = 36;
    if (max_slots < 4) return E4S_ERR_SHAPE;
    p.resident = (!UP2 && p.n_tiles == 1 && planes <= max_slots) ? 1 : 0;     // parity items change weights with every item
    // streamed weights: a deep ring (TMA latency ~2000 cycles against 4 MMAs per slot pair on the small-N layers);
    // even, so that (hi, lo) pairs never straddle the wrap
    p.nslot_b = p.resident ? planes : (max_slots > 16 ? 16 : (max_slots & ~1));
    const size_t smem = 1024 + A_BYTES + (size_t)p.nslot_b * B_SLOT + tab_bytes + (size_t)(2 * NSTAGE_A + 4 + 2 * p.nslot_b + 2 * NXS) * 8 + 64;
    static E4sSmemOptIn optin;
    if (const int rc = e4s_smem_optin(optin, modconv3x3_tcr_kernel<NTC, KC, NPH, XS, UP2, STK>, smem)) return rc;
    const int grid = p.items < num_sms() ? p.items : num_sms();
    modconv3x3_tcr_kernel<NTC, KC, NPH, XS, UP2, STK><<<grid, XS ? NUM_THREADS_XS : NUM_THREADS, smem, st>>>(map, xmap, p);
    return e4s_launch_status();
}

static int pick_ntile(int channels, int widest, int64_t pixel_tiles) {
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
        if (pixel_tiles * (channels / c) >= num_sms() / 2) return c;
    }
    return last;
}

int dispatch(const void* w_hilo_bf16, Params p, int up, cudaStream_t st) {
    const int cin = p.cin, cout = p.cout;
    if (cin <= 64) {                 // small K: HBM-bound layers -> TMA-staged activations, 32-channel chunks
        if (!up) {
            if (cout % 128 == 0) return launch<128, 32, 1, true>(w_hilo_bf16, p, st);
            // N = 32: w_hi and w_lo stacked along N (STK) - two MMAs per (tap, K step) instead of three: c15 2.25 -> 1.95 ms
            // (the layer is bound by operand fetches from shared memory).  At N = 64 it is a wash (c13: 1.13 vs 1.15 ms:
            // that layer is bound by re-streaming 147 KB of weights per tile), so only E4S_B200_STK=1 stacks there; =0: never.
            const char* fstk = getenv("E4S_B200_STK");
            const int stk = fstk ? atoi(fstk) : -1;
            if (cout % 64 == 0) return stk == 1 ? launch<64, 32, 1, true, false, true>(w_hilo_bf16, p, st) : launch<64, 32, 1, true>(w_hilo_bf16, p, st);
            return stk != 0 ? launch<32, 32, 1, true, false, true>(w_hilo_bf16, p, st) : launch<32, 32, 1, true>(w_hilo_bf16, p, st);
        }
        if (cout % 64 == 0) return launch<64, 32, 4, true>(w_hilo_bf16, p, st);
        return launch<32, 32, 4, true>(w_hilo_bf16, p, st);
    }
    // N-tile width by occupancy (same rule and override as the gradient kernel, modconv_dgrad_tc.cu:pick_ntile): the
    // low-resolution 512-channel layers have a handful of pixel tiles, and wide N tiles left most SMs without work.
    const bool k64 = (cin % 64) == 0;
    const int64_t pixel_tiles = e4s_ceil_div(p.w, TW) * e4s_ceil_div(p.h, TH) * p.batch;
    if (!up) {
        if (k64) {
            int nt = pick_ntile(cout, 256, pixel_tiles);
            if (nt == 256) {
                const int rc = launch<256, 64, 1>(w_hilo_bf16, p, st);
                if (rc != E4S_ERR_SHAPE) return rc;          // too many regions for the style table next to 32-KB weight slots
                nt = 128;
            }
            if (nt == 128) return launch<128, 64, 1>(w_hilo_bf16, p, st);
            if (nt == 64) return launch<64, 64, 1>(w_hilo_bf16, p, st);
            return launch<32, 64, 1>(w_hilo_bf16, p, st);
        }
        if (pick_ntile(cout, 64, pixel_tiles) == 64) return launch<64, 32, 1>(w_hilo_bf16, p, st);
        return launch<32, 32, 1>(w_hilo_bf16, p, st);
    }
    // Up-sampling layer.  Parity work items (UP2, N tiles up to 256 wide) when the layer is wide enough for them to be no
    // worse on region-pure tiles (Cout >= 256: same MMA shape and operand re-use as four parities x 64 channels) - i.e. the
    // 512-channel layers up to 128x128, whose tiles mostly mix >= 3 regions; E4S_B200_UP2=1|0 forces / forbids it.
    if (k64) {
        int want = (cin >= 128 && cout >= 256) ? 1 : 0;
        if (const char* f = getenv("E4S_B200_UP2")) want = atoi(f) != 0;
        if (want) {
            const int ntu = pick_ntile(cout, 256, pixel_tiles * 4);
            int rc = E4S_ERR_SHAPE;
            if (ntu == 256) rc = launch<256, 64, 1, false, true>(w_hilo_bf16, p, st);
            if (rc == E4S_ERR_SHAPE && ntu >= 128 && cout % 128 == 0) rc = launch<128, 64, 1, false, true>(w_hilo_bf16, p, st);
            if (rc == E4S_ERR_SHAPE && ntu >= 64 && cout % 64 == 0) rc = launch<64, 64, 1, false, true>(w_hilo_bf16, p, st);
            if (rc == E4S_ERR_SHAPE) rc = launch<32, 64, 1, false, true>(w_hilo_bf16, p, st);
            return rc;
        }
    }
    const int nt = pick_ntile(cout, 64, pixel_tiles);
    if (k64) {
        if (nt == 64) {
            const int rc = launch<64, 64, 4>(w_hilo_bf16, p, st);
            if (rc != E4S_ERR_SHAPE) return rc;              // > 19 regions: the style table does not fit next to 32-KB weight slots
        }
        return launch<32, 64, 4>(w_hilo_bf16, p, st);
    }
    return launch<32, 32, 4>(w_hilo_bf16, p, st);
}

}  // namespace tcr

extern "C" int e4s_modconv3x3_tcr_fwd(const float* x, const void* w_hilo_bf16, const float* s, const float* demod,
                                      const uint8_t* label, const float* noise, const float* noise_w, const float* bias,
                                      float* y, int batch, int h, int w, int cin, int cout, int ncls, int up, int noise_b,
                                      int act, void* stream) {
    E4S_REQUIRE(x && w_hilo_bf16 && s && y, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && ncls > 0 && ncls <= 32, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0 && (cout % 32) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE((int64_t)h * w * cin < (1ll << 31), E4S_ERR_SHAPE);      // per-image element offsets are 32-bit in the transform
    E4S_REQUIRE(label || ncls == 1, E4S_ERR_ARG);
    E4S_REQUIRE(!noise || (noise_w && (noise_b == 1 || noise_b == batch)), E4S_ERR_ARG);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w_hilo_bf16) && e4s_aligned16(s) && e4s_aligned16(y) &&
                    (!demod || e4s_aligned16(demod)) && (!bias || e4s_aligned16(bias)),
                E4S_ERR_ALIGN);
    E4S_REQUIRE((reinterpret_cast<uintptr_t>(y) & 31) == 0, E4S_ERR_ALIGN);          // 256-bit stores
    tcr::Params p{x, s, demod, label, noise, noise_w, bias, y, batch, h, w, cin, cout, ncls, noise_b, act ? 1 : 0,
                  0, 0, 0, 0, 0, 0, nullptr, nullptr, 1};
    return tcr::dispatch(w_hilo_bf16, p, up, (cudaStream_t)stream);
}

extern "C" int e4s_conv3x3_tcr_f32(const float* x, const void* w_hilo_bf16, const float* scale, const float* shift,
                                   const float* prelu_slope, float* y, int batch, int h, int w, int cin, int cout,
                                   int out_stride, int tap_mask, void* stream) {
    E4S_REQUIRE(x && w_hilo_bf16 && y, E4S_ERR_ARG);
    E4S_REQUIRE(tap_mask >= 0 && tap_mask <= 0x1FF, E4S_ERR_ARG);
    E4S_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 32) == 0 && (cout % 32) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(out_stride == 1 || ((out_stride == 2 || out_stride == 4) && (h % 2) == 0 && (w % 2) == 0), E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w_hilo_bf16) && e4s_aligned16(y) && (!scale || e4s_aligned16(scale)) &&
                    (!shift || e4s_aligned16(shift)) && (!prelu_slope || e4s_aligned16(prelu_slope)),
                E4S_ERR_ALIGN);
    E4S_REQUIRE((reinterpret_cast<uintptr_t>(y) & 31) == 0, E4S_ERR_ALIGN);          // 256-bit stores
    tcr::Params p{x, scale, nullptr, nullptr, nullptr, nullptr, nullptr, y, batch, h, w, cin, cout, 1, 1, prelu_slope ? 2 : 0,
                  0, 0, 0, 0, 0, 0, shift, prelu_slope, out_stride};
    p.tap_mask = tap_mask;
    return tcr::dispatch(w_hilo_bf16, p, 0, (cudaStream_t)stream);
}

// Diagnostic: per-role stall attribution of CTA 0 of every following tcr launch ([5 roles][4] int64 cycle counters in
// device memory: role time, then the cycles it spent in its barrier waits).  nullptr switches it off (the default).
extern "C" int e4s_tcr_set_profile(long long* device_counters) {
#ifdef E4S_TCR_PROFILE
    tcr::g_prof = device_counters;
    return E4S_OK;
#else
    return device_counters ? E4S_ERR_ARG : E4S_OK;      // production build carries no counters
#endif
}
