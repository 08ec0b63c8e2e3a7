// Small fp32 GEMMs of the path: the EqualLinear style modulations (reference model.py:135-169, :276; 26 per synthesis forward,
// [B * regions, 512] x [512, Cin]) and the per-region LocalMLPs (networks.py:15-39, :135-158; 12 x ([B, 1280] x [1280, 512],
// leaky ReLU 0.01, [B, 512] x [512, 13 * 512])).  Round 1 ran them on cuBLAS (38 sgemm launches per two steps); they are tiny
// (< 0.2 GFLOP) or weight-streaming (163 MB of MLP weights for a handful of rows), so a plain tiled SIMT kernel is at their
// roofline class and keeps library nodes out of the captured inversion step.
//
//   y[g, m, n] = act( sum_k x[g, m, k] * w[g, n, k] + bias[g, n] )            (TN: w in nn.Linear layout, K contiguous)
//   y[g, m, n] = sum_k x[g, m, k] * w[g, k, n]                                (NN: the input gradient of the above)
//
// fp32 FMA throughout (parity with the reference's fp32 linears).  Tile 32 rows x 32 columns x 32 deep, 128 threads, each
// 2 x 4 outputs; operands staged through shared memory with 128-bit loads, next chunk prefetched into registers.
#include "common.cuh"

namespace {

constexpr int BM = 32, BN = 32, BK = 32, NT = 128;

struct LinParams {
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    int m, n, k;
    long long xg, wg, bg, yg;      // batch strides in elements (0 = shared)
    float slope;                   // leaky-ReLU slope; 1 = no activation
    float rsqrt_eps;               // >= 0: demodulation form y = rsqrt(sum_k x^2 w + eps) (x squared on load); < 0: plain
};

// 32 x 32 output tile per CTA of 128 threads (2 rows x 4 columns each), K in chunks of 32 through shared memory.  These
// GEMMs are tiny, so a launch is bound by the serial chain of K chunks, not by FLOPs: the next chunk's global loads are
// issued into registers BEFORE the current chunk is multiplied (their L2 latency hides behind the 256 FMAs), and the tile
// is small so that even a [192, 512] x [512, 512] modulation spreads over 96 CTAs.
template <bool NN>
__global__ void __launch_bounds__(NT) linear_kernel(LinParams p) {
    __shared__ __align__(16) float xs[BK][BM + 4];
    __shared__ __align__(16) float ws[BK][BN + 4];
    const int g = blockIdx.z;
    const float* x = p.x + g * p.xg;
    const float* w = p.w + g * p.wg;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    float acc[2][4] = {};
    // per chunk every thread loads two float4 of x (rows r, r + 16; 4 consecutive k) and two of w
    const int xr = t >> 3, xk = (t & 7) * 4;              // x: row 0..15 (+16), k offset
    float4 xv[2], wv[2];
    auto load = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = xr + 16 * i;
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < p.m && k0 + xk < p.k) xv[i] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)(m0 + r) * p.k + k0 + xk));
            wv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NN) {       // w[k][n], n contiguous: 32 k x 32 n = 256 float4; thread e = t + 128 i -> k row e / 8, n quad e % 8
                const int e = t + NT * i, kr = e >> 3, nq = (e & 7) * 4;
                if (k0 + kr < p.k && n0 + nq < p.n) wv[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(k0 + kr) * p.n + n0 + nq));
            } else {        // w[n][k], k contiguous: 32 n x 32 k
                const int e = t + NT * i, nr = e >> 3, kq = (e & 7) * 4;
                if (n0 + nr < p.n && k0 + kq < p.k) wv[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + nr) * p.k + k0 + kq));
            }
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = xr + 16 * i;
            float4 v = xv[i];
            if (p.rsqrt_eps >= 0.f) v.x *= v.x, v.y *= v.y, v.z *= v.z, v.w *= v.w;
            xs[xk][r] = v.x, xs[xk + 1][r] = v.y, xs[xk + 2][r] = v.z, xs[xk + 3][r] = v.w;
            const int e = t + NT * i;
            if (NN) {
                *reinterpret_cast<float4*>(&ws[e >> 3][(e & 7) * 4]) = wv[i];
            } else {
                const int nr = e >> 3, kq = (e & 7) * 4;
                ws[kq][nr] = wv[i].x, ws[kq + 1][nr] = wv[i].y, ws[kq + 2][nr] = wv[i].z, ws[kq + 3][nr] = wv[i].w;
            }
        }
    };
    load(0);
    for (int k0 = 0; k0 < p.k; k0 += BK) {
        stash();
        __syncthreads();
        if (k0 + BK < p.k) load(k0 + BK);                  // in flight while this chunk is multiplied
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float a0 = xs[kk][ty], a1 = xs[kk][ty + 16];
            const float4 b = *reinterpret_cast<const float4*>(&ws[kk][tx * 4]);
            acc[0][0] = fmaf(a0, b.x, acc[0][0]), acc[0][1] = fmaf(a0, b.y, acc[0][1]), acc[0][2] = fmaf(a0, b.z, acc[0][2]), acc[0][3] = fmaf(a0, b.w, acc[0][3]);
            acc[1][0] = fmaf(a1, b.x, acc[1][0]), acc[1][1] = fmaf(a1, b.y, acc[1][1]), acc[1][2] = fmaf(a1, b.z, acc[1][2]), acc[1][3] = fmaf(a1, b.w, acc[1][3]);
        }
        __syncthreads();
    }
    const int n = n0 + tx * 4;
    if (n >= p.n) return;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + g * p.bg + n);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= p.m) continue;
        float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (p.rsqrt_eps >= 0.f) o.x = rsqrtf(o.x + p.rsqrt_eps), o.y = rsqrtf(o.y + p.rsqrt_eps), o.z = rsqrtf(o.z + p.rsqrt_eps), o.w = rsqrtf(o.w + p.rsqrt_eps);
        if (p.slope != 1.f) {
            o.x = o.x > 0.f ? o.x : o.x * p.slope, o.y = o.y > 0.f ? o.y : o.y * p.slope;
            o.z = o.z > 0.f ? o.z : o.z * p.slope, o.w = o.w > 0.f ? o.w : o.w * p.slope;
        }
        *reinterpret_cast<float4*>(p.y + g * p.yg + (int64_t)m * p.n + n) = o;
    }
}

// Skinny case (at most 16 rows: the LocalMLPs of a batch of faces, B <= 16; the one-face inversion loop has ONE row): the op
// is a weight stream - 163 MB for layer 2 of the twelve MLPs - so the kernel is organised around the weight matrix, not
// around output tiles.  W is [I][J] with J contiguous; a thread owns four consecutive j and walks a slab of i rows with
// 128-bit coalesced loads (8 in flight per thread, 64 threads per CTA so that even one class of one layer spreads over 26 CTAs), all M rows of x for that slab sit in shared memory and are read as broadcasts, the
// M x 4 partial sums stay in registers.  One slab: direct store with bias and activation (deterministic).  Several slabs (the
// input gradient through layer 2: I = 13 x 512): partial sums meet by red.global.add (y zeroed by a memset on the stream).
struct SkinnyParams {
    const float* x;     // [G][M][I]
    const float* w;     // [G][I][J]
    const float* bias;  // [G][J] or null (single slab only)
    float* y;           // [G][M][J]
    int m, i_total, j_total, slab;
    float slope;
};

template <int MR>
__global__ void __launch_bounds__(64) linear_skinny_kernel(SkinnyParams p) {
    extern __shared__ float xs[];                       // [MR][slab]
    const int g = blockIdx.z, i0 = blockIdx.y * p.slab;
    const int ni = min(p.slab, p.i_total - i0);
    const float* x = p.x + (int64_t)g * p.m * p.i_total;
    for (int e = threadIdx.x; e < MR * ni; e += 64) {
        const int m = e / ni, i = e - m * ni;
        xs[m * p.slab + i] = m < p.m ? x[(int64_t)m * p.i_total + i0 + i] : 0.f;
    }
    __syncthreads();
    const int j = (blockIdx.x * 64 + threadIdx.x) * 4;
    if (j >= p.j_total) return;
    const float* w = p.w + ((int64_t)g * p.i_total + i0) * p.j_total + j;
    float acc[MR][4];
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
    int i = 0;
    for (; i + 8 <= ni; i += 8) {
        float4 wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = ld_stream_f4(w + (int64_t)(i + u) * p.j_total);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const float a = xs[m * p.slab + i + u];
                acc[m][0] = fmaf(a, wv[u].x, acc[m][0]), acc[m][1] = fmaf(a, wv[u].y, acc[m][1]);
                acc[m][2] = fmaf(a, wv[u].z, acc[m][2]), acc[m][3] = fmaf(a, wv[u].w, acc[m][3]);
            }
        }
    }
    for (; i < ni; ++i) {
        const float4 wv = ld_stream_f4(w + (int64_t)i * p.j_total);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            const float a = xs[m * p.slab + i];
            acc[m][0] = fmaf(a, wv.x, acc[m][0]), acc[m][1] = fmaf(a, wv.y, acc[m][1]), acc[m][2] = fmaf(a, wv.z, acc[m][2]), acc[m][3] = fmaf(a, wv.w, acc[m][3]);
        }
    }
    float* y = p.y + (int64_t)g * p.m * p.j_total + j;
    if (gridDim.y == 1) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + (int64_t)g * p.j_total + j);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            if (m >= p.m) break;
            float4 o = make_float4(acc[m][0] + bv.x, acc[m][1] + bv.y, acc[m][2] + bv.z, acc[m][3] + bv.w);
            if (p.slope != 1.f) {
                o.x = o.x > 0.f ? o.x : o.x * p.slope, o.y = o.y > 0.f ? o.y : o.y * p.slope;
                o.z = o.z > 0.f ? o.z : o.z * p.slope, o.w = o.w > 0.f ? o.w : o.w * p.slope;
            }
            *reinterpret_cast<float4*>(y + (int64_t)m * p.j_total) = o;
        }
    } else {
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            if (m >= p.m) break;
            float* d = y + (int64_t)m * p.j_total;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(acc[m][0]), "f"(acc[m][1]), "f"(acc[m][2]), "f"(acc[m][3]) : "memory");
        }
    }
}

template <int MR>
int launch_skinny(const SkinnyParams& p, int groups, cudaStream_t st) {
    const int nslab = (int)e4s_ceil_div(p.i_total, p.slab);
    dim3 grid((unsigned)e4s_ceil_div(p.j_total, 256), (unsigned)nslab, (unsigned)groups);      // 64 threads x 4 columns per CTA
    if (nslab > 1 && cudaMemsetAsync(p.y, 0, sizeof(float) * (size_t)groups * p.m * p.j_total, st) != cudaSuccess) return (int)cudaGetLastError();
    const size_t smem = sizeof(float) * MR * p.slab;
    static E4sSmemOptIn optin;
    if (const int rc = e4s_smem_optin(optin, linear_skinny_kernel<MR>, smem)) return rc;
    linear_skinny_kernel<MR><<<grid, 64, smem, st>>>(p);
    return e4s_launch_status();
}

}  // namespace

// y[g, m, j] = act( sum_i x[g, m, i] * w[g, i, j] + bias[g, j] ) for m <= 16 rows (weight-streaming form, see above).
// w: [G, I, J] with J contiguous; j % 4 == 0.  The reduction is cut into slabs of 512 rows of w when that is needed to fill
// the GPU (I >= 2048): y is then zeroed by a memset enqueued on `stream` and accumulated with red.global.add (the last bits
// of y may differ between runs); bias / activation require the single-slab case.
extern "C" int e4s_linear_skinny_f32(const float* x, const float* w_ij, const float* bias, float* y, int groups, int m, int i, int j,
                                     float act_slope, void* stream) {
    E4S_REQUIRE(x && w_ij && y && groups > 0 && m > 0 && m <= 16 && i > 0 && j > 0 && groups <= 65535, E4S_ERR_ARG);
    E4S_REQUIRE((j % 4) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(w_ij) && e4s_aligned16(y) && (!bias || e4s_aligned16(bias)), E4S_ERR_ALIGN);
    const int slab = i >= 2048 ? 512 : i;
    E4S_REQUIRE(slab == i || (!bias && act_slope == 1.f), E4S_ERR_ARG);
    E4S_REQUIRE((size_t)16 * slab * sizeof(float) <= 96 * 1024, E4S_ERR_SHAPE);
    SkinnyParams p{x, w_ij, bias, y, m, i, j, slab, act_slope};
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 1) return launch_skinny<1>(p, groups, st);
    if (m <= 2) return launch_skinny<2>(p, groups, st);
    if (m <= 4) return launch_skinny<4>(p, groups, st);
    if (m <= 8) return launch_skinny<8>(p, groups, st);
    return launch_skinny<16>(p, groups, st);
}

extern "C" int e4s_linear_f32(const float* x, const float* w, const float* bias, float* y, int groups, int m, int n, int k,
                              long long x_gstride, long long w_gstride, long long bias_gstride, long long y_gstride, int w_is_kn,
                              float act_slope, void* stream) {
    E4S_REQUIRE(x && w && y && groups > 0 && m > 0 && n > 0 && k > 0, E4S_ERR_ARG);
    E4S_REQUIRE((n % 4) == 0 && (k % 4) == 0 && groups <= 65535, E4S_ERR_SHAPE);
    E4S_REQUIRE((x_gstride % 4) == 0 && (w_gstride % 4) == 0 && (bias_gstride % 4) == 0 && (y_gstride % 4) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w) && e4s_aligned16(y) && (!bias || e4s_aligned16(bias)), E4S_ERR_ALIGN);
    E4S_REQUIRE(!w_is_kn || !bias, E4S_ERR_ARG);
    LinParams p{x, w, bias, y, m, n, k, x_gstride, w_gstride, bias_gstride, y_gstride, act_slope, -1.f};
    dim3 grid((unsigned)e4s_ceil_div(n, BN), (unsigned)e4s_ceil_div(m, BM), (unsigned)groups);
    E4S_REQUIRE(grid.y <= 65535, E4S_ERR_SHAPE);
    if (w_is_kn) linear_kernel<true><<<grid, NT, 0, (cudaStream_t)stream>>>(p);
    else linear_kernel<false><<<grid, NT, 0, (cudaStream_t)stream>>>(p);
    return e4s_launch_status();
}

// Demodulation coefficients on the same tiled kernel: demod[r, o] = rsqrt(sum_i s[r, i]^2 * wsq[o, i] + eps)
// (model.py:279-281 in the shared-weight form).  s: [rows, cin], wsq: [cout, cin], demod: [rows, cout]; cin % 4 == 0, cout % 4 == 0.
extern "C" int e4s_demod_gemm_f32(const float* s, const float* wsq, float* demod, int rows, int cin, int cout, float eps, void* stream) {
    E4S_REQUIRE(s && wsq && demod && rows > 0 && cin > 0 && cout > 0 && eps >= 0.f, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 4) == 0 && (cout % 4) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(s) && e4s_aligned16(wsq) && e4s_aligned16(demod), E4S_ERR_ALIGN);
    LinParams p{s, wsq, nullptr, demod, rows, cout, cin, 0, 0, 0, 0, 1.f, eps};
    dim3 grid((unsigned)e4s_ceil_div(cout, BN), (unsigned)e4s_ceil_div(rows, BM), 1);
    E4S_REQUIRE(grid.y <= 65535, E4S_ERR_SHAPE);
    linear_kernel<false><<<grid, NT, 0, (cudaStream_t)stream>>>(p);
    return e4s_launch_status();
}
