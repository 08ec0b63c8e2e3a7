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
    int ksplit, kper;              // K cut into ksplit slices of kper (multiple of BK) handled by different CTAs ...
    float* part;                   // ... whose raw partial sums go to part[ks][g][m][n]; linear_reduce_kernel finishes
    int groups;
};

// 32 x 32 output tile per CTA of 128 threads (2 rows x 4 columns each), K in chunks of 32 through shared memory.  These
// GEMMs are tiny, so a launch is bound by the serial chain of K chunks, not by FLOPs: the next chunk's global loads are
// issued into registers BEFORE the current chunk is multiplied (their L2 latency hides behind the 256 FMAs), and the tile
// is small so that even a [192, 512] x [512, 512] modulation spreads over 96 CTAs.
template <bool NN>
__global__ void __launch_bounds__(NT) linear_kernel(LinParams p) {
    __shared__ __align__(16) float xs[BK][BM + 4];
    __shared__ __align__(16) float ws[BK][BN + 4];
    const int g = blockIdx.z / p.ksplit, ks = blockIdx.z - g * p.ksplit;
    const float* x = p.x + g * p.xg;
    const float* w = p.w + g * p.wg;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = ks * p.kper, k_end = min(p.k, k_begin + p.kper);
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
    load(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        stash();
        __syncthreads();
        if (k0 + BK < k_end) load(k0 + BK);                // in flight while this chunk is multiplied
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
    if (p.ksplit > 1) {                                    // raw partial sums; bias / activation happen in linear_reduce_kernel
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = m0 + ty + 16 * i;
            if (m < p.m)
                *reinterpret_cast<float4*>(p.part + (((int64_t)ks * p.groups + g) * p.m + m) * p.n + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        return;
    }
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

// second stage of a K-split product: y = act(sum_ks part[ks] + bias)
__global__ void __launch_bounds__(256) linear_reduce_kernel(LinParams p) {
    const int64_t per_group = (int64_t)p.m * p.n, total4 = (int64_t)p.groups * per_group / 4;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total4; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t flat = 4 * e;
        const int g = (int)(flat / per_group);
        const int64_t rem = flat - (int64_t)g * per_group;
        const int n = (int)(rem % p.n);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ks = 0; ks < p.ksplit; ++ks) {
            const float4 v = *reinterpret_cast<const float4*>(p.part + (int64_t)ks * p.groups * per_group + flat);
            o.x += v.x, o.y += v.y, o.z += v.z, o.w += v.w;
        }
        if (p.bias) {
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + g * p.bg + n);
            o.x += bv.x, o.y += bv.y, o.z += bv.z, o.w += bv.w;
        }
        if (p.rsqrt_eps >= 0.f) o.x = rsqrtf(o.x + p.rsqrt_eps), o.y = rsqrtf(o.y + p.rsqrt_eps), o.z = rsqrtf(o.z + p.rsqrt_eps), o.w = rsqrtf(o.w + p.rsqrt_eps);
        if (p.slope != 1.f) {
            o.x = o.x > 0.f ? o.x : o.x * p.slope, o.y = o.y > 0.f ? o.y : o.y * p.slope;
            o.z = o.z > 0.f ? o.z : o.z * p.slope, o.w = o.w > 0.f ? o.w : o.w * p.slope;
        }
        *reinterpret_cast<float4*>(p.y + g * p.yg + rem) = o;
    }
}

// ---- many small products in one launch ------------------------------------------------------------------------------------
// One synthesis forward needs 26 style modulations ([B * regions, 512] x [512, Cin], different Cin, different latent) and 17
// demodulation products; launched one by one they are ~85 launches of 6-8 us each (latency, not work).  Here the problems of
// one kind travel BY VALUE in the kernel parameters (<= E4S_LINEAR_MULTI_MAX per launch) and a CTA finds its (problem, tile)
// from a prefix of tile counts: all modulations are one launch, all demodulations a second one.  Together the tiles fill the
// GPU, so K is not split (deterministic, no workspace).  x rows may be strided (ldx): a latent slice is read in place.
struct MultiParams {
    E4sLinearProblem prob[E4S_LINEAR_MULTI_MAX];
    int tile0[E4S_LINEAR_MULTI_MAX + 1];
    int nprob;
};

__global__ void __launch_bounds__(NT) linear_multi_kernel(const __grid_constant__ MultiParams mp) {
    __shared__ __align__(16) float xs[BK][BM + 4];
    __shared__ __align__(16) float ws[BK][BN + 4];
    int pi = 0;
    while (pi + 1 < mp.nprob && (int)blockIdx.x >= mp.tile0[pi + 1]) ++pi;
    const E4sLinearProblem& p = mp.prob[pi];
    const int tile = blockIdx.x - mp.tile0[pi], tiles_n = (p.n + BN - 1) / BN;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const float* __restrict__ x = p.x;
    const float* __restrict__ w = p.w;
    const int M = p.m, N = p.n, Kd = p.k, ldx = p.ldx;
    const bool demod = p.rsqrt_eps >= 0.f;
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    const int xr = t >> 3, xk = (t & 7) * 4;
    float acc[2][4] = {};
    float4 xv[2], wv[2];
    auto load = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = xr + 16 * i;
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M && k0 + xk < Kd) xv[i] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)(m0 + r) * ldx + k0 + xk));
            wv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int e = t + NT * i, nr = e >> 3, kq = (e & 7) * 4;
            if (n0 + nr < N && k0 + kq < Kd) wv[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + nr) * Kd + k0 + kq));
        }
    };
    load(0);
    for (int k0 = 0; k0 < Kd; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = xr + 16 * i;
            float4 v = xv[i];
            if (demod) v.x *= v.x, v.y *= v.y, v.z *= v.z, v.w *= v.w;
            xs[xk][r] = v.x, xs[xk + 1][r] = v.y, xs[xk + 2][r] = v.z, xs[xk + 3][r] = v.w;
            const int e = t + NT * i, nr = e >> 3, kq = (e & 7) * 4;
            ws[kq][nr] = wv[i].x, ws[kq + 1][nr] = wv[i].y, ws[kq + 2][nr] = wv[i].z, ws[kq + 3][nr] = wv[i].w;
        }
        __syncthreads();
        if (k0 + BK < Kd) load(k0 + BK);
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
    if (n >= N) return;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
        float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (demod) o.x = rsqrtf(o.x + p.rsqrt_eps), o.y = rsqrtf(o.y + p.rsqrt_eps), o.z = rsqrtf(o.z + p.rsqrt_eps), o.w = rsqrtf(o.w + p.rsqrt_eps);
        *reinterpret_cast<float4*>(p.y + (int64_t)m * N + n) = o;
    }
}

// These GEMMs are bound by the serial chain of K chunks in one CTA, not by FLOPs: when the output tiles alone cannot fill the
// GPU, K is cut into slices handled by different CTAs (deterministic: fixed slices, summed in a fixed order by a second
// tiny kernel).  Returns the slice count; 1 = no split, no workspace.
int choose_ksplit(int groups, int m, int n, int k) {
    const int64_t tiles = e4s_ceil_div(n, BN) * e4s_ceil_div(m, BM) * groups;
    const int sms = e4s_num_sms();
    if (k < 4 * BK) return 1;
    const int64_t chunks = e4s_ceil_div(k, BK);
    int64_t want = tiles >= 2 * sms ? 1 : e4s_ceil_div(3 * (int64_t)sms, tiles);       // fill the GPU ...
    const int64_t by_chain = e4s_ceil_div(chunks, 16);                                  // ... and keep a CTA's chain <= 16 chunks
    if (by_chain > want && tiles * by_chain <= 24 * (int64_t)sms) want = by_chain;
    const int64_t max_by_k = chunks / 2;                   // at least two chunks per slice
    if (want > max_by_k) want = max_by_k;
    if (want > 16) want = 16;
    return want < 2 ? 1 : (int)want;
}

template <bool NN>
int launch_linear(LinParams p, int groups, cudaStream_t st) {
    dim3 grid((unsigned)e4s_ceil_div(p.n, BN), (unsigned)e4s_ceil_div(p.m, BM), (unsigned)(groups * p.ksplit));
    if (grid.y > 65535 || grid.z > 65535) return E4S_ERR_SHAPE;
    linear_kernel<NN><<<grid, NT, 0, st>>>(p);
    if (p.ksplit > 1) {
        const int64_t total4 = (int64_t)groups * p.m * p.n / 4;
        int64_t blocks = e4s_ceil_div(total4, 256);
        const int64_t cap = (int64_t)e4s_num_sms() * 8;
        linear_reduce_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(p);
    }
    return e4s_launch_status();
}

}  // namespace

// Host-only: workspace (in floats) e4s_linear_f32 / e4s_demod_gemm_f32 need for a shape; 0 = none.
extern "C" long long e4s_linear_workspace_floats(int groups, int m, int n, int k) {
    if (groups <= 0 || m <= 0 || n <= 0 || k <= 0) return 0;
    const int ks = choose_ksplit(groups, m, n, k);
    return ks > 1 ? (long long)ks * groups * m * n : 0;
}

extern "C" int e4s_linear_f32(const float* x, const float* w, const float* bias, float* y, int groups, int m, int n, int k,
                              long long x_gstride, long long w_gstride, long long bias_gstride, long long y_gstride, int w_is_kn,
                              float act_slope, float* workspace, void* stream) {
    E4S_REQUIRE(x && w && y && groups > 0 && m > 0 && n > 0 && k > 0, E4S_ERR_ARG);
    E4S_REQUIRE((n % 4) == 0 && (k % 4) == 0 && groups <= 4095, E4S_ERR_SHAPE);
    E4S_REQUIRE((x_gstride % 4) == 0 && (w_gstride % 4) == 0 && (bias_gstride % 4) == 0 && (y_gstride % 4) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w) && e4s_aligned16(y) && (!bias || e4s_aligned16(bias)), E4S_ERR_ALIGN);
    E4S_REQUIRE(!w_is_kn || !bias, E4S_ERR_ARG);
    const int ks = choose_ksplit(groups, m, n, k);
    E4S_REQUIRE(ks == 1 || (workspace && e4s_aligned16(workspace) && y_gstride == (long long)m * n), E4S_ERR_ARG);
    const int kper = (int)(e4s_ceil_div(e4s_ceil_div(k, ks), BK) * BK);
    LinParams p{x, w, bias, y, m, n, k, x_gstride, w_gstride, bias_gstride, y_gstride, act_slope, -1.f, ks, kper, workspace, groups};
    return w_is_kn ? launch_linear<true>(p, groups, (cudaStream_t)stream) : launch_linear<false>(p, groups, (cudaStream_t)stream);
}

// Demodulation coefficients on the same tiled kernel: demod[r, o] = rsqrt(sum_i s[r, i]^2 * wsq[o, i] + eps)
// (model.py:279-281 in the shared-weight form).  s: [rows, cin], wsq: [cout, cin], demod: [rows, cout]; cin % 4 == 0, cout % 4 == 0.
extern "C" int e4s_demod_gemm_f32(const float* s, const float* wsq, float* demod, int rows, int cin, int cout, float eps,
                                  float* workspace, void* stream) {
    E4S_REQUIRE(s && wsq && demod && rows > 0 && cin > 0 && cout > 0 && eps >= 0.f, E4S_ERR_ARG);
    E4S_REQUIRE((cin % 4) == 0 && (cout % 4) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(s) && e4s_aligned16(wsq) && e4s_aligned16(demod), E4S_ERR_ALIGN);
    const int ks = choose_ksplit(1, rows, cout, cin);
    E4S_REQUIRE(ks == 1 || (workspace && e4s_aligned16(workspace)), E4S_ERR_ARG);
    const int kper = (int)(e4s_ceil_div(e4s_ceil_div(cin, ks), BK) * BK);
    LinParams p{s, wsq, nullptr, demod, rows, cout, cin, 0, 0, 0, (long long)rows * cout, 1.f, eps, ks, kper, workspace, 1};
    return launch_linear<false>(p, 1, (cudaStream_t)stream);
}

// Many independent small products in ONE launch per <= E4S_LINEAR_MULTI_MAX problems: y_i = x_i w_i^T + bias_i, or with
// rsqrt_eps >= 0 the demodulation form y_i = rsqrt((x_i * x_i) w_i^T + rsqrt_eps).  `problems` is a HOST array (copied into the
// kernel parameters; nothing is read from it after the call returns); the pointers inside are device pointers.
extern "C" int e4s_linear_multi_f32(const E4sLinearProblem* problems, int nproblems, void* stream) {
    E4S_REQUIRE(problems && nproblems > 0, E4S_ERR_ARG);
    for (int i = 0; i < nproblems; ++i) {
        const E4sLinearProblem& q = problems[i];
        E4S_REQUIRE(q.x && q.w && q.y && q.m > 0 && q.n > 0 && q.k > 0 && q.ldx >= q.k, E4S_ERR_ARG);
        E4S_REQUIRE((q.n % 4) == 0 && (q.k % 4) == 0 && (q.ldx % 4) == 0, E4S_ERR_SHAPE);
        E4S_REQUIRE(e4s_aligned16(q.x) && e4s_aligned16(q.w) && e4s_aligned16(q.y) && (!q.bias || e4s_aligned16(q.bias)), E4S_ERR_ALIGN);
    }
    for (int first = 0; first < nproblems; first += E4S_LINEAR_MULTI_MAX) {
        MultiParams mp;
        mp.nprob = nproblems - first < E4S_LINEAR_MULTI_MAX ? nproblems - first : E4S_LINEAR_MULTI_MAX;
        int64_t tiles = 0;
        for (int i = 0; i < mp.nprob; ++i) {
            mp.prob[i] = problems[first + i];
            mp.tile0[i] = (int)tiles;
            tiles += e4s_ceil_div(mp.prob[i].m, BM) * e4s_ceil_div(mp.prob[i].n, BN);
            E4S_REQUIRE(tiles < (1ll << 30), E4S_ERR_SHAPE);
        }
        mp.tile0[mp.nprob] = (int)tiles;
        linear_multi_kernel<<<(unsigned)tiles, NT, 0, (cudaStream_t)stream>>>(mp);
        const int rc = e4s_launch_status();
        if (rc) return rc;
    }
    return 0;
}
