// Small fp32 GEMMs of the path: the EqualLinear style modulations (reference model.py:135-169, :276; 26 per synthesis forward,
// [B * regions, 512] x [512, Cin]) and the per-region LocalMLPs (networks.py:15-39, :135-158; 12 x ([B, 1280] x [1280, 512],
// leaky ReLU 0.01, [B, 512] x [512, 13 * 512])).  Round 1 ran them on cuBLAS (38 sgemm launches per two steps); they are tiny
// (< 0.2 GFLOP) or weight-streaming (163 MB of MLP weights for a handful of rows), so a plain tiled SIMT kernel is at their
// roofline class and keeps library nodes out of the captured inversion step.
//
//   y[g, m, n] = act( sum_k x[g, m, k] * w[g, n, k] + bias[g, n] )            (TN: w in nn.Linear layout, K contiguous)
//   y[g, m, n] = sum_k x[g, m, k] * w[g, k, n]                                (NN: the input gradient of the above)
//
// fp32 FMA throughout (parity with the reference's fp32 linears).  Tile 32 rows x 64 columns x 32 deep, 256 threads, each
// 2 x 4 outputs; operands staged through shared memory with 128-bit loads.
#include "common.cuh"

namespace {

constexpr int BM = 32, BN = 64, BK = 32;

struct LinParams {
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    int m, n, k;
    long long xg, wg, bg, yg;      // batch strides in elements (0 = shared)
    float slope;                   // leaky-ReLU slope; 1 = no activation
};

template <bool NN>
__global__ void __launch_bounds__(256) linear_kernel(LinParams p) {
    __shared__ __align__(16) float xs[BK][BM + 4];
    __shared__ __align__(16) float ws[BK][BN + 4];
    const int g = blockIdx.z;
    const float* x = p.x + g * p.xg;
    const float* w = p.w + g * p.wg;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    float acc[2][4] = {};
    for (int k0 = 0; k0 < p.k; k0 += BK) {
        {   // x tile: 32 rows x 32 k; thread loads 4 consecutive k of one row
            const int r = t >> 3, kq = (t & 7) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < p.m && k0 + kq < p.k) v = *reinterpret_cast<const float4*>(x + (int64_t)(m0 + r) * p.k + k0 + kq);
            xs[kq][r] = v.x, xs[kq + 1][r] = v.y, xs[kq + 2][r] = v.z, xs[kq + 3][r] = v.w;
        }
        if (NN) {   // w[k][n]: 32 k x 64 n, n contiguous: two 128-bit loads per thread
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int e = t + 256 * i, kr = e >> 4, nq = (e & 15) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + kr < p.k && n0 + nq < p.n) v = *reinterpret_cast<const float4*>(w + (int64_t)(k0 + kr) * p.n + n0 + nq);
                *reinterpret_cast<float4*>(&ws[kr][nq]) = v;
            }
        } else {    // w[n][k]: 64 n x 32 k, k contiguous
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int e = t + 256 * i, nr = e >> 3, kq = (e & 7) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + nr < p.n && k0 + kq < p.k) v = *reinterpret_cast<const float4*>(w + (int64_t)(n0 + nr) * p.k + k0 + kq);
                ws[kq][nr] = v.x, ws[kq + 1][nr] = v.y, ws[kq + 2][nr] = v.z, ws[kq + 3][nr] = v.w;
            }
        }
        __syncthreads();
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
        if (p.slope != 1.f) {
            o.x = o.x > 0.f ? o.x : o.x * p.slope, o.y = o.y > 0.f ? o.y : o.y * p.slope;
            o.z = o.z > 0.f ? o.z : o.z * p.slope, o.w = o.w > 0.f ? o.w : o.w * p.slope;
        }
        *reinterpret_cast<float4*>(p.y + g * p.yg + (int64_t)m * p.n + n) = o;
    }
}

}  // namespace

extern "C" int e4s_linear_f32(const float* x, const float* w, const float* bias, float* y, int groups, int m, int n, int k,
                              long long x_gstride, long long w_gstride, long long bias_gstride, long long y_gstride, int w_is_kn,
                              float act_slope, void* stream) {
    E4S_REQUIRE(x && w && y && groups > 0 && m > 0 && n > 0 && k > 0, E4S_ERR_ARG);
    E4S_REQUIRE((n % 4) == 0 && (k % 4) == 0 && groups <= 65535, E4S_ERR_SHAPE);
    E4S_REQUIRE((x_gstride % 4) == 0 && (w_gstride % 4) == 0 && (bias_gstride % 4) == 0 && (y_gstride % 4) == 0, E4S_ERR_SHAPE);
    E4S_REQUIRE(e4s_aligned16(x) && e4s_aligned16(w) && e4s_aligned16(y) && (!bias || e4s_aligned16(bias)), E4S_ERR_ALIGN);
    E4S_REQUIRE(!w_is_kn || !bias, E4S_ERR_ARG);
    LinParams p{x, w, bias, y, m, n, k, x_gstride, w_gstride, bias_gstride, y_gstride, act_slope};
    dim3 grid((unsigned)e4s_ceil_div(n, BN), (unsigned)e4s_ceil_div(m, BM), (unsigned)groups);
    E4S_REQUIRE(grid.y <= 65535, E4S_ERR_SHAPE);
    if (w_is_kn) linear_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    else linear_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return e4s_launch_status();
}
