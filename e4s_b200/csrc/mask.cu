// Mask / index operations (bit-exact) and layout shuffles.
//
// The reference carries region masks as float one-hot tensors [B, ncls, H, W] and multiplies conv
// outputs by them (src/models/stylegan2/model.py:391-398, 430-437).  Here a mask becomes a uint8
// label map [B, H, W] plus a nearest-resized label pyramid: 1 byte per pixel instead of 4*ncls,
// and the per-region sum turns into a per-pixel selection (bit-identical for one-hot masks,
// SURVEY.md App. A).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) onehot_to_label_kernel(const float* __restrict__ onehot,
                                                              uint8_t* __restrict__ label, int* __restrict__ flag,
                                                              int ncls, int64_t hw, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / hw, p = i - b * hw;
        const float* src = onehot + b * ncls * hw + p;
        int best = 0, ones = 0;
        bool clean = true;
        for (int c = 0; c < ncls; ++c) {
            float v = src[(int64_t)c * hw];
            if (v == 1.0f) {
                if (ones == 0) best = c;
                ++ones;
            } else if (v != 0.0f) {
                clean = false;
            }
        }
        if (!clean || ones != 1) *flag = 1;  // benign race: every writer stores 1
        label[i] = (uint8_t)best;
    }
}

__global__ void __launch_bounds__(256) label_to_onehot_kernel(const uint8_t* __restrict__ label,
                                                              float* __restrict__ onehot, int ncls, int64_t hw,
                                                              int64_t total) {
    // total = B * ncls * hw output elements; consecutive threads -> consecutive pixels of one class plane
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = i % hw;
        int64_t t = i / hw;
        int c = (int)(t % ncls);
        int64_t b = t / ncls;
        onehot[i] = (label[b * hw + p] == c) ? 1.0f : 0.0f;
    }
}

__global__ void __launch_bounds__(256) label_resize_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           int in_h, int in_w, int out_h, int out_w, float sy, float sx,
                                                           int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ox = (int)(i % out_w);
        int64_t t = i / out_w;
        int oy = (int)(t % out_h);
        int64_t b = t / out_h;
        // ATen nearest_neighbor_compute_source_index: min(floorf(dst * scale), in - 1)
        int iy = min((int)floorf(oy * sy), in_h - 1);
        int ix = min((int)floorf(ox * sx), in_w - 1);
        dst[i] = src[(b * in_h + iy) * in_w + ix];
    }
}

__global__ void __launch_bounds__(256) label_remap_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          const uint8_t* __restrict__ lut, int64_t n) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = s_lut[src[i]];
}

// Region mean: grid (channel chunks of 32, B).  Each warp walks pixels; lane = channel inside the
// 32-channel chunk (pixel-major: the 32 lanes read 128 contiguous bytes).  Per-class partial sums
// live in shared memory [ncls][32] per warp, reduced across warps at the end.  Area counts are
// integers; sum order is fixed -> deterministic.
constexpr int RM_WARPS = 8;
__global__ void __launch_bounds__(32 * RM_WARPS) region_mean_kernel(const float* __restrict__ feats,
                                                                     const uint8_t* __restrict__ label,
                                                                     float* __restrict__ out, int* __restrict__ area,
                                                                     int ncls, int hw, int c) {
    extern __shared__ float sm[];  // [RM_WARPS][ncls][32] sums, then [RM_WARPS][ncls] counts (as int)
    float* sums = sm;
    int* cnts = reinterpret_cast<int*>(sm + RM_WARPS * ncls * 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, ch = blockIdx.x * 32 + lane;
    for (int i = threadIdx.x; i < RM_WARPS * ncls * 32; i += blockDim.x) sums[i] = 0.f;
    for (int i = threadIdx.x; i < RM_WARPS * ncls; i += blockDim.x) cnts[i] = 0;
    __syncthreads();
    const float* fb = feats + (int64_t)b * hw * c;
    const uint8_t* lb = label + (int64_t)b * hw;
    float* my = sums + warp * ncls * 32;
    int* myc = cnts + warp * ncls;
    for (int p = warp; p < hw; p += RM_WARPS) {
        int cls = lb[p];
        if (cls >= ncls) continue;
        if (ch < c) my[cls * 32 + lane] += fb[(int64_t)p * c + ch];
        if (lane == 0) myc[cls] += 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncls * 32; i += blockDim.x) {
        int cls = i >> 5, l = i & 31;
        float s = 0.f;
        int n = 0;
        for (int w = 0; w < RM_WARPS; ++w) s += sums[w * ncls * 32 + i], n += cnts[w * ncls + cls];
        int chn = blockIdx.x * 32 + l;
        if (chn < c) out[((int64_t)b * ncls + cls) * c + chn] = n > 0 ? s / (float)n : 0.f;
        if (blockIdx.x == 0 && l == 0) area[b * ncls + cls] = n;
    }
}

// Tiled transpose between planar [B, C, HW] and pixel-major [B, HW, C].
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int rows, int cols) {
    // src: [batch][rows][cols] -> dst: [batch][cols][rows]
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const float* s = src + (int64_t)b * rows * cols;
    float* d = dst + (int64_t)b * rows * cols;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) t[j][tx] = s[(int64_t)r * cols + c];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, r = r0 + tx;
        if (r < rows && c < cols) d[(int64_t)c * rows + r] = t[tx][j];
    }
}

inline unsigned gs_grid(int64_t total) {
    int64_t want = e4s_ceil_div(total, 256), cap = (int64_t)E4S_NUM_SMS * 16;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace

extern "C" int e4s_onehot_to_label_u8(const float* onehot, uint8_t* label, int* flag, int batch, int ncls, int h, int w,
                                      void* stream) {
    E4S_REQUIRE(onehot && label && flag && batch > 0 && ncls > 0 && ncls <= 255 && h > 0 && w > 0, E4S_ERR_ARG);
    int64_t hw = (int64_t)h * w, total = hw * batch;
    onehot_to_label_kernel<<<gs_grid(total), 256, 0, (cudaStream_t)stream>>>(onehot, label, flag, ncls, hw, total);
    return e4s_launch_status();
}

extern "C" int e4s_label_to_onehot_f32(const uint8_t* label, float* onehot, int batch, int ncls, int h, int w,
                                       void* stream) {
    E4S_REQUIRE(onehot && label && batch > 0 && ncls > 0 && ncls <= 255 && h > 0 && w > 0, E4S_ERR_ARG);
    int64_t hw = (int64_t)h * w, total = hw * batch * ncls;
    label_to_onehot_kernel<<<gs_grid(total), 256, 0, (cudaStream_t)stream>>>(label, onehot, ncls, hw, total);
    return e4s_launch_status();
}

extern "C" int e4s_label_resize_nearest_u8(const uint8_t* src, uint8_t* dst, int batch, int in_h, int in_w, int out_h,
                                           int out_w, void* stream) {
    E4S_REQUIRE(src && dst && batch > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, E4S_ERR_ARG);
    float sy = (float)in_h / (float)out_h, sx = (float)in_w / (float)out_w;
    int64_t total = (int64_t)batch * out_h * out_w;
    label_resize_kernel<<<gs_grid(total), 256, 0, (cudaStream_t)stream>>>(src, dst, in_h, in_w, out_h, out_w, sy, sx,
                                                                           total);
    return e4s_launch_status();
}

extern "C" int e4s_label_remap_u8(const uint8_t* src, uint8_t* dst, const uint8_t* lut256, int64_t n, void* stream) {
    E4S_REQUIRE(src && dst && lut256 && n > 0, E4S_ERR_ARG);
    label_remap_kernel<<<gs_grid(n), 256, 0, (cudaStream_t)stream>>>(src, dst, lut256, n);
    return e4s_launch_status();
}

extern "C" int e4s_region_mean_f32(const float* feats, const uint8_t* label, float* out, int* area, int batch, int ncls,
                                   int h, int w, int c, void* stream) {
    E4S_REQUIRE(feats && label && out && area && batch > 0 && ncls > 0 && h > 0 && w > 0 && c > 0, E4S_ERR_ARG);
    E4S_REQUIRE(ncls <= 64, E4S_ERR_SHAPE);
    dim3 grid((unsigned)e4s_ceil_div(c, 32), batch);
    size_t smem = (size_t)RM_WARPS * ncls * 32 * sizeof(float) + (size_t)RM_WARPS * ncls * sizeof(int);
    region_mean_kernel<<<grid, 32 * RM_WARPS, smem, (cudaStream_t)stream>>>(feats, label, out, area, ncls, h * w, c);
    return e4s_launch_status();
}

extern "C" int e4s_planar_to_pixel_f32(const float* x, float* y, int batch, int c, int h, int w, void* stream) {
    E4S_REQUIRE(x && y && batch > 0 && c > 0 && h > 0 && w > 0, E4S_ERR_ARG);
    int hw = h * w;
    dim3 grid((unsigned)e4s_ceil_div(hw, 32), (unsigned)e4s_ceil_div(c, 32), batch);
    transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, c, hw);
    return e4s_launch_status();
}

extern "C" int e4s_pixel_to_planar_f32(const float* x, float* y, int batch, int c, int h, int w, void* stream) {
    E4S_REQUIRE(x && y && batch > 0 && c > 0 && h > 0 && w > 0, E4S_ERR_ARG);
    int hw = h * w;
    dim3 grid((unsigned)e4s_ceil_div(c, 32), (unsigned)e4s_ceil_div(hw, 32), batch);
    transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, hw, c);
    return e4s_launch_status();
}
