// Shared helpers for the e4s_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/e4s_b200.h"

#define E4S_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define E4S_REQUIRE(cond, code) \
    do {                        \
        if (!(cond)) return (code); \
    } while (0)

// Launch-error check that never synchronises: reports bad launch configuration only.
static inline int e4s_launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear
        return (int)e;
    }
    return E4S_OK;
}

static inline bool e4s_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline int64_t e4s_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Streaming 128-bit accesses: data touched once should not pollute L1.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// 256-bit variant (sm_100+): one full 32-byte sector per lane and request.  p must be 32-byte aligned.
__device__ __forceinline__ void ld_stream_f8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float lrelu_scaled(float v, float alpha, float scale) {
    return (v > 0.f ? v : v * alpha) * scale;
}
