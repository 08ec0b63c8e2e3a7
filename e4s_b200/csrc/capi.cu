// Library identity / capability entry points of the C ABI (include/e4s_b200.h).
#include "common.cuh"

extern "C" int e4s_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* e4s_build_arch(void) { return "sm_100a"; }

extern "C" int e4s_device_ok(void) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return E4S_ERR_ARCH;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return E4S_ERR_ARCH;
    }
    return major == 10 ? E4S_OK : E4S_ERR_ARCH;
}

// ---- run-to-run bit reproducibility of the tensor-core convolutions -------------------------------------------------------
// Default (0): three warps issue the three split-precision products concurrently; their MMAs accumulate in no fixed order,
// so results are reproducible to fp32 rounding (~2e-6 relative), not bit for bit.  1: one warp issues the products in a
// fixed order (csrc/modconv_tcr.cu, modconv_tch.cu) - every output bit is the same in every run, at a lower issue rate.
// Initial value: environment variable E4S_B200_DETERMINISTIC (1 / 0).
#include <atomic>
#include <cstdlib>
static std::atomic<int> g_deterministic{-1};

extern "C" int e4s_get_deterministic(void) {
    int v = g_deterministic.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("E4S_B200_DETERMINISTIC");
        v = (e && atoi(e) != 0) ? 1 : 0;
        g_deterministic.store(v, std::memory_order_relaxed);
    }
    return v;
}

extern "C" int e4s_set_deterministic(int on) {
    g_deterministic.store(on ? 1 : 0, std::memory_order_relaxed);
    return E4S_OK;
}
