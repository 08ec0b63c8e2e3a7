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
