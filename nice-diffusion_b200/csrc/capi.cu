// Library-level entry points of the C-ABI (version / capability queries).
#include "common.cuh"

static char g_last_error[256] = "";

extern "C" int ndb_version(void) { return 1; }

extern "C" int ndb_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (major == 10 && minor == 0) ? 1 : 0;
}

extern "C" const char* ndb_last_error(void) {
    cudaError_t e = cudaPeekAtLastError();
    std::snprintf(g_last_error, sizeof(g_last_error), "%s", e == cudaSuccess ? "" : cudaGetErrorString(e));
    return g_last_error;
}
