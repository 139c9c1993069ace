// Shared host-side helpers: status codes, launch checks, TMA tensor-map encoding.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/ndb200.h"

namespace ndb {

#define NDB_CHECK_ARG(cond)                                                                       \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            std::fprintf(stderr, "[ndb200] invalid argument: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            return NDB_ERR_INVALID;                                                               \
        }                                                                                         \
    } while (0)

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[ndb200] %s: CUDA error %s\n", what, cudaGetErrorString(e));
        return NDB_ERR_CUDA;
    }
    return NDB_OK;
}

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 16-bit element tensor map, 128-byte swizzle, zero fill out of bounds. dims/box innermost first;
// strides_bytes has rank-1 entries (dimension 0 is dense).
inline int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) {
        std::fprintf(stderr, "[ndb200] cuTensorMapEncodeTiled unavailable\n");
        return NDB_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base),
                    reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                    reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::fprintf(stderr, "[ndb200] cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u)\n", (int)r,
                     rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return NDB_ERR_CUDA;
    }
    return NDB_OK;
}

}  // namespace ndb
