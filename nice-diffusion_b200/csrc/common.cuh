// Shared host-side helpers: status codes, launch checks, TMA tensor-map encoding.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// Per-device state (SM count, "kernel attributes set" flags) is indexed by the CURRENT device ordinal: function attributes
// and SM counts belong to a device, and one process may drive several (the engines are keyed by device index).
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

inline int num_sms() {
    static int n[kMaxDevices] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// `static PerDeviceOnce once; if (once.pending()) { ...; once.done(); }` -- runs a block once per device
struct PerDeviceOnce {
    bool flag[kMaxDevices] = {};
    bool pending() const { return !flag[current_device()]; }
    void done() { flag[current_device()] = true; }
};

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

// Tiled tensor map, zero fill out of bounds. dims/box innermost first; strides_bytes has rank-1 entries (dimension 0 is
// dense). elem_bytes 2 (uint16 payload) or 4 (float32); swizzle_bytes 128 or 64.
inline int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) {
        std::fprintf(stderr, "[ndb200] cuTensorMapEncodeTiled unavailable\n");
        return NDB_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                    const_cast<void*>(base), reinterpret_cast<const cuuint64_t*>(dims),
                    reinterpret_cast<const cuuint64_t*>(strides_bytes), reinterpret_cast<const cuuint32_t*>(box), estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::fprintf(stderr, "[ndb200] cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u)\n", (int)r,
                     rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return NDB_ERR_CUDA;
    }
    return NDB_OK;
}

// 16-bit elements, 128-byte swizzle: the tensor-core operand tiles
inline int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap(out, base, 2, 128, rank, dims, strides_bytes, box);
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Every kernel of this library is launched with programmaticStreamSerialization: its CTAs may start while the previous
// kernel of the stream is still draining and run their prologue (barrier init, TMEM allocation, tensor-map prefetch,
// index math); `pdl_wait()` then blocks until the previous grid has completed and its writes are visible. It must precede
// the first global-memory access. `pdl_launch_dependents()` lets the NEXT kernel start early in the same way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) on = std::getenv("NDB_DISABLE_PDL") ? 0 : 1;
    return on != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace ndb
