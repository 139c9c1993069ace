// Does TMA multicast lift the L2 -> SM delivery cap (~12 TB/s measured on the conv GEMMs)?
// Every CTA keeps kStages bulk copies of kTile bytes in flight from an L2-resident buffer into its shared memory.
//   mode 0: unicast, every CTA streams its own tiles
//   mode 1: unicast, the CTAs of a cluster stream the SAME tiles (what the GEMM does today with shared operands)
//   mode 2: multicast, each CTA of a cluster fetches 1/csz of the shared tile and broadcasts it to all peers
// Reported: bytes landed in shared memory per second, summed over the CTAs.
// The copies are tiled tensor-map TMA ({64 x 16-bit, 128 rows} boxes, SWIZZLE_128B) like the GEMM operand loads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mcast_probe mcast_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

constexpr int kTile = 16384;
constexpr int kStages = 6;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap full, const __grid_constant__ CUtensorMap part, size_t ntiles, int iters, int mode, int csz, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[kStages];
    uint32_t rank, cid;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (csz > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        const size_t stream = (mode == 0) ? blockIdx.x : cid;                  // which tile sequence this CTA follows
        const uint32_t slice = (mode == 2) ? kTile / csz : kTile;
        const uint16_t mask = (uint16_t)((1u << csz) - 1);
        for (int it = 0; it < iters + kStages; ++it) {
            const int s = it % kStages;
            if (it >= kStages) {                                                 // the copy that used this stage must have landed
                const uint32_t parity = ((it / kStages) - 1) & 1;
                uint32_t done = 0;
                while (!done)
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                                 : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
            }
            if (it < iters) {
                const size_t tile = (stream * 7919u + (size_t)it * 131u) % ntiles;
                const int row = (int)tile * 128;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"((uint32_t)kTile) : "memory");
                if (mode == 2) {
                    const int rows = 128 / csz;
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                 ::"r"(smem_u32(smem + s * kTile + rank * slice)), "l"(reinterpret_cast<uint64_t>(&part)), "r"(smem_u32(&bar[s])), "r"(0), "r"(row + (int)rank * rows), "h"(mask) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(smem_u32(smem + s * kTile)), "l"(reinterpret_cast<uint64_t>(&full)), "r"(smem_u32(&bar[s])), "r"(0), "r"(row) : "memory");
                }
            }
        }
        if (out) atomicAdd(out, (unsigned long long)iters * kTile);
    }
    __syncthreads();
    if (csz > 1) {     // nobody leaves while a peer may still multicast into its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

int main() {
    const size_t bytes = 48ull << 20;                 // L2 resident
    const size_t ntiles = bytes / kTile;
    uint8_t* src;
    cudaMalloc(&src, bytes);
    cudaMemset(src, 1, bytes);
    unsigned long long* out;
    cudaMalloc(&out, 8);
    const int smem = kStages * kTile;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 4000;
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    auto make = [&](CUtensorMap* m, int rows) {
        cuuint64_t dims[2] = {64, (cuuint64_t)(bytes / 128)}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    CUtensorMap full; make(&full, 128);
    for (int csz : {1, 2, 4, 8}) {
        for (int mode = 0; mode < 3; ++mode) {
            if (csz == 1 && mode) continue;
            CUtensorMap part; make(&part, 128 / csz);
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
            int nclusters = 0;
            cfg.gridDim = dim3(296 / csz * csz);
            cudaOccupancyMaxActiveClusters(&nclusters, probe, &cfg);
            int grid = nclusters * csz; if (grid > 296) grid = 296 / csz * csz;
            cfg.gridDim = dim3(grid);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaMemset(out, 0, 8);
                cudaEventRecord(e0);
                cudaLaunchKernelEx(&cfg, probe, full, part, ntiles, iters, mode, csz, out);
                cudaEventRecord(e1);
                cudaError_t e = cudaEventSynchronize(e1);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            unsigned long long total; cudaMemcpy(&total, out, 8, cudaMemcpyDeviceToHost);
            printf("cluster %d mode %d: max active clusters %d, grid %d, %.3f ms, %.2f TB/s landed in smem (%.1f B/clk/SM at 1.9 GHz)\n",
                   csz, mode, nclusters, grid, best, total / best / 1e9, total / best / 1e9 * 1e3 / grid / 1.9);
            fflush(stdout);
        }
    }
    return 0;
}
