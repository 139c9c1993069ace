// What limits the operand feed of one CTA? One CTA per SM streams {64 x 16-bit, 128 rows} TMA boxes (16 KB, SWIZZLE_128B)
// from an L2-resident buffer. Variants: W issuing warps per CTA (each with its own stage ring and barriers), one or two
// tensor maps used alternately, boxes of 128 or 64 rows.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate_probe tma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

constexpr int kTile = 16384;
constexpr int kStagesTotal = 12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                                             size_t ntiles, int iters, int nwarps, int nmaps, int box_rows, int shared_every, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[kStagesTotal];
    if (threadIdx.x == 0)
        for (int s = 0; s < kStagesTotal; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp < nwarps && (threadIdx.x & 31) == 0) {
        const int stages = kStagesTotal / nwarps;
        uint64_t* mybar = bar + warp * stages;
        uint8_t* mysm = smem + warp * stages * kTile;
        const int my_iters = iters / nwarps;
        const int per = 128 / box_rows;                       // boxes per stage
        for (int it = 0; it < my_iters + stages; ++it) {
            const int s = it % stages;
            if (it >= stages) {
                const uint32_t parity = ((it / stages) - 1) & 1;
                uint32_t done = 0;
                while (!done)
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                                 : "=r"(done) : "r"(smem_u32(&mybar[s])), "r"(parity) : "memory");
            }
            if (it < my_iters) {
                // shared_every = n: every n-th load of a warp fetches a tile that ALL CTAs fetch at that step (the weight tile of
                // a GEMM whose CTAs walk K together); the others are private to the CTA (its activation tile)
                const bool shared = shared_every > 0 && (it % shared_every) == 0;
                const size_t tile = shared ? ((size_t)warp * 977u + (size_t)it * 131u) % ntiles
                                           : ((size_t)(blockIdx.x * 4 + warp) * 7919u + (size_t)it * 131u) % ntiles;
                const int row = (int)tile * 128;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mybar[s])), "r"((uint32_t)kTile) : "memory");
                for (int b = 0; b < per; ++b) {
                    const CUtensorMap* m = (nmaps == 2 && ((it * per + b) & 1)) ? &m1 : &m0;
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(smem_u32(mysm + s * kTile + b * box_rows * 128)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(&mybar[s])), "r"(0), "r"(row + b * box_rows) : "memory");
                }
            }
        }
        if (out) atomicAdd(out, (unsigned long long)my_iters * kTile);
    }
    __syncthreads();
}

int main() {
    const size_t bytes = 48ull << 20;
    const size_t ntiles = bytes / kTile;
    uint8_t* src; cudaMalloc(&src, 2 * bytes); cudaMemset(src, 1, 2 * bytes);
    unsigned long long* out; cudaMalloc(&out, 8);
    const int smem = kStagesTotal * kTile;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    auto make = [&](CUtensorMap* m, void* base, int rows) {
        cuuint64_t dims[2] = {64, (cuuint64_t)(bytes / 128)}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    const int iters = 6000;
    for (int box_rows : {128}) {
        CUtensorMap m0, m1; make(&m0, src, box_rows); make(&m1, src + bytes, box_rows);
        for (int nwarps : {1, 2, 3, 4}) {
            for (int shared_every : {0, 3, 2, 1}) { const int nmaps = 1;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                float best = 1e9;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaMemset(out, 0, 8);
                    cudaEventRecord(e0);
                    probe<<<148, 128, smem>>>(m0, m1, ntiles, iters, nwarps, nmaps, box_rows, shared_every, out);
                    cudaEventRecord(e1);
                    cudaError_t e = cudaEventSynchronize(e1);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (rep && ms < best) best = ms;
                }
                unsigned long long total; cudaMemcpy(&total, out, 8, cudaMemcpyDeviceToHost);
                printf("box rows %3d, issuing warps %d, every %d-th load shared by all CTAs: %.3f ms, %.2f TB/s, %.1f GB/s per SM\n", box_rows, nwarps, shared_every, best,
                       total / best / 1e9, total / best / 1e6 / 148);
                fflush(stdout);
            }
        }
    }
    return 0;
}
