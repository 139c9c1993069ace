// Micro-benchmark: special-function throughput per SM for the softmax inner loop candidates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe scripts/mufu_probe.cu && /tmp/mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters) {
    float a[8];
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B400u + threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 5) asm volatile("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; ex2.approx.f16x2 t, t; mov.b32 %0, t; }" : "+f"(a[i]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
    float* out; cudaMalloc(&out, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 4000, threads = 512, blocks = sms * 4;
    probe<MODE><<<blocks, threads>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8 * per_op;
    printf("%-34s %8.3f ms  %8.1f Gop/s  = %6.2f results/clk/SM at %d MHz nominal\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<4>("ex2.approx.ftz.bf16x2", 2);
    run<2>("tanh.approx.f32", 1);
    run<3>("tanh.approx.f16x2", 2);
    run<5>("cvt.f16x2 + ex2.f16x2 (1 result)", 1);
    return 0;
}
