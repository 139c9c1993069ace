// Small bandwidth / latency bound kernels around the GEMMs: timestep + class embedding front end,
// stem convolution (Cin = 1..4), standalone resampling, layout casts, strided-conv gather, uint8 export.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ndb {

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
    if (fmt == 0) {
        __half h = __float2half_rn(v);
        return *reinterpret_cast<uint16_t*>(&h);
    }
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------------
// embedding front end, two launches, one warp per output feature (weight row held in registers, looped over
// the batch so every weight is read from HBM once per forward):
//   L1: h1 = SiLU(W0 [cos(t f) | sin(t f)] + b0)      L2: emb = W2 h1 + b2 (+ class_emb[y]);  act = SiLU(emb)
// ------------------------------------------------------------------------------------------------------
constexpr int kEmbMaxMcPerLane = 16;   // model_channels <= 512
constexpr int kEmbMaxEPerLane = 64;    // 4 * model_channels <= 2048

__global__ void __launch_bounds__(256)
embed_l1_kernel(const float* __restrict__ t, const float* __restrict__ freqs, int B, int mc,
                const float* __restrict__ w0, const float* __restrict__ b0, float* __restrict__ h1) {
    const int E = 4 * mc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= E) return;
    const int half = mc / 2;
    float w[kEmbMaxMcPerLane], f[kEmbMaxMcPerLane];
#pragma unroll
    for (int i = 0; i < kEmbMaxMcPerLane; ++i) {
        const int k = lane + 32 * i;
        w[i] = (k < mc) ? __ldg(w0 + (size_t)j * mc + k) : 0.f;
        f[i] = (k < 2 * half) ? __ldg(freqs + (k < half ? k : k - half)) : 0.f;
    }
    const float bj = b0[j];
    for (int b = 0; b < B; ++b) {
        const float tb = t[b];
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kEmbMaxMcPerLane; ++i) {
            const int k = lane + 32 * i;
            if (k < 2 * half) {
                const float ang = tb * f[i];
                acc += w[i] * (k < half ? cosf(ang) : sinf(ang));   // zero pad for odd mc (model.py:521-522)
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const float v = acc + bj;
            h1[(size_t)b * E + j] = v / (1.f + expf(-v));
        }
    }
}

__global__ void __launch_bounds__(256)
embed_l2_kernel(const float* __restrict__ h1, const int64_t* __restrict__ y, int B, int E,
                const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ class_emb,
                float* __restrict__ emb_out, uint16_t* __restrict__ act_out, int fmt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= E) return;
    float w[kEmbMaxEPerLane];
#pragma unroll
    for (int i = 0; i < kEmbMaxEPerLane; ++i) {
        const int k = lane + 32 * i;
        w[i] = (k < E) ? __ldg(w2 + (size_t)j * E + k) : 0.f;
    }
    const float bj = b2[j];
    for (int b = 0; b < B; ++b) {
        const float* hb = h1 + (size_t)b * E;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kEmbMaxEPerLane; ++i) {
            const int k = lane + 32 * i;
            if (k < E) acc += w[i] * __ldg(hb + k);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = acc + bj;
            if (y != nullptr) v += __ldg(class_emb + (size_t)y[b] * E + j);
            if (emb_out != nullptr) emb_out[(size_t)b * E + j] = v;
            if (act_out != nullptr) act_out[(size_t)b * E + j] = to16(v / (1.f + expf(-v)), fmt);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// stem conv 3x3 pad 1, NCHW fp32 in (Cin <= 4) -> NHWC fp32 out
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out, int N, int Cin, int H, int W, int Cout) {
    extern __shared__ float sw[];  // [Cin*9][Cout] transposed weights, then bias[Cout]
    const int K = Cin * 9;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int co = i % Cout, k = i / Cout;
        sw[i] = w[(size_t)co * K + k];  // OIHW: k = ci*9 + ky*3 + kx
    }
    float* sb = sw + K * Cout;
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int Q = Cout >> 2;
    const long long total = (long long)N * H * W * Q;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % Q);
        long long pix = i / Q;
        const int wx = (int)(pix % W);
        const int hy = (int)((pix / W) % H);
        const int n = (int)(pix / ((long long)W * H));
        float4 acc = *reinterpret_cast<const float4*>(sb + 4 * q);
        for (int ci = 0; ci < Cin; ++ci) {
            const float* xp = x + ((size_t)n * Cin + ci) * H * W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = hy + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = wx + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    const float v = __ldg(xp + (size_t)yy * W + xx);
                    const float4 w4 = *reinterpret_cast<const float4*>(sw + (ci * 9 + ky * 3 + kx) * Cout + 4 * q);
                    acc.x += v * w4.x; acc.y += v * w4.y; acc.z += v * w4.z; acc.w += v * w4.w;
                }
            }
        }
        *reinterpret_cast<float4*>(out + pix * Cout + 4 * q) = acc;
    }
}

// ------------------------------------------------------------------------------------------------------
// standalone resample (fp32 NHWC), cast, strided-conv gather, uint8 export
// ------------------------------------------------------------------------------------------------------
__global__ void resample_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C,
                                    int mode) {
    const int Ho = mode == NDB_RESAMPLE_UP2 ? 2 * H : H / 2;
    const int Wo = mode == NDB_RESAMPLE_UP2 ? 2 * W : W / 2;
    const int V = C >> 2;
    const long long total = (long long)N * Ho * Wo * V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % V);
        long long p = i / V;
        const int wo = (int)(p % Wo);
        const int ho = (int)((p / Wo) % Ho);
        const int n = (int)(p / ((long long)Wo * Ho));
        float4 r;
        if (mode == NDB_RESAMPLE_UP2) {
            r = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + (ho >> 1)) * W + (wo >> 1)) * C + 4 * v));
        } else {
            const float* b = x + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + 4 * v;
            const float4 a = __ldg(reinterpret_cast<const float4*>(b));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b + C));
            const float4 c = __ldg(reinterpret_cast<const float4*>(b + (size_t)W * C));
            const float4 d = __ldg(reinterpret_cast<const float4*>(b + (size_t)W * C + C));
            r = make_float4((a.x + bb.x + c.x + d.x) * 0.25f, (a.y + bb.y + c.y + d.y) * 0.25f,
                            (a.z + bb.z + c.z + d.z) * 0.25f, (a.w + bb.w + c.w + d.w) * 0.25f);
        }
        *reinterpret_cast<float4*>(out + p * C + 4 * v) = r;
    }
}

__global__ void cast16_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, long long numel, int fmt) {
    const long long n4 = numel >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
        uint2 o;
        o.x = (uint32_t)to16(a.x, fmt) | ((uint32_t)to16(a.y, fmt) << 16);
        o.y = (uint32_t)to16(a.z, fmt) | ((uint32_t)to16(a.w, fmt) << 16);
        reinterpret_cast<uint2*>(out)[i] = o;
    }
}

// out[tap][n][ho][wo][c] = x[n][2ho+ky-1][2wo+kx-1][c] (zero outside), tap = ky*3+kx
__global__ void im2col_s2_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, int N, int H, int W, int C,
                                 int Ho, int Wo, int fmt) {
    const int V = C >> 2;
    const long long per_tap = (long long)N * Ho * Wo * V;
    const long long total = per_tap * 9;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(i / per_tap);
        long long r = i % per_tap;
        const int v = (int)(r % V);
        long long p = r / V;
        const int wo = (int)(p % Wo);
        const int ho = (int)((p / Wo) % Ho);
        const int n = (int)(p / ((long long)Wo * Ho));
        const int yy = 2 * ho + tap / 3 - 1, xx = 2 * wo + tap % 3 - 1;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
            a = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + yy) * W + xx) * C + 4 * v));
        uint2 o;
        o.x = (uint32_t)to16(a.x, fmt) | ((uint32_t)to16(a.y, fmt) << 16);
        o.y = (uint32_t)to16(a.z, fmt) | ((uint32_t)to16(a.w, fmt) << 16);
        *reinterpret_cast<uint2*>(out + ((size_t)tap * N * Ho * Wo + p) * C + 4 * v) = o;
    }
}

__global__ void to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int C, int H, int W) {
    const long long total = (long long)B * C * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int w = (int)(p % W);
        const int h = (int)((p / W) % H);
        const int b = (int)(p / ((long long)W * H));
        float v = (x[(((size_t)b * C + c) * H + h) * W + w] + 1.f) * 127.5f;
        v = fminf(fmaxf(v, 0.f), 255.f);
        out[i] = (uint8_t)v;  // truncation, like tensor.to(torch.uint8)
    }
}

static inline int grid_for(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_embed(const float* t, const int64_t* y, const float* freqs, int B, int mc, const float* w0,
                         const float* b0, const float* w2, const float* b2, const float* class_emb, float* h1_scratch,
                         float* emb_out, void* act_out, int operand_format, void* stream) {
    NDB_CHECK_ARG(t && freqs && w0 && b0 && w2 && b2 && h1_scratch && B > 0 && mc > 0);
    NDB_CHECK_ARG(mc % 4 == 0 && mc <= 32 * kEmbMaxMcPerLane);
    NDB_CHECK_ARG((y == nullptr) == (class_emb == nullptr));
    const int E = 4 * mc;
    const int grid = (E + 7) / 8;
    embed_l1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, freqs, B, mc, w0, b0, h1_scratch);
    embed_l2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h1_scratch, y, B, E, w2, b2, class_emb, emb_out,
                                                           (uint16_t*)act_out, operand_format);
    return launch_status("ndb_embed");
}

extern "C" int ndb_stem_conv(const float* x_nchw, const float* w_oihw, const float* bias, float* out_nhwc, int N,
                             int Cin, int H, int W, int Cout, void* stream) {
    NDB_CHECK_ARG(x_nchw && w_oihw && out_nhwc);
    NDB_CHECK_ARG(Cin >= 1 && Cin <= 16 && Cout % 4 == 0);
    const size_t smem = ((size_t)Cin * 9 * Cout + Cout) * sizeof(float);
    NDB_CHECK_ARG(smem <= 200 * 1024);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    const long long work = (long long)N * H * W * (Cout / 4);
    int grid = grid_for(work, 256);
    stem_conv_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x_nchw, w_oihw, bias, out_nhwc, N, Cin, H, W, Cout);
    return launch_status("ndb_stem_conv");
}

extern "C" int ndb_resample_f32(const float* x, float* out, int N, int H, int W, int C, int mode, void* stream) {
    NDB_CHECK_ARG(x && out && C % 4 == 0);
    NDB_CHECK_ARG(mode == NDB_RESAMPLE_UP2 || mode == NDB_RESAMPLE_DOWN2);
    if (mode == NDB_RESAMPLE_DOWN2) NDB_CHECK_ARG(H % 2 == 0 && W % 2 == 0);
    const long long work = (mode == NDB_RESAMPLE_UP2 ? 4LL : 1LL) * N * H * W * (C / 4) / (mode == NDB_RESAMPLE_UP2 ? 1 : 4);
    resample_f32_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(x, out, N, H, W, C, mode);
    return launch_status("ndb_resample_f32");
}

extern "C" int ndb_cast_nhwc16(const float* x, void* out, int64_t numel, int operand_format, void* stream) {
    NDB_CHECK_ARG(x && out && numel % 4 == 0);
    cast16_kernel<<<grid_for(numel / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)out, numel, operand_format);
    return launch_status("ndb_cast_nhwc16");
}

extern "C" int ndb_im2col_s2(const float* x, void* out, int N, int H, int W, int C, int operand_format, void* stream) {
    NDB_CHECK_ARG(x && out && C % 4 == 0);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long work = 9LL * N * Ho * Wo * (C / 4);
    im2col_s2_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)out, N, H, W, C, Ho, Wo,
                                                                           operand_format);
    return launch_status("ndb_im2col_s2");
}

extern "C" int ndb_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int B, int C, int H, int W, void* stream) {
    NDB_CHECK_ARG(x_nchw && out_nhwc);
    to_uint8_kernel<<<grid_for((long long)B * C * H * W, 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, out_nhwc, B, C, H, W);
    return launch_status("ndb_to_uint8");
}
