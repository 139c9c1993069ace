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
constexpr int kEmbBatchPerBlock = 4;   // samples per CTA (grid.y covers the batch)

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
    const int b_lo = blockIdx.y * kEmbBatchPerBlock, b_hi = min(B, b_lo + kEmbBatchPerBlock);
    for (int b = b_lo; b < b_hi; ++b) {
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
    const int b_lo = blockIdx.y * kEmbBatchPerBlock, b_hi = min(B, b_lo + kEmbBatchPerBlock);
    for (int b = b_lo; b < b_hi; ++b) {
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
// Block = 64 consecutive pixels x 4 channel groups. A thread owns one pixel and CPT = Cout/4 consecutive output channels:
// it reads its Cin*9 input taps once (L1-resident, shared with its neighbours), keeps the CPT accumulators in registers
// and streams the weights [tap][cout] from shared memory as warp-wide broadcasts (one LDS.128 feeds four FMAs per lane).
// The 64 x Cout result tile is staged in shared memory so that (a) it leaves as one contiguous, fully coalesced 64*Cout*4
// byte store (thread-per-pixel stores are 16 bytes every Cout*4 bytes) and (b) per-channel (sum, sum of squares) over the 64
// pixels come from conflict-free column reads: with `stat_parts` the kernel emits the GroupNorm statistics of its output
// as one slot per 64-pixel block, float2 [N][H*W/64][Cout] (needs H*W % 64 == 0), fixed summation order.
template <int CIN, int CPT>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out, float2* __restrict__ stat_parts, int N, int H, int W) {
    constexpr int K = CIN * 9;
    constexpr int COUT = 4 * CPT;
    constexpr int LD = COUT + 4;                 // padded tile row: conflict-free 16-byte stores by consecutive pixels
    extern __shared__ __align__(16) float stem_smem[];
    float* sw = stem_smem;                       // [K][COUT]
    float* tile = stem_smem + K * COUT;          // [64][LD]
    for (int i = threadIdx.x; i < K * COUT; i += blockDim.x) {
        const int co = i % COUT, k = i / COUT;
        sw[i] = w[(size_t)co * K + k];  // OIHW: k = ci*9 + ky*3 + kx
    }
    __syncthreads();
    const int HW = H * W;
    const int total = N * HW;
    const int cg = threadIdx.x >> 6;           // channel group 0..3 (warp-uniform)
    const int pl = threadIdx.x & 63;
    for (int pix0 = blockIdx.x * 64; pix0 < total; pix0 += gridDim.x * 64) {
        const int pix = pix0 + pl;
        const bool live = pix < total;
        float acc[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
        if (live) {
            const int n = pix / HW;
            const int rem = pix - n * HW;
            const int hy = rem / W, wx = rem - hy * W;
            float v[K];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float* xp = x + ((size_t)n * CIN + ci) * HW;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int yy = hy + ky - 1;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int xx = wx + kx - 1;
                        v[ci * 9 + ky * 3 + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xp + yy * W + xx) : 0.f;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = bias ? __ldg(bias + cg * CPT + j) : 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4* wrow = reinterpret_cast<const float4*>(sw + k * COUT + cg * CPT);
#pragma unroll
                for (int j = 0; j < CPT / 4; ++j) {
                    const float4 w4 = wrow[j];
                    acc[4 * j] += v[k] * w4.x; acc[4 * j + 1] += v[k] * w4.y;
                    acc[4 * j + 2] += v[k] * w4.z; acc[4 * j + 3] += v[k] * w4.w;
                }
            }
        }
        // stage the tile (rows of dead pixels are zero: they add nothing to the statistics and are not stored)
        float4* trow = reinterpret_cast<float4*>(tile + pl * LD + cg * CPT);
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) trow[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        __syncthreads();
        const int rows = min(64, total - pix0);
        float4* og = reinterpret_cast<float4*>(out + (size_t)pix0 * COUT);
        for (int i = threadIdx.x; i < rows * (COUT / 4); i += blockDim.x) {
            const int r = i / (COUT / 4), c4 = i - r * (COUT / 4);
            og[i] = *reinterpret_cast<const float4*>(tile + r * LD + 4 * c4);
        }
        if (stat_parts != nullptr && threadIdx.x < COUT) {
            float sx = 0.f, sy = 0.f;
#pragma unroll 8
            for (int r = 0; r < 64; ++r) {
                const float t = tile[r * LD + threadIdx.x];
                sx += t;
                sy += t * t;
            }
            stat_parts[(size_t)(pix0 >> 6) * COUT + threadIdx.x] = make_float2(sx, sy);   // (n * HW/64 + slot) == pix0 / 64
        }
        __syncthreads();
    }
}

template <int CIN>
static int launch_stem(const float* x, const float* w, const float* bias, float* out, float2* stat_parts, int N, int H, int W,
                       int Cout, cudaStream_t st) {
    long long blocks = ((long long)N * H * W + 63) / 64;
    const long long cap = (long long)num_sms() * 3;
    if (blocks > cap) blocks = cap;
    const size_t smem = ((size_t)CIN * 9 * Cout + 64 * (size_t)(Cout + 4)) * sizeof(float);
#define NDB_STEM_CASE(CPT)                                                                                              \
    {                                                                                                                   \
        static PerDeviceOnce attr;                                                                                      \
        if (attr.pending()) {                                                                                           \
            cudaFuncSetAttribute(stem_conv_kernel<CIN, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);  \
            attr.done();                                                                                                \
        }                                                                                                               \
        stem_conv_kernel<CIN, CPT><<<(int)blocks, 256, smem, st>>>(x, w, bias, out, stat_parts, N, H, W);               \
    }
    switch (Cout) {
        case 32: NDB_STEM_CASE(8) break;
        case 64: NDB_STEM_CASE(16) break;
        case 96: NDB_STEM_CASE(24) break;
        case 128: NDB_STEM_CASE(32) break;
        case 160: NDB_STEM_CASE(40) break;
        case 192: NDB_STEM_CASE(48) break;
        case 224: NDB_STEM_CASE(56) break;
        case 256: NDB_STEM_CASE(64) break;
        default: return NDB_ERR_UNSUPPORTED;
    }
#undef NDB_STEM_CASE
    return NDB_OK;
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

// fp32 -> (hi, lo) fp16 pair with hi + lo == x to ~2^-22 relative (the operands of the fp32-class mode: three tensor-core
// products a_hi w_hi + a_lo w_hi + a_hi w_lo stand in for one fp32 product). silu: apply SiLU (accurate expf) first.
__global__ void split16_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, long long numel,
                               int silu) {
    const long long n4 = numel >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
        float v[4] = {a.x, a.y, a.z, a.w};
        uint16_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (silu) v[k] = v[k] / (1.f + expf(-v[k]));
            const __half hh = __float2half_rn(v[k]);
            h[k] = __half_as_ushort(hh);
            l[k] = __half_as_ushort(__float2half_rn(v[k] - __half2float(hh)));
        }
        reinterpret_cast<uint2*>(hi)[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        reinterpret_cast<uint2*>(lo)[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

// Stem conv as a GEMM: out[n][h][w][tap*CIN + c] = x[n][c][h+ky-1][w+kx-1] (zero outside the image), tap = ky*3+kx, the
// 9*CIN <= 64 gathered values of a pixel zero-padded to ONE 64-element K block. x is NCHW fp32 (the model input); thread =
// pixel, consecutive threads read consecutive w (coalesced per (c, ky, kx)), each writes its 128-byte row as 8 x 16 bytes.
template <int CIN>
__global__ void stem_im2col_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, int N, int H, int W, int fmt) {
    const long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const int n = (int)(i / ((long long)W * H));
        const float* xn = x + (long long)n * CIN * H * W;
        uint32_t pk[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) pk[k] = 0u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
            const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float f = in ? __ldg(xn + ((long long)c * H + hh) * W + ww) : 0.f;
                const int k = tap * CIN + c;
                pk[k >> 1] |= (uint32_t)to16(f, fmt) << ((k & 1) * 16);
            }
        }
        uint4* o = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
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

// gray_invert reproduces what the reference CLI does to one-channel images between the sampler and the file: the float
// image is inverted (255 - v, scripts/sample.py:98-100), truncated to uint8 (:168) and inverted back (:170-171), i.e. the
// stored byte is 255 - trunc(255 - v) = ceil(v), not trunc(v).
__global__ void to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int C, int H, int W,
                                int gray_invert) {
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
        if (gray_invert) {
            const float a = 255.f - v;
            out[i] = (uint8_t)(255 - (int)(uint8_t)a);
        } else {
            out[i] = (uint8_t)v;  // truncation, like tensor.to(torch.uint8)
        }
    }
}

// %globaltimer stamp between two launches of a stream (scripts/timeline.py: in-graph, warm-cache duration of every kernel
// of a denoising step = difference of the stamps around it, minus the stamp kernel's own cost)
__global__ void timestamp_kernel(unsigned long long* dst) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *dst = t;
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
    const dim3 grid((E + 7) / 8, (B + kEmbBatchPerBlock - 1) / kEmbBatchPerBlock);
    embed_l1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, freqs, B, mc, w0, b0, h1_scratch);
    embed_l2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h1_scratch, y, B, E, w2, b2, class_emb, emb_out,
                                                           (uint16_t*)act_out, operand_format);
    return launch_status("ndb_embed");
}

extern "C" int ndb_stem_conv(const float* x_nchw, const float* w_oihw, const float* bias, float* out_nhwc, void* stat_parts,
                             int N, int Cin, int H, int W, int Cout, void* stream) {
    NDB_CHECK_ARG(x_nchw && w_oihw && out_nhwc);
    NDB_CHECK_ARG(Cin >= 1 && Cin <= 4);
    NDB_CHECK_ARG(stat_parts == nullptr || (H * W) % 64 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    float2* sp = reinterpret_cast<float2*>(stat_parts);
    int rc;
    switch (Cin) {
        case 1: rc = launch_stem<1>(x_nchw, w_oihw, bias, out_nhwc, sp, N, H, W, Cout, st); break;
        case 2: rc = launch_stem<2>(x_nchw, w_oihw, bias, out_nhwc, sp, N, H, W, Cout, st); break;
        case 3: rc = launch_stem<3>(x_nchw, w_oihw, bias, out_nhwc, sp, N, H, W, Cout, st); break;
        default: rc = launch_stem<4>(x_nchw, w_oihw, bias, out_nhwc, sp, N, H, W, Cout, st); break;
    }
    if (rc != NDB_OK) {
        std::fprintf(stderr, "[ndb200] ndb_stem_conv: unsupported width %d (model_channels must be 64, 128, 192 or 256)\n", Cout);
        return rc;
    }
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

extern "C" int ndb_stem_im2col(const float* x_nchw, void* out, int N, int Cin, int H, int W, int operand_format, void* stream) {
    NDB_CHECK_ARG(x_nchw && out && N > 0 && H > 0 && W > 0);
    NDB_CHECK_ARG(Cin >= 1 && Cin <= 7);
    const int grid = grid_for((long long)N * H * W, 128);
    cudaStream_t st = (cudaStream_t)stream;
    uint16_t* o = (uint16_t*)out;
    switch (Cin) {
        case 1: stem_im2col_kernel<1><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
        case 2: stem_im2col_kernel<2><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
        case 3: stem_im2col_kernel<3><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
        case 4: stem_im2col_kernel<4><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
        case 5: stem_im2col_kernel<5><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
        case 6: stem_im2col_kernel<6><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
        default: stem_im2col_kernel<7><<<grid, 128, 0, st>>>(x_nchw, o, N, H, W, operand_format); break;
    }
    return launch_status("ndb_stem_im2col");
}

extern "C" int ndb_im2col_s2(const float* x, void* out, int N, int H, int W, int C, int operand_format, void* stream) {
    NDB_CHECK_ARG(x && out && C % 4 == 0);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long work = 9LL * N * Ho * Wo * (C / 4);
    im2col_s2_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)out, N, H, W, C, Ho, Wo,
                                                                           operand_format);
    return launch_status("ndb_im2col_s2");
}

extern "C" int ndb_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int B, int C, int H, int W, int gray_invert,
                            void* stream) {
    NDB_CHECK_ARG(x_nchw && out_nhwc);
    NDB_CHECK_ARG(!gray_invert || C == 1);
    to_uint8_kernel<<<grid_for((long long)B * C * H * W, 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, out_nhwc, B, C, H, W,
                                                                                               gray_invert);
    return launch_status("ndb_to_uint8");
}

extern "C" int ndb_split16(const float* x, void* hi, void* lo, int64_t numel, int silu, void* stream) {
    NDB_CHECK_ARG(x && hi && lo && numel > 0 && numel % 4 == 0);
    split16_kernel<<<grid_for(numel / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)hi, (uint16_t*)lo, numel, silu);
    return launch_status("ndb_split16");
}

extern "C" int ndb_timestamp(uint64_t* dst, void* stream) {
    NDB_CHECK_ARG(dst);
    // plain launch (no programmatic serialization): starts only after the previous kernel has completed
    timestamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(dst));
    return launch_status("ndb_timestamp");
}
