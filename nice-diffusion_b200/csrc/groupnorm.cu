// GroupNorm(32) for NHWC fp32 activations: statistics pass + fused apply pass.
//
// Bandwidth-bound kernels (roofline = HBM): the stats pass reads the tensor once (float4, fully coalesced
// along C) and leaves fp64 (sum, sum of squares) per (sample, group); the apply pass reads it a second
// time, folds mean / rstd / gamma / beta / AdaGN (1+scale), shift into one per-channel FMA, applies SiLU,
// optionally 2x nearest-upsamples or 2x2 average-pools, optionally concatenates two sources, and writes the
// 16-bit NHWC tensor-core operand (16-byte stores).
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ndb {

constexpr int kGroups = 32;
constexpr int kGnThreads = 256;
constexpr int kMaxVecPerThread = 4;  // supports C <= 4 * 4 * 256 = 4096

// ------------------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, int HW,
                float* __restrict__ stats0, float* __restrict__ stats1,
                double* __restrict__ partials, unsigned int* __restrict__ counters, int pix_per_block,
                int threads_per_row, int rows_per_block) {
    // Deterministic two-level reduction (no floating-point atomics): per-thread partials -> fixed-order block
    // reduction in shared memory -> per-block fp64 partials in global scratch -> the last block of each sample
    // (ticket counter) sums the per-block partials in block order and writes the final (sum, sumsq).
    extern __shared__ float sm[];  // [2][rows_per_block][C]
    __shared__ bool is_last;
    pdl_launch_dependents();
    pdl_wait();
    const int C = C0 + C1;
    const int V = C >> 2;
    const int n = blockIdx.y;
    const int chunks = gridDim.x;
    const int p_begin = blockIdx.x * pix_per_block;
    int p_end = p_begin + pix_per_block;
    if (p_end > HW) p_end = HW;
    float* red_s = sm;
    float* red_ss = sm + rows_per_block * C;

    const int r = threadIdx.x / threads_per_row;
    const int tv = threadIdx.x - r * threads_per_row;
    if (r < rows_per_block) {
#pragma unroll
        for (int j = 0; j < kMaxVecPerThread; ++j) {
            const int v = tv + j * threads_per_row;
            if (v < V) {
                const int c = v << 2;
                const float* src;
                int ldc, cc;
                if (c < C0) { src = x0; ldc = C0; cc = c; } else { src = x1; ldc = C1; cc = c - C0; }
                const float* base = src + ((size_t)n * HW) * ldc + cc;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = p_begin + r; p < p_end; p += rows_per_block) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * ldc));
                    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
                    ss.x += a.x * a.x; ss.y += a.y * a.y; ss.z += a.z * a.z; ss.w += a.w * a.w;
                }
                *reinterpret_cast<float4*>(red_s + r * C + c) = s;
                *reinterpret_cast<float4*>(red_ss + r * C + c) = ss;
            }
        }
    }
    __syncthreads();
    // per-channel block partial (fixed order over the row slots), fp64
    double* my_part = partials + ((size_t)n * chunks + blockIdx.x) * (2 * (size_t)C);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        const int which = i / C, c = i - which * C;
        const float* red = which ? red_ss : red_s;
        double acc = 0.0;
        for (int rr = 0; rr < rows_per_block; ++rr) acc += (double)red[rr * C + c];
        my_part[i] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&counters[n], 1u);
        is_last = (ticket == (unsigned int)(chunks - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
            const int which = i / C, c = i - which * C;
            double tot = 0.0;
            for (int ch = 0; ch < chunks; ++ch) tot += __ldcg(partials + ((size_t)n * chunks + ch) * (2 * (size_t)C) + i);
            // split across the two sources so that each tensor owns its [N][C_src][2] statistics
            if (c < C0) stats0[((size_t)n * C0 + c) * 2 + which] = (float)tot;
            else stats1[((size_t)n * C1 + (c - C0)) * 2 + which] = (float)tot;
        }
        if (threadIdx.x == 0) counters[n] = 0u;  // self-resetting: the workspace needs zeroing only once
    }
}

// ------------------------------------------------------------------------------------------------------
// apply
// ------------------------------------------------------------------------------------------------------
// SiLU(v) = v * sigmoid(v) = h + h * tanh(h) with h = v / 2: ONE special-function op (tanh.approx, |error| <= 2^-10.99
// relative -- the size of the 16-bit rounding applied to the result right after) and two FMA-pipe ops. The exact
// v / (1 + exp(-v)) form needs two MUFU ops per element (ex2 + rcp), which kept the XU pipe at > 50 % and this kernel
// issue-bound; the IEEE reciprocal before that cost a guarded slow-path call per element.
// takes h = v / 2 (the factor is folded into the per-channel affine coefficients)
__device__ __forceinline__ float silu_half(float h) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2h(float a, float b) {
    if (FMT == 0) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

struct GnApplyArgs {
    const float* x0; int C0;
    const float* x1; int C1;
    int N, H, W, Ho, Wo;
    const float2* parts0; const int* counts0; int smax0;
    const float2* parts1; const int* counts1; int smax1;
    const float* gamma; const float* beta;
    const float* scale; const float* shift; int ld_ss;
    const float* pre_add; int ld_pre;
    int silu, resample;
    uint16_t* out; uint16_t* raw16;
    uint16_t* out_lo; uint16_t* raw_lo;   // fp32-class mode: the fp16 remainders (SPLIT kernels)
    int cpg;         // channels per group; 0 = (C0 + C1) / 32. A slice of a wider normalised tensor passes the parent's value
    int x0_is16;     // x0 holds 16-bit values (operand_format) instead of fp32
    float* pool32;   // DOWN2 only: fp32 2x2 mean of the raw input (the x_resample operand of a down ResBlock)
    float eps;
    int pix_per_block;
    int chunk_ch;    // channels per block (blockIdx.z selects the chunk): whole groups, a multiple of 8; C when not chunked
};

// Block = one sample n (blockIdx.y), a range of output pixels (blockIdx.x) and a group-aligned range of channels
// (blockIdx.z). Large feature maps split over pixels only; the low-resolution levels (8x8, 16x16: a few hundred KB per
// sample) split over CHANNELS as well, so that the machine is filled with blocks whose statistics prologue covers only
// their own 64..144 channels -- with pixel-only splitting every block folded the slots of all C channels, which read more
// bytes than the tensor itself and made those launches pure latency (8 us for 9 MB).
// Thread (r, v) owns the 8-channel vector v of the chunk and pixels p_begin + r, + rows_per_block, ...  Order of business,
// arranged so that the dependent-load chains overlap:
//   0. before griddepcontrol.wait: slot counts (static per plan);
//   1. request the first four pixels of this thread's data and, one channel per thread, ALL statistics slots of that
//      channel (independent loads, 8 in flight) -> per-channel fp64 (sum, sumsq) in shared memory;
//   2. 8 lanes per group fold the channels of a group into (mean, rstd) -> smem;
//   3. one channel per thread folds mean / rstd / gamma / beta / AdaGN (1+scale), shift / pre-add into one FMA (a_c, b_c);
//   4. stream: y = SiLU(a x + b) (+ resample) -> 16-bit operand, 16-byte stores, four pixels in flight per thread.
// The summation order (slots in order per channel; channels sub, sub+8, ... per lane; xor-shuffles) is fixed: results are
// bit-reproducible and independent of the chunking.
// IN16: x0 is a 16-bit tensor (the fp16 `h` between the two convs of a ResBlock) instead of fp32.
// SPLIT (fp32-class mode): SiLU is evaluated with an accurate exponential and a true division, and every output value v is
// stored as the fp16 pair (hi = fp16(v), lo = fp16(v - hi)) in two tensors of the same layout.
__device__ __forceinline__ float silu_exact(float v) { return v / (1.f + expf(-v)); }
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int FMT, int RESAMPLE, bool SILU, bool RAW, bool IN16, bool SPLIT = false>
__global__ void __launch_bounds__(kGnThreads, 3)
gn_apply_kernel(const GnApplyArgs p) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ float s_mean[kGroups], s_rstd[kGroups];
    const int C = p.C0 + p.C1;
    const int n = blockIdx.y;
    const int cpg = p.cpg > 0 ? p.cpg : C / kGroups;
    const int c_lo = blockIdx.z * p.chunk_ch;
    const int Cb = (C - c_lo) < p.chunk_ch ? (C - c_lo) : p.chunk_ch;   // channels of this block (whole groups, % 8 == 0)
    const int Vb = Cb >> 3;
    const int ngroups = Cb / cpg;         // <= 32
    // shared memory: [2][Cb] doubles during the statistics phase, then (aliased, after a barrier) a[Cb], b[Cb], pre[Cb]
    double* dS = reinterpret_cast<double*>(sm_raw);
    double* dSS = dS + Cb;
    float* sa = reinterpret_cast<float*>(sm_raw);
    float* sb = sa + Cb;
    float* sp = sa + 2 * Cb;
    const int threads_per_row = Vb < kGnThreads ? Vb : kGnThreads;
    const int rows_per_block = kGnThreads / threads_per_row;
    const int r = threadIdx.x / threads_per_row;
    const int tv = threadIdx.x - r * threads_per_row;
    const bool worker = (r < rows_per_block) && (tv < Vb);
    const int npix = p.Ho * p.Wo;
    const int p_begin = blockIdx.x * p.pix_per_block;
    int p_end = p_begin + p.pix_per_block;
    if (p_end > npix) p_end = npix;
    constexpr int U = 4;

    // element pointer of channel c of sample n (in units of the source element: 4 bytes, or 2 when IN16)
    auto src_of = [&](int c, const float*& src, int& ldc) {
        int cc;
        if (c < p.C0) { src = p.x0; ldc = p.C0; cc = c; } else { src = p.x1; ldc = p.C1; cc = c - p.C0; }
        if (IN16) src = reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(src) + (size_t)n * p.H * p.W * ldc + cc);
        else src += (size_t)n * p.H * p.W * ldc + cc;
    };
    // eight consecutive channels of pixel `pix` -> two 16-byte registers (fp32) or one (16-bit)
    auto load8 = [&](const float* src, int ldc, int pix, uint4& a, uint4& b) {
        if (IN16) {
            a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(src) + (size_t)pix * ldc));
        } else {
            const float* px = src + (size_t)pix * ldc;
            a = __ldg(reinterpret_cast<const uint4*>(px));
            b = __ldg(reinterpret_cast<const uint4*>(px + 4));
        }
    };
    auto unpack8 = [&](const uint4& a, const uint4& b, float* xv) {
        if (IN16) {
            const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 f;
                if (FMT == 0) f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                xv[2 * i] = f.x; xv[2 * i + 1] = f.y;
            }
        } else {
            xv[0] = __uint_as_float(a.x); xv[1] = __uint_as_float(a.y); xv[2] = __uint_as_float(a.z); xv[3] = __uint_as_float(a.w);
            xv[4] = __uint_as_float(b.x); xv[5] = __uint_as_float(b.y); xv[6] = __uint_as_float(b.z); xv[7] = __uint_as_float(b.w);
        }
    };
    auto in_pixel = [&](int pu) {
        if (RESAMPLE == NDB_RESAMPLE_UP2) {
            const int ho = pu / p.Wo, wo = pu - ho * p.Wo;
            return (ho >> 1) * p.W + (wo >> 1);
        }
        return pu;
    };

    // eight results -> one 16-byte store (two in SPLIT mode)
    auto store8 = [&](uint16_t* dst, uint16_t* dst_lo, const float* y) {
        if (SPLIT) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split2(y[2 * k], y[2 * k + 1], h[k], l[k]);
            *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst_lo) = make_uint4(l[0], l[1], l[2], l[3]);
        } else {
            *reinterpret_cast<uint4*>(dst) =
                make_uint4(pack2h<FMT>(y[0], y[1]), pack2h<FMT>(y[2], y[3]), pack2h<FMT>(y[4], y[5]), pack2h<FMT>(y[6], y[7]));
        }
    };
    auto act = [&](float t) { return !SILU ? t : (SPLIT ? silu_exact(t) : silu_half(t)); };

    // ---- 0. static data (written when the plan was built, never by a kernel of the step): legal before the wait ----
    const int cnt0 = __ldg(p.counts0 + n);
    const int cnt1 = p.parts1 ? __ldg(p.counts1 + n) : 0;
    pdl_launch_dependents();
    pdl_wait();
    // ---- 1a. first trip's data ----
    uint4 lo[U], hi[U];
    if (RESAMPLE != NDB_RESAMPLE_DOWN2 && worker) {
        const float* src; int ldc;
        src_of(c_lo + (tv << 3), src, ldc);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pu = p_begin + r + u * rows_per_block;
            if (pu < p_end) load8(src, ldc, in_pixel(pu), lo[u], hi[u]);
        }
    }
    // ---- 1b. per-channel statistics: the slots of a channel are fetched eight at a time (independent loads) and added in
    //          slot order. With `pre_add` e_c (non-adaptive ResBlock: h += emb before the norm) the sums of (x + e_c) follow
    //          analytically: S' = S + P e, SS' = SS + 2 e S + P e^2. ----
    const double P = (double)p.H * p.W;
    for (int cl = threadIdx.x; cl < Cb; cl += kGnThreads) {
        const int c = c_lo + cl;
        const bool first = c < p.C0;
        const float2* src = first ? p.parts0 + (size_t)n * p.smax0 * p.C0 + c
                                  : p.parts1 + (size_t)n * p.smax1 * p.C1 + (c - p.C0);
        const size_t ldp = first ? (size_t)p.C0 : (size_t)p.C1;
        const int cnt = first ? cnt0 : cnt1;
        double s1 = 0.0, s2 = 0.0;
        for (int sl = 0; sl < cnt; sl += 8) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (sl + u < cnt) ? __ldg(src + (size_t)(sl + u) * ldp) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { s1 += (double)v[u].x; s2 += (double)v[u].y; }
        }
        if (p.pre_add != nullptr) {
            const double e = (double)p.pre_add[(size_t)n * p.ld_pre + c];
            s2 += 2.0 * e * s1 + P * e * e;
            s1 += P * e;
        }
        dS[cl] = s1;
        dSS[cl] = s2;
    }
    __syncthreads();
    // ---- 2. groups ----
    {
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        double S = 0.0, SS = 0.0;
        if (g < ngroups) {
            for (int k = sub; k < cpg; k += 8) { S += dS[g * cpg + k]; SS += dSS[g * cpg + k]; }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            S += __shfl_xor_sync(0xffffffffu, S, o);
            SS += __shfl_xor_sync(0xffffffffu, SS, o);
        }
        if (sub == 0 && g < ngroups) {
            const double inv_cnt = 1.0 / ((double)cpg * P);
            const double mean = S * inv_cnt;
            double var = SS * inv_cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean[g] = (float)mean;
            s_rstd[g] = rsqrtf((float)var + p.eps);
        }
    }
    __syncthreads();      // also: every read of dS / dSS is done, the region is re-used for a / b / pre below
    // ---- 3. per-channel affine ----
    for (int cl = threadIdx.x; cl < Cb; cl += kGnThreads) {
        const int c = c_lo + cl;
        const int g = cl / cpg;
        float ak = s_rstd[g] * __ldg(p.gamma + c);
        float bk = __ldg(p.beta + c) - s_mean[g] * ak;
        const float pr = p.pre_add ? __ldg(p.pre_add + (size_t)n * p.ld_pre + c) : 0.f;
        bk += ak * pr;
        if (p.scale != nullptr) {
            const float s1 = 1.f + __ldg(p.scale + (size_t)n * p.ld_ss + c);
            ak *= s1;
            bk = bk * s1 + __ldg(p.shift + (size_t)n * p.ld_ss + c);
        }
        if (SILU && !SPLIT) { ak *= 0.5f; bk *= 0.5f; }   // SiLU is evaluated from v / 2 (silu_half)
        sa[cl] = ak; sb[cl] = bk;
        if (RAW) sp[cl] = pr;
    }
    __syncthreads();
    if (!worker) return;

    // ---- 4. stream ----
    for (int v = tv; v < Vb; v += threads_per_row) {
        const int cl = v << 3;
        const int c = c_lo + cl;
        float a[8], b[8], pre[8];
        {
            const float4 a0 = *reinterpret_cast<const float4*>(sa + cl), a1 = *reinterpret_cast<const float4*>(sa + cl + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(sb + cl), b1 = *reinterpret_cast<const float4*>(sb + cl + 4);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            if (RAW) {
#pragma unroll
                for (int k = 0; k < 8; ++k) pre[k] = sp[cl + k];
            }
        }
        const float* src; int ldc;
        src_of(c, src, ldc);
        const float* __restrict__ srcp = src;
        uint16_t* __restrict__ outq = p.out + (size_t)n * npix * C + c;
        uint16_t* __restrict__ rawq = RAW ? p.raw16 + (size_t)n * npix * C + c : nullptr;
        uint16_t* __restrict__ outq_lo = SPLIT ? p.out_lo + (size_t)n * npix * C + c : nullptr;
        uint16_t* __restrict__ rawq_lo = (SPLIT && RAW) ? p.raw_lo + (size_t)n * npix * C + c : nullptr;
        if (RESAMPLE == NDB_RESAMPLE_DOWN2) {
            for (int po = p_begin + r; po < p_end; po += rows_per_block) {
                float y[8], raw[8];
                const int ho = po / p.Wo, wo = po - ho * p.Wo;
                const float* px = srcp + ((size_t)(2 * ho) * p.W + 2 * wo) * ldc;
                const float4 l0 = __ldg(reinterpret_cast<const float4*>(px)), h0 = __ldg(reinterpret_cast<const float4*>(px + 4));
                const float4 l1 = __ldg(reinterpret_cast<const float4*>(px + ldc)), h1 = __ldg(reinterpret_cast<const float4*>(px + ldc + 4));
                const float* py = px + (size_t)p.W * ldc;
                const float4 l2 = __ldg(reinterpret_cast<const float4*>(py)), h2 = __ldg(reinterpret_cast<const float4*>(py + 4));
                const float4 l3 = __ldg(reinterpret_cast<const float4*>(py + ldc)), h3 = __ldg(reinterpret_cast<const float4*>(py + ldc + 4));
                const float xv[4][8] = {{l0.x, l0.y, l0.z, l0.w, h0.x, h0.y, h0.z, h0.w},
                                        {l1.x, l1.y, l1.z, l1.w, h1.x, h1.y, h1.z, h1.w},
                                        {l2.x, l2.y, l2.z, l2.w, h2.x, h2.y, h2.z, h2.w},
                                        {l3.x, l3.y, l3.z, l3.w, h3.x, h3.y, h3.z, h3.w}};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float acc = 0.f, racc = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float t = a[k] * xv[q][k] + b[k];
                        t = act(t);
                        acc += t;
                        if (RAW) racc += xv[q][k] + pre[k];
                    }
                    y[k] = acc * 0.25f;
                    raw[k] = racc * 0.25f;
                }
                const size_t o = (size_t)po * C;
                store8(outq + o, SPLIT ? outq_lo + o : nullptr, y);
                if (RAW) store8(rawq + o, SPLIT ? rawq_lo + o : nullptr, raw);
                if (p.pool32 != nullptr) {
                    // same association as F.avg_pool2d: sum of the four taps, then * 0.25 (pre_add is not part of x)
                    float pl[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) pl[k] = (xv[0][k] + xv[1][k] + xv[2][k] + xv[3][k]) * 0.25f;
                    float* pp = p.pool32 + ((size_t)n * npix + po) * C + c;
                    *reinterpret_cast<float4*>(pp) = make_float4(pl[0], pl[1], pl[2], pl[3]);
                    *reinterpret_cast<float4*>(pp + 4) = make_float4(pl[4], pl[5], pl[6], pl[7]);
                }
            }
        } else {
            // this thread's pixels: p_begin + r, + rows_per_block, ... ; element offsets advance by constant strides (no
            // per-pixel 64-bit index arithmetic in the loop)
            const int rows = rows_per_block;
            int cnt = (p_end - (p_begin + r) + rows - 1) / rows;           // > 0 for workers with a pixel
            const size_t esz = IN16 ? 2 : 4;
            const uint32_t out_stride = (uint32_t)rows * (uint32_t)C;       // elements between two of this thread's pixels
            uint16_t* op = outq + (size_t)(p_begin + r) * C;
            uint16_t* rp = RAW ? rawq + (size_t)(p_begin + r) * C : nullptr;
            const ptrdiff_t lo_off = SPLIT ? (p.out_lo - p.out) : 0;            // hi -> lo tensor, in elements
            const ptrdiff_t raw_lo_off = (SPLIT && RAW) ? (p.raw_lo - p.raw16) : 0;
            bool first = (v == tv);                                        // the very first trip was requested before the prologue
            if (RESAMPLE == NDB_RESAMPLE_NONE) {
                const uint32_t in_stride = (uint32_t)rows * (uint32_t)ldc * (uint32_t)esz;   // bytes
                const char* ip = reinterpret_cast<const char*>(srcp) + (size_t)(p_begin + r) * ldc * esz;
                for (; cnt > 0; cnt -= U, ip += (size_t)U * in_stride, op += (size_t)U * out_stride, rp += RAW ? (size_t)U * out_stride : 0) {
                    if (!first) {
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (u < cnt) {
                                const char* px = ip + u * in_stride;
                                lo[u] = __ldg(reinterpret_cast<const uint4*>(px));
                                if (!IN16) hi[u] = __ldg(reinterpret_cast<const uint4*>(px + 16));
                            }
                        }
                    }
                    first = false;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (u < cnt) {
                            float xv[8];
                            unpack8(lo[u], hi[u], xv);
                            float y[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) y[k] = act(fmaf(a[k], xv[k], b[k]));
                            store8(op + u * out_stride, op + u * out_stride + lo_off, y);
                            if (RAW) {
                                float rw[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) rw[k] = xv[k] + pre[k];
                                store8(rp + u * out_stride, rp + u * out_stride + raw_lo_off, rw);
                            }
                        }
                    }
                }
            } else {
                // nearest-neighbour x2: the source pixel of an output pixel is not an affine function of its index
                for (int po = p_begin + r; po < p_end; po += U * rows) {
                    if (!first) {
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int pu = po + u * rows;
                            if (pu < p_end) load8(srcp, ldc, in_pixel(pu), lo[u], hi[u]);
                        }
                    }
                    first = false;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int pu = po + u * rows;
                        if (pu < p_end) {
                            float xv[8];
                            unpack8(lo[u], hi[u], xv);
                            float y[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) y[k] = act(fmaf(a[k], xv[k], b[k]));
                            const size_t o = (size_t)pu * C;
                            store8(outq + o, SPLIT ? outq_lo + o : nullptr, y);
                            if (RAW) {
                                float rw[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) rw[k] = xv[k] + pre[k];
                                store8(rawq + o, SPLIT ? rawq_lo + o : nullptr, rw);
                            }
                        }
                    }
                }
            }
        }
    }
}

// shared memory of a block: max(2 doubles, 3 floats) per channel of its chunk
static size_t gn_apply_smem(const GnApplyArgs& a) { return 16 * (size_t)a.chunk_ch; }

template <typename K>
static void launch_gn(K kernel, const GnApplyArgs& a, dim3 grid, cudaStream_t st) {
    const size_t smem = gn_apply_smem(a);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // C > 3072, unchunked
    launch_kernel(kernel, grid, dim3(kGnThreads), smem, st, 1, a);
}

template <int FMT, bool RAW>
static void launch_gn_apply2(const GnApplyArgs& a, dim3 grid, cudaStream_t st) {
#define NDB_GN_CASE(RS, SL) launch_gn(gn_apply_kernel<FMT, RS, SL, RAW, false>, a, grid, st)
    if (a.resample == NDB_RESAMPLE_NONE) { if (a.silu) NDB_GN_CASE(NDB_RESAMPLE_NONE, true); else NDB_GN_CASE(NDB_RESAMPLE_NONE, false); }
    else if (a.resample == NDB_RESAMPLE_UP2) { if (a.silu) NDB_GN_CASE(NDB_RESAMPLE_UP2, true); else NDB_GN_CASE(NDB_RESAMPLE_UP2, false); }
    else { if (a.silu) NDB_GN_CASE(NDB_RESAMPLE_DOWN2, true); else NDB_GN_CASE(NDB_RESAMPLE_DOWN2, false); }
#undef NDB_GN_CASE
}

template <bool RAW>
static void launch_gn_apply_split(const GnApplyArgs& a, dim3 grid, cudaStream_t st) {
#define NDB_GN_SCASE(RS, SL) launch_gn(gn_apply_kernel<0, RS, SL, RAW, false, true>, a, grid, st)
    if (a.resample == NDB_RESAMPLE_NONE) { if (a.silu) NDB_GN_SCASE(NDB_RESAMPLE_NONE, true); else NDB_GN_SCASE(NDB_RESAMPLE_NONE, false); }
    else if (a.resample == NDB_RESAMPLE_UP2) { if (a.silu) NDB_GN_SCASE(NDB_RESAMPLE_UP2, true); else NDB_GN_SCASE(NDB_RESAMPLE_UP2, false); }
    else { if (a.silu) NDB_GN_SCASE(NDB_RESAMPLE_DOWN2, true); else NDB_GN_SCASE(NDB_RESAMPLE_DOWN2, false); }
#undef NDB_GN_SCASE
}

template <int FMT>
static void launch_gn_apply(const GnApplyArgs& a, dim3 grid, cudaStream_t st) {
    if (a.out_lo != nullptr) {
        if (a.raw16 != nullptr) launch_gn_apply_split<true>(a, grid, st);
        else launch_gn_apply_split<false>(a, grid, st);
    } else if (a.x0_is16) {
        // 16-bit source: the out_norm of a ResBlock (single source, no resampling, no raw copy)
        if (a.silu) launch_gn(gn_apply_kernel<FMT, NDB_RESAMPLE_NONE, true, false, true>, a, grid, st);
        else launch_gn(gn_apply_kernel<FMT, NDB_RESAMPLE_NONE, false, false, true>, a, grid, st);
    } else if (a.raw16 != nullptr) launch_gn_apply2<FMT, true>(a, grid, st);
    else launch_gn_apply2<FMT, false>(a, grid, st);
}

}  // namespace ndb

using namespace ndb;

struct GnStatsGeom {
    int threads_per_row, rows_per_block, pix_per_block, chunks;
};

static GnStatsGeom gn_stats_geom(int N, int HW, int C) {
    GnStatsGeom g;
    const int V = C / 4;
    g.threads_per_row = V < kGnThreads ? V : kGnThreads;
    g.rows_per_block = kGnThreads / g.threads_per_row;
    // enough blocks to fill the machine, but at least ~8 pixels per row-slot
    int target_blocks_per_sample = (4 * num_sms() + N - 1) / N;
    if (target_blocks_per_sample < 1) target_blocks_per_sample = 1;
    g.pix_per_block = (HW + target_blocks_per_sample - 1) / target_blocks_per_sample;
    const int min_pix = g.rows_per_block * 8;
    if (g.pix_per_block < min_pix) g.pix_per_block = min_pix;
    g.chunks = (HW + g.pix_per_block - 1) / g.pix_per_block;
    return g;
}

static size_t gn_counter_bytes(int N) { return (((size_t)N * sizeof(unsigned int)) + 255) / 256 * 256; }

extern "C" int64_t ndb_gn_stats_workspace_bytes(int N, int HW, int C) {
    if (N <= 0 || HW <= 0 || C <= 0) return 0;
    const GnStatsGeom g = gn_stats_geom(N, HW, C);
    return (int64_t)(gn_counter_bytes(N) + (size_t)N * g.chunks * 2 * C * sizeof(double));
}

extern "C" int ndb_gn_stats(const float* x0, int C0, const float* x1, int C1, int N, int HW, void* stats0,
                            void* stats1, void* workspace, void* stream) {
    NDB_CHECK_ARG(x0 && stats0 && workspace && N > 0 && HW > 0);
    if (x1 == nullptr) C1 = 0;
    NDB_CHECK_ARG(C1 == 0 || stats1 != nullptr);
    const int C = C0 + C1;
    NDB_CHECK_ARG(C0 % 4 == 0 && C1 % 4 == 0);
    const int V = C / 4;
    NDB_CHECK_ARG(V <= kMaxVecPerThread * kGnThreads);
    const GnStatsGeom g = gn_stats_geom(N, HW, C);
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + gn_counter_bytes(N));
    dim3 grid(g.chunks, N);
    const size_t smem = 2 * (size_t)g.rows_per_block * C * sizeof(float);
    gn_stats_kernel<<<grid, kGnThreads, smem, (cudaStream_t)stream>>>(
        x0, C0, x1, C1, HW, (float*)stats0, (float*)stats1, partials, counters, g.pix_per_block, g.threads_per_row, g.rows_per_block);
    return launch_status("ndb_gn_stats");
}

extern "C" int ndb_gn_apply(const ndb_gn_apply_params* q, void* stream) {
    NDB_CHECK_ARG(q && q->x0 && q->parts0 && q->counts0 && q->smax0 >= 1 && q->gamma && q->beta && q->out);
    NDB_CHECK_ARG(q->x1 == nullptr || (q->parts1 != nullptr && q->counts1 != nullptr && q->smax1 >= 1));
    GnApplyArgs a{};
    a.x0 = q->x0; a.C0 = q->C0;
    a.x1 = q->x1; a.C1 = q->x1 ? q->C1 : 0;
    const int C = a.C0 + a.C1;
    NDB_CHECK_ARG(C % kGroups == 0 && a.C0 % 8 == 0 && a.C1 % 8 == 0);
    NDB_CHECK_ARG((q->scale == nullptr) == (q->shift == nullptr));
    a.N = q->N; a.H = q->H; a.W = q->W;
    a.Ho = q->H; a.Wo = q->W;
    if (q->resample == NDB_RESAMPLE_UP2) { a.Ho = 2 * q->H; a.Wo = 2 * q->W; }
    if (q->resample == NDB_RESAMPLE_DOWN2) {
        NDB_CHECK_ARG(q->H % 2 == 0 && q->W % 2 == 0);
        a.Ho = q->H / 2; a.Wo = q->W / 2;
    }
    a.parts0 = (const float2*)q->parts0; a.counts0 = q->counts0; a.smax0 = q->smax0;
    a.parts1 = q->x1 ? (const float2*)q->parts1 : nullptr; a.counts1 = q->counts1; a.smax1 = q->smax1;
    a.gamma = q->gamma; a.beta = q->beta;
    a.scale = q->scale; a.shift = q->shift; a.ld_ss = q->ld_ss;
    a.pre_add = q->pre_add; a.ld_pre = q->ld_pre;
    a.silu = q->silu; a.resample = q->resample;
    a.out = (uint16_t*)q->out; a.raw16 = (uint16_t*)q->raw16;
    a.out_lo = (uint16_t*)q->out_lo; a.raw_lo = (uint16_t*)q->raw_lo;
    if (q->out_lo != nullptr) {
        NDB_CHECK_ARG(q->operand_format == 0 && !q->x0_is16);
        NDB_CHECK_ARG((q->raw16 == nullptr) == (q->raw_lo == nullptr));
    }
    a.pool32 = q->pool32;
    a.x0_is16 = q->x0_is16;
    a.cpg = q->cpg;
    if (q->cpg > 0) NDB_CHECK_ARG(C % q->cpg == 0 && C / q->cpg <= kGroups);
    if (q->x0_is16) NDB_CHECK_ARG(q->x1 == nullptr && q->resample == NDB_RESAMPLE_NONE && q->raw16 == nullptr);
    NDB_CHECK_ARG(q->pool32 == nullptr || q->resample == NDB_RESAMPLE_DOWN2);
    a.eps = q->eps;
    const int npix = a.Ho * a.Wo;
    const int cpg = q->cpg > 0 ? q->cpg : C / kGroups;
    // target: one full wave at the kernel's occupancy (3 blocks per SM): no tail wave, prologue amortised over more pixels
    int target = (3 * num_sms()) / q->N;
    if (target < 1) target = 1;
    auto pixel_split = [&](int chunk_ch, int want_blocks, int& ppb) {
        const int v8 = chunk_ch / 8;
        const int rows = kGnThreads / (v8 < kGnThreads ? v8 : kGnThreads);
        ppb = (npix + want_blocks - 1) / want_blocks;
        if (ppb < rows * 4) ppb = rows * 4;
        return (npix + ppb - 1) / ppb;
    };
    int ppb = 0;
    int chunks = 1;
    a.chunk_ch = C;
    int pblocks = pixel_split(C, target, ppb);
    static int no_chunk = -1;
    if (no_chunk < 0) no_chunk = std::getenv("NDB_GN_NO_CHUNK") != nullptr ? 1 : 0;
    if (pblocks < target && !no_chunk) {
        // small feature map: split the channels too. A chunk is a whole number of groups and of 8-channel vectors
        // (granule = lcm(cpg, 8) channels) and at least 64 channels wide (256-byte rows keep the accesses coalesced).
        int gran = cpg;
        while (gran % 8 != 0) gran += cpg;
        int granules = C / gran;                               // C % gran == 0: C is a multiple of 32 * cpg ... and of 8
        if (C % gran != 0) granules = 0;
        int per = 1;                                           // granules per chunk
        while (per * gran < 64 && per < granules) ++per;
        if (granules > 0) {
            int max_chunks = granules / per;
            if (max_chunks < 1) max_chunks = 1;
            int want = target < max_chunks ? target : max_chunks;
            per = (granules + want - 1) / want;
            a.chunk_ch = per * gran;
            chunks = (C + a.chunk_ch - 1) / a.chunk_ch;
            pblocks = pixel_split(a.chunk_ch, (target + chunks - 1) / chunks, ppb);
        }
    }
    a.pix_per_block = ppb;
    dim3 grid(pblocks, q->N, chunks);
    if (q->operand_format == 0)
        launch_gn_apply<0>(a, grid, (cudaStream_t)stream);
    else
        launch_gn_apply<1>(a, grid, (cudaStream_t)stream);
    return launch_status("ndb_gn_apply");
}
