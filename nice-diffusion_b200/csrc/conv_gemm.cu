// Implicit-GEMM convolution (3x3 pad 1 / 1x1) on the 5th-gen tensor cores.
//
//   M = N*H*W output pixels, tiled as boxes of TN images x TH rows x TW columns (<= 128 pixels),
//   N = output channels (BLOCK_N <= 256 per tile), K = sum over segments of taps * C.
//
// Data path: activations stay NHWC fp16 in HBM; for every (tap, 64-channel chunk) one TMA 4-D box
// {64ch, TW, TH, TN} shifted by the tap offset lands in shared memory as a 128-row K-major SWIZZLE_128B
// operand tile (image borders = TMA out-of-bounds zero fill, so there is no im2col buffer and no padding
// pass). Weights are [tap][cout][cin] fp16 and arrive as {64, BLOCK_N, 1} boxes. One elected thread issues
// tcgen05.mma (128 x BLOCK_N x 16, fp32 accumulate in TMEM); accumulators are double-buffered in TMEM so
// the epilogue of tile i overlaps the main loop of tile i+1. Persistent CTAs, one per SM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias / residual -> global).
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ndb {

constexpr int kMaxSeg = NDB_MAX_SEGMENTS;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                        // 64 x 16-bit = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct TmapPack {
    CUtensorMap a[kMaxSeg];
    CUtensorMap b[kMaxSeg];
};

struct ConvArgs {
    int N, H, W;
    int TW, TH, TN;
    int tiles_w, tiles_h, tiles_n;
    int num_m_tiles, num_n_tiles;
    int block_n;
    int cout;
    int num_seg;
    int seg_taps[kMaxSeg];
    int seg_chunks[kMaxSeg];
    int total_kblocks;
    int stages;
    uint32_t a_box_bytes;
    uint32_t tmem_cols;
    uint32_t idesc;
    const float* bias;
    const float* residual;
    int res_mode;
    int out_kind;
    void* out;
    int fmt;
};

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (FMT == 0) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ TmapPack maps, const ConvArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int b_tile_bytes = p.block_n * 128;
    const int stage_bytes = kATileBytes + b_tile_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.num_seg; ++s) {
            prefetch_tmap(&maps.a[s]);
            prefetch_tmap(&maps.b[s]);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.num_n_tiles;
                int m_tile = tile / p.num_n_tiles;
                const int tw_i = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th_i = m_tile % p.tiles_h;
                const int tn_i = m_tile / p.tiles_h;
                const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
                for (int s = 0; s < p.num_seg; ++s) {
                    const int taps = p.seg_taps[s];
                    const int chunks = p.seg_chunks[s];
                    for (int tap = 0; tap < taps; ++tap) {
                        const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
                        const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
                        for (int ch = 0; ch < chunks; ++ch) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* sa = smem + stage * stage_bytes;
                            uint8_t* sb = sa + kATileBytes;
                            mbar_arrive_expect_tx(&full_bar[stage], p.a_box_bytes + (uint32_t)b_tile_bytes);
                            tma_load_4d(sa, &maps.a[s], &full_bar[stage], ch * kBlockK, w0 + dx, h0 + dy, n0);
                            tma_load_3d(sb, &maps.b[s], &full_bar[stage], ch * kBlockK, n_tile * p.block_n, tap);
                            if (++stage == p.stages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.block_n;
                for (int kb = 0; kb < p.total_kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                    const uint64_t adesc = make_desc_k_sw128(sa);
                    const uint64_t bdesc = make_desc_k_sw128(sa + kATileBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the >>4 address field
                        umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;     // accumulator row == pixel index inside the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        const int pix_per_img = p.TH * p.TW;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.num_n_tiles;
            int m_tile = tile / p.num_n_tiles;
            const int tw_i = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th_i = m_tile % p.tiles_h;
            const int tn_i = m_tile / p.tiles_h;
            const int tn = row / pix_per_img;
            const int rem = row - tn * pix_per_img;
            const int th = rem / p.TW;
            const int tw = rem - th * p.TW;
            const int n = tn_i * p.TN + tn, h = th_i * p.TH + th, w = tw_i * p.TW + tw;
            const bool valid = (tn < p.TN) && (n < p.N) && (h < p.H) && (w < p.W);
            const int col0 = n_tile * p.block_n;
            const long long pix = ((long long)n * p.H + h) * p.W + w;

            const float* res0 = nullptr;
            const float* res1 = nullptr;
            const float* res2 = nullptr;
            const float* res3 = nullptr;
            if (valid && p.residual != nullptr) {
                if (p.res_mode == NDB_RES_SAME) {
                    res0 = p.residual + pix * p.cout;
                } else if (p.res_mode == NDB_RES_UP2) {
                    const int Hh = p.H >> 1, Wh = p.W >> 1;
                    res0 = p.residual + (((long long)n * Hh + (h >> 1)) * Wh + (w >> 1)) * p.cout;
                } else if (p.res_mode == NDB_RES_DOWN2) {
                    const int H2 = p.H * 2, W2 = p.W * 2;
                    res0 = p.residual + (((long long)n * H2 + 2 * h) * W2 + 2 * w) * p.cout;
                    res1 = res0 + p.cout;
                    res2 = res0 + (long long)W2 * p.cout;
                    res3 = res2 + p.cout;
                }
            }

            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * p.block_n;
            for (int c0 = 0; c0 < p.block_n; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_row + c0, r);
                tmem_ld_wait();
                if (valid) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                    const int c = col0 + c0;
                    if (c + 16 <= p.cout) {
                        if (p.bias != nullptr) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c + i));
                                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                            }
                        }
                        if (res0 != nullptr) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                float4 a4 = __ldg(reinterpret_cast<const float4*>(res0 + c + i));
                                if (res1 != nullptr) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(res1 + c + i));
                                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(res2 + c + i));
                                    const float4 d4 = __ldg(reinterpret_cast<const float4*>(res3 + c + i));
                                    // same association as F.avg_pool2d: sum of the 4 taps, then * 0.25
                                    a4.x = (a4.x + b4.x + c4.x + d4.x) * 0.25f;
                                    a4.y = (a4.y + b4.y + c4.y + d4.y) * 0.25f;
                                    a4.z = (a4.z + b4.z + c4.z + d4.z) * 0.25f;
                                    a4.w = (a4.w + b4.w + c4.w + d4.w) * 0.25f;
                                }
                                v[i] += a4.x; v[i + 1] += a4.y; v[i + 2] += a4.z; v[i + 3] += a4.w;
                            }
                        }
                        if (p.out_kind == NDB_OUT_F32_NHWC) {
                            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.cout + c);
#pragma unroll
                            for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        } else if (p.out_kind == NDB_OUT_F16_NHWC) {
                            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + pix * p.cout + c);
                            uint32_t w[8];
                            if (p.fmt == 0) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) w[i] = pack2<0>(v[2 * i], v[2 * i + 1]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) w[i] = pack2<1>(v[2 * i], v[2 * i + 1]);
                            }
                            o[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            o[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        } else {  // NCHW fp32
                            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                o[(((long long)n * p.cout + c + i) * p.H + h) * p.W + w] = v[i];
                        }
                    } else {
                        // ragged channel tail (e.g. the 6-channel head in a 16-wide tile): scalar path
                        for (int i = 0; i < 16; ++i) {
                            const int cc = c + i;
                            if (cc >= p.cout) break;
                            float val = v[i];
                            if (p.bias != nullptr) val += __ldg(p.bias + cc);
                            if (res0 != nullptr) {
                                float rr = __ldg(res0 + cc);
                                if (res1 != nullptr) rr = (rr + __ldg(res1 + cc) + __ldg(res2 + cc) + __ldg(res3 + cc)) * 0.25f;
                                val += rr;
                            }
                            if (p.out_kind == NDB_OUT_F32_NHWC) {
                                reinterpret_cast<float*>(p.out)[pix * p.cout + cc] = val;
                            } else if (p.out_kind == NDB_OUT_F16_NHWC) {
                                if (p.fmt == 0) reinterpret_cast<__half*>(p.out)[pix * p.cout + cc] = __float2half_rn(val);
                                else reinterpret_cast<__nv_bfloat16*>(p.out)[pix * p.cout + cc] = __float2bfloat16_rn(val);
                            } else {
                                reinterpret_cast<float*>(p.out)[(((long long)n * p.cout + cc) * p.H + h) * p.W + w] = val;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [tap][cout_pad][cin_len] 16-bit
// ------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ dst, int cout, int cin_total,
                                   int taps, int cout_pad, int cin_off, int cin_len, int fmt) {
    const long long total = (long long)taps * cout_pad * cin_len;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_len);
        long long r = i / cin_len;
        const int co = (int)(r % cout_pad);
        const int tap = (int)(r / cout_pad);
        float v = 0.f;
        if (co < cout) v = w[((long long)co * cin_total + cin_off + ci) * taps + tap];
        uint16_t bits;
        if (fmt == 0) {
            __half h = __float2half_rn(v);
            bits = *reinterpret_cast<uint16_t*>(&h);
        } else {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            bits = *reinterpret_cast<uint16_t*>(&h);
        }
        dst[i] = bits;
    }
}

static int pick_block_n(int cout_pad) {
    for (int bn = 256; bn >= 16; bn -= 16)
        if (cout_pad % bn == 0) return bn;
    return 0;
}

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_pack_conv_weight(const float* w_oihw, void* dst, int cout, int cin_total, int ksize, int cout_pad,
                                    int cin_off, int cin_len, int operand_format, void* stream) {
    NDB_CHECK_ARG(w_oihw && dst);
    NDB_CHECK_ARG(ksize == 1 || ksize == 3);
    NDB_CHECK_ARG(cout_pad >= cout && cin_off >= 0 && cin_off + cin_len <= cin_total);
    const int taps = ksize * ksize;
    const long long total = (long long)taps * cout_pad * cin_len;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, (uint16_t*)dst, cout, cin_total, taps, cout_pad,
                                                                  cin_off, cin_len, operand_format);
    return launch_status("ndb_pack_conv_weight");
}

extern "C" int ndb_conv_gemm(const ndb_conv_params* prm, void* stream) {
    NDB_CHECK_ARG(prm != nullptr);
    NDB_CHECK_ARG(prm->N > 0 && prm->H > 0 && prm->W > 0);
    NDB_CHECK_ARG(prm->num_segments >= 1 && prm->num_segments <= kMaxSeg);
    NDB_CHECK_ARG(prm->cout >= 1 && prm->cout_pad >= prm->cout && prm->cout_pad % 16 == 0);
    NDB_CHECK_ARG(prm->out != nullptr);
    NDB_CHECK_ARG(prm->operand_format == 0 || prm->operand_format == 1);
    if (prm->gn_stats != nullptr) return NDB_ERR_UNSUPPORTED;
    if (prm->out_kind != NDB_OUT_F32_NCHW) NDB_CHECK_ARG(prm->cout % 4 == 0 || prm->cout < 16);
    if (prm->residual_mode == NDB_RES_UP2) NDB_CHECK_ARG(prm->H % 2 == 0 && prm->W % 2 == 0);

    ConvArgs a{};
    a.N = prm->N; a.H = prm->H; a.W = prm->W;
    a.TW = prm->W < kBlockM ? prm->W : kBlockM;
    a.TH = kBlockM / a.TW;
    if (a.TH > prm->H) a.TH = prm->H;
    if (a.TH < 1) a.TH = 1;
    a.TN = 1;
    if (a.TH == prm->H && a.TW == prm->W) {
        a.TN = kBlockM / (prm->H * prm->W);
        if (a.TN < 1) a.TN = 1;
        if (a.TN > prm->N) a.TN = prm->N;
    }
    a.tiles_w = (prm->W + a.TW - 1) / a.TW;
    a.tiles_h = (prm->H + a.TH - 1) / a.TH;
    a.tiles_n = (prm->N + a.TN - 1) / a.TN;
    a.num_m_tiles = a.tiles_w * a.tiles_h * a.tiles_n;
    a.block_n = prm->block_n > 0 ? prm->block_n : pick_block_n(prm->cout_pad);
    NDB_CHECK_ARG(a.block_n >= 16 && a.block_n <= 256 && a.block_n % 16 == 0 && prm->cout_pad % a.block_n == 0);
    a.num_n_tiles = prm->cout_pad / a.block_n;
    a.cout = prm->cout;
    a.num_seg = prm->num_segments;
    a.a_box_bytes = (uint32_t)(a.TN * a.TH * a.TW * 128);
    a.tmem_cols = a.block_n <= 16 ? 32 : a.block_n <= 32 ? 64 : a.block_n <= 64 ? 128 : a.block_n <= 128 ? 256 : 512;
    a.idesc = make_idesc_f16(kBlockM, a.block_n, prm->operand_format, 0, 0);
    a.bias = prm->bias;
    a.residual = prm->residual;
    a.res_mode = prm->residual ? prm->residual_mode : NDB_RES_NONE;
    if (prm->residual) NDB_CHECK_ARG(prm->residual_mode >= NDB_RES_SAME && prm->residual_mode <= NDB_RES_DOWN2);
    a.out_kind = prm->out_kind;
    a.out = prm->out;
    a.fmt = prm->operand_format;

    TmapPack maps;
    a.total_kblocks = 0;
    for (int s = 0; s < prm->num_segments; ++s) {
        const ndb_conv_segment& sg = prm->seg[s];
        NDB_CHECK_ARG(sg.act && sg.weight);
        NDB_CHECK_ARG(sg.C > 0 && sg.C % kBlockK == 0);
        NDB_CHECK_ARG(sg.taps == 1 || sg.taps == 9);
        a.seg_taps[s] = sg.taps;
        a.seg_chunks[s] = sg.C / kBlockK;
        a.total_kblocks += sg.taps * a.seg_chunks[s];
        {
            uint64_t dims[4] = {(uint64_t)sg.C, (uint64_t)prm->W, (uint64_t)prm->H, (uint64_t)prm->N};
            uint64_t strides[3] = {(uint64_t)sg.C * 2, (uint64_t)prm->W * sg.C * 2,
                                   (uint64_t)prm->H * prm->W * sg.C * 2};
            uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)a.TW, (uint32_t)a.TH, (uint32_t)a.TN};
            int rc = encode_tmap_16bit(&maps.a[s], sg.act, 4, dims, strides, box);
            if (rc != NDB_OK) return rc;
        }
        {
            uint64_t dims[3] = {(uint64_t)sg.C, (uint64_t)prm->cout_pad, (uint64_t)sg.taps};
            uint64_t strides[2] = {(uint64_t)sg.C * 2, (uint64_t)prm->cout_pad * sg.C * 2};
            uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)a.block_n, 1u};
            int rc = encode_tmap_16bit(&maps.b[s], sg.weight, 3, dims, strides, box);
            if (rc != NDB_OK) return rc;
        }
    }

    const int stage_bytes = kATileBytes + a.block_n * 128;
    const int smem_budget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/;
    int stages = smem_budget / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > a.total_kblocks) stages = a.total_kblocks < 2 ? 2 : a.total_kblocks;
    NDB_CHECK_ARG(stages >= 2);
    a.stages = stages;
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 256;

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return launch_status("cudaFuncSetAttribute(conv_gemm_kernel)");
        attr_set = true;
    }
    const int num_tiles = a.num_m_tiles * a.num_n_tiles;
    int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    conv_gemm_kernel<<<grid, kThreads, smem_bytes, (cudaStream_t)stream>>>(maps, a);
    return launch_status("ndb_conv_gemm");
}
