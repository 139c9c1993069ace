// Fused multi-head self-attention on tcgen05: S = Q K^T and O += P V run on the tensor cores with fp32
// accumulators in TMEM; the online softmax runs in registers, two threads per query row (each owns 64 of the 128 keys
// of a block, read out of TMEM once). The [T,T] score matrix never touches HBM (the reference materialises it:
// model.py:273-277).
//
// One CTA = one (batch, head, 128-query block). Q/K/V tiles are read straight out of the qkv 1x1-conv
// output ([B][T][3C], fp16) with 4-D TMA boxes {d, column-block, rows, batch}; the inner extent is the true
// head dim d, so head dims that are not a multiple of 64 are zero-filled up to the next 64 by TMA.
//   warps 0-7 : softmax + correction + epilogue. Warp w works on TMEM lane quarter (w & 3) -- query rows 32 (w & 3) .. +31
//               -- and key half (w >> 2); the two warps of a quarter exchange their row maxima through shared memory
//               (one 64-thread named barrier per block) and keep separate partial row sums until the end.
//   warp  8   : one elected thread issues TMA loads and tcgen05.mma
// TMEM: S at columns [0,128), O at [128,128+Dp).
// K/V tiles are double-buffered; with d <= 64 two CTAs fit per SM so one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace ndb {

constexpr int kAttnBM = 128;  // queries per CTA
constexpr int kAttnBN = 128;  // keys per iteration
constexpr int kChunkBytes = 128 * 128;  // one [128 rows x 64 x 16-bit] swizzled tile
constexpr int kAttnThreads = 288;       // 8 softmax warps + 1 TMA/MMA warp

struct AttnArgs {
    int B, T, heads, d, order;
    int dp_chunks;   // ceil(d / 64)
    int kv_stages;
    int fmt;
    uint32_t tmem_cols;
    uint32_t idesc_qk, idesc_pv;
    float scale_log2e;
    uint16_t* out;
    int C;
    int dbg;       // -DNDB_BRINGUP only (env NDB_ATTN_DBG): 1 = pass 1 without its TMEM reads, 2 = pass 1 reads twice (timing probes)
};

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2a(float a, float b) {
    if (FMT == 0) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

// SPLIT (fp32-class mode, head dim <= 64): q, k, v arrive as fp16 (hi, lo) pairs in two tensors of the same layout and every
// product runs as three MMAs (hi hi + lo hi + hi lo); P is written as a (hi, lo) pair too; the output is fp32. One CTA per
// SM (160 KB of operand tiles), single-buffered K / V.
template <int FMT, bool SPLIT>
// Two CTAs per SM (shared memory: 2 x ~114 KB). Registers are allocated for an even number of warps, so the budget is
// 65536 / (2 * 10 * 32) = 102 -> 96 registers per thread: at 111 (measured) only one CTA is resident and T = 1024 takes 145 us
// instead of 108.
__global__ void __launch_bounds__(kAttnThreads, SPLIT ? 1 : 2)
attention_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap qkv_lo_map, const AttnArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qb * kAttnBM;
    const int nkv = (p.T + kAttnBN - 1) / kAttnBN;
    const int dpc = p.dp_chunks;
    const int Dp = dpc * 64;

    // smem carve-up
    // SPLIT: every tile is followed by its "lo" twin (kv_stages == 1)
    constexpr int kTw = SPLIT ? 2 : 1;
    uint8_t* sQ = smem;                                         // dpc chunks (x kTw)
    uint8_t* sK = sQ + kTw * dpc * kChunkBytes;                 // kv_stages * dpc chunks (x kTw)
    uint8_t* sV = sK + kTw * p.kv_stages * dpc * kChunkBytes;   // kv_stages * dpc chunks (x kTw)
    uint8_t* sP = sV + kTw * p.kv_stages * dpc * kChunkBytes;   // 2 chunks (128 keys) (x kTw)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kTw * 2 * kChunkBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* v_full = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* pv_done = bars + 7;
    uint64_t* s_free = bars + 8;     // all softmax threads have pulled S_j out of TMEM: S_{j+1} may overwrite it
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);
    uint16_t* xchg = reinterpret_cast<uint16_t*>(bars + 10);   // [2 key halves][128 rows]: row maxima (bf16, rounded up)

    if (threadIdx.x == 256) {
        prefetch_tmap(&qkv_map);
        if (SPLIT) prefetch_tmap(&qkv_lo_map);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 256);
        mbar_init(pv_done, 1);
        mbar_init(s_free, 256);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr_smem, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + kAttnBN;

    // column-block indices of q / k / v for this head inside the 3C channel axis (units of d)
    int cq, ck, cv;
    if (p.order == 0) { cq = h; ck = p.heads + h; cv = 2 * p.heads + h; }
    else { cq = 3 * h; ck = 3 * h + 1; cv = 3 * h + 2; }

    if (warp == 8) {
        // The whole warp runs this warp-uniform loop and one elected lane issues the TMA / MMA instructions, so their
        // descriptor operands stay in uniform registers (no per-instruction R2UR waterfall).
        {
            const uint32_t tile_bytes = (uint32_t)dpc * kChunkBytes;
            const uint32_t lo_off = (uint32_t)dpc * kChunkBytes;      // SPLIT: the lo twin sits right behind its hi tile
            auto load_tile = [&](uint8_t* dst, uint64_t* bar, int colblk, int row0) {
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar, kTw * tile_bytes);
                    for (int c = 0; c < dpc; ++c) {
                        tma_load_4d(dst + c * kChunkBytes, &qkv_map, bar, c * 64, colblk, row0, b);
                        if (SPLIT) tma_load_4d(dst + lo_off + c * kChunkBytes, &qkv_lo_map, bar, c * 64, colblk, row0, b);
                    }
                }
                __syncwarp();
            };
            load_tile(sQ, q_full, cq, q0);
            load_tile(sK, &k_full[0], ck, 0);
            load_tile(sV, &v_full[0], cv, 0);
            const bool two = p.kv_stages == 2;
            if (two && nkv > 1) load_tile(sK + kTw * dpc * kChunkBytes, &k_full[1], ck, kAttnBN);
            mbar_wait(q_full, 0);
            auto issue_S = [&](int j) {      // S = Q K_j^T into TMEM (K tile j must have been requested)
                const int st = two ? (j & 1) : 0;
                const uint32_t use_parity = two ? ((j >> 1) & 1) : (j & 1);
                mbar_wait(&k_full[st], use_parity);
                tc_fence_after();
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + kTw * st * dpc * kChunkBytes);
                if (elect_one()) {
                    // SPLIT: q_hi k_hi + q_lo k_hi + q_hi k_lo
                    for (int pass = 0; pass < (SPLIT ? 3 : 1); ++pass) {
                        const uint32_t qo = (pass == 1) ? lo_off : 0u, ko = (pass == 2) ? lo_off : 0u;
                        for (int i = 0; i < Dp / 16; ++i) {
                            const uint32_t off = (i >> 2) * kChunkBytes + (i & 3) * 32;
                            umma_f16(tmem_S, make_desc_k_sw128(qa + qo + off), make_desc_k_sw128(ka + ko + off), p.idesc_qk,
                                     (pass | i) ? 1u : 0u);
                        }
                    }
                    umma_commit(s_full);
                }
                __syncwarp();
            };
            issue_S(0);
            // Double-buffered K / V with SEPARATE schedules: a K stage is free as soon as the S MMA that read it has
            // completed (one block earlier than its V stage, which PV reads last), so K_{j+2} is requested right after
            // S_{j+1} is issued -- a block and a half before S_{j+2} needs it -- and the S MMA never waits on a TMA round trip;
            // V_{j+1} is requested when PV_{j-1} has retired, a full block before PV_{j+1}.
            for (int j = 0; j < nkv; ++j) {
                const int st = two ? (j & 1) : 0;
                const uint32_t use_parity = two ? ((j >> 1) & 1) : (j & 1);
                if (two && j + 1 < nkv) {
                    if (j >= 1) mbar_wait(pv_done, (j - 1) & 1);          // V stage (j+1)&1 was last read by PV_{j-1}
                    load_tile(sV + kTw * ((j + 1) & 1) * dpc * kChunkBytes, &v_full[(j + 1) & 1], cv, (j + 1) * kAttnBN);
                    // S_{j+1} as soon as the softmax threads hold S_j in registers: it runs under their exp phase
                    mbar_wait(s_free, j & 1);
                    issue_S(j + 1);
                    if (j + 2 < nkv) {
                        // K stage j&1 was read by S_j, complete since before s_free(j)
                        load_tile(sK + kTw * (j & 1) * dpc * kChunkBytes, &k_full[j & 1], ck, (j + 2) * kAttnBN);
                    }
                }
                // ---- O += P_j V_j ----
                mbar_wait(p_full, j & 1);
                mbar_wait(&v_full[st], use_parity);
                tc_fence_after();
                {
                    const uint32_t pa = smem_u32(sP), va = smem_u32(sV + kTw * st * dpc * kChunkBytes);
                    if (elect_one()) {
                        // SPLIT: p_hi v_hi + p_lo v_hi + p_hi v_lo
                        for (int pass = 0; pass < (SPLIT ? 3 : 1); ++pass) {
                            const uint32_t po = (pass == 1) ? 2u * kChunkBytes : 0u, vo = (pass == 2) ? lo_off : 0u;
                            for (int i = 0; i < kAttnBN / 16; ++i) {
                                const uint64_t adesc = make_desc_k_sw128(pa + po + (i >> 2) * kChunkBytes + (i & 3) * 32);
                                // V tile: rows = keys (K dim), 64 d-elements contiguous per row (MN-major), 16 keys = 2048 B
                                const uint64_t bdesc = make_desc_mn_sw128(va + vo + i * 2048, kChunkBytes);
                                umma_f16(tmem_O, adesc, bdesc, p.idesc_pv, (j | pass | i) ? 1u : 0u);
                            }
                        }
                        umma_commit(pv_done);
                    }
                    __syncwarp();
                }
                if (j + 1 < nkv && !two) {
                    // single-buffered K/V (large head dims): reload only after both MMAs of block j retired
                    mbar_wait(pv_done, j & 1);
                    load_tile(sK, &k_full[0], ck, (j + 1) * kAttnBN);
                    load_tile(sV, &v_full[0], cv, (j + 1) * kAttnBN);
                    mbar_wait(s_free, j & 1);
                    issue_S(j + 1);
                }
            }
        }
    } else {
        // ===================== softmax / correction / epilogue =====================
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = uint32_t(q * 32) << 16;
        const int key0 = hf * 64;                      // this thread's keys inside a block
        const int oc_half = Dp >> 1;                   // this warp's O columns: [hf * Dp/2, +Dp/2)
        float m_run = -INFINITY, l_run = 0.f;          // l_run: partial sum over this thread's keys
        // 64-thread named barrier of this row quarter (immediate ids: a register id makes ptxas reserve all 16 barriers)
        auto pair_sync = [&]() {
            if (q == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
            else if (q == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
            else if (q == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
            else asm volatile("bar.sync 4, 64;" ::: "memory");
        };
        // Lazy rescaling: the running shift m_run of a row only moves when the block maximum exceeds it by more than
        // kLazy in the exp2 domain; until then p = exp2(s - m_run) may reach 2^kLazy (exact in fp32, representable in both
        // 16-bit formats) and O / l keep accumulating against the same shift. The O correction (a TMEM round trip) then
        // happens on the first block and almost never again, instead of on nearly every block of every warp.
        constexpr float kLazy = 8.f;
        const float lazy_raw = kLazy / p.scale_log2e;       // threshold in raw score units
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = p.T - j * kAttnBN - key0;  // this thread's keys >= kv_valid are padding
            const bool full = kv_valid >= 64;
            // pass 1: maximum over this thread's 64 keys (two 32-column TMEM reads; the scores are read again in pass 2 --
            // holding all 64 in registers across the exchange / correction raises the register count past what two
            // resident CTAs allow)
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r1[32];
#ifdef NDB_BRINGUP
                if (p.dbg == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) r1[i] = 0u;
                } else {
                    tmem_ld32(tmem_S + lane_addr + key0 + c0, r1);
                    tmem_ld_wait();
                    if (p.dbg == 2) {
                        tmem_ld32(tmem_S + lane_addr + key0 + c0, r1);
                        tmem_ld_wait();
                    }
                }
#else
                tmem_ld32(tmem_S + lane_addr + key0 + c0, r1);
                tmem_ld_wait();
#endif
                if (full) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r1[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i < kv_valid) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r1[i]));
                }
            }
            // Both warps of a row must use the SAME block maximum. They exchange 16-bit values (the two-CTAs-per-SM shared
            // memory budget has no room for 1 KB of floats): each side rounds its own maximum UP to bf16 first -- any
            // upper bound of the scores is a valid softmax shift.
            uint32_t mu = __float_as_uint(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
            mu = (mu & 0x80000000u) ? (mu & 0xffff0000u) : ((mu + 0xffffu) & 0xffff0000u);
            xchg[hf * 128 + row] = (uint16_t)(mu >> 16);
            pair_sync();
            const float m_blk = fmaxf(__uint_as_float(mu), __uint_as_float((uint32_t)xchg[(hf ^ 1) * 128 + row] << 16));
            const bool move = m_blk > m_run + lazy_raw;        // always true on the first block (m_run = -inf)
            const float m_new = move ? m_blk : m_run;
            const float alpha = move ? ex2_fast((m_run - m_new) * p.scale_log2e) : 1.f;   // 0 when m_run = -inf
            const bool need_fix = (j > 0) && move;
            // previous block's P V must have retired before P (smem) and O (TMEM) are touched
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
            }
            if (__any_sync(0xffffffffu, need_fix)) {
                // rare: eight columns at a time, so that the 64 live scores are not spilled around it
#pragma unroll 1
                for (int c0 = hf * oc_half; c0 < (hf + 1) * oc_half; c0 += 8) {
                    uint32_t o[8];
                    tmem_ld8(tmem_O + lane_addr + c0, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st8(tmem_O + lane_addr + c0, o);
                }
                tmem_st_wait();
            }
            l_run *= alpha;
            m_run = m_new;
            // p = exp2((s - m) * scale*log2e) -> 16-bit P tile in the K-major SWIZZLE_128B layout; this thread's 64 keys are
            // exactly the 128-byte line of row `row` in K-chunk `hf`
            const float mb = m_new * p.scale_log2e;
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
            uint8_t* line = sP + hf * kChunkBytes + row * 128;
            // pass 2 pulls all 64 scores into registers and releases S at once, so the S_{j+1} MMA overlaps the exp phase.
            uint32_t r[64];
            tmem_ld32(tmem_S + lane_addr + key0, r);
            tmem_ld32(tmem_S + lane_addr + key0 + 32, r + 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            // (Two copies of the exp loop: with the padding test folded in as predication every element pays three extra
            //  issue slots even when the block is full -- the kernel is issue-bound.)
            if (full) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t pk[4], pl[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = c * 8 + 2 * k;
                        const float p0 = ex2_fast(fmaf(__uint_as_float(r[i]), p.scale_log2e, -mb));
                        const float p1 = ex2_fast(fmaf(__uint_as_float(r[i + 1]), p.scale_log2e, -mb));
                        ls[k] += p0 + p1;
                        pk[k] = pack2a<FMT>(p0, p1);
                        if (SPLIT) {
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&pk[k]));
                            pl[k] = pack2a<0>(p0 - hf.x, p1 - hf.y);
                        }
                    }
                    *reinterpret_cast<uint4*>(line + ((c ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (SPLIT)
                        *reinterpret_cast<uint4*>(line + 2 * kChunkBytes + ((c ^ (row & 7)) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t pk[4], pl[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = c * 8 + 2 * k;
                        float p0 = ex2_fast(fmaf(__uint_as_float(r[i]), p.scale_log2e, -mb));
                        float p1 = ex2_fast(fmaf(__uint_as_float(r[i + 1]), p.scale_log2e, -mb));
                        if (i >= kv_valid) p0 = 0.f;
                        if (i + 1 >= kv_valid) p1 = 0.f;
                        ls[k] += p0 + p1;
                        pk[k] = pack2a<FMT>(p0, p1);
                        if (SPLIT) {
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&pk[k]));
                            pl[k] = pack2a<0>(p0 - hf.x, p1 - hf.y);
                        }
                    }
                    *reinterpret_cast<uint4*>(line + ((c ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (SPLIT)
                        *reinterpret_cast<uint4*>(line + 2 * kChunkBytes + ((c ^ (row & 7)) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                }
            }
            l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> 16-bit [B][T][C]; the row sum is the two key halves' partial sums ----
        mbar_wait(pv_done, (nkv - 1) & 1);   // the last P V has retired: O is final and the P tile is free
        tc_fence_after();
        float* lx = reinterpret_cast<float*>(sP);
        lx[hf * 128 + row] = l_run;
        pair_sync();
        const float inv_l = 1.f / (lx[row] + lx[128 + row]);
        const int qrow = q0 + row;
        uint16_t* orow = p.out + ((size_t)b * p.T + qrow) * p.C + (size_t)h * p.d;
        float* orow32 = reinterpret_cast<float*>(p.out) + ((size_t)b * p.T + qrow) * p.C + (size_t)h * p.d;   // SPLIT: fp32 output
#pragma unroll 1
        for (int c0 = hf * oc_half; c0 < (hf + 1) * oc_half; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_addr + c0, o);
            tmem_ld_wait();
            if (SPLIT) {
                if (qrow < p.T) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (c0 + i + 4 <= p.d)
                            *reinterpret_cast<float4*>(orow32 + c0 + i) =
                                make_float4(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l,
                                            __uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
                }
            } else if (qrow < p.T) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    if (c0 + i + 8 <= p.d) {
                        uint32_t w[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            w[k] = pack2a<FMT>(__uint_as_float(o[i + 2 * k]) * inv_l, __uint_as_float(o[i + 2 * k + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c0 + i) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

}  // namespace ndb

using namespace ndb;

static int launch_attention(const void* qkv, const void* qkv_lo, void* out, int B, int T, int heads, int d, int order,
                            int operand_format, void* stream) {
    NDB_CHECK_ARG(qkv && out && B > 0 && T > 0 && heads > 0);
    NDB_CHECK_ARG(d % 8 == 0 && d >= 8 && d <= 256);
    NDB_CHECK_ARG(order == 0 || order == 1);
    NDB_CHECK_ARG(operand_format == 0 || operand_format == 1);
    const bool split = qkv_lo != nullptr;
    if (split) {
        NDB_CHECK_ARG(operand_format == 0);
        if (d > 64) {
            std::fprintf(stderr, "[ndb200] fp32-class attention supports head dims up to 64 (got %d)\n", d);
            return NDB_ERR_UNSUPPORTED;
        }
    }
    const int C = heads * d;
    AttnArgs a{};
    a.B = B; a.T = T; a.heads = heads; a.d = d; a.order = order;
    a.dp_chunks = (d + 63) / 64;
    const int Dp = a.dp_chunks * 64;
    const int tw = split ? 2 : 1;
    a.kv_stages = (a.dp_chunks <= 2 && !split) ? 2 : 1;
    // smem: Q + kv_stages*(K+V) + P(2 chunks) (all doubled in split mode) + barriers
    size_t smem = (size_t)tw * (a.dp_chunks * (1 + 2 * a.kv_stages) + 2) * kChunkBytes + 640;
    if (smem > 227 * 1024) {
        a.kv_stages = 1;
        smem = (size_t)tw * (a.dp_chunks * 3 + 2) * kChunkBytes + 640;
    }
    NDB_CHECK_ARG(smem <= 227 * 1024);
    a.fmt = operand_format;
    const int cols = kAttnBN + Dp;
    a.tmem_cols = cols <= 256 ? 256 : 512;
    a.idesc_qk = make_idesc_f16(kAttnBM, kAttnBN, operand_format, 0, 0);
    a.idesc_pv = make_idesc_f16(kAttnBM, Dp, operand_format, 0, 1);
    a.scale_log2e = (float)(1.4426950408889634 / sqrt((double)d));
    a.out = (uint16_t*)out;
    a.C = C;
#ifdef NDB_BRINGUP
    a.dbg = std::getenv("NDB_ATTN_DBG") ? std::atoi(std::getenv("NDB_ATTN_DBG")) : 0;
#endif

    CUtensorMap map, map_lo;
    {
        // view [B][T][3*heads][d]
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)(3 * heads), (uint64_t)T, (uint64_t)B};
        uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)3 * C * 2, (uint64_t)T * 3 * C * 2};
        uint32_t box[4] = {64u, 1u, (uint32_t)kAttnBM, 1u};
        int rc = encode_tmap_16bit(&map, qkv, 4, dims, strides, box);
        if (rc != NDB_OK) return rc;
        rc = encode_tmap_16bit(&map_lo, split ? qkv_lo : qkv, 4, dims, strides, box);
        if (rc != NDB_OK) return rc;
    }
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        cudaFuncSetAttribute(attention_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attention_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(attention_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_once.done();
    }
    dim3 grid((T + kAttnBM - 1) / kAttnBM, heads, B);
    if (split)
        launch_kernel(attention_kernel<0, true>, grid, dim3(kAttnThreads), smem, (cudaStream_t)stream, 1, map, map_lo, a);
    else if (operand_format == 0)
        launch_kernel(attention_kernel<0, false>, grid, dim3(kAttnThreads), smem, (cudaStream_t)stream, 1, map, map_lo, a);
    else
        launch_kernel(attention_kernel<1, false>, grid, dim3(kAttnThreads), smem, (cudaStream_t)stream, 1, map, map_lo, a);
    return launch_status("ndb_attention");
}

extern "C" int ndb_attention(const void* qkv, void* out, int B, int T, int heads, int d, int order, int operand_format,
                             void* stream) {
    return launch_attention(qkv, nullptr, out, B, T, heads, d, order, operand_format, stream);
}

extern "C" int ndb_attention_split(const void* qkv_hi, const void* qkv_lo, float* out, int B, int T, int heads, int d, int order,
                                   void* stream) {
    NDB_CHECK_ARG(qkv_lo != nullptr);
    return launch_attention(qkv_hi, qkv_lo, out, B, T, heads, d, order, 0, stream);
}
