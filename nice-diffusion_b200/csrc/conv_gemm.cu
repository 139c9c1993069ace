// Implicit-GEMM convolution (3x3 pad 1 / 1x1) on the 5th-gen tensor cores.
//
//   M = N*H*W output pixels, tiled as boxes of TN images x TH rows x TW columns (<= 128 pixels),
//   N = output channels (BLOCK_N <= 256 per tile), K = sum over segments of taps * C.
//
// Data path: activations stay NHWC fp16 in HBM and reach shared memory as TMA 4-D boxes {64ch, w, h, n} in the K-major
// SWIZZLE_128B operand layout (image borders = TMA out-of-bounds zero fill, so there is no im2col buffer and no padding
// pass). Three operand loops, chosen per layer (ndb_conv_gemm):
//   halo     (3x3, tiles of 8 columns x 16 rows): ONE box of (TH+2) x (TW+2) pixels per 64-channel block serves all nine
//            taps -- the MMA of tap (dy, dx) reads 128 rows through a descriptor that starts (dy*(TW+2) + dx) * 128 bytes
//            into the box (off the 1024-byte swizzle atom: the swizzle is a function of the smem address, which TMA write
//            and MMA read share) with a stride byte offset of (TW+2) * 128;
//   vertical (3x3, other tile widths that are multiples of 8): one (TH+2)-row box per (dx, channel block), the three dy
//            windows start whole atoms apart;
//   classic  (1x1 layers, tiles that hold several images, fp32-class mode): one {64ch, TW, TH, TN} box per (tap, channel
//            block), activation and weight tile of a K block share a pipeline stage.
// In the first two, boxes and weight tiles ({64, BLOCK_N, 1} boxes of the [tap][cout][cin] weights, one per K block) travel
// through two independent rings and the K loop walks a host-built list of units (= boxes). One elected thread issues
// tcgen05.mma (128 x BLOCK_N x 16, fp32 accumulate in TMEM); accumulators are double-buffered in TMEM so the epilogue of
// tile i overlaps the main loop of tile i+1. Persistent CTAs, one per SM.
//
// Warp roles (384 threads): warp 0 = activation producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue,
// warp 10 = GroupNorm-statistics reducer (only active when chan_stats is requested), warp 11 = weight producer.
// Epilogue: each warp owns 32 accumulator rows (its TMEM lane quarter) and half of the tile's 32-column chunks.
//   bulk path (NHWC outputs; tiles have power-of-two extents, so the 32 rows of a warp always form a {w, h} pixel box): the
//   chunk is pulled out of TMEM (thread = row; the read of chunk i+1 is in flight while chunk i is processed), bias and
//   -- when there is one -- the residual, which a TMA load has placed in the same 4 KB SWIZZLE_128B staging tile, are added
//   in that thread = row form with conflict-free 16-byte accesses; when GroupNorm statistics are wanted the warp then
//   walks the tile with lanes along channels for the column sums; ONE TMA store writes the box back.
//   legacy path (NCHW head, 2x up/down-sampled residuals): padded transpose tile + per-row stores.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <vector>

namespace ndb {

#ifdef NDB_BRINGUP
#define NDB_DBG_BIT(p, b) ((p).dbg & (b))
// phase trace of the last launch (bring-up builds only): g_trace[cta][phase] = %globaltimer when that CTA reached the phase
__device__ unsigned long long g_trace[160][12];
#define NDB_TRACE(ph)                                                                            \
    do {                                                                                         \
        unsigned long long t_;                                                                   \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                   \
        if (blockIdx.x < 160) g_trace[blockIdx.x][ph] = t_;                                      \
    } while (0)
#else
#define NDB_DBG_BIT(p, b) 0
#define NDB_TRACE(ph) do {} while (0)
#endif

constexpr int kMaxSeg = NDB_MAX_SEGMENTS;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                        // 64 x 16-bit = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kThreads = 384;                       // 12 warps: TMA (A), MMA, 8 epilogue, 1 statistics, TMA (B)
constexpr int kEpiWarps = 8;
constexpr int kStageFloats = 32 * 33;               // padded 32x32 transpose tile per epilogue warp
constexpr int kMaxStages = 12;
constexpr int kMaxUnits = 96;
constexpr int kBarBytes = 384;                      // mbarriers + TMEM pointer, in front of the epilogue row tables

struct TmapPack {
    CUtensorMap a[kMaxSeg];
    CUtensorMap b[kMaxSeg];
    CUtensorMap out;    // bulk epilogue: {32 channels, box_w, box_h, 1} boxes of the NHWC output
    CUtensorMap res;    // bulk epilogue with a same-size residual: the same boxes of the fp32 NHWC residual tensor
};

struct ConvArgs {
    int N, H, W;
    int TW, TH, TN;
    int tiles_w, tiles_h, tiles_n;
    int num_m_tiles, num_n_tiles;
    int block_n;
    int n_stride;              // output channels between consecutive N tiles (= block_n, or less: overlapping tiles)
    int cout;
    int num_seg;
    int seg_taps[kMaxSeg];
    int seg_chunks[kMaxSeg];
    int total_kblocks;
    int stages;
    uint32_t a_box_bytes;
    uint32_t tmem_cols;
    int acc_stride;   // TMEM columns between the two accumulator stages
    uint32_t idesc;
    const float* bias;
    const float* residual;
    int res_mode;
    int out_kind;
    void* out;
    int fmt;
    float out_scale;           // accumulators are multiplied by this (a power of two: exact) before bias / residual
    int acc_chunk;             // kChunk kernels: K blocks per accumulation chunk
    // fused GroupNorm statistics: per (image, CTA slot, channel) partial (sum, sum of squares) of the stored output
    float2* stat_parts;        // [N][stat_smax][cout] or nullptr
    int stat_smax;             // partial slots per image
    int tiles_per_group;       // tiles (m x n) covering one group of TN images
    int b_rows;                // weight rows per CTA per stage (block_n, or block_n / 2 in CTA-pair mode)
    int epi_bulk;              // 1: staged TMA-store epilogue
    int epi_bufs;              // staging tiles per epilogue warp (2 for short-K GEMMs whose epilogue is on the critical path)
    int epi_region;            // bytes of the staging region
    int split_prod;            // 1: warp 11 issues the weight (B) loads, warp 0 only the activation (A) loads
    // vertical tap reuse (3x3 segments): one activation box of TH + 2 image rows per (dx, channel block) serves the three
    // dy taps -- their windows start TW pixel rows (a multiple of the 1024-byte swizzle atom) apart inside it
    int vreuse;
    int a_slots, b_slots;      // ring depths (activation boxes / weight tiles)
    int a_slot_bytes;          // ring pitch: (TH + 2) * TW * 128 (halo: (TH + 2) * (TW + 2) * 128 rounded up to 1024)
    uint32_t a_k3_bytes;       // bytes of a 3x3 segment's activation box
    int ring_bytes;            // operand rings in front of the epilogue region
    // reuse modes: the K loop as a list of UNITS = one activation box each: (segment, channel block, dx index); a 3x3 unit is
    // 9 (halo) or 3 (vertical) K blocks, a 1x1 unit one. The host interleaves the 1x1 units among the 3x3 ones: a plain
    // tile occupies a box slot for one K block only, so a run of them is bound by the TMA round trip over the few slots;
    // spread out, they are requested a 3x3 unit's worth of MMA time ahead.
    int n_units;
    uint8_t unit_seg[kMaxUnits], unit_ch[kMaxUnits], unit_dx[kMaxUnits];
    int dbg;                   // bring-up switches, compiled in with -DNDB_BRINGUP only (env NDB_DBG; results are garbage):
                               // 1 no TMA loads, 2 no MMAs, 4 no TMA stores, 8 no staging writes
};

// Persistent CTAs own CONTIGUOUS tile ranges: CTA c processes tiles [tile_start(c), tile_start(c+1)). Tiles are ordered
// (image group, row block, column block, n tile), so a CTA sweeps an image's tiles back to back and can keep that
// image's per-channel sums on chip; an image is shared by few CTAs (its "slots").
__host__ __device__ inline int tile_start(int c, int T, int G) { return (int)(((long long)c * T) / G); }
__host__ __device__ inline int cta_of_tile(int t, int T, int G) {
    int c = (int)((((long long)t + 1) * G - 1) / T);
    while (c > 0 && tile_start(c, T, G) > t) --c;
    while (c + 1 < G && tile_start(c + 1, T, G) <= t) ++c;
    return c;
}

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

// k2 = CTA-pair mode (cta_group::2): two CTAs of a cluster compute adjacent 128-row M tiles of the same N tile with ONE
// 256 x BLOCK_N tcgen05.mma issued by the leader; each CTA stages its own A tile and only HALF of the weight tile, which
// halves the B traffic per SM (L2 -> smem writes and smem -> tensor-core reads) -- the single-CTA kernel is shared-memory
// bandwidth bound at ~55 % tensor utilisation.
// kChunk (fp32-class mode): the tensor core accumulates in fp32 with TRUNCATION -- measured on B200: the result shrinks by
// 5.05e-9 per element of K, -7e-5 relative at K = 13824 -- so the K range is cut into chunks of p.acc_chunk K blocks; every
// chunk starts from a zeroed accumulator (the two TMEM stages alternate) and the epilogue warps add the finished chunks into
// registers with ordinary round-to-nearest adds while the next chunk is being multiplied. BLOCK_N <= 192 (3 x 32 registers).
template <bool kStats, bool k2, bool kChunk = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ TmapPack maps, const ConvArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    // (offset arithmetic on the __shared__ pointer, not an integer round trip: keeps the accesses below LDS/STS instead of
    //  generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int b_tile_bytes = p.b_rows * 128;
    const int stage_bytes = kATileBytes + b_tile_bytes;
    uint8_t* epi_base = smem + p.ring_bytes;   // 1024-aligned: per-warp staging tiles (bulk 4 KB / legacy 33x32 floats)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_base + p.epi_region);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* stat_full_bar = tmem_empty_bar + 2;
    uint64_t* stat_empty_bar = stat_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(stat_empty_bar + 2);
    uint64_t* a_full_bar = stat_empty_bar + 3;       // vreuse: activation ring (<= 4 slots)
    uint64_t* a_empty_bar = a_full_bar + 4;
    long long* epi_rows = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(full_bar) + kBarBytes);
    float2* part_smem = reinterpret_cast<float2*>(epi_rows + kEpiWarps * 128);  // [2 buffers][4 quarters][block_n]
    float2* stat_acc = part_smem + 2 * 4 * p.block_n;                           // [TN][cout] (statistics warp only)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) NDB_TRACE(0);
    // work items: (m tile, n tile) per CTA, or (pair of adjacent m tiles, n tile) per CTA pair
    const uint32_t rank = k2 ? cluster_ctarank() : 0u;
    const int worker = k2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_workers = k2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int num_tiles = (k2 ? ((p.num_m_tiles + 1) >> 1) : p.num_m_tiles) * p.num_n_tiles;
    const int tile_begin = tile_start(worker, num_tiles, n_workers);
    const int tile_end = tile_start(worker + 1, num_tiles, n_workers);
    auto m_tile_of = [&](int item) { return k2 ? 2 * (item / p.num_n_tiles) + (int)rank : item / p.num_n_tiles; };

    if (warp == 0) {
        // descriptor prefetches and barrier initialisation spread over the lanes (one thread doing all ~50 of them was
        // 0.5 us of every launch's prologue)
        if (lane < p.num_seg) {
            prefetch_tmap(&maps.a[lane]);
            prefetch_tmap(&maps.b[lane]);
        }
        if (lane == 31 && p.epi_bulk) {
            prefetch_tmap(&maps.out);
            if (p.res_mode == NDB_RES_SAME) prefetch_tmap(&maps.res);
        }
        if (lane < p.stages) {
            // pair mode: both CTAs' producers arrive on the leader's barrier; split producers arrive separately
            mbar_init(&full_bar[lane], (k2 ? 2 : 1) * ((p.split_prod && !p.vreuse) ? 2 : 1));
            mbar_init(&empty_bar[lane], 1);
        }
        if (p.vreuse && lane >= 16 && lane < 16 + p.a_slots) {
            mbar_init(&a_full_bar[lane - 16], k2 ? 2 : 1);
            mbar_init(&a_empty_bar[lane - 16], 1);
        }
        if (lane >= 8 && lane < 10) {
            const int s = lane - 8;
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], k2 ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
            mbar_init(&stat_full_bar[s], kEpiWarps);
            mbar_init(&stat_empty_bar[s], 1);
        }
        if (p.epi_bulk && lane >= 8) {        // 3 * kEpiWarps = 24 barriers on lanes 8..31
            // per epilogue warp: three "residual box has landed" barriers (inside its row-table slot)
            const int i = lane - 8;
            mbar_init(reinterpret_cast<uint64_t*>(epi_rows + (i / 3) * 128 + 64) + (i % 3), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (k2) { tmem_alloc_2sm(tmem_ptr_smem, p.tmem_cols); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_ptr_smem, p.tmem_cols); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (k2) cluster_sync_all();    // the peer's barriers must be initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) NDB_TRACE(1);
    pdl_launch_dependents();   // the next kernel may begin its prologue as SM resources free up
    pdl_wait();                // everything above overlapped the previous kernel's tail; its outputs are visible from here
    if (threadIdx.x == 0) NDB_TRACE(2);

    if (warp == 0 || warp == kThreads / 32 - 1) {
        // ===================== TMA producer(s) =====================
        // The WHOLE warp runs the (warp-uniform) loop and one elected lane issues: with the loop inside `if (lane == 0)`
        // ptxas cannot prove the descriptor / coordinate operands uniform and wraps every UTMALDG / UTCHMMA in an
        // ELECT + R2UR "waterfall" loop (~150 cycles per instruction, measured).
        // Two producer warps: warp 0 streams the activation boxes, the last warp the weight tiles (in the reuse modes below
        // these are two independent rings with their own barriers; in the classic mode both arrive on the stage's barrier
        // with their own byte count -- 0.6 % faster than one warp issuing both, measured on the whole step).
        const bool do_a = !p.split_prod || warp == 0;
        const bool do_b = !p.split_prod || warp != 0;
        if (p.vreuse) {
            // Two independent rings. Warp 0 streams activation boxes: per 3x3 segment one box of TH + 2 rows per (dx, channel
            // block), per 1x1 segment the plain tile. The last warp streams one weight tile per K block, in the order the MMA
            // warp consumes them: segment, dx, channel block, dy.
            int slot = 0;
            uint32_t phase = 0;
            uint8_t* ring_b = smem + p.a_slots * p.a_slot_bytes;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int n_tile = tile % p.num_n_tiles;
                int m_tile = m_tile_of(tile);
                const int tw_i = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th_i = m_tile % p.tiles_h;
                const int tn_i = m_tile / p.tiles_h;
                const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
                const int b_row0 = n_tile * p.n_stride + (int)rank * p.b_rows;
                const bool halo = p.vreuse == 2;      // one (TW + 2) x (TH + 2) box per channel block serves all nine taps
                for (int u = 0; u < p.n_units; ++u) {
                    const int s = p.unit_seg[u], ch = p.unit_ch[u], dxi = p.unit_dx[u];
                    const bool k3 = p.seg_taps[s] == 9;
                    if (warp == 0) {
                        const uint32_t a_bytes = k3 ? (uint32_t)p.a_k3_bytes : p.a_box_bytes;
                        mbar_wait(&a_empty_bar[slot], phase ^ 1);
                        if (elect_one()) {
                            uint8_t* sa = smem + slot * p.a_slot_bytes;
                            const int cw = w0 + (k3 ? (halo ? -1 : dxi - 1) : 0), chh = h0 - (k3 ? 1 : 0);
                            if (k2) {
                                const uint32_t fb = mapa_u32(smem_u32(&a_full_bar[slot]), 0);
                                mbar_arrive_expect_tx_cluster(fb, a_bytes);
                                tma_load_4d_2sm(sa, &maps.a[s], fb, ch * kBlockK, cw, chh, n0);
                            } else {
                                mbar_arrive_expect_tx(&a_full_bar[slot], a_bytes);
                                tma_load_4d(sa, &maps.a[s], &a_full_bar[slot], ch * kBlockK, cw, chh, n0);
                            }
                        }
                        __syncwarp();
                        if (++slot == p.a_slots) { slot = 0; phase ^= 1; }
                    } else {
                        for (int t = 0; t < (k3 ? (halo ? 9 : 3) : 1); ++t) {
                            mbar_wait(&empty_bar[slot], phase ^ 1);
                            if (elect_one()) {
                                uint8_t* sb = ring_b + slot * b_tile_bytes;
                                const int tap = k3 ? (halo ? t : t * 3 + dxi) : 0;
                                if (k2) {
                                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[slot]), 0);
                                    mbar_arrive_expect_tx_cluster(fb, (uint32_t)b_tile_bytes);
                                    tma_load_3d_2sm(sb, &maps.b[s], fb, ch * kBlockK, b_row0, tap);
                                } else {
                                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)b_tile_bytes);
                                    tma_load_3d(sb, &maps.b[s], &full_bar[slot], ch * kBlockK, b_row0, tap);
                                }
                            }
                            __syncwarp();
                            if (++slot == p.b_slots) { slot = 0; phase ^= 1; }
                        }
                    }
                }
            }
        } else if (p.split_prod || warp == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int n_tile = tile % p.num_n_tiles;
                int m_tile = m_tile_of(tile);
                const int tw_i = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th_i = m_tile % p.tiles_h;
                const int tn_i = m_tile / p.tiles_h;
                const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;   // n0 >= N for the odd leftover tile: all zero
                const int b_row0 = n_tile * p.n_stride + (int)rank * p.b_rows;
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
                            if (elect_one()) {
                                const uint32_t tx = (do_a ? p.a_box_bytes : 0u) + (do_b ? (uint32_t)b_tile_bytes : 0u);
                                if (NDB_DBG_BIT(p, 1)) {
                                    if (k2) mbar_arrive_cluster(mapa_u32(smem_u32(&full_bar[stage]), 0));
                                    else mbar_arrive(&full_bar[stage]);
                                } else if (k2) {
                                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);   // leader's barrier
                                    mbar_arrive_expect_tx_cluster(fb, tx);
                                    if (do_a) tma_load_4d_2sm(sa, &maps.a[s], fb, ch * kBlockK, w0 + dx, h0 + dy, n0);
                                    if (do_b) tma_load_3d_2sm(sb, &maps.b[s], fb, ch * kBlockK, b_row0, tap);
                                } else {
                                    mbar_arrive_expect_tx(&full_bar[stage], tx);
                                    if (do_a) tma_load_4d(sa, &maps.a[s], &full_bar[stage], ch * kBlockK, w0 + dx, h0 + dy, n0);
                                    if (do_b) tma_load_3d(sb, &maps.b[s], &full_bar[stage], ch * kBlockK, b_row0, tap);
                                }
                            }
                            __syncwarp();
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
        // ===================== MMA issuer (pair mode: the leader CTA only) =====================
        if (rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            // K-major SWIZZLE_128B descriptor (see make_desc_k_sw128): constant high word, low word = (address >> 4) | LBO
            constexpr uint64_t kDescHi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61) | (uint64_t(1) << 16);
            const uint32_t desc_lo0 = (smem_u32(smem) & 0x3FFFF) >> 4;
            const uint32_t desc_step = (uint32_t)stage_bytes >> 4;
            const int chunk_kb = kChunk ? p.acc_chunk : p.total_kblocks;
            if (!kChunk && p.vreuse) {
                int as = 0, bs = 0;
                uint32_t aph = 0, bph = 0;
                const uint32_t a_step = (uint32_t)p.a_slot_bytes >> 4;
                const uint32_t b_lo0 = desc_lo0 + (uint32_t)p.a_slots * a_step;
                const uint32_t b_step = (uint32_t)b_tile_bytes >> 4;
                const bool halo = p.vreuse == 2;
                // one image row of the box: TW (halo: TW + 2) pixel rows of 128 B. Halo boxes: the eight pixels of an image
                // row are one swizzle atom, atoms are (TW + 2) * 128 bytes apart (the descriptor's stride byte offset) and a
                // tap's window starts dx pixel rows into the box, i.e. NOT on a 1024-byte boundary: the swizzle is a
                // function of the shared-memory address, which the TMA write and the MMA read agree on.
                const uint32_t dy_step = (uint32_t)((p.TW + (halo ? 2 : 0)) * 128) >> 4;
                const uint64_t kDescHiA = halo ? ((uint64_t(dy_step) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61) | (uint64_t(1) << 16)) : kDescHi;
                for (int tile = tile_begin; tile < tile_end; ++tile) {
                    mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
                    uint32_t started = 0;
                    for (int u = 0; u < p.n_units; ++u) {
                        const bool k3 = p.seg_taps[p.unit_seg[u]] == 9;
                        const int ntap = k3 ? (halo ? 9 : 3) : 1;
                        const uint64_t hiA = k3 ? kDescHiA : kDescHi;
                        mbar_wait(&a_full_bar[as], aph);
                        const uint32_t a_lo = desc_lo0 + (uint32_t)as * a_step;
                        for (int dyi = 0; dyi < ntap; ++dyi) {
                            mbar_wait(&full_bar[bs], bph);
                            tc_fence_after();
                            if (lane == 0 && tile == tile_begin && !started) NDB_TRACE(3);
                            if (elect_one()) {
                                // halo: tap t = (dy, dx) starts dy image rows + dx pixels into the box
                                const uint32_t a_off = halo ? (uint32_t)(dyi / 3) * dy_step + (uint32_t)(dyi % 3) * 8u : (uint32_t)dyi * dy_step;
                                const uint64_t adesc = hiA | (a_lo + a_off);
                                const uint64_t bdesc = kDescHi | (b_lo0 + (uint32_t)bs * b_step);
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k) {
                                    if (k2) umma_f16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (started | k) ? 1u : 0u);
                                    else umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (started | k) ? 1u : 0u);
                                }
                                if (k2) umma_commit_2sm(&empty_bar[bs]); else umma_commit(&empty_bar[bs]);
                                if (dyi == ntap - 1) {      // the box has served its last tap
                                    if (k2) umma_commit_2sm(&a_empty_bar[as]); else umma_commit(&a_empty_bar[as]);
                                }
                            }
                            __syncwarp();
                            started = 1;
                            if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
                        }
                        if (++as == p.a_slots) { as = 0; aph ^= 1; }
                    }
                    if (elect_one()) {
                        if (k2) umma_commit_2sm(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
                    }
                    __syncwarp();
                    if (lane == 0 && tile + 1 == tile_end) NDB_TRACE(4);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            } else
            for (int tile = tile_begin; tile < tile_end; ++tile)
            for (int kb0 = 0; kb0 < p.total_kblocks; kb0 += chunk_kb) {
                const int kb1 = (kb0 + chunk_kb < p.total_kblocks) ? kb0 + chunk_kb : p.total_kblocks;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
                // The issue loop is the critical path for narrow N tiles (4 MMAs of 48..96 cycles per K block): the operand
                // descriptors advance by one multiply-add per stage.
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0 && tile == tile_begin && kb == 0) NDB_TRACE(3);
                    const uint32_t a_lo = desc_lo0 + (uint32_t)stage * desc_step;
                    if (elect_one()) {
                        const uint64_t adesc = kDescHi | a_lo;
                        const uint64_t bdesc = kDescHi | (a_lo + (uint32_t)(kATileBytes >> 4));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            if (NDB_DBG_BIT(p, 2)) break;
                            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the >>4 address field
                            if (k2) umma_f16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, ((kb - kb0) | k) ? 1u : 0u);
                            else umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, ((kb - kb0) | k) ? 1u : 0u);
                        }
                        // frees the smem slot (in both CTAs of a pair) once these MMAs retire
                        if (k2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if (elect_one()) {
                    if (k2) umma_commit_2sm(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
                if (lane == 0 && tile + 1 == tile_end && kb1 == p.total_kblocks) NDB_TRACE(4);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp < 2 + kEpiWarps) {
        // ===================== epilogue warps =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;       // which half of the column chunks
        const int row = q * 32 + lane;          // accumulator row == pixel index inside the tile
        const int nch = (p.block_n + 31) >> 5;
        const int ch_begin = half ? (nch + 1) / 2 : 0;
        const int ch_end = half ? nch : (nch + 1) / 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        int stat_buf = 0;
        uint32_t stat_phase = 0;
        const int pix_per_img = p.TH * p.TW;
        const float* __restrict__ resp = p.residual;
        // position of this thread's row inside a tile
        const int tn = row / pix_per_img;
        const int rem = row - tn * pix_per_img;
        const int th = rem / p.TW;
        const int tw = rem - th * p.TW;
        // kChunk: running sums of this warp's (<= 3) column chunks; all chunks of a tile but the last are drained here
        float csum[kChunk ? 3 : 1][32];
        const int n_acc_chunks = kChunk ? (p.total_kblocks + p.acc_chunk - 1) / p.acc_chunk : 1;
        auto drain_chunks = [&]() {
            if (!kChunk) return;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) csum[c][i] = 0.f;
            for (int cc = 0; cc + 1 < n_acc_chunks; ++cc) {
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * p.acc_stride;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (ch_begin + c < ch_end) {
                        uint32_t v[32];
                        tmem_ld32(t_row + (ch_begin + c) * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) csum[c][i] += __uint_as_float(v[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (k2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
                    else mbar_arrive(&tmem_empty_bar[acc]);
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        };
        // adds the running sum of column chunk `ch` to the last chunk's accumulator values
        auto add_csum = [&](uint32_t (&racc)[32], int ch) {
            if (!kChunk) return;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (ch - ch_begin == c) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) racc[i] = __float_as_uint(__uint_as_float(racc[i]) + csum[c][i]);
                }
        };
        if (p.epi_bulk) {
            // ------------------------------------------------------------------------------------------------
            // bulk path
            // ------------------------------------------------------------------------------------------------
            // Per warp and 32-column chunk: the accumulator rows come out of TMEM (thread = row), bias (stashed in shared
            // memory once per tile, before the accumulator is waited for) and -- if there is one -- the residual are added in
            // that same thread = row form, the result sits in a 4 KB SWIZZLE_128B staging tile and ONE TMA store writes the
            // {32 ch, w, h} box. The RESIDUAL box arrives by TMA as well, straight into the staging tile it will be added in:
            // three tiles per warp rotate through [loading chunk i+1 | being processed: chunk i | being stored: chunk i-1],
            // so two boxes per warp are always in flight without holding a register (the per-row __ldg version kept 64
            // registers busy, could only run one chunk ahead and capped the residual stream at ~3 TB/s: the proj GEMMs of the
            // attention blocks took twice their HBM time).
            const bool has_res = p.res_mode == NDB_RES_SAME;
            const float osc = p.out_scale;
            const int nbuf = p.epi_bufs;        // residual: 3 (short K: the epilogue is the critical path) or 2 (long K: the
                                                // staging space goes to the operand pipeline instead)
            uint8_t* sbuf0 = epi_base + (warp - 2) * 4096 * nbuf;
            float* bias_s = reinterpret_cast<float*>(epi_rows + (warp - 2) * 128);              // [<= 4 chunks][32]
            uint64_t* res_bar = reinterpret_cast<uint64_t*>(epi_rows + (warp - 2) * 128 + 64);   // [3]
            // the box covered by this warp's 32 rows
            const int r0 = q * 32;
            const int btn = r0 / pix_per_img;
            const int brem = r0 - btn * pix_per_img;
            const int bth = brem / p.TW;
            const int btw = brem - bth * p.TW;
            const bool box_on = btn < p.TN;
            const bool use_res = has_res && box_on;
            const bool any = tile_begin < tile_end && ch_begin < ch_end;

            // residual box of chunk `ch` of work item `tile` -> staging tile b (elected lane)
            auto load_res = [&](int tile, int ch, int b) {
                int m_tile = m_tile_of(tile);
                const int tw_i = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th_i = m_tile % p.tiles_h;
                const int tn_i = m_tile / p.tiles_h;
                mbar_arrive_expect_tx(&res_bar[b], 4096u);
                tma_load_4d(sbuf0 + b * 4096, &maps.res, &res_bar[b], (tile % p.num_n_tiles) * p.n_stride + ch * 32,
                            tw_i * p.TW + btw, th_i * p.TH + bth, tn_i * p.TN + btn);
            };
            // position of the chunk two ahead of (tile, ch) in this warp's sequence
            auto advance = [&](int& tile, int& ch) {
                if (++ch == ch_end) { ch = ch_begin; ++tile; }
            };
            if (use_res && any) {
                if (elect_one()) {
                    int t1 = tile_begin, c1 = ch_begin;
                    load_res(t1, c1, 0);
                    advance(t1, c1);
                    if (t1 < tile_end) load_res(t1, c1, 1);
                }
                __syncwarp();
            }
            int ci = 0;                 // chunks processed by this warp
            uint32_t res_phase = 0;     // bit b: parity of the next completion of res_bar[b]
            int sb = 0;
            // bias values of this lane's columns, fetched one tile ahead (an L2 round trip per tile sat on the epilogue's
            // critical path of the short-K GEMMs)
            float bias_r[4];
            auto fetch_bias = [&](int tile) {
                const int col0n = (tile % p.num_n_tiles) * p.n_stride;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = col0n + (ch_begin + c) * 32 + lane;
                    bias_r[c] = (p.bias != nullptr && ch_begin + c < ch_end && tile < tile_end && col < p.cout) ? __ldg(p.bias + col) : 0.f;
                }
            };
            fetch_bias(tile_begin);
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int col0 = (tile % p.num_n_tiles) * p.n_stride;
                int m_tile = m_tile_of(tile);
                const int tw_i = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th_i = m_tile % p.tiles_h;
                const int tn_i = m_tile / p.tiles_h;
                const int bn = tn_i * p.TN + btn, bh = th_i * p.TH + bth, bw = tw_i * p.TW + btw;
                unsigned vmask = 0;
                if (kStats) {
                    const int n = tn_i * p.TN + tn, h = th_i * p.TH + th, w = tw_i * p.TW + tw;
                    vmask = __ballot_sync(0xffffffffu, (tn < p.TN) && (n < p.N) && (h < p.H) && (w < p.W));
                }
                // this tile's bias values for this warp's chunks -> shared memory, while the MMAs of the tile still run
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (ch_begin + c < ch_end) bias_s[c * 32 + lane] = bias_r[c];
                __syncwarp();
                fetch_bias(tile + 1);

                drain_chunks();
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                if (warp == 2 && lane == 0 && tile + 1 == tile_end) NDB_TRACE(5);
                if (kStats) mbar_wait(&stat_empty_bar[stat_buf], stat_phase ^ 1);   // part_smem buffer free
                const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * p.acc_stride;
                // one chunk: `racc` was requested from TMEM by the previous call (or before the loop); the request for the
                // next chunk goes out as soon as this one has landed, so its latency hides under this chunk's arithmetic,
                // and the accumulator stage is handed back to the MMA warp the moment its last chunk is in registers
                auto do_chunk = [&](uint32_t (&racc)[32], uint32_t (&rnext)[32], const int ch) {
                    const int c0 = col0 + ch * 32;
                    if (kChunk) tmem_ld32(t_row + ch * 32, racc);      // (no look-ahead next to the 96 running-sum registers)
                    int b;
                    if (has_res) {
                        b = ci % nbuf;
                        if (use_res) {
                            mbar_wait(&res_bar[b], (res_phase >> b) & 1u);    // the residual box is in the staging tile
                            res_phase ^= 1u << b;
                        }
                    } else {
                        b = sb;
                        // the TMA store that last used this staging tile must have drained it
                        if (elect_one()) {
                            if (nbuf == 2) bulk_wait_read_1(); else bulk_wait_read_all();
                        }
                        __syncwarp();
                        if (nbuf == 2) sb ^= 1;
                    }
                    tmem_ld_wait();
                    if (ch + 1 < ch_end) {
                        if (!kChunk) tmem_ld32(t_row + (ch + 1) * 32, rnext);
                    } else {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (k2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));   // the leader's MMA thread waits
                            else mbar_arrive(&tmem_empty_bar[acc]);
                        }
                    }
                    add_csum(racc, ch);
                    uint8_t* sbuf = sbuf0 + b * 4096;
                    float* sf = reinterpret_cast<float*>(sbuf);
                    const float4* bias4 = reinterpret_cast<const float4*>(bias_s + (ch - ch_begin) * 32);
                    if (NDB_DBG_BIT(p, 8)) {
                    } else if (p.out_kind == NDB_OUT_F32_NHWC) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = bias4[j];
                            float4* dst = reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4));
                            // acc * 2^k is exact, so the fma rounds exactly like the plain add it replaces (out_scale = 1)
                            float4 v = make_float4(fmaf(__uint_as_float(racc[4 * j]), osc, b4.x), fmaf(__uint_as_float(racc[4 * j + 1]), osc, b4.y),
                                                   fmaf(__uint_as_float(racc[4 * j + 2]), osc, b4.z), fmaf(__uint_as_float(racc[4 * j + 3]), osc, b4.w));
                            if (use_res) {
                                const float4 rv = *dst;
                                v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                            }
                            *dst = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b0 = bias4[2 * j], b1 = bias4[2 * j + 1];
                            const float v0 = fmaf(__uint_as_float(racc[8 * j]), osc, b0.x), v1 = fmaf(__uint_as_float(racc[8 * j + 1]), osc, b0.y);
                            const float v2 = fmaf(__uint_as_float(racc[8 * j + 2]), osc, b0.z), v3 = fmaf(__uint_as_float(racc[8 * j + 3]), osc, b0.w);
                            const float v4 = fmaf(__uint_as_float(racc[8 * j + 4]), osc, b1.x), v5 = fmaf(__uint_as_float(racc[8 * j + 5]), osc, b1.y);
                            const float v6 = fmaf(__uint_as_float(racc[8 * j + 6]), osc, b1.z), v7 = fmaf(__uint_as_float(racc[8 * j + 7]), osc, b1.w);
                            uint4 w4;
                            if (p.fmt == 0) w4 = make_uint4(pack2<0>(v0, v1), pack2<0>(v2, v3), pack2<0>(v4, v5), pack2<0>(v6, v7));
                            else w4 = make_uint4(pack2<1>(v0, v1), pack2<1>(v2, v3), pack2<1>(v4, v5), pack2<1>(v6, v7));
                            // 64-byte rows, SWIZZLE_64B: 16-byte chunk index ^ bits [1,2] of the row
                            *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = w4;
                        }
                    }
                    if (kStats && !NDB_DBG_BIT(p, 16)) {
                        // the staged values are final: a lanes-along-channels pass READS them for the column sums (for a
                        // 16-bit output these are the sums of the ROUNDED values -- exactly what the consuming norm will read)
                        __syncwarp();
                        const int col = c0 + lane;
                        const unsigned okmask = (col < p.cout) ? vmask : 0u;
                        float cs = 0.f, css = 0.f;
                        if (p.out_kind == NDB_OUT_F32_NHWC) {
                            const int lswz = lane >> 2, lw = lane & 3;
#pragma unroll
                            for (int r = 0; r < 32; ++r) {
                                const float t = sf[r * 32 + (((lswz ^ (r & 7)) << 2) | lw)];
                                const float v = ((okmask >> r) & 1u) ? t : 0.f;
                                cs += v;
                                css += v * v;
                            }
                        } else {
                            const uint16_t* sh = reinterpret_cast<const uint16_t*>(sbuf);
                            const int lchunk = lane >> 3, lw = lane & 7;
#pragma unroll
                            for (int r = 0; r < 32; ++r) {
                                const uint16_t bits = sh[r * 32 + (((lchunk ^ ((r >> 1) & 3)) << 3) | lw)];
                                float t;
                                if (p.fmt == 0) t = __half2float(__ushort_as_half(bits));
                                else t = __bfloat162float(__ushort_as_bfloat16(bits));
                                const float v = ((okmask >> r) & 1u) ? t : 0.f;
                                cs += v;
                                css += v * v;
                            }
                        }
                        part_smem[(stat_buf * 4 + q) * p.block_n + ch * 32 + lane] = make_float2(cs, css);
                    }
                    fence_proxy_async();      // generic-proxy writes -> visible to the TMA engine
                    __syncwarp();
                    if (elect_one()) {
                        if (box_on && !NDB_DBG_BIT(p, 4)) {
                            tma_store_4d(&maps.out, sbuf, c0, bw, bh, bn);
                            bulk_commit();
                        }
                        if (use_res) {
                            // chunk i+2 goes into the tile chunk i-1 was stored from: all but the store just committed
                            // must have finished reading shared memory
                            int t2 = tile, c2 = ch;
                            advance(t2, c2);
                            advance(t2, c2);
                            if (t2 < tile_end) {
                                if (nbuf == 3) bulk_wait_read_1(); else bulk_wait_read_all();
                                load_res(t2, c2, (ci + 2) % nbuf);
                            }
                        }
                    }
                    __syncwarp();
                    ++ci;
                };
                if (kChunk) {
                    uint32_t r0[32];
                    for (int ch = ch_begin; ch < ch_end; ++ch) do_chunk(r0, r0, ch);
                } else {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(t_row + ch_begin * 32, r0);
                    for (int ch = ch_begin; ch < ch_end; ch += 2) {
                        do_chunk(r0, r1, ch);
                        if (ch + 1 < ch_end) do_chunk(r1, r0, ch + 1);
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                if (kStats) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&stat_full_bar[stat_buf]);   // this warp's column sums are in part_smem
                    if ((stat_buf ^= 1) == 0) stat_phase ^= 1;
                }
            }
            if (warp == 2 && lane == 0) NDB_TRACE(6);
            if (elect_one()) bulk_wait_all();     // shared memory must outlive the last store's reads
            __syncwarp();
            if (warp == 2 && lane == 0) NDB_TRACE(7);
        } else {
        // ----------------------------------------------------------------------------------------------------
        // legacy path
        // ----------------------------------------------------------------------------------------------------
        float* stg = reinterpret_cast<float*>(epi_base) + (warp - 2) * kStageFloats;
        long long* row_out = epi_rows + (warp - 2) * 128;
        long long* row_res = row_out + 32;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int n_tile = tile % p.num_n_tiles;
            int m_tile = m_tile_of(tile);
            const int tw_i = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th_i = m_tile % p.tiles_h;
            const int tn_i = m_tile / p.tiles_h;
            const int n = tn_i * p.TN + tn, h = th_i * p.TH + th, w = tw_i * p.TW + tw;
            const bool valid = (tn < p.TN) && (n < p.N) && (h < p.H) && (w < p.W);
            const int col0 = n_tile * p.n_stride;
            // element offsets of this row's pixel (valid rows only)
            long long out_off = 0, res_off = 0;
            if (valid) {
                const long long pix = ((long long)n * p.H + h) * p.W + w;
                out_off = (p.out_kind == NDB_OUT_F32_NCHW) ? ((long long)n * p.cout * p.H + h) * p.W + w : pix * p.cout;
                if (p.res_mode == NDB_RES_SAME) res_off = pix * p.cout;
                else if (p.res_mode == NDB_RES_UP2)
                    res_off = (((long long)n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1)) * p.cout;
                else if (p.res_mode == NDB_RES_DOWN2)
                    res_off = (((long long)n * (2 * p.H) + 2 * h) * (2 * p.W) + 2 * w) * p.cout;
            }
            const long long down_dx = p.cout, down_dy = (long long)(2 * p.W) * p.cout;
            const unsigned valid_mask = __ballot_sync(0xffffffffu, valid);
            // per-row element offsets, published to the warp through smem (invalid rows point at offset 0 and are
            // masked) so the inner loops are free of shuffles and branches
            __syncwarp();
            row_out[lane] = out_off;
            row_res[lane] = res_off;
            __syncwarp();

            drain_chunks();
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (kStats) mbar_wait(&stat_empty_bar[stat_buf], stat_phase ^ 1);   // part_smem buffer free
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * p.acc_stride;
            for (int ch = ch_begin; ch < ch_end; ++ch) {
                const int cl = ch * 32 + lane;             // column inside the tile owned by this lane (transposed view)
                const int col = col0 + cl;
                const bool colok = (cl < p.block_n) && (col < p.cout);
                const unsigned okmask = colok ? valid_mask : 0u;
                // (1) residual, lanes along channels: 32 independent coalesced loads in flight
                float res[32];
                if (p.res_mode == NDB_RES_NONE) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) res[r] = 0.f;
                } else if (p.res_mode != NDB_RES_DOWN2) {
                    const float* base = resp + col;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        float v = 0.f;
                        if ((okmask >> r) & 1u) v = __ldg(base + row_res[r]);
                        res[r] = v;
                    }
                } else {
                    const float* base = resp + col;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                        if ((okmask >> r) & 1u) {
                            const float* a = base + row_res[r];
                            v0 = __ldg(a); v1 = __ldg(a + down_dx); v2 = __ldg(a + down_dy); v3 = __ldg(a + down_dy + down_dx);
                        }
                        // same association as F.avg_pool2d: sum of the four taps, then * 0.25
                        res[r] = (v0 + v1 + v2 + v3) * 0.25f;
                    }
                }
                // (2) accumulator chunk out of TMEM (thread = row)
                uint32_t racc[32];
                tmem_ld32(t_row + ch * 32, racc);
                tmem_ld_wait();
                add_csum(racc, ch);
                // (3) transpose through the padded per-warp tile
#pragma unroll
                for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = __uint_as_float(racc[i]);
                __syncwarp();
                const float bias_v = (colok && p.bias != nullptr) ? __ldg(p.bias + col) : 0.f;
                // (4) bias + residual, then store with lanes along channels
#pragma unroll
                for (int r = 0; r < 32; ++r) res[r] = fmaf(stg[r * 33 + lane], p.out_scale, bias_v) + res[r];
                if (p.out_kind == NDB_OUT_F32_NHWC) {
                    float* ob = reinterpret_cast<float*>(p.out) + col;
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        if ((okmask >> r) & 1u) ob[row_out[r]] = res[r];
                } else if (p.out_kind == NDB_OUT_F16_NHWC) {
                    uint16_t* ob = reinterpret_cast<uint16_t*>(p.out) + col;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        uint16_t bits;
                        if (p.fmt == 0) { __half hv = __float2half_rn(res[r]); bits = *reinterpret_cast<uint16_t*>(&hv); }
                        else { __nv_bfloat16 hv = __float2bfloat16_rn(res[r]); bits = *reinterpret_cast<uint16_t*>(&hv); }
                        if ((okmask >> r) & 1u) ob[row_out[r]] = bits;
                    }
                } else {
                    float* ob = reinterpret_cast<float*>(p.out) + (long long)col * p.H * p.W;
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        if ((okmask >> r) & 1u) ob[row_out[r]] = res[r];
                }
                if (kStats) {
                    // column sums over this warp's 32 rows come for free in the transposed view (fixed order)
                    float cs = 0.f, css = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float v = ((okmask >> r) & 1u) ? res[r] : 0.f;
                        cs += v;
                        css += v * v;
                    }
                    if (cl < p.block_n) part_smem[(stat_buf * 4 + q) * p.block_n + cl] = make_float2(cs, css);
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (k2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));   // the leader's MMA thread waits
                else mbar_arrive(&tmem_empty_bar[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
            if (kStats) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&stat_full_bar[stat_buf]);   // this warp's column sums are in part_smem
                if ((stat_buf ^= 1) == 0) stat_phase ^= 1;
            }
        }
        }
    } else if (kStats) {
        // ===================== statistics warp =====================
        // Folds the four row-quarters of every finished tile into per-(image, channel) running sums kept in shared
        // memory while this CTA sweeps the tiles of one image group, and writes them out ONCE when it moves on to
        // the next group: slot (this CTA - first CTA of the group) of stat_parts[image]. No atomics, no fences, no
        // cross-CTA traffic inside the kernel; the consuming gn_apply adds the (few) slots of its image in slot order.
        int stat_buf = 0;
        uint32_t stat_phase = 0;
        const int warps_per_slot = 4 / p.TN;                  // TN in {1, 2, 4} (host-checked)
        int cur_group = -1;
        int cur_slot = 0;        // slot this CTA writes for cur_group: computed when the group starts, off the critical path
        // group of images a work item belongs to, and the slot this CTA writes for that group (see stats_slots())
        auto group_of = [&](int item) {
            if (!k2) return item / p.tiles_per_group;
            if (p.TN == 1) return item / (p.tiles_per_group >> 1);      // both CTAs of a pair sweep the same images
            return m_tile_of(item);                                      // one tile == one group of TN images
        };
        auto slot_of = [&](int group) {
            if (!k2) return worker - cta_of_tile(group * p.tiles_per_group, num_tiles, n_workers);
            if (p.TN == 1) return 2 * (worker - cta_of_tile(group * (p.tiles_per_group >> 1), num_tiles, n_workers)) + (int)rank;
            return worker - cta_of_tile((group >> 1) * p.num_n_tiles, num_tiles, n_workers);
        };
        // the tail of the kernel waits for this: two channels (16 bytes) per lane and trip, trips independent
        auto flush = [&](int group, int slot_idx) {
            for (int sl = 0; sl < p.TN; ++sl) {
                const int img = group * p.TN + sl;
                if (img >= p.N) break;
                float2* dst = p.stat_parts + ((size_t)img * p.stat_smax + slot_idx) * p.cout;
                const float2* src = stat_acc + sl * p.cout;
                if ((p.cout & 1) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
#pragma unroll 4
                    for (int c = 2 * lane; c < p.cout; c += 64)
                        *reinterpret_cast<float4*>(dst + c) = *reinterpret_cast<const float4*>(src + c);
                } else {
                    for (int c = lane; c < p.cout; c += 32) dst[c] = src[c];
                }
            }
        };
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int n_tile = tile % p.num_n_tiles;
            const int group = group_of(tile);
            const int col0 = n_tile * p.n_stride;
            if (group != cur_group) {
                if (cur_group >= 0) flush(cur_group, cur_slot);
                for (int i = lane; i < p.TN * p.cout; i += 32) stat_acc[i] = make_float2(0.f, 0.f);
                cur_group = group;
                cur_slot = slot_of(group);
                __syncwarp();
            }
            mbar_wait(&stat_full_bar[stat_buf], stat_phase);
            if (lane == 0 && tile + 1 == tile_end) NDB_TRACE(10);
            const float2* ps = part_smem + stat_buf * 4 * p.block_n;
            for (int sl = 0; sl < p.TN; ++sl) {
                float2* acc_row = stat_acc + sl * p.cout + col0;
                for (int c = lane; c < p.n_stride; c += 32) {      // (columns past n_stride are the next N tile's)
                    float sx = 0.f, sy = 0.f;
                    for (int qq = sl * warps_per_slot; qq < (sl + 1) * warps_per_slot; ++qq) {
                        const float2 v = ps[qq * p.block_n + c];
                        sx += v.x; sy += v.y;
                    }
                    if (col0 + c < p.cout) {
                        float2 a = acc_row[c];
                        a.x += sx; a.y += sy;
                        acc_row[c] = a;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&stat_empty_bar[stat_buf]);
            if ((stat_buf ^= 1) == 0) stat_phase ^= 1;
        }
        if (cur_group >= 0) flush(cur_group, cur_slot);
        if (lane == 0) NDB_TRACE(11);
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) NDB_TRACE(8);
    if (k2) cluster_sync_all();    // the peer may still be reading this CTA's half of the weights / signalling its barriers
    if (warp == 1) {
        tc_fence_after();
        if (k2) tmem_dealloc_2sm(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (threadIdx.x == 32) NDB_TRACE(9);
}

// ------------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [tap][cout_pad][cin_len] 16-bit
// ------------------------------------------------------------------------------------------------------
// part 0: the 16-bit rounding of w * scale; part 1 (fp32-class mode): the fp16 rounding of what part 0 lost
__global__ void pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ dst, int cout, int cin_total,
                                   int taps, int cout_pad, int cin_off, int cin_len, int fmt, float scale, int part) {
    const long long total = (long long)taps * cout_pad * cin_len;
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_len);
        long long r = i / cin_len;
        const int co = (int)(r % cout_pad);
        const int tap = (int)(r / cout_pad);
        float v = 0.f;
        if (co < cout) v = w[((long long)co * cin_total + cin_off + ci) * taps + tap] * scale;
        if (part == 1) v -= __half2float(__float2half_rn(v));
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

// Number of partial slots an image needs (max over image groups of the CTAs sweeping its tiles) and, optionally, the
// per-image slot counts. Mirrors the device-side group_of()/flush() logic of the statistics warp.
//   single CTA : work item = tile; a group of TN images spans `tiles_per_group` items; one slot per CTA
//   pair, TN==1: work item = pair of m tiles x n tile; a group spans tiles_per_group/2 items; two slots per pair
//   pair, TN>1 : each tile is its own group; the pair's items (all n tiles) give one slot per pair for each of its tiles
static int stats_slots(int groups, int tiles_per_group, int num_n_tiles, int TN, bool pair, int T, int G, int32_t* counts,
                       int N) {
    int smax = 1;
    for (int g = 0; g < groups; ++g) {
        int first_item, last_item, mult;
        if (!pair) { first_item = g * tiles_per_group; last_item = (g + 1) * tiles_per_group - 1; mult = 1; }
        else if (TN == 1) { first_item = g * (tiles_per_group / 2); last_item = (g + 1) * (tiles_per_group / 2) - 1; mult = 2; }
        else { first_item = (g / 2) * num_n_tiles; last_item = (g / 2 + 1) * num_n_tiles - 1; mult = 1; }
        const int n = mult * (cta_of_tile(last_item, T, G) - cta_of_tile(first_item, T, G) + 1);
        if (n > smax) smax = n;
        if (counts != nullptr)
            for (int sl = 0; sl < TN; ++sl)
                if (g * TN + sl < N) counts[g * TN + sl] = n;
    }
    return smax;
}

struct ConvGeom {
    int TW, TH, TN, tiles_w, tiles_h, tiles_n, num_m_tiles;
};

// M tiles are boxes of TN images x TH rows x TW columns with POWER-OF-TWO extents (<= 128 pixels): an image whose width or
// height is not a power of two (EMNIST: 28, 14, 7) is covered by boxes that overhang its border -- TMA zero-fills the
// overhang on load and clips it on store, the epilogue masks it out of the statistics -- so every 32-row group of a tile is
// a {w, h} pixel box and all sizes run the same staged TMA-store epilogue with fused statistics (at most 12.5 % / 23 % of the
// rows are padding at 28 / 7 pixels).
static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static ConvGeom conv_geom(int N, int H, int W, bool allow_tall = true) {
    ConvGeom g;
    const int Wv = next_pow2(W), Hv = next_pow2(H);     // virtual (padded) image extents
    g.TW = Wv < kBlockM ? Wv : kBlockM;
    g.TH = kBlockM / g.TW;
    if (g.TH > Hv) g.TH = Hv;
    if (g.TH < 1) g.TH = 1;
    // tall tiles (8 columns x 16 rows) when the image allows: the vertical tap reuse of the 3x3 layers then loads
    // (16 + 2) / (3 * 16) of the activation bytes; 8 columns = one 1024-byte swizzle atom per image row
    // (the fp32-class GEMMs keep one box per tap and prefer the wide tiles: 825 vs 908 TFLOP/s)
    static const int tall = std::getenv("NDB_TILE_TALL") ? std::atoi(std::getenv("NDB_TILE_TALL")) : 1;
    if (tall && allow_tall && Wv >= 8 && Hv >= 16) {
        g.TW = 8;
        g.TH = 16;
    }
    g.TN = 1;
    if (g.TH == Hv && g.TW == Wv) {
        g.TN = kBlockM / (Hv * Wv);
        if (g.TN < 1) g.TN = 1;
        if (g.TN > N) g.TN = N;
    }
    g.tiles_w = (W + g.TW - 1) / g.TW;
    g.tiles_h = (H + g.TH - 1) / g.TH;
    g.tiles_n = (N + g.TN - 1) / g.TN;
    g.num_m_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
    return g;
}

// image slots inside a tile must be whole warps (statistics warp: 4 / TN row quarters per image)
static bool stats_geom_ok(const ConvGeom& g) {
    if (g.TN == 1) return true;
    return (g.TN == 2 || g.TN == 4) && (g.TH * g.TW) % 32 == 0 && g.TN * g.TH * g.TW == kBlockM;
}

// CTA-pair mode is used whenever two M tiles exist and, when statistics are requested, both CTAs of a pair always sweep
// the same images (even number of m tiles per image) or each tile holds whole images.
// Minimum reduction length (in 64-element K blocks) for the CTA-pair kernel. Back-to-back replays of one short-K GEMM
// (scripts/blockn_probe.py with NDB_DISABLE_2CTA; warm L2) favour the single-CTA kernel below ~24 K blocks (16x16 576->1728:
// 18.5 vs 23.6 us), but inside the denoising step (cold L2, different neighbours) pairs are equal or faster on every qkv /
// proj layer (23.0 vs 25.4 us for the same shape), so the threshold stays at 1. The statistics layout depends on this
// choice, which is why ndb_conv_stats_layout() takes the reduction length.
constexpr int kPairMinKBlocks = 1;
static bool use_pair_mode(const ConvGeom& g, int block_n, bool want_stats, int kblocks) {
    static int disabled = -1;
    if (disabled < 0) disabled = std::getenv("NDB_DISABLE_2CTA") != nullptr ? 1 : 0;
    if (disabled) return false;
    if (kblocks < kPairMinKBlocks) return false;
    if (g.num_m_tiles < 2 || (block_n / 2) % 8 != 0) return false;
    if (want_stats && g.TN == 1 && ((g.tiles_w * g.tiles_h) % 2 != 0)) return false;
    return true;
}

struct WorkGeom {
    bool pair;
    int items, workers, grid;
};

static WorkGeom work_geom(const ConvGeom& g, int num_n_tiles, int block_n, bool want_stats, int kblocks) {
    WorkGeom w;
    w.pair = use_pair_mode(g, block_n, want_stats, kblocks);
    if (w.pair) {
        w.items = ((g.num_m_tiles + 1) / 2) * num_n_tiles;
        const int max_workers = num_sms() / 2;
        w.workers = w.items < max_workers ? w.items : max_workers;
        w.grid = 2 * w.workers;
    } else {
        w.items = g.num_m_tiles * num_n_tiles;
        w.workers = w.items < num_sms() ? w.items : num_sms();
        w.grid = w.workers;
    }
    return w;
}

// BLOCK_N by a small cost model: rounds of tiles per worker x tile cost. Narrow tiles run the tensor pipe at a lower rate
// (shared-memory operand bandwidth per MMA grows relative to its work; rates measured with scripts/blockn_probe.py), but
// they fill the machine when a layer has few tiles (the 8x8 levels: 16 M tiles).
// One more shape: OVERLAPPING tiles, 160 columns wide every 144 -- for widths like 576 = 4 x 144 whose 3 x 192 tiling leaves a
// nearly empty last round (16x16 level: 96 work items on 74 CTA pairs). The 16 columns two neighbours share are computed
// by both with identical operands in identical order, i.e. stored twice with the same bits, and counted in the
// statistics by their owner only; whole 32-column chunks keep the epilogue unchanged. `*stride` = channels between tiles.
static int pick_block_n(int cout_pad, const ConvGeom& g, bool want_stats, int kblocks, int max_bn = 256, int* stride = nullptr) {
    static const int cand[6] = {256, 192, 160, 128, 96, 64};
    static const int cstride[6] = {256, 192, 144, 128, 96, 64};
    static const double rate[6] = {1.0, 1.0, 0.95, 0.78, 0.61, 0.5};
    static const int no_overlap = std::getenv("NDB_NO_OVERLAP_TILES") != nullptr ? 1 : 0;
    int best = 0, best_stride = 0;
    double best_cost = 0.0;
    for (int i = 0; i < 6; ++i) {
        const int bn = cand[i], st = cstride[i];
        if (cout_pad % st != 0 || bn > max_bn) continue;
        if (bn != st && (no_overlap || stride == nullptr || max_bn < 256 || cout_pad < 2 * st)) continue;
        const WorkGeom w = work_geom(g, cout_pad / st, bn, want_stats, kblocks);
        const int rounds = (w.items + w.workers - 1) / w.workers;
        double cost = rounds * (double)bn / rate[i];
        if (bn == 256) cost *= 1.02;    // 192 leaves room for one more smem stage: prefer it on ties
        if (best == 0 || cost < best_cost) { best = bn; best_stride = st; best_cost = cost; }
    }
    if (!best)
        for (int bn = max_bn; bn >= 16; bn -= 16)
            if (cout_pad % bn == 0) { best = best_stride = bn; break; }
    if (stride) *stride = best_stride;
    return best;
}

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_pack_conv_weight(const float* w_oihw, void* dst, int cout, int cin_total, int ksize, int cout_pad,
                                    int cin_off, int cin_len, int operand_format, float scale, int part, void* stream) {
    NDB_CHECK_ARG(w_oihw && dst);
    NDB_CHECK_ARG(part == 0 || (part == 1 && operand_format == 0));
    if (scale == 0.f) scale = 1.f;
    NDB_CHECK_ARG(ksize == 1 || ksize == 3);
    NDB_CHECK_ARG(cout_pad >= cout && cin_off >= 0 && cin_off + cin_len <= cin_total);
    const int taps = ksize * ksize;
    const long long total = (long long)taps * cout_pad * cin_len;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_kernel(pack_weight_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, w_oihw, (uint16_t*)dst, cout,
                  cin_total, taps, cout_pad, cin_off, cin_len, operand_format, scale, part);
    return launch_status("ndb_pack_conv_weight");
}

extern "C" int ndb_conv_gemm(const ndb_conv_params* prm, void* stream) {
    NDB_CHECK_ARG(prm != nullptr);
    NDB_CHECK_ARG(prm->N > 0 && prm->H > 0 && prm->W > 0);
    NDB_CHECK_ARG(prm->num_segments >= 1 && prm->num_segments <= kMaxSeg);
    NDB_CHECK_ARG(prm->cout >= 1 && prm->cout_pad >= prm->cout && prm->cout_pad % 16 == 0);
    NDB_CHECK_ARG(prm->out != nullptr);
    NDB_CHECK_ARG(prm->operand_format == 0 || prm->operand_format == 1);
    if (prm->residual_mode == NDB_RES_UP2) NDB_CHECK_ARG(prm->H % 2 == 0 && prm->W % 2 == 0);

    ConvArgs a{};
    a.N = prm->N; a.H = prm->H; a.W = prm->W;
    const ConvGeom cg = conv_geom(prm->N, prm->H, prm->W, prm->acc_chunk_kblocks == 0);
    a.TW = cg.TW; a.TH = cg.TH; a.TN = cg.TN;
    a.tiles_w = cg.tiles_w; a.tiles_h = cg.tiles_h; a.tiles_n = cg.tiles_n;
    a.num_m_tiles = cg.num_m_tiles;
    int kblocks_total = 0;
    for (int s = 0; s < prm->num_segments; ++s) kblocks_total += prm->seg[s].taps * ((prm->seg[s].C + kBlockK - 1) / kBlockK);
    a.n_stride = prm->block_n;
    a.block_n = prm->block_n > 0 ? prm->block_n
                                 : pick_block_n(prm->cout_pad, cg, prm->stat_parts != nullptr, kblocks_total,
                                                prm->acc_chunk_kblocks > 0 ? 192 : 256, &a.n_stride);
    NDB_CHECK_ARG(a.block_n >= 16 && a.block_n <= 256 && a.block_n % 16 == 0 && a.n_stride > 0 && a.n_stride <= a.block_n &&
                  prm->cout_pad % a.n_stride == 0);
    a.num_n_tiles = prm->cout_pad / a.n_stride;
    a.cout = prm->cout;
    a.num_seg = prm->num_segments;
    a.a_box_bytes = (uint32_t)(a.TN * a.TH * a.TW * 128);
    a.acc_stride = (a.block_n + 31) / 32 * 32;
    a.tmem_cols = a.acc_stride <= 32 ? 64 : a.acc_stride <= 64 ? 128 : a.acc_stride <= 128 ? 256 : 512;
    a.idesc = make_idesc_f16(kBlockM, a.block_n, prm->operand_format, 0, 0);
    {
        static const int split = std::getenv("NDB_SPLIT_PRODUCER") ? std::atoi(std::getenv("NDB_SPLIT_PRODUCER")) : 1;
        a.split_prod = split;
    }
    a.bias = prm->bias;
    a.residual = prm->residual;
    a.res_mode = prm->residual ? prm->residual_mode : NDB_RES_NONE;
    if (prm->residual) NDB_CHECK_ARG(prm->residual_mode >= NDB_RES_SAME && prm->residual_mode <= NDB_RES_DOWN2);
    a.out_kind = prm->out_kind;
    a.out = prm->out;
    a.fmt = prm->operand_format;
    a.out_scale = prm->out_scale == 0.f ? 1.f : prm->out_scale;
    a.acc_chunk = prm->acc_chunk_kblocks;
    NDB_CHECK_ARG(a.acc_chunk >= 0);
    a.stat_parts = reinterpret_cast<float2*>(prm->stat_parts);
    a.tiles_per_group = a.tiles_w * a.tiles_h * a.num_n_tiles;
    const WorkGeom wg = work_geom(cg, a.num_n_tiles, a.block_n, prm->stat_parts != nullptr, kblocks_total);
    a.b_rows = wg.pair ? a.block_n / 2 : a.block_n;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = std::getenv("NDB_DBG"); dbg = e ? std::atoi(e) : 0; }
        a.dbg = dbg;
    }
    if (wg.pair) a.idesc = make_idesc_f16(2 * kBlockM, a.block_n, prm->operand_format, 0, 0);
    if (prm->stat_parts != nullptr) {
        NDB_CHECK_ARG(prm->out_kind != NDB_OUT_F32_NCHW);
        NDB_CHECK_ARG(stats_geom_ok(cg));
        a.stat_smax = stats_slots(a.tiles_n, a.tiles_per_group, a.num_n_tiles, a.TN, wg.pair, wg.items, wg.workers, nullptr, 0);
        NDB_CHECK_ARG(prm->stat_smax == a.stat_smax);
    }

    {
        static const int off = std::getenv("NDB_DISABLE_VREUSE") != nullptr ? 1 : 0;
        bool any3 = false;
        for (int s = 0; s < prm->num_segments; ++s) any3 = any3 || prm->seg[s].taps == 9;
        // worth it from 4 image rows per tile up (box of TH + 2 rows instead of three of TH); the boxes must leave room
        // for a weight ring (wide flat tiles of short images do not)
        a.a_slot_bytes = (a.TH + 2) * a.TW * 128;
        a.vreuse = (!off && any3 && a.TN == 1 && a.TW % 8 == 0 && a.TH >= 4 && a.acc_chunk == 0 && a.a_slot_bytes <= 24 * 1024) ? 1 : 0;
        a.a_k3_bytes = (uint32_t)a.a_slot_bytes;
        static const int halo = std::getenv("NDB_HALO") ? std::atoi(std::getenv("NDB_HALO")) : 1;
        if (a.vreuse && halo && a.TW == 8) {
            a.vreuse = 2;
            a.a_k3_bytes = (uint32_t)((a.TH + 2) * (a.TW + 2) * 128);
            a.a_slot_bytes = ((int)a.a_k3_bytes + 1023) / 1024 * 1024;
        }
        if (a.vreuse) {
            // unit list: 3x3 units in segment / dx / channel-block order, the 1x1 units spread evenly among them
            static const int no_mix = std::getenv("NDB_NO_INTERLEAVE") != nullptr ? 1 : 0;
            struct U { uint8_t s, ch, dx; };
            std::vector<U> u3, u1;
            for (int s = 0; s < prm->num_segments; ++s) {
                const int chunks = (prm->seg[s].C + kBlockK - 1) / kBlockK;
                if (prm->seg[s].taps == 9) {
                    for (int dx = 0; dx < (a.vreuse == 2 ? 1 : 3); ++dx)
                        for (int ch = 0; ch < chunks; ++ch) u3.push_back({(uint8_t)s, (uint8_t)ch, (uint8_t)dx});
                } else {
                    for (int ch = 0; ch < chunks; ++ch) u1.push_back({(uint8_t)s, (uint8_t)ch, 0});
                }
            }
            if ((int)(u3.size() + u1.size()) > kMaxUnits) {
                a.vreuse = 0;       // (very long reductions: the classic path)
            } else {
                a.n_units = 0;
                auto put = [&](const U& u) { a.unit_seg[a.n_units] = u.s; a.unit_ch[a.n_units] = u.ch; a.unit_dx[a.n_units] = u.dx; ++a.n_units; };
                size_t done1 = 0;
                for (size_t i = 0; i < u3.size(); ++i) {
                    put(u3[i]);
                    const size_t want = no_mix ? 0 : u1.size() * (i + 1) / u3.size();      // 1x1 units due after i + 1 of the 3x3 ones
                    while (done1 < want) put(u1[done1++]);
                }
                while (done1 < u1.size()) put(u1[done1++]);
            }
        }
    }
    TmapPack maps;
    a.total_kblocks = 0;
    for (int s = 0; s < prm->num_segments; ++s) {
        const ndb_conv_segment& sg = prm->seg[s];
        NDB_CHECK_ARG(sg.act && sg.weight);
        // channels need not fill whole 64-wide K blocks: TMA zero-fills the rest of the last block of BOTH operands (the
        // 16-byte stride rule of tensor maps asks for a multiple of 8)
        NDB_CHECK_ARG(sg.C > 0 && sg.C % 8 == 0);
        NDB_CHECK_ARG(sg.taps == 1 || sg.taps == 9);
        a.seg_taps[s] = sg.taps;
        a.seg_chunks[s] = (sg.C + kBlockK - 1) / kBlockK;
        a.total_kblocks += sg.taps * a.seg_chunks[s];
        {
            uint64_t dims[4] = {(uint64_t)sg.C, (uint64_t)prm->W, (uint64_t)prm->H, (uint64_t)prm->N};
            uint64_t strides[3] = {(uint64_t)sg.C * 2, (uint64_t)prm->W * sg.C * 2,
                                   (uint64_t)prm->H * prm->W * sg.C * 2};
            const bool k3 = a.vreuse && sg.taps == 9;
            uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(a.TW + ((k3 && a.vreuse == 2) ? 2 : 0)), (uint32_t)(a.TH + (k3 ? 2 : 0)), (uint32_t)a.TN};
            int rc = encode_tmap_16bit(&maps.a[s], sg.act, 4, dims, strides, box);
            if (rc != NDB_OK) return rc;
        }
        {
            uint64_t dims[3] = {(uint64_t)sg.C, (uint64_t)prm->cout_pad, (uint64_t)sg.taps};
            uint64_t strides[2] = {(uint64_t)sg.C * 2, (uint64_t)prm->cout_pad * sg.C * 2};
            uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)a.b_rows, 1u};
            int rc = encode_tmap_16bit(&maps.b[s], sg.weight, 3, dims, strides, box);
            if (rc != NDB_OK) return rc;
        }
    }

    // staged TMA-store epilogue: NHWC output, plain or same-size residual, whole 32-column chunks, and the 32 rows of a
    // warp must form a {box_w, box_h} pixel box inside one image
    {
        static int off = -1;
        if (off < 0) off = std::getenv("NDB_DISABLE_BULK_EPILOGUE") != nullptr ? 1 : 0;
        const bool stats = prm->stat_parts != nullptr;
        const int box_w = a.TW < 32 ? a.TW : 32, box_h = 32 / box_w;
        const bool geom_ok = (a.TW % 32 == 0 || (32 % a.TW == 0 && a.TH % box_h == 0)) && (a.TH * a.TW) % 32 == 0;
        const bool kind_ok = prm->out_kind == NDB_OUT_F32_NHWC ||
                             (prm->out_kind == NDB_OUT_F16_NHWC && a.res_mode == NDB_RES_NONE);
        (void)stats;
        const int es = prm->out_kind == NDB_OUT_F32_NHWC ? 4 : 2;
        a.epi_bulk = (!off && geom_ok && kind_ok && (a.res_mode == NDB_RES_NONE || a.res_mode == NDB_RES_SAME) &&
                      a.block_n % 32 == 0 && (prm->cout * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(prm->out) & 15) == 0)
                         ? 1 : 0;
        if (a.epi_bulk) {
            uint64_t dims[4] = {(uint64_t)prm->cout, (uint64_t)prm->W, (uint64_t)prm->H, (uint64_t)prm->N};
            uint64_t strides[3] = {(uint64_t)prm->cout * es, (uint64_t)prm->W * prm->cout * es,
                                   (uint64_t)prm->H * prm->W * prm->cout * es};
            uint32_t box[4] = {32u, (uint32_t)box_w, (uint32_t)box_h, 1u};
            int rc = encode_tmap(&maps.out, prm->out, es, es == 4 ? 128 : 64, 4, dims, strides, box);
            if (rc != NDB_OK) return rc;
            if (a.res_mode == NDB_RES_SAME) {
                if ((reinterpret_cast<uintptr_t>(prm->residual) & 15) != 0) a.epi_bulk = 0;     // TMA needs 16-byte alignment
                else {
                    rc = encode_tmap(&maps.res, prm->residual, 4, 128, 4, dims, strides, box);
                    if (rc != NDB_OK) return rc;
                }
            }
        }
    }

    const int stage_bytes = kATileBytes + a.b_rows * 128;
    // staging tiles per epilogue warp. Residual ring: 3 deep when the epilogue is the critical path (short K) or when the
    // operand pipeline still gets 5 stages next to it, else 2 (the smem goes to the pipeline: the step time is flat from 5
    // stages up and drops below)
    auto epi_bytes_for = [&](int bufs) {
        const int region = bufs >= 2 ? kEpiWarps * bufs * 4096 : kEpiWarps * kStageFloats * (int)sizeof(float);
        return region + kEpiWarps * 128 * (int)sizeof(long long) + 2 * 4 * a.block_n * (int)sizeof(float2) +
               (prm->stat_parts ? a.TN * prm->cout * (int)sizeof(float2) : 0);
    };
    const int b_tile = a.b_rows * 128;
    // vreuse: three activation boxes + as many weight tiles as fit (two boxes when that leaves fewer than five tiles)
    auto vreuse_slots = [&](int bufs, int* a_slots) {
        const int budget = 227 * 1024 - 1024 - kBarBytes - epi_bytes_for(bufs);
        // a halo box lasts nine K blocks, so two of them are plenty and the rest goes to the weight ring (one tile per K
        // block: its depth is what covers the TMA round trip)
        // ... unless the layer also has 1x1 segments (the fused skip convolution): their plain tiles occupy a box slot for a
        // single K block each, so the box ring is their whole look-ahead
        // (64x64 3x3 192->192 + 1x1 skip over 384 channels: 96.0 / 91.6 / 88.3 us with 2 / 3 / 4 box slots)
        bool any1 = false;
        for (int s = 0; s < prm->num_segments; ++s) any1 = any1 || prm->seg[s].taps == 1;
        int as = a.vreuse == 2 ? (any1 ? 4 : 2) : 3, bs = (budget - as * a.a_slot_bytes) / b_tile;
        while (bs < 5 && as > 2) { --as; bs = (budget - as * a.a_slot_bytes) / b_tile; }
        if (a_slots) *a_slots = as;
        return bs > kMaxStages ? kMaxStages : bs;
    };
    if (a.epi_bulk && a.res_mode == NDB_RES_SAME) {
        const int stages3 = a.vreuse ? vreuse_slots(3, nullptr) - 2 : (227 * 1024 - 1024 - kBarBytes - epi_bytes_for(3)) / stage_bytes;
        a.epi_bufs = (a.total_kblocks <= 24 || stages3 >= 5) ? 3 : 2;
    } else {
        // (two tiles for the long-K layers as well, at the price of one pipeline stage: measured, no gain)
        a.epi_bufs = (a.epi_bulk && a.total_kblocks <= 24) ? 2 : 1;
    }
    a.epi_region = a.epi_bufs >= 2 ? kEpiWarps * a.epi_bufs * 4096 : kEpiWarps * kStageFloats * (int)sizeof(float);
    const int epi_bytes = epi_bytes_for(a.epi_bufs);
    const int smem_budget = 227 * 1024 - 1024 /*align slack*/ - kBarBytes - epi_bytes;
    int stages = smem_budget / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > a.total_kblocks) stages = a.total_kblocks < 2 ? 2 : a.total_kblocks;
    NDB_CHECK_ARG(stages >= 2);
    a.stages = stages;
    a.ring_bytes = stages * stage_bytes;
    if (a.vreuse) {
        a.b_slots = vreuse_slots(a.epi_bufs, &a.a_slots);
        if (a.b_slots < 3) a.vreuse = 0;      // (does not happen with the tile shapes above; the maps were encoded for vreuse)
        NDB_CHECK_ARG(a.vreuse);
        a.stages = a.b_slots;                 // the weight ring uses full_bar / empty_bar
        a.ring_bytes = a.a_slots * a.a_slot_bytes + a.b_slots * b_tile;
    }
    const size_t smem_bytes = (size_t)a.ring_bytes + 1024 + kBarBytes + epi_bytes;

    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return launch_status("cudaFuncSetAttribute(conv_gemm_kernel)");
        attr_once.done();
    }
    const bool stats = a.stat_parts != nullptr;
    cudaError_t le;
    if (a.acc_chunk > 0) {
        NDB_CHECK_ARG(a.block_n <= 192);      // three 32-column running-sum chunks per epilogue warp
        static PerDeviceOnce chunk_attr;
        if (chunk_attr.pending()) {
            cudaFuncSetAttribute(conv_gemm_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            cudaFuncSetAttribute(conv_gemm_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            cudaFuncSetAttribute(conv_gemm_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            cudaFuncSetAttribute(conv_gemm_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            chunk_attr.done();
        }
        if (!wg.pair)
            le = stats ? launch_kernel(conv_gemm_kernel<true, false, true>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 1, maps, a)
                       : launch_kernel(conv_gemm_kernel<false, false, true>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 1, maps, a);
        else
            le = stats ? launch_kernel(conv_gemm_kernel<true, true, true>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 2, maps, a)
                       : launch_kernel(conv_gemm_kernel<false, true, true>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 2, maps, a);
    } else if (!wg.pair)
        le = stats ? launch_kernel(conv_gemm_kernel<true, false>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 1, maps, a)
                   : launch_kernel(conv_gemm_kernel<false, false>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 1, maps, a);
    else
        le = stats ? launch_kernel(conv_gemm_kernel<true, true>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 2, maps, a)
                   : launch_kernel(conv_gemm_kernel<false, true>, dim3(wg.grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, 2, maps, a);
    if (le != cudaSuccess) {
        std::fprintf(stderr, "[ndb200] launch of conv_gemm_kernel failed: %s\n", cudaGetErrorString(le));
        return NDB_ERR_CUDA;
    }
    return launch_status("ndb_conv_gemm");
}

// Layout of the fused statistics output for a conv with this geometry and reduction length k_total (= sum over the
// segments of taps * C; it decides between the CTA-pair and the single-CTA kernel): returns the number of partial slots per image
// (stat_smax) and fills counts[N] (host memory) with the slots actually written for each image; <0 on error.
// stat_parts must hold N * stat_smax * cout float2.
extern "C" int ndb_conv_stats_layout(int N, int H, int W, int cout_pad, int block_n, int k_total, int acc_chunk_kblocks,
                                     int32_t* counts) {
    if (N <= 0 || H <= 0 || W <= 0 || cout_pad <= 0 || k_total <= 0) return NDB_ERR_INVALID;
    const int kblocks = (k_total + kBlockK - 1) / kBlockK;      // only "at least one" matters to the plan
    const ConvGeom g = conv_geom(N, H, W, acc_chunk_kblocks == 0);
    int st = block_n;
    const int bn = block_n > 0 ? block_n : pick_block_n(cout_pad, g, true, kblocks, acc_chunk_kblocks > 0 ? 192 : 256, &st);
    if (bn <= 0 || st <= 0 || cout_pad % st != 0) return NDB_ERR_INVALID;
    const int n_tiles = cout_pad / st;
    const WorkGeom w = work_geom(g, n_tiles, bn, true, kblocks);
    return stats_slots(g.tiles_n, g.tiles_w * g.tiles_h * n_tiles, n_tiles, g.TN, w.pair, w.items, w.workers, counts, N);
}

#ifdef NDB_BRINGUP
// bring-up builds only (not declared in ndb200.h): copies the phase trace of the last conv launch to host memory [160][12]
extern "C" int ndb_debug_conv_trace(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(g_trace)) == cudaSuccess ? NDB_OK : NDB_ERR_CUDA;
}
#endif

// 1 when the epilogue can produce per-image channel statistics for this geometry (else use ndb_gn_stats).
extern "C" int ndb_conv_stats_supported(int N, int H, int W) {
    return stats_geom_ok(conv_geom(N, H, W)) ? 1 : 0;
}
