/* ndb200 -- C-ABI of the B200-native nice-diffusion sampling hot path.
 *
 * The reference (edogariu/nice-diffusion) is pure Python/PyTorch and has NO FFI / plugin layer
 * (SURVEY.md section 8b); what the reference "binds" for this path is a set of torch ops called from
 * nicediffusion/model.py and nicediffusion/diffusion.py. Each entry point below replaces the group of
 * reference ops cited next to it. The Python host (nice-diffusion_b200/*.py) mirrors the reference's
 * classes and reaches these symbols through ctypes with raw device pointers.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; `stream` is a cudaStream_t passed as void*;
 *   - kernels are enqueued on `stream`, never synchronise, never allocate: CUDA-graph capturable;
 *   - return value: NDB_OK (0) or a negative NDB_ERR_* code; no C++ exception crosses the boundary;
 *   - activations inside the UNet are NHWC ("channels last"): fp32 for the residual stream,
 *     fp16 for tensor-core operands. Model inputs/outputs are NCHW fp32 like the reference's.
 */
#ifndef NDB200_H
#define NDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDB_OK 0
#define NDB_ERR_INVALID (-1)
#define NDB_ERR_CUDA (-2)
#define NDB_ERR_UNSUPPORTED (-3)

#define NDB_MAX_SEGMENTS 9

/* library / device queries ------------------------------------------------------------------ */
int ndb_version(void);                     /* ABI version, currently 1 */
int ndb_device_supported(void);            /* 1 when the current CUDA device is sm_100 (B200), else 0 */
const char* ndb_last_error(void);          /* last CUDA error string seen by the library (host pointer) */

/* ---- implicit-GEMM convolution on tcgen05 / TMEM / TMA -------------------------------------------
 * Replaces nn.Conv2d 3x3 s1 p1 and 1x1 (model.py:72,103,166-178,367,448), nn.Conv1d k=1 qkv / proj_out
 * (model.py:247,253), the step-embedding nn.Linear of every ResBlock (model.py:180-182,197) and, through
 * the `residual` / extra-segment arguments, the ops fused around them: bias add, `h + self.skip(x)`
 * (model.py:211), `h + x` (model.py:291), x_resample (model.py:193) and torch.cat skip concat (model.py:474).
 *
 *   out[n,h,w,co] = bias[co] + residual(n,h,w,co)
 *                 + sum_seg sum_tap sum_ci  a_seg[n, h+dy(tap), w+dx(tap), ci] * w_seg[tap][co][ci]
 *
 * Every segment is one K-range of the GEMM: its own NHWC fp16 activation tensor (all segments share
 * N,H,W) and its own packed weight block [taps][cout_pad][C] (fp16, K-major). Zero padding comes from TMA
 * out-of-bounds fill. C must be a multiple of 64, cout_pad a multiple of 16.
 */
typedef struct {
    const void* act;     /* fp16 NHWC [N,H,W,C] */
    const void* weight;  /* fp16 [taps][cout_pad][C] (ndb_pack_conv_weight) */
    int32_t C;
    int32_t taps;        /* 1 (1x1) or 9 (3x3, pad 1) */
} ndb_conv_segment;

#define NDB_RES_NONE 0
#define NDB_RES_SAME 1   /* residual is fp32 NHWC [N,H,W,cout] */
#define NDB_RES_UP2 2    /* residual is [N,H/2,W/2,cout], nearest-neighbour x2 (F.interpolate, model.py:77) */
#define NDB_RES_DOWN2 3  /* residual is [N,2H,2W,cout], 2x2 mean (F.avg_pool2d, model.py:111) */

#define NDB_OUT_F32_NHWC 0
#define NDB_OUT_F16_NHWC 1
#define NDB_OUT_F32_NCHW 2

typedef struct {
    int32_t N, H, W;
    int32_t cout;        /* valid output channels */
    int32_t cout_pad;    /* rows of every weight block (>= cout, multiple of 16) */
    int32_t num_segments;
    ndb_conv_segment seg[NDB_MAX_SEGMENTS];
    const float* bias;       /* [cout] or NULL */
    const float* residual;   /* see NDB_RES_* or NULL */
    int32_t residual_mode;
    int32_t out_kind;        /* NDB_OUT_* */
    void* out;
    int32_t operand_format;  /* 0 = fp16, 1 = bf16 (both: fp32 accumulation in TMEM) */
    int32_t block_n;         /* 0 = auto */
    void* stat_parts;        /* optional float2 [N][stat_smax][cout]: GroupNorm statistics of `out`, produced by the
                                epilogue: per image and channel, (sum, sum of squares) of the stored values, split
                                over the (few) persistent CTAs that swept the image's tiles -- slot s of image n is
                                written by the s-th such CTA, ndb_conv_stats_layout() says how many there are. The
                                consumer (ndb_gn_apply) adds the slots in order: deterministic, no atomics, no fences,
                                no extra pass over the tensor. NULL = off. */
    int32_t stat_smax;       /* slots per image, must equal ndb_conv_stats_layout(...) */
    float out_scale;         /* the accumulators are multiplied by this before bias / residual; 0 = 1. The fp32-class mode packs
                                its weights times 2^8 (so that the low halves stay normal fp16 numbers) and passes 2^-8 here;
                                powers of two are exact */
    int32_t acc_chunk_kblocks; /* 0 = one accumulation over the whole K range. > 0 (fp32-class mode): the tensor core's fp32
                                accumulation truncates (measured: -5e-9 relative per element of K), so K is accumulated in chunks
                                of this many 64-element blocks, each from a zeroed TMEM accumulator, and the chunks are added with
                                round-to-nearest adds in the epilogue warps' registers. Limits BLOCK_N to 192. */
} ndb_conv_params;

int ndb_conv_gemm(const ndb_conv_params* p, void* stream);
/* Returns stat_smax for this geometry and fills counts[N] (HOST pointer) with the slots written per image. */
int ndb_conv_stats_layout(int N, int H, int W, int cout_pad, int block_n, int k_total, int acc_chunk_kblocks,
                          int32_t* counts);
int ndb_conv_stats_supported(int N, int H, int W);   /* 1 if stat_parts is available for this geometry */

/* OIHW fp32 (nn.Conv2d / Conv1d / Linear weight, kernel 1 or 3) -> [taps][cout_pad][cin] 16-bit, rows >= cout zero.
 * `cin_off`/`cin_len` select an input-channel range (used to split a weight over concat segments).
 * The stored value is the 16-bit rounding of w * scale (scale 0 = 1; part 0), or -- part 1, fp16 only -- the fp16 rounding of
 * the remainder w * scale - part0: the (hi, lo) weight pair of the fp32-class mode. */
int ndb_pack_conv_weight(const float* w_oihw, void* dst, int cout, int cin_total, int ksize, int cout_pad,
                         int cin_off, int cin_len, int operand_format, float scale, int part, void* stream);

/* ---- GroupNorm(32) statistics + apply ------------------------------------------------------------------
 * Replaces nn.GroupNorm(32, C, eps=1e-5) + SiLU + AdaGN scale/shift + h_resample + torch.cat
 * (model.py:172,176,190-193,199-207,250,264,446-447).
 * Statistics travel as PER-CHANNEL partial sums: float2 parts[n][slot][c] = (sum, sum of squares) over a subset of
 * the pixels of image n; the slots of an image add up to the whole image (counts[n] of the smax slots are in use).
 * They are normally produced for free by the conv epilogue (ndb_conv_params.stat_parts); ndb_gn_stats is the
 * standalone producer (one slot per image) for tensors that do not come out of a GEMM (the stem) or geometries the
 * epilogue does not cover. The consumer (ndb_gn_apply) folds slots and channels into the 32 groups itself in fp64, so
 * one tensor's statistics serve every norm that reads it -- also the norms over a channel concat [x0 ++ x1]
 * (model.py:474), whose groups straddle the two tensors. */
int ndb_gn_stats(const float* x0, int C0, const float* x1, int C1, int N, int HW,
                 void* parts0 /* float2 [N][1][C0] */, void* parts1 /* float2 [N][1][C1] or NULL */,
                 void* workspace, void* stream);
/* Scratch for the deterministic cross-block reduction (no floating-point atomics: results are bit-identical from
 * run to run). Must be zero-filled ONCE before first use; it is left clean by every call and may be shared by all
 * GroupNorm launches of a stream. */
int64_t ndb_gn_stats_workspace_bytes(int N, int HW, int C);

#define NDB_RESAMPLE_NONE 0
#define NDB_RESAMPLE_UP2 1
#define NDB_RESAMPLE_DOWN2 2

typedef struct {
    const float* x0; int32_t C0;
    const float* x1; int32_t C1;       /* optional second concat source */
    int32_t N, H, W;                   /* input spatial size */
    const void* parts0; const int32_t* counts0; int32_t smax0;   /* statistics of x0: float2 [N][smax0][C0], int [N] */
    const void* parts1; const int32_t* counts1; int32_t smax1;   /* statistics of x1 (when x1 != NULL) */
    const float* gamma; const float* beta;   /* [C0+C1] */
    const float* scale; const float* shift;  /* optional AdaGN [N][ld_ss] rows: y = gn*(1+scale)+shift */
    int32_t ld_ss;
    const float* pre_add; int32_t ld_pre;    /* optional [N][ld_pre]: added per (n, channel) BEFORE the norm (the
                                                non-adaptive `h += embedding` path, model.py:205); the statistics are
                                                corrected analytically */
    int32_t silu;                      /* apply SiLU after the affine */
    int32_t resample;                  /* NDB_RESAMPLE_* applied after the activation */
    void* out;                         /* 16-bit NHWC [N,H',W',C0+C1] */
    void* raw16;                       /* optional 16-bit copy of the un-normalised (resampled) input, same shape */
    int32_t x0_is16;                   /* x0 is a 16-bit tensor (operand_format) instead of fp32: the `h` between the two
                                          convs of a ResBlock; single source, no resampling, no raw copy */
    float* pool32;                     /* optional, DOWN2 only: fp32 2x2 mean of the raw input [N,H/2,W/2,C] -- the
                                          x_resample operand of a down ResBlock (model.py:193), read back by the conv
                                          epilogue as a plain residual */
    int32_t operand_format;
    float eps;
    void* out_lo; void* raw_lo;        /* fp32-class mode (operand_format 0, fp32 input): when out_lo != NULL every result v is
                                          stored as the fp16 pair out = fp16(v), out_lo = fp16(v - out) (same for raw16 /
                                          raw_lo), and SiLU is evaluated with an accurate exponential and a true division */
    int32_t cpg;                       /* channels per group; 0 = (C0 + C1) / 32. Non-zero: this call normalises a channel SLICE
                                          of a wider GroupNorm(32) input (the skip half of a concatenated tensor) whose groups
                                          do not straddle the slice: (C0 + C1) / cpg <= 32 groups, gamma / beta already offset */
} ndb_gn_apply_params;

int ndb_gn_apply(const ndb_gn_apply_params* p, void* stream);

/* ---- fused multi-head attention (tcgen05) ------------------------------------------------------------------
 * Replaces q@k^T * scale, softmax, w@v and the head split / merge reshapes (model.py:266-287).
 * qkv: fp16 [B][T][3C] as produced by the qkv 1x1 conv; out: fp16 [B][T][C], channel = head*d + j.
 * order 0: channels laid out [3][heads][d] (split_qkv_first=True); order 1: [heads][3][d] (legacy). */
int ndb_attention(const void* qkv, void* out, int B, int T, int heads, int d, int order, int operand_format,
                  void* stream);
/* fp32-class variant (head dim <= 64): q, k, v as fp16 (hi, lo) pairs in two tensors of that layout (ndb_split16 of the fp32
 * qkv GEMM output), every product evaluated as hi*hi + lo*hi + hi*lo, probabilities split the same way; out: fp32 [B][T][C]. */
int ndb_attention_split(const void* qkv_hi, const void* qkv_lo, float* out, int B, int T, int heads, int d, int order,
                        void* stream);

/* ---- timestep / class embedding front end ------------------------------------------------------------------
 * Replaces timestep_embedding + step_embed MLP + class_embedding add + the SiLU in front of every
 * ResBlock's step_embedding Linear (model.py:456-459,514-523,197).
 * freqs: [mc/2] fp32 (built by the host with the reference formula); t: [B] fp32 timesteps;
 * y: [B] int64 labels or NULL. h1_scratch: [B][4mc] fp32 workspace. emb_out: [B][4mc] fp32 (optional);
 * act_out: 16-bit SiLU(emb) [B][4mc]. Two launches (one per Linear layer). */
int ndb_embed(const float* t, const int64_t* y, const float* freqs, int B, int mc,
              const float* w0, const float* b0, const float* w2, const float* b2, const float* class_emb,
              float* h1_scratch, float* emb_out, void* act_out, int operand_format, void* stream);

/* ---- stem conv / small direct convs ------------------------------------------------------------------
 * conv3x3 s1 p1 with tiny Cin (model.py:367), exact fp32: x NCHW fp32 [N,Cin,H,W] -> NHWC fp32 [N,H,W,Cout].
 * stat_parts (nullable; needs H*W % 64 == 0): the GroupNorm statistics of the output in the slot format of
 * ndb_gn_apply, float2 [N][H*W/64][Cout], slot s of image n = (sum, sum of squares) over its pixels [64 s, 64 s + 64);
 * every slot is written (counts[n] = H*W/64). */
int ndb_stem_conv(const float* x_nchw, const float* w_oihw, const float* bias, float* out_nhwc, void* stat_parts, int N,
                  int Cin, int H, int W, int Cout, void* stream);

/* the same conv as ONE tensor-core K block: gathers the 3x3 neighbourhood of every pixel of an NCHW fp32 image with
 * Cin <= 7 channels into a 64-wide 16-bit operand row out[n][h][w][tap*Cin + c] (zero padded; tap = ky*3+kx), so that
 * ndb_conv_gemm runs the stem as a 1x1 GEMM with weight[o][tap*Cin + c] = w[o][c][ky][kx] -- and gets the fused bias,
 * TMA-store epilogue and GroupNorm statistics of every other conv (model.py:367). */
int ndb_stem_im2col(const float* x_nchw, void* out, int N, int Cin, int H, int W, int operand_format, void* stream);

/* generic fp32 NHWC resample (standalone Upsample / Downsample without conv, model.py:77,111) and layout casts */
int ndb_resample_f32(const float* x, float* out, int N, int H, int W, int C, int mode, void* stream);
int ndb_cast_nhwc16(const float* x, void* out, int64_t numel, int operand_format, void* stream);
/* fp32 -> fp16 pair (hi, lo), hi + lo == x to ~2^-22 relative: the activation operands of the fp32-class mode. silu != 0
 * applies SiLU (accurate expf) first (the step-embedding activation, model.py:197). numel % 4 == 0. */
int ndb_split16(const float* x, void* hi, void* lo, int64_t numel, int silu, void* stream);
/* strided conv3x3 s2 p1 input gather for Downsample(with_conv) (model.py:103-105): builds the 9-tap
 * tap-major operand [9][N,Ho,Wo,C] (Ho=(H-1)/2+1) so the GEMM runs it as nine 1x1 K-segments. */
int ndb_im2col_s2(const float* x, void* out, int N, int H, int W, int C, int operand_format, void* stream);

/* ---- fused sampler step ------------------------------------------------------------------------------
 * Replaces get_eps_and_log_var, the CFG combine, pred_x0 + clamp, posterior mean, classifier-guidance
 * shift, noise add and the t==0 mask (diffusion.py:242-316) or the DDIM update (diffusion.py:318-369).
 * All image tensors NCHW fp32. model_out: [B][Cm][H][W] with Cm = 2C when the variance is learned, else C.
 * null_out: unconditional model output for classifier-free guidance or NULL. t: [B] fp32 respaced step
 * index per sample. tables: [NDB_TAB_COUNT][R] fp32 (fp64 numpy tables cast to fp32 as diffusion.py:492).
 * noise: injected N(0,1) tensor, or NULL to draw Philox4x32-10 + Box-Muller noise in-kernel keyed by
 * (seed, sample_ids[b], step t, element). grad: classifier gradient d log p(y|x)/dx or NULL. */
#define NDB_TAB_SQRT_RECIP_ABAR 0
#define NDB_TAB_SQRT_RECIPM1_ABAR 1
#define NDB_TAB_POST_COEF_X0 2
#define NDB_TAB_POST_COEF_XT 3
#define NDB_TAB_LOGVAR_MIN 4   /* log_posterior_var_clipped (interp) or the fixed log-variance (small/large) */
#define NDB_TAB_LOGVAR_MAX 5   /* log(betas) */
#define NDB_TAB_ABAR 6
#define NDB_TAB_ABAR_PREV 7
#define NDB_TAB_SQRT_1M_ABAR 8
#define NDB_TAB_LOG_POST_VAR_CLIPPED 9   /* log_posterior_var_clipped for every variance type (the true log-variance of the
                                            variational bound, diffusion.py:409); sampling does not read it */
#define NDB_TAB_COUNT 10

#define NDB_VAR_FIXED 0          /* 'small' / 'large': log_var = tables[LOGVAR_MIN][t] */
#define NDB_VAR_LEARNED 1        /* log_var = v */
#define NDB_VAR_LEARNED_INTERP 2 /* log_var = frac*max + (1-frac)*min, frac=(v+1)/2 */

typedef struct {
    const float* x_t;
    const float* model_out;
    const float* null_out;
    const float* t;
    const float* tables;
    int32_t R;                /* table row length */
    int32_t B, C, HW, Cm;
    int32_t var_type;
    int32_t ddim;             /* 0 = DDPM step, 1 = DDIM step */
    double eta;               /* DDIM eta (python float in the reference) */
    double cfg_strength;      /* used when null_out != NULL */
    const float* grad;        /* classifier guidance gradient or NULL */
    double clf_strength;
    int32_t clip_x0;
    const float* noise;       /* NULL -> Philox */
    uint64_t seed;
    const int64_t* sample_ids; /* [B] global sample indices for Philox keying (NULL -> b) */
    float* x_prev;            /* out (may alias x_t) */
    float* x_prev2;           /* optional second copy of the output (the unconditional half of a doubled CFG batch) */
    float* pred_x0;           /* out, optional */
} ndb_step_params;

int ndb_sampler_step(const ndb_step_params* p, void* stream);

/* Device-side loop counter for CUDA-graph replay of whole steps: t_step[b] <- max(t_step[0]-1, 0) for b < B and
 * t_model[b] <- timestep_map_f32[t_step] for b < Bm (diffusion.py:215-216,246). One tiny launch per step. */
int ndb_advance_step(float* t_step, int B, float* t_model, int Bm, const float* timestep_map_f32, void* stream);

/* x_T ~ N(0,1) from the same Philox stream family: element e of sample id s at "step" = R (diffusion.py:202). */
int ndb_philox_normal(float* out, int B, int64_t per_sample, uint64_t seed, const int64_t* sample_ids,
                      int32_t step, void* stream);

/* q-sample: x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) noise (diffusion.py:232-240); coefficient tables fp32 [R]. */
int ndb_qsample(const float* x0, const float* noise, const float* t, const float* sqrt_abar,
                const float* sqrt_1m_abar, float* out, int B, int64_t per_sample, void* stream);

/* [-1,1] float NCHW -> uint8 NHWC image, ((x+1)*127.5).clamp(0,255) truncated (sample.py:94,168).
 * gray_invert (C == 1 only): the byte the reference CLI ends up storing for one-channel models -- it inverts the float
 * image (sample.py:98-100), truncates (:168) and inverts again (:170-171): 255 - trunc(255 - v). */
int ndb_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int B, int C, int H, int W, int gray_invert, void* stream);

/* ---- fused training loss, forward + gradient ------------------------------------------------------------------------
 * Replaces Diffusion.loss / variational_lower_bound / kl_div / log_likelihood / approx_cdf / mean_flat
 * (diffusion.py:375-438, 499-549) GIVEN the model output: per-sample loss[b] and grad_out = d loss[b] / d model_out.
 * x0, x_t, noise: [B][C][H][W]; model_out: [B][Cm][H][W] (Cm = 2C with a learned variance); t: [B] fp32 step index into
 * the tables (the trainer draws it in [0, original_num_steps), diffusion.py:88 of trainer.py -- tables must be that long).
 * loss_type: 0 simple, 1 KL, 2 KL_rescaled, 3 hybrid (eps detached inside the bound, weight 0.001 * rescaled_num_steps). */
typedef struct {
    const float* x0; const float* x_t; const float* noise; const float* model_out;
    const float* t; const float* tables;
    int32_t R, B, C, HW, Cm;
    int32_t var_type;              /* NDB_VAR_* */
    int32_t loss_type;
    int32_t rescaled_num_steps;
    float* loss;                   /* [B] */
    float* grad_out;               /* [B][Cm][HW] or NULL */
} ndb_loss_params;
int ndb_diffusion_loss(const ndb_loss_params* p, void* stream);
/* g[r][:] *= upstream[r]: chain rule of the per-sample loss vector (backward of ndb_diffusion_loss) */
int ndb_scale_rows(float* g, const float* upstream, int64_t rows, int64_t per_row, void* stream);

/* measurement aid: writes the device's %globaltimer (ns) to *dst when everything enqueued before it on `stream` has
 * completed (scripts/timeline.py brackets every launch of a step with these, inside the step's CUDA graph). */
int ndb_timestamp(uint64_t* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NDB200_H */
