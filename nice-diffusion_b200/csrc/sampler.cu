// Fused per-step sampler kernels (HBM-bound, elementwise over NCHW fp32 images).
//
// One launch per denoising step replaces ~200 ATen calls of the reference: variance interpolation,
// classifier-free-guidance combine, x0 prediction + clamp, posterior mean / DDIM update, classifier-guidance
// shift, Gaussian noise (injected or generated in-kernel with Philox4x32-10 + Box-Muller, keyed by the
// GLOBAL sample index so results do not depend on how the batch is sharded over GPUs) and the t==0 mask.
// Arithmetic is written with explicit round-to-nearest mul/add (no FMA contraction) so that, given the same
// model output and noise, it matches the reference's separate fp32 torch ops.
#include "common.cuh"

namespace ndb {

// ------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) -- counter-based, stateless.
// ------------------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// four N(0,1) values for (sample id, step, group-of-4 index)
__device__ __forceinline__ void philox_normal4(uint64_t seed, int64_t sample_id, int step, uint32_t group, float* z) {
    const Philox4 r = philox4x32_10(group, (uint32_t)step, (uint32_t)sample_id, (uint32_t)((uint64_t)sample_id >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = ((float)r.v[0] + 0.5f) * k, u1 = ((float)r.v[1] + 0.5f) * k;
    const float u2 = ((float)r.v[2] + 0.5f) * k, u3 = ((float)r.v[3] + 0.5f) * k;
    // (x+0.5)*2^-32 can round to 1.0f for x close to 2^32; clamp away from 0 on the log side
    const float a0 = fminf(u0, 0.99999994f), a2 = fminf(u2, 0.99999994f);
    const float r0 = sqrtf(-2.f * logf(a0)), r1 = sqrtf(-2.f * logf(a2));
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

struct StepArgs {
    ndb_step_params p;
    int step_for_rng;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__global__ void __launch_bounds__(256)
sampler_step_kernel(const ndb_step_params p) {
    const int per = p.C * p.HW;          // elements per sample (x)
    const int groups = (per + 3) >> 2;   // groups of 4 elements share one Philox call
    const long long total = (long long)p.B * groups;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total;
         gi += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(gi / groups);
        const int g = (int)(gi - (long long)b * groups);
        const float tf = p.t[b];
        const int ti = (int)tf;  // t.long()
        const float* T = p.tables;
        const float srac = T[NDB_TAB_SQRT_RECIP_ABAR * p.R + ti];
        const float srm1 = T[NDB_TAB_SQRT_RECIPM1_ABAR * p.R + ti];
        const float mask = (tf == 0.f) ? 0.f : 1.f;

        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.noise == nullptr) {
            const int64_t sid = p.sample_ids ? p.sample_ids[b] : (int64_t)b;
            philox_normal4(p.seed, sid, ti, (uint32_t)g, z);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = g * 4 + k;
            if (e >= per) break;
            const int c = e / p.HW;
            const int hw = e - c * p.HW;
            const size_t xi = (size_t)b * per + e;
            const size_t mi = ((size_t)b * p.Cm + c) * p.HW + hw;
            const float x = p.x_t[xi];
            float eps = p.model_out[mi];
            float log_var;
            if (p.var_type == NDB_VAR_FIXED) {
                log_var = T[NDB_TAB_LOGVAR_MIN * p.R + ti];
            } else {
                const float v = p.model_out[mi + (size_t)p.C * p.HW];
                if (p.var_type == NDB_VAR_LEARNED) {
                    log_var = v;
                } else {
                    const float min_log = T[NDB_TAB_LOGVAR_MIN * p.R + ti];
                    const float max_log = T[NDB_TAB_LOGVAR_MAX * p.R + ti];
                    const float frac = add(v, 1.f) / 2.f;
                    log_var = add(mul(frac, max_log), mul(sub(1.f, frac), min_log));
                }
            }
            const float nz = p.noise ? p.noise[xi] : z[k];
            float out, x0;
            if (!p.ddim) {
                if (p.null_out != nullptr)
                    eps = sub(mul((float)(1.0 + p.cfg_strength), eps), mul((float)p.cfg_strength, p.null_out[mi]));
                x0 = sub(mul(srac, x), mul(srm1, eps));
                if (p.clip_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
                float mean = add(mul(T[NDB_TAB_POST_COEF_X0 * p.R + ti], x0), mul(T[NDB_TAB_POST_COEF_XT * p.R + ti], x));
                if (p.grad != nullptr) mean = add(mean, mul(mul((float)p.clf_strength, p.grad[xi]), expf(log_var)));
                out = add(mean, mul(mul(mask, expf(mul(0.5f, log_var))), nz));
            } else {
                if (p.grad != nullptr)
                    eps = sub(eps, mul(mul((float)p.clf_strength, p.grad[xi]), T[NDB_TAB_SQRT_1M_ABAR * p.R + ti]));
                else if (p.null_out != nullptr)
                    eps = sub(mul((float)(1.0 + p.cfg_strength), eps), mul((float)p.cfg_strength, p.null_out[mi]));
                x0 = sub(mul(srac, x), mul(srm1, eps));
                if (p.clip_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
                const float abar = T[NDB_TAB_ABAR * p.R + ti];
                const float abar_prev = T[NDB_TAB_ABAR_PREV * p.R + ti];
                // eta**2 is a python float (fp64) multiplied into an fp32 tensor
                const float eta2 = (float)(p.eta * p.eta);
                const float var = mul(mul(eta2, sub(1.f, abar_prev)), sub(1.f, abar / abar_prev)) / sub(1.f, abar);
                const float mean = add(mul(x0, sqrtf(abar_prev)), mul(sqrtf(sub(sub(1.f, abar_prev), var)), eps));
                out = add(mean, mul(mul(mask, sqrtf(var)), nz));
            }
            p.x_prev[xi] = out;
            if (p.x_prev2 != nullptr) p.x_prev2[xi] = out;
            if (p.pred_x0 != nullptr) p.pred_x0[xi] = x0;
        }
    }
}

__global__ void philox_normal_kernel(float* __restrict__ out, int B, long long per, uint64_t seed,
                                     const int64_t* __restrict__ sample_ids, int step) {
    const long long groups = (per + 3) >> 2;
    const long long total = (long long)B * groups;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total;
         gi += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(gi / groups);
        const long long g = gi - (long long)b * groups;
        float z[4];
        philox_normal4(seed, sample_ids ? sample_ids[b] : (int64_t)b, step, (uint32_t)g, z);
        for (int k = 0; k < 4; ++k) {
            const long long e = g * 4 + k;
            if (e < per) out[(size_t)b * per + e] = z[k];
        }
    }
}

__global__ void qsample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                               const float* __restrict__ t, const float* __restrict__ sa, const float* __restrict__ s1,
                               float* __restrict__ out, int B, long long per) {
    const long long total = (long long)B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const int ti = (int)t[b];
        out[i] = __fadd_rn(__fmul_rn(sa[ti], x0[i]), __fmul_rn(s1[ti], noise[i]));
    }
}

__global__ void advance_step_kernel(float* t_step, int B, float* t_model, int Bm, const float* __restrict__ tmap) {
    const float cur = t_step[0];
    __syncthreads();
    const float nxt = cur > 0.f ? cur - 1.f : 0.f;
    const float tm = tmap[(int)nxt];
    for (int i = threadIdx.x; i < B; i += blockDim.x) t_step[i] = nxt;
    for (int i = threadIdx.x; i < Bm; i += blockDim.x) t_model[i] = tm;
}

static inline int grid1d(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_sampler_step(const ndb_step_params* p, void* stream) {
    NDB_CHECK_ARG(p && p->x_t && p->model_out && p->t && p->tables && p->x_prev);
    NDB_CHECK_ARG(p->B > 0 && p->C > 0 && p->HW > 0 && p->R > 0);
    NDB_CHECK_ARG(p->var_type >= NDB_VAR_FIXED && p->var_type <= NDB_VAR_LEARNED_INTERP);
    if (p->var_type != NDB_VAR_FIXED && !p->ddim) NDB_CHECK_ARG(p->Cm == 2 * p->C);
    NDB_CHECK_ARG(p->Cm == p->C || p->Cm == 2 * p->C);
    const long long groups = (long long)p->B * (((long long)p->C * p->HW + 3) / 4);
    sampler_step_kernel<<<grid1d(groups, 256), 256, 0, (cudaStream_t)stream>>>(*p);
    return launch_status("ndb_sampler_step");
}

extern "C" int ndb_philox_normal(float* out, int B, int64_t per_sample, uint64_t seed, const int64_t* sample_ids,
                                 int32_t step, void* stream) {
    NDB_CHECK_ARG(out && B > 0 && per_sample > 0);
    const long long groups = (long long)B * ((per_sample + 3) / 4);
    philox_normal_kernel<<<grid1d(groups, 256), 256, 0, (cudaStream_t)stream>>>(out, B, per_sample, seed, sample_ids, step);
    return launch_status("ndb_philox_normal");
}

extern "C" int ndb_qsample(const float* x0, const float* noise, const float* t, const float* sqrt_abar,
                           const float* sqrt_1m_abar, float* out, int B, int64_t per_sample, void* stream) {
    NDB_CHECK_ARG(x0 && noise && t && sqrt_abar && sqrt_1m_abar && out && B > 0 && per_sample > 0);
    qsample_kernel<<<grid1d((long long)B * per_sample, 256), 256, 0, (cudaStream_t)stream>>>(x0, noise, t, sqrt_abar,
                                                                                           sqrt_1m_abar, out, B, per_sample);
    return launch_status("ndb_qsample");
}

extern "C" int ndb_advance_step(float* t_step, int B, float* t_model, int Bm, const float* timestep_map_f32,
                                void* stream) {
    NDB_CHECK_ARG(t_step && t_model && timestep_map_f32 && B > 0 && Bm > 0);
    advance_step_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(t_step, B, t_model, Bm, timestep_map_f32);
    return launch_status("ndb_advance_step");
}
