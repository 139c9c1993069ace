// Fused training loss of the diffusion model, forward AND gradient w.r.t. the model output, in one pass over the batch
// (reference: nicediffusion/diffusion.py:375-438 loss / variational_lower_bound, :499-549 kl_div / approx_cdf /
// log_likelihood / mean_flat). The reference evaluates ~60 elementwise torch ops per step here, each a launch and a full
// read + write of the [B, C, H, W] tensors, and autograd replays them backwards; this kernel reads x_0, x_t, noise and the
// model output once and writes the per-sample loss plus d loss_b / d model_out.
//
//   simple      : loss_b = mean((eps - noise)^2)
//   KL          : loss_b = vlb_b                       (gradient through the mean AND the variance)
//   KL_rescaled : loss_b = vlb_b * R
//   hybrid      : loss_b = mean((eps - noise)^2) + 0.001 * R * vlb_b(eps detached: gradient through the variance only)
//   vlb_b = (t_b == 0 ? mean(-discretised log-likelihood of x_0) : mean(KL(q(x_{t-1}|x_t,x_0) || p(x_{t-1}|x_t)))) / ln 2
#include "common.cuh"

namespace ndb {

constexpr int kLossThreads = 256;

struct LossArgs {
    const float* x0; const float* xt; const float* noise; const float* out;
    const float* t; const float* tab;
    int R, B, C, HW, Cm, var_type, loss_type, steps;
    float* loss; float* grad;
};

__device__ __forceinline__ float cdf_approx(float z, float& dcdf) {
    // Page (1977): 0.5 (1 + tanh(sqrt(2/pi) (z + 0.0444715 z^3)))
    const float k = 0.7978845608028654f, a = 0.0444715f;
    const float th = tanhf(k * (z + a * (z * z * z)));
    dcdf = 0.5f * (1.f - th * th) * k * (1.f + 3.f * a * z * z);
    return 0.5f * (1.f + th);
}

__global__ void __launch_bounds__(kLossThreads)
diffusion_loss_kernel(const LossArgs p) {
    const int b = blockIdx.x;
    const int D = p.C * p.HW;
    const int ti = (int)p.t[b];
    const float* T = p.tab;
    const float c0 = T[NDB_TAB_POST_COEF_X0 * p.R + ti], ct = T[NDB_TAB_POST_COEF_XT * p.R + ti];
    const float sra = T[NDB_TAB_SQRT_RECIP_ABAR * p.R + ti], srm1 = T[NDB_TAB_SQRT_RECIPM1_ABAR * p.R + ti];
    const float tlv = T[NDB_TAB_LOG_POST_VAR_CLIPPED * p.R + ti];
    const float lv_min = T[NDB_TAB_LOGVAR_MIN * p.R + ti], lv_max = T[NDB_TAB_LOGVAR_MAX * p.R + ti];
    const bool learned = p.var_type != NDB_VAR_FIXED;
    const bool want_simple = p.loss_type == 0 || p.loss_type == 3;
    const bool want_vlb = p.loss_type != 0;
    const bool eps_in_vlb = p.loss_type == 1 || p.loss_type == 2;      // hybrid: eps is detached inside the bound
    const float inv_ln2 = 1.f / 0.6931471805599453f;
    const float vlb_w = (p.loss_type == 2 ? (float)p.steps : p.loss_type == 3 ? 0.001f * (float)p.steps : 1.f);
    const float inv_D = 1.f / (float)D;
    const size_t xoff = (size_t)b * D, ooff = (size_t)b * p.Cm * p.HW;

    float s_simple = 0.f, s_vlb = 0.f;
    for (int e = threadIdx.x; e < D; e += kLossThreads) {
        const float x0 = p.x0[xoff + e], xt = p.xt[xoff + e], nz = p.noise[xoff + e];
        const float eps = p.out[ooff + e];
        float g_eps = 0.f, g_v = 0.f;
        if (want_simple) {
            const float dlt = eps - nz;
            s_simple += dlt * dlt;
            g_eps += 2.f * dlt * inv_D;
        }
        if (want_vlb) {
            float lv, dlv_dv = 0.f;
            if (p.var_type == NDB_VAR_LEARNED) { lv = p.out[ooff + D + e]; dlv_dv = 1.f; }
            else if (p.var_type == NDB_VAR_LEARNED_INTERP) {
                const float frac = (p.out[ooff + D + e] + 1.f) / 2.f;
                lv = frac * lv_max + (1.f - frac) * lv_min;
                dlv_dv = 0.5f * (lv_max - lv_min);
            } else lv = lv_min;
            const float tm = c0 * x0 + ct * xt;
            const float px0 = sra * xt - srm1 * eps;
            const float mean = c0 * px0 + ct * xt;
            float term, d_mean, d_lv;
            if (ti != 0) {
                const float dm = tm - mean, iv = expf(-lv), ev = expf(tlv - lv);
                term = ((lv - tlv) + ev + (dm * dm) * iv - 1.0f) / 2.f;
                d_lv = 0.5f * (1.f - ev - dm * dm * iv);
                d_mean = -dm * iv;
            } else {
                const float s = expf(-0.5f * lv), cen = x0 - mean;
                const float zp = (cen + 1.0f / 255.0f) * s, zm = (cen - 1.0f / 255.0f) * s;
                float dcp, dcm;
                const float cp = cdf_approx(zp, dcp), cm = cdf_approx(zm, dcm);
                float ll, dll_dzp = 0.f, dll_dzm = 0.f;
                if (x0 < -0.999f) {
                    ll = logf(fmaxf(cp, 1e-12f));
                    if (cp >= 1e-12f) dll_dzp = dcp / cp;
                } else if (x0 > 0.999f) {
                    const float om = 1.0f - cm;
                    ll = logf(fmaxf(om, 1e-12f));
                    if (om >= 1e-12f) dll_dzm = -dcm / om;
                } else {
                    const float dl = cp - cm;
                    ll = logf(fmaxf(dl, 1e-12f));
                    if (dl >= 1e-12f) { dll_dzp = dcp / dl; dll_dzm = -dcm / dl; }
                }
                term = -ll;
                // dz/dmean = -s, dz/dlv = -z/2
                d_mean = -(dll_dzp + dll_dzm) * (-s);
                d_lv = -(dll_dzp * (-0.5f * zp) + dll_dzm * (-0.5f * zm));
            }
            s_vlb += term;
            const float w = vlb_w * inv_ln2 * inv_D;
            if (learned) g_v = w * d_lv * dlv_dv;
            if (eps_in_vlb) g_eps += w * d_mean * (-c0 * srm1);
        }
        if (p.grad != nullptr) {
            p.grad[ooff + e] = g_eps;
            if (p.Cm == 2 * p.C) p.grad[ooff + D + e] = g_v;
        }
    }
    // fixed-order block reduction (deterministic)
    __shared__ float red[2][kLossThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_simple += __shfl_xor_sync(0xffffffffu, s_simple, o);
        s_vlb += __shfl_xor_sync(0xffffffffu, s_vlb, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_simple; red[1][threadIdx.x >> 5] = s_vlb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, v = 0.f;
        for (int w = 0; w < kLossThreads / 32; ++w) { a += red[0][w]; v += red[1][w]; }
        const float simple = a * inv_D;
        const float vlb = (v * inv_D) * inv_ln2;
        float l = 0.f;
        if (p.loss_type == 0) l = simple;
        else if (p.loss_type == 1) l = vlb;
        else if (p.loss_type == 2) l = vlb * (float)p.steps;
        else l = simple + 0.001f * (vlb * (float)p.steps);
        p.loss[b] = l;
    }
}

// grad_out[b][...] *= upstream[b] (backward of the per-sample loss vector)
__global__ void scale_rows_kernel(float* __restrict__ g, const float* __restrict__ up, long long per, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        g[i] *= up[i / per];
}

}  // namespace ndb

using namespace ndb;

extern "C" int ndb_diffusion_loss(const ndb_loss_params* q, void* stream) {
    NDB_CHECK_ARG(q && q->x0 && q->x_t && q->noise && q->model_out && q->t && q->tables && q->loss);
    NDB_CHECK_ARG(q->B > 0 && q->C > 0 && q->HW > 0 && q->R > 0);
    NDB_CHECK_ARG(q->Cm == q->C || q->Cm == 2 * q->C);
    NDB_CHECK_ARG(q->loss_type >= 0 && q->loss_type <= 3);
    NDB_CHECK_ARG(q->var_type >= NDB_VAR_FIXED && q->var_type <= NDB_VAR_LEARNED_INTERP);
    if (q->var_type != NDB_VAR_FIXED) NDB_CHECK_ARG(q->Cm == 2 * q->C);
    LossArgs a{q->x0, q->x_t, q->noise, q->model_out, q->t, q->tables, q->R, q->B, q->C, q->HW, q->Cm, q->var_type,
               q->loss_type, q->rescaled_num_steps, q->loss, q->grad_out};
    diffusion_loss_kernel<<<q->B, kLossThreads, 0, (cudaStream_t)stream>>>(a);
    return launch_status("ndb_diffusion_loss");
}

extern "C" int ndb_scale_rows(float* g, const float* upstream, int64_t rows, int64_t per_row, void* stream) {
    NDB_CHECK_ARG(g && upstream && rows > 0 && per_row > 0);
    const long long total = rows * per_row;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scale_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, upstream, per_row, total);
    return launch_status("ndb_scale_rows");
}
