// fusion.cu — the fused MoPoE kernel (forward + backward).
//
// One warp per sample row; lanes stride the latent dimension with float4 loads, so every HBM access
// is a coalesced 512 B row segment.  For each (b, d) the thread keeps the M expert precisions in
// registers and walks all 2^M-1 subsets: precision-weighted product (or positional mixture), KL term,
// joint-mixture selection by batch row, reparameterisation.  Per-subset KL sums are reduced
// warp -> per-sample fp64 partial -> fixed-order final sum (deterministic, no atomics).
//
// Reference: utils/BaseMMVae.py:101-196, evaluation/divergence_measures/mm_div.py:10-17,
// kl_div.py:8-16, utils/utils.py:45-77 (math in SURVEY.md Appendix C).
#include "common.cuh"

#define POE_EPS 1e-8f
constexpr int FUS_WARPS = 4;
constexpr int MAXM = 4;

struct FusionPtrs {
    const float* mu[MAXM];
    const float* lv[MAXM];
    float* dmu[MAXM];
    float* dlv[MAXM];
};

struct FusionDev {
    int M, B, D, nsub, S, fuse_mode, prior_expert;
    int members[16], stacked[16], sel_end[16];
    int mem_cnt[16];
    int mem_idx[16][MAXM];
    int mem_end[16][MAXM];
    float inv_norm;
};

__device__ __forceinline__ int joint_component(const FusionDev& c, int b) {
    int k = 0;
    while (k < c.S - 1 && b >= c.sel_end[k]) ++k;
    return c.stacked[k];
}
// which member of subset s supplies row b in mixture mode (utils.mixture_component_selection)
__device__ __forceinline__ int moe_member(const FusionDev& c, int s, int b) {
    const int n = c.mem_cnt[s];
    for (int j = 0; j < n - 1; ++j)
        if (b < c.mem_end[s][j]) return c.mem_idx[s][j];
    return c.mem_idx[s][n - 1];
}
__device__ __forceinline__ float pick(const float4 (&v)[MAXM], int i, int q) {
    float r = 0.f;
#pragma unroll
    for (int ii = 0; ii < MAXM; ++ii)
        if (ii == i) r = (&v[ii].x)[q];
    return r;
}

__global__ void __launch_bounds__(FUS_WARPS * 32) fusion_fwd_kernel(FusionDev c, FusionPtrs p, const float* __restrict__ eps,
                                                                    float* __restrict__ sub_mu, float* __restrict__ sub_lv,
                                                                    float* __restrict__ jmu, float* __restrict__ jlv,
                                                                    float* __restrict__ z, double* __restrict__ kl_part,
                                                                    int* __restrict__ nan_flag) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * FUS_WARPS + (threadIdx.x >> 5);
    if (b >= c.B) return;
    const int kj = joint_component(c, b);
    float klacc[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) klacc[s] = 0.f;
    bool bad = false;
    const long long BD = (long long)c.B * c.D;
    for (int d = lane * 4; d < c.D; d += 128) {
        const long long off = (long long)b * c.D + d;
        float4 mu[MAXM], lv[MAXM], T[MAXM];
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            if (i < c.M) {
                mu[i] = *reinterpret_cast<const float4*>(p.mu[i] + off);
                lv[i] = *reinterpret_cast<const float4*>(p.lv[i] + off);
                const float* m_ = &mu[i].x; const float* l_ = &lv[i].x; float* t_ = &T[i].x;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    t_[q] = 1.f / (expf(l_[q]) + POE_EPS);
                    bad |= isnan(m_[q]) | isnan(l_[q]);
                }
            }
        }
        const float4 e4 = *reinterpret_cast<const float4*>(eps + off);
        const float tprior = 1.f / (1.f + POE_EPS);   // expert N(0,I): exp(0) + eps
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            if (s < c.nsub) {
                float4 om, ol;
                float* om_ = &om.x; float* ol_ = &ol.x;
                if (c.fuse_mode == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float P = 0.f, num = 0.f;     // torch.sum(dim=0) order = stacking order, prior last
#pragma unroll
                        for (int j = 0; j < MAXM; ++j)
                            if (j < c.mem_cnt[s]) {
                                const int i = c.mem_idx[s][j];
                                const float t = pick(T, i, q);
                                P += t;
                                num += pick(mu, i, q) * t;
                            }
                        if (c.prior_expert) P += tprior;
                        om_[q] = num / P;
                        ol_[q] = logf(1.f / P);
                    }
                } else {
                    const int im = moe_member(c, s, b);
#pragma unroll
                    for (int i = 0; i < MAXM; ++i)
                        if (i == im) { om = mu[i]; ol = lv[i]; }
                }
                *reinterpret_cast<float4*>(sub_mu + (long long)s * BD + off) = om;
                *reinterpret_cast<float4*>(sub_lv + (long long)s * BD + off) = ol;
#pragma unroll
                for (int q = 0; q < 4; ++q) klacc[s] += 1.f - expf(ol_[q]) - om_[q] * om_[q] + ol_[q];
                if (s == kj) {
                    *reinterpret_cast<float4*>(jmu + off) = om;
                    *reinterpret_cast<float4*>(jlv + off) = ol;
                    float4 zz;
                    zz.x = e4.x * expf(0.5f * ol.x) + om.x;
                    zz.y = e4.y * expf(0.5f * ol.y) + om.y;
                    zz.z = e4.z * expf(0.5f * ol.z) + om.z;
                    zz.w = e4.w * expf(0.5f * ol.w) + om.w;
                    *reinterpret_cast<float4*>(z + off) = zz;
                }
            }
        }
        if (kj < 0) {       // jsd: this batch row belongs to the prior component N(0, I) of the mixture (BaseMMVae.py:180-186)
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(jmu + off) = zero;
            *reinterpret_cast<float4*>(jlv + off) = zero;
            *reinterpret_cast<float4*>(z + off) = e4;               // eps * exp(0) + 0
        }
    }
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        if (s < c.nsub) {
            double v = warp_sum((double)klacc[s]);
            if (lane == 0) kl_part[(long long)s * c.B + b] = v;
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(nan_flag, 1);
}

__global__ void fusion_kl_finalize(const double* kl_part, int B, int nsub, float inv_norm, float* kl) {
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int b0 = 0; b0 < B; b0 += 32) {          // fixed order: 32-wide strips, lanes reduced by shuffle
        int b = b0 + threadIdx.x;
        double v = b < B ? kl_part[(long long)s * B + b] : 0.0;
        acc += warp_sum(v);
    }
    if (threadIdx.x == 0) kl[s] = (float)(-0.5 * acc * (double)inv_norm);
}

static int fill_dev(const mopoe_fusion_cfg_t* cfg, FusionDev& d) {
    MOPOE_REQUIRE(cfg->M >= 1 && cfg->M <= MAXM, "fusion: M=%d (max %d)", cfg->M, MAXM);
    MOPOE_REQUIRE(cfg->nsub >= 1 && cfg->nsub <= 15, "fusion: nsub=%d", cfg->nsub);
    MOPOE_REQUIRE(cfg->S >= 1 && cfg->S <= 16, "fusion: S=%d", cfg->S);
    MOPOE_REQUIRE(cfg->D % 4 == 0, "fusion: D=%d must be a multiple of 4", cfg->D);
    d.M = cfg->M; d.B = cfg->B; d.D = cfg->D; d.nsub = cfg->nsub; d.S = cfg->S;
    d.fuse_mode = cfg->fuse_mode; d.prior_expert = cfg->prior_expert;
    d.inv_norm = 1.f / cfg->norm;
    for (int i = 0; i < 16; ++i) {
        d.members[i] = cfg->members[i];
        d.stacked[i] = cfg->stacked[i];
        d.sel_end[i] = cfg->sel_end[i];
        d.mem_cnt[i] = cfg->mem_cnt[i];
        for (int j = 0; j < MAXM; ++j) { d.mem_end[i][j] = cfg->mem_end[i][j]; d.mem_idx[i][j] = cfg->mem_idx[i][j]; }
    }
    for (int j = 0; j < cfg->S; ++j)
        MOPOE_REQUIRE(cfg->stacked[j] >= -1 && cfg->stacked[j] < cfg->nsub, "fusion: stacked[%d]=%d", j, cfg->stacked[j]);
    return 0;
}

extern "C" int mopoe_fusion_fwd(const mopoe_fusion_cfg_t* cfg, const float* const* mu, const float* const* logvar,
                                const float* eps, float* sub_mu, float* sub_lv, float* joint_mu, float* joint_lv,
                                float* z, float* kl, int32_t* nan_flag, double* ws, void* stream) {
    FusionDev d;
    if (fill_dev(cfg, d)) return 1;
    FusionPtrs p = {};
    for (int i = 0; i < cfg->M; ++i) { p.mu[i] = mu[i]; p.lv[i] = logvar[i]; }
    cudaStream_t st = (cudaStream_t)stream;
    fusion_fwd_kernel<<<(cfg->B + FUS_WARPS - 1) / FUS_WARPS, FUS_WARPS * 32, 0, st>>>(d, p, eps, sub_mu, sub_lv, joint_mu,
                                                                                      joint_lv, z, ws, nan_flag);
    MOPOE_CHECK_LAUNCH("fusion_fwd");
    fusion_kl_finalize<<<cfg->nsub, 32, 0, st>>>(ws, cfg->B, cfg->nsub, d.inv_norm, kl);
    MOPOE_CHECK_LAUNCH("fusion_kl_finalize");
    return 0;
}

__global__ void __launch_bounds__(FUS_WARPS * 32) fusion_bwd_kernel(FusionDev c, FusionPtrs p, const float* __restrict__ eps,
                                                                    const float* __restrict__ sub_mu, const float* __restrict__ sub_lv,
                                                                    const float* __restrict__ d_z, const float* __restrict__ d_jmu,
                                                                    const float* __restrict__ d_jlv, const float* __restrict__ d_smu,
                                                                    const float* __restrict__ d_slv, const float* __restrict__ d_kl) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * FUS_WARPS + (threadIdx.x >> 5);
    if (b >= c.B) return;
    const int kj = joint_component(c, b);
    const long long BD = (long long)c.B * c.D;
    const float tprior = 1.f / (1.f + POE_EPS);
    for (int d = lane * 4; d < c.D; d += 128) {
        const long long off = (long long)b * c.D + d;
        float4 mu[MAXM], lv[MAXM], T[MAXM], gmu[MAXM], glv[MAXM];
#pragma unroll
        for (int i = 0; i < MAXM; ++i) {
            gmu[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            glv[i] = gmu[i];
            if (i < c.M) {
                mu[i] = *reinterpret_cast<const float4*>(p.mu[i] + off);
                lv[i] = *reinterpret_cast<const float4*>(p.lv[i] + off);
#pragma unroll
                for (int q = 0; q < 4; ++q) (&T[i].x)[q] = 1.f / (expf((&lv[i].x)[q]) + POE_EPS);
            }
        }
        const float4 e4 = *reinterpret_cast<const float4*>(eps + off);
        float4 gz = d_z ? *reinterpret_cast<const float4*>(d_z + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 gjm = d_jmu ? *reinterpret_cast<const float4*>(d_jmu + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 gjl = d_jlv ? *reinterpret_cast<const float4*>(d_jlv + off) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            if (s < c.nsub) {
                const float4 sm = *reinterpret_cast<const float4*>(sub_mu + (long long)s * BD + off);
                const float4 sl = *reinterpret_cast<const float4*>(sub_lv + (long long)s * BD + off);
                float4 Gm = d_smu ? *reinterpret_cast<const float4*>(d_smu + (long long)s * BD + off)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 Gl = d_slv ? *reinterpret_cast<const float4*>(d_slv + (long long)s * BD + off)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                const float ck = d_kl ? d_kl[s] * c.inv_norm : 0.f;
                const int im = c.fuse_mode ? moe_member(c, s, b) : -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float m_s = (&sm.x)[q], l_s = (&sl.x)[q];
                    float g_m = (&Gm.x)[q] + ck * m_s;                       // d(-0.5*sum(1-e^lv-mu^2+lv))/dmu = mu
                    float g_l = (&Gl.x)[q] + ck * 0.5f * (expf(l_s) - 1.f);
                    if (s == kj) {
                        g_m += (&gjm.x)[q] + (&gz.x)[q];
                        g_l += (&gjl.x)[q] + (&gz.x)[q] * 0.5f * (&e4.x)[q] * expf(0.5f * l_s);
                    }
                    if (c.fuse_mode == 0) {
                        float P = c.prior_expert ? tprior : 0.f;
#pragma unroll
                        for (int i = 0; i < MAXM; ++i)
                            if (i < c.M && ((c.members[s] >> i) & 1)) P += (&T[i].x)[q];
                        const float invP = 1.f / P;
#pragma unroll
                        for (int i = 0; i < MAXM; ++i)
                            if (i < c.M && ((c.members[s] >> i) & 1)) {
                                const float t = (&T[i].x)[q];
                                (&gmu[i].x)[q] += g_m * t * invP;
                                const float dT = g_m * ((&mu[i].x)[q] - m_s) * invP - g_l * invP;
                                (&glv[i].x)[q] += dT * (-t * t * expf((&lv[i].x)[q]));
                            }
                    } else {
#pragma unroll
                        for (int i = 0; i < MAXM; ++i)
                            if (i == im) {
                                (&gmu[i].x)[q] += g_m;
                                (&glv[i].x)[q] += g_l;
                            }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MAXM; ++i)
            if (i < c.M) {
                *reinterpret_cast<float4*>(p.dmu[i] + off) = gmu[i];
                *reinterpret_cast<float4*>(p.dlv[i] + off) = glv[i];
            }
    }
}

extern "C" int mopoe_fusion_bwd(const mopoe_fusion_cfg_t* cfg, const float* const* mu, const float* const* logvar,
                                const float* eps, const float* sub_mu, const float* sub_lv, const float* d_z,
                                const float* d_joint_mu, const float* d_joint_lv, const float* d_sub_mu,
                                const float* d_sub_lv, const float* d_kl, float* const* d_mu, float* const* d_lv,
                                void* stream) {
    FusionDev d;
    if (fill_dev(cfg, d)) return 1;
    FusionPtrs p = {};
    for (int i = 0; i < cfg->M; ++i) { p.mu[i] = mu[i]; p.lv[i] = logvar[i]; p.dmu[i] = d_mu[i]; p.dlv[i] = d_lv[i]; }
    fusion_bwd_kernel<<<(cfg->B + FUS_WARPS - 1) / FUS_WARPS, FUS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d, p, eps, sub_mu, sub_lv, d_z, d_joint_mu, d_joint_lv, d_sub_mu, d_sub_lv, d_kl);
    MOPOE_CHECK_LAUNCH("fusion_bwd");
    return 0;
}

// ---- alpha-JSD divergence with a dynamic prior (jsd mode) ----------------------------------------------------------------
// Reference: BaseMMVae.divergence_dynamic_prior (utils/BaseMMVae.py:87-99) -> calc_alphaJSD_modalities (mm_div.py:67-87)
// -> alpha_poe (mm_div.py:20-32) + calc_kl_divergence two-Gaussian branch (kl_div.py:11-13).
//   T_i = 1/(exp(lv_i) + 1e-8);  V = 1/sum_i a_i T_i;  m = V * sum_i a_i mu_i T_i;  L = log V     (the dynamic prior)
//   kl_k = -0.5 * sum_{b,d} (1 - exp(lv_k)/exp(L) - (mu_k - m)^2/exp(L) + lv_k - L) / norm
constexpr int JSD_MAXK = 5;
struct JsdArgs {
    const float* mu[JSD_MAXK];
    const float* lv[JSD_MAXK];
    float* dmu[JSD_MAXK];
    float* dlv[JSD_MAXK];
    float alpha[JSD_MAXK];
    int K, B, D;
    float inv_norm;
};

__global__ void __launch_bounds__(FUS_WARPS * 32) jsd_fwd_kernel(JsdArgs a, float* __restrict__ dyn_mu, float* __restrict__ dyn_lv,
                                                                 double* __restrict__ kl_part) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * FUS_WARPS + (threadIdx.x >> 5);
    if (b >= a.B) return;
    float acc[JSD_MAXK];
#pragma unroll
    for (int k = 0; k < JSD_MAXK; ++k) acc[k] = 0.f;
    for (int d = lane; d < a.D; d += 32) {
        const long long off = (long long)b * a.D + d;
        float mu[JSD_MAXK], lv[JSD_MAXK];
        float S = 0.f, N = 0.f;
#pragma unroll
        for (int k = 0; k < JSD_MAXK; ++k)
            if (k < a.K) {
                mu[k] = a.mu[k][off];
                lv[k] = a.lv[k][off];
                const float T = 1.f / (expf(lv[k]) + POE_EPS);
                S += a.alpha[k] * T;                       // torch.sum(dim=0): stacking order
                N += a.alpha[k] * mu[k] * T;
            }
        const float V = 1.f / S, m = V * N, Lg = logf(V), eL = expf(Lg);
        dyn_mu[off] = m;
        dyn_lv[off] = Lg;
#pragma unroll
        for (int k = 0; k < JSD_MAXK; ++k)
            if (k < a.K) {
                const float dm = mu[k] - m;
                acc[k] += 1.f - expf(lv[k]) / eL - dm * dm / eL + lv[k] - Lg;
            }
    }
#pragma unroll
    for (int k = 0; k < JSD_MAXK; ++k)
        if (k < a.K) {
            const double v = warp_sum((double)acc[k]);
            if (lane == 0) kl_part[(long long)k * a.B + b] = v;
        }
}

__global__ void __launch_bounds__(FUS_WARPS * 32) jsd_bwd_kernel(JsdArgs a, const float* __restrict__ d_kl) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * FUS_WARPS + (threadIdx.x >> 5);
    if (b >= a.B) return;
    float ck[JSD_MAXK];
#pragma unroll
    for (int k = 0; k < JSD_MAXK; ++k) ck[k] = k < a.K ? d_kl[k] * a.inv_norm : 0.f;
    for (int d = lane; d < a.D; d += 32) {
        const long long off = (long long)b * a.D + d;
        float mu[JSD_MAXK], ev[JSD_MAXK], T[JSD_MAXK];
        float S = 0.f, N = 0.f;
#pragma unroll
        for (int k = 0; k < JSD_MAXK; ++k)
            if (k < a.K) {
                mu[k] = a.mu[k][off];
                ev[k] = expf(a.lv[k][off]);
                T[k] = 1.f / (ev[k] + POE_EPS);
                S += a.alpha[k] * T[k];
                N += a.alpha[k] * mu[k] * T[k];
            }
        const float V = 1.f / S, m = V * N, iV = S;
        // J = sum_k ck * KL_k:  gradients through the dynamic prior (m, V) and the direct terms
        float Gm = 0.f, GV = 0.f;
#pragma unroll
        for (int k = 0; k < JSD_MAXK; ++k)
            if (k < a.K) {
                const float dm = mu[k] - m;
                Gm += ck[k] * (-dm * iV);
                GV += ck[k] * (-0.5f) * (ev[k] * iV * iV + dm * dm * iV * iV - iV);
            }
#pragma unroll
        for (int k = 0; k < JSD_MAXK; ++k)
            if (k < a.K) {
                const float dm = mu[k] - m;
                const float g_mu = ck[k] * dm * iV + Gm * V * a.alpha[k] * T[k];
                const float dT = Gm * a.alpha[k] * V * dm - GV * a.alpha[k] * V * V;
                const float g_lv = 0.5f * ck[k] * (ev[k] * iV - 1.f) + dT * (-T[k] * T[k] * ev[k]);
                a.dmu[k][off] = g_mu;
                a.dlv[k][off] = g_lv;
            }
    }
}

static int fill_jsd(int K, int B, int D, const float* const* mu, const float* const* logvar, const float* alpha, float norm,
                    JsdArgs& a) {
    MOPOE_REQUIRE(K >= 1 && K <= JSD_MAXK, "jsd_divergence: K=%d (max %d)", K, JSD_MAXK);
    MOPOE_REQUIRE(B >= 1 && D >= 1 && norm > 0.f, "jsd_divergence: bad sizes");
    a = JsdArgs{};
    a.K = K; a.B = B; a.D = D; a.inv_norm = 1.f / norm;
    for (int k = 0; k < K; ++k) {
        MOPOE_REQUIRE(mu[k] && logvar[k], "jsd_divergence: null expert %d", k);
        a.mu[k] = mu[k]; a.lv[k] = logvar[k]; a.alpha[k] = alpha[k];
    }
    return 0;
}
extern "C" int mopoe_jsd_divergence_fwd(int K, int B, int D, const float* const* mu, const float* const* logvar,
                                        const float* alpha, float norm, float* dyn_mu, float* dyn_lv, float* kl, double* ws,
                                        void* stream) {
    JsdArgs a;
    if (fill_jsd(K, B, D, mu, logvar, alpha, norm, a)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    jsd_fwd_kernel<<<(B + FUS_WARPS - 1) / FUS_WARPS, FUS_WARPS * 32, 0, st>>>(a, dyn_mu, dyn_lv, ws);
    MOPOE_CHECK_LAUNCH("jsd_divergence_fwd");
    fusion_kl_finalize<<<K, 32, 0, st>>>(ws, B, K, a.inv_norm, kl);
    MOPOE_CHECK_LAUNCH("jsd_kl_finalize");
    return 0;
}
extern "C" int mopoe_jsd_divergence_bwd(int K, int B, int D, const float* const* mu, const float* const* logvar,
                                        const float* alpha, float norm, const float* d_kl, float* const* d_mu,
                                        float* const* d_lv, void* stream) {
    JsdArgs a;
    if (fill_jsd(K, B, D, mu, logvar, alpha, norm, a)) return 1;
    for (int k = 0; k < K; ++k) {
        MOPOE_REQUIRE(d_mu[k] && d_lv[k], "jsd_divergence_bwd: null gradient buffer %d", k);
        a.dmu[k] = d_mu[k]; a.dlv[k] = d_lv[k];
    }
    jsd_bwd_kernel<<<(B + FUS_WARPS - 1) / FUS_WARPS, FUS_WARPS * 32, 0, (cudaStream_t)stream>>>(a, d_kl);
    MOPOE_CHECK_LAUNCH("jsd_divergence_bwd");
    return 0;
}
