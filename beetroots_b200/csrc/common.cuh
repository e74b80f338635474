// Shared device helpers: error handling, Philox4x32-10, special functions, the per-line
// mixing-noise likelihood and the prior terms.  All formulas cite the reference file:line
// (relative to /root/reference/beetroots/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#include "../../include/beetroots_b200.h"

#define BT_MAX_T 4          // forward-mode tangents = sampled NN inputs

// ------------------------------------------------------------------------------------------
// Per (pixel, line) observation constants, precomputed on the host in float64 and stored as
// 3 x float4 (48 B).  All quantities that are combined with ln f are shifted by the per-line
// output bias `shift_l` of the network (the MLP kernel returns lf' = ln f - shift_l, |lf'| ~ 1,
// so that float32 keeps ~1e-7 absolute accuracy instead of ulp(20.7) ~ 2e-6), and intensities
// are expressed in units of sigma_a (1e-10 W/m2/sr squared would underflow float32).
// ------------------------------------------------------------------------------------------
struct __align__(16) ObsConst {
    float lsa;     // ln sigma_a - shift
    float sm2;     // sigma_m^2
    float c;       // exp(sigma_m^2) - 1
    float yr;      // y / sigma_a
    float ly;      // ln y - shift        (NaN if y <= 0)
    float wr;      // omega / sigma_a
    float lw;      // ln omega - shift
    float fm1;     // log_fm1 - shift
    float fp1;     // log_fp1 - shift
    float lsa0;    // ln sigma_a   (unshifted: enters ln s_a and the final constant, :560)
    float lsm;     // ln sigma_m
    uint32_t flags;  // bit0: y <= omega (value branch, :1042)  bit1: y == omega (gradient branch, :655)
};
static_assert(sizeof(ObsConst) == 48, "ObsConst must be 48 bytes");

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: noise is a pure function of
// (seed, iteration, purpose|site, global pixel id, sub-index) -> identical chains for any
// sharding of the map.
// ------------------------------------------------------------------------------------------
#define BT_PURPOSE_PMALA_Z 1u
#define BT_PURPOSE_PMALA_U 2u
#define BT_PURPOSE_MTM_Z 3u
#define BT_PURPOSE_MTM_SUBSET 4u
#define BT_PURPOSE_MTM_USEL 5u
#define BT_PURPOSE_MTM_UACC 6u
#define BT_PURPOSE_MTM_IND 9u     /* 7, 8: model check (sampler_kernels.cuh) */

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c[0] = hi1 ^ c[1] ^ k0;
    c[1] = lo1;
    c[2] = hi0 ^ c[3] ^ k1;
    c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint64_t seed) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// counter layout: c0 = global pixel id (low 32), c1 = sub-index (candidate k, draw block),
// c2 = purpose | site << 8 | pixel id high bits << 16, c3 = iteration t (low 32); t high bits fold in the key.
__device__ __forceinline__ void bt_philox(uint32_t (&out)[4], uint64_t seed, int64_t t, uint32_t purpose,
                                          int site, int64_t gid, uint32_t sub) {
    out[0] = (uint32_t)gid;
    out[1] = sub;
    out[2] = purpose | ((uint32_t)site << 8) | ((uint32_t)((uint64_t)gid >> 32) << 16);
    out[3] = (uint32_t)t;
    philox4x32_10(out, seed ^ ((uint64_t)((uint64_t)t >> 32) << 32));
}

// uniform in (0, 1): 24 random bits, never 0 or 1, exactly representable
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Box-Muller: two uniforms -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float r = sqrtf(-2.0f * logf(u01(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// D (<= BT_MAX_D) standard normals for (purpose, site, gid, sub)
__device__ __forceinline__ void bt_normals(float (&z)[BT_MAX_D], int D, uint64_t seed, int64_t t,
                                           uint32_t purpose, int site, int64_t gid, uint32_t sub) {
    uint32_t r[4];
    bt_philox(r, seed, t, purpose, site, gid, sub * 2u);
    float a, b;
    box_muller(r[0], r[1], a, b);
    z[0] = a;
    z[1] = b;
    box_muller(r[2], r[3], a, b);
    z[2] = a;
    z[3] = b;
    if (D > 4) {
        bt_philox(r, seed, t, purpose, site, gid, sub * 2u + 1u);
        box_muller(r[0], r[1], a, b);
        z[4] = a;
    } else {
        z[4] = 0.f;
    }
}

__device__ __forceinline__ float bt_uniform(uint64_t seed, int64_t t, uint32_t purpose, int site, int64_t gid) {
    uint32_t r[4];
    bt_philox(r, seed, t, purpose, site, gid, 0u);
    return u01(r[0]);
}

// ------------------------------------------------------------------------------------------
// special functions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float nan_to_num(float x) {  // numpy.nan_to_num in float32
    if (isnan(x)) return 0.f;
    if (isinf(x)) return x > 0.f ? FLT_MAX : -FLT_MAX;
    return x;
}

// log Phi(z) (scipy.special.log_ndtr), stable for very negative z through erfcx
__device__ __forceinline__ float log_ndtr_f(float z) {
    const float t = -z * 0.70710678118654752f;
    if (z > 0.f) return log1pf(-0.5f * erfcf(-t));          // Phi = 1 - erfc(z/sqrt2)/2
    if (z > -1.f) return logf(0.5f * erfcf(t));
    return logf(0.5f * erfcxf(t)) - t * t;                   // erfc(t) = erfcx(t) exp(-t^2)
}

// Mills-type ratio phi(x)/Phi(x) (likelihoods/utils.py:45-58) = sqrt(2/pi) / erfcx(-x/sqrt2)
__device__ __forceinline__ float norm_pdf_cdf_ratio_f(float x) {
    return 0.79788456080286536f / erfcxf(-x * 0.70710678118654752f);
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_prime_f(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// ------------------------------------------------------------------------------------------
// Mixing-noise likelihood for one (row, line):  value and d nll / d ln f.
// Every theta-gradient of the reference factors as  (scalar) x grad ln f  because all
// intermediate quantities depend on theta only through ln f (approx_censored_add_mult.py
// :1072-1136, :385-405), so the kernel returns the scalar G and the caller contracts it with
// the Jacobian:  grad_d = sum_l nan_to_num(G_l J_dl).
// lf is the SHIFTED log-intensity (see ObsConst).
// ------------------------------------------------------------------------------------------
template <bool WITH_GRAD>
__device__ __forceinline__ void mixing_line(float lf, const ObsConst& o, float& nll, float& G) {
    const float a = lf - o.lsa;                   // ln(f / sigma_a)
    const float r = expf(a);                    // f / sigma_a
    const float cr2 = o.c * r * r;
    const float sar2 = cr2 + 1.0f;                // (s_a / sigma_a)^2      :1056-1061
    const float sar = sqrtf(sar2);
    const float q = expf(-2.0f * a - o.sm2);    // exp(2 log_combination) = 1/E   :1047-1050
    const float sm2_ = o.sm2 + log1pf(q);         // s_m^2 = -2 m_m          :1063-1066
    const float m_m = -0.5f * sm2_;
    const float s_m = sqrtf(sm2_);
    // mixing weight lambda :341-352
    float lam, dlam = 0.f;
    if (lf <= o.fm1) {
        lam = 1.0f;
    } else if (lf >= o.fp1) {
        lam = 0.0f;
    } else {
        const float inv_w = 1.0f / (o.fp1 - o.fm1);
        const float u = (lf - o.fm1) * inv_w;
        lam = ((((-6.0f * u + 15.0f) * u - 10.0f) * u) * u) * u + 1.0f;
        dlam = ((((-30.0f * u + 60.0f) * u - 30.0f) * u) * u) * inv_w;     // :385-405 (per unit grad ln f)
    }
    const float da = o.yr - r;                    // (y - f - m_a) / sigma_a, m_a = 0  :1055
    const float za = da / sar;
    // The reference evaluates the four negative log-densities (additive / multiplicative x uncensored / censored) and
    // SELECTS two of them by y <= omega (value) and two by y == omega (gradient, :655).  The flags are constants of the
    // (pixel, line), i.e. uniform over the warp's candidates: the pair that is not selected is never computed (exactly the
    // same result; the censored pair costs two log_ndtr evaluations).
    const bool cens = o.flags & 1u, eq = o.flags & 2u;
    const bool need_c = cens || (WITH_GRAD && eq), need_u = !cens || (WITH_GRAD && !eq);
    float nll_au = 0.f, nll_mu = 0.f, nll_ac = 0.f, nll_mc = 0.f;
    if (need_u) {
        const float ln_sa = o.lsa0 + 0.5f * logf(sar2);    // ln s_a (true scale)
        nll_au = 0.5f * za * za + ln_sa;                                // :14-19
        const float dm = o.ly - lf - m_m;
        nll_mu = 0.5f * dm * dm / sm2_ + 0.5f * logf(sm2_);             // :22-33
    }
    if (need_c) {
        nll_ac = nan_to_num(-log_ndtr_f(za));                           // :574-585 with y in the omega slot (:1212)
        const float zmc = (o.lw - lf - m_m) / s_m;
        nll_mc = nan_to_num(-log_ndtr_f(zmc));                          // :604-616
    }
    const float val = lam * (cens ? nll_ac : nll_au) + (1.0f - lam) * (cens ? nll_mc : nll_mu);   // :553-557
    nll = nan_to_num(val) - (o.lsa0 + o.lsm);                                                     // :559-560
    if (WITH_GRAD) {
        // all terms are coefficients of grad ln f
        const float g_sa2_over = 2.0f * cr2 / sar2;                 // grad_s_a2 / s_a2 = 2 f^2 c / s_a2
        const float gmm = q / (1.0f + q);                           // multiplicative model :1107-1136: grad m_m = q/(1+q) grad lnf
        if (!eq) {
            // gradient_cost_au :36-59, rescaled by sigma_a:  r_ = (f - y)/sigma_a = -da
            const float g_au = 0.5f * g_sa2_over + (-da) * r * (sar2 + da * r * o.c) / (sar2 * sar2);
            const float rm = lf - m_m - o.ly;                       // sign quirk of :73,78
            float g_mu = -gmm / sm2_;                               // 0.5 grad_s_m2 / s_m2
            if (!isnan(o.ly)) g_mu += 0.5f * rm / (sm2_ * sm2_) * (2.0f * sm2_ * (1.0f + gmm) + rm * 2.0f * gmm);
            G = lam * g_au + dlam * nll_au + (1.0f - lam) * g_mu - dlam * nll_mu;
        } else {
            // censored branches :676-692, :709-728 (selected by y == omega :655)
            const float ta = r - o.wr;                              // (f + m_a - omega)/sigma_a
            const float g_ac = norm_pdf_cdf_ratio_f(-ta / sar) * (r * sar - ta * cr2 / sar) / sar2;
            const float tm = lf + m_m - o.lw;
            const float g_mc = norm_pdf_cdf_ratio_f(-tm / s_m) * ((1.0f + gmm) * s_m + tm * gmm / s_m) / sm2_;
            G = lam * g_ac + dlam * nll_ac + (1.0f - lam) * g_mc - dlam * nll_mc;
        }
    }
}
