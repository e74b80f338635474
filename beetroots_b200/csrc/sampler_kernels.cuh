// Likelihood / prior / sampler-update kernels (float32, HBM-bound, one thread or one warp per
// pixel).  Replaces, per colour class, the bodies of
//   MySampler.generate_new_sample_pmala_rmsprop  (sampler/my_sampler.py:752-903)
//   MySampler.generate_new_sample_mtm            (sampler/my_sampler.py:1006-1197)
// and the pieces of Posterior.compute_all they call (modelling/posterior.py:332-446).
#pragma once
#include "common.cuh"
#include "mlp_simt.cuh"

struct PriorDev {
    int has_spatial, has_indicator;
    float tau[BT_MAX_D];        // prior_spatial.weights
    float tau_prop[BT_MAX_D];   // min(initial_weights, weights)  (my_sampler.py:149-152, 1097-1100)
    float lo[BT_MAX_D], hi[BT_MAX_D];
    float inv_delta;            // 1 / indicator_margin_scale
    float g4;                   // 4 / delta^4
};

// ---- priors -------------------------------------------------------------------------------
// sum over the neighbour set V_n of (x_d - theta_id)^2 and (x_d - theta_id)
// (l22_discrete_grad_prior.py:9-80 evaluated through the neighbour table)
__device__ __forceinline__ void neighbour_sums(const float* __restrict__ theta, const int* __restrict__ nbr_n,
                                               int max_degree, int D, const float (&x)[BT_MAX_D],
                                               float (&had)[BT_MAX_D], float (&lap)[BT_MAX_D]) {
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) had[d] = lap[d] = 0.f;
    for (int j = 0; j < max_degree; ++j) {
        const int m = nbr_n[j];
        if (m < 0) continue;
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d)
            if (d < D) {
                const float diff = x[d] - theta[(size_t)m * D + d];
                had[d] = fmaf(diff, diff, had[d]);
                lap[d] += diff;
            }
    }
}

// smooth indicator: penalty (smooth_indicator_prior.py:35-54) and gradient (:57-82)
__device__ __forceinline__ float indicator_terms(const PriorDev& pr, int D, const float (&x)[BT_MAX_D],
                                                 float (&g)[BT_MAX_D]) {
    float pen = 0.f;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) {
        g[d] = 0.f;
        if (d < D) {
            float dist = 0.f;
            if (x[d] > pr.hi[d]) dist = x[d] - pr.hi[d];
            else if (x[d] < pr.lo[d]) dist = x[d] - pr.lo[d];
            const float sc = fabsf(dist) * pr.inv_delta;
            const float sc2 = sc * sc;
            pen += sc2 * sc2;
            g[d] = pr.g4 * dist * dist * dist;
        }
    }
    return pen;
}

// prior part of the per-pixel objective (chromatic convention, factor 1) and of the gradient
__device__ __forceinline__ float prior_eval(const PriorDev& pr, const float* __restrict__ theta,
                                            const int* __restrict__ nbr_n, int max_degree, int D,
                                            const float (&x)[BT_MAX_D], float (&grad)[BT_MAX_D],
                                            float* spatial_out = nullptr) {
    float obj = 0.f, sp = 0.f;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) grad[d] = 0.f;
    if (pr.has_spatial) {
        float had[BT_MAX_D], lap[BT_MAX_D];
        neighbour_sums(theta, nbr_n, max_degree, D, x, had, lap);
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d)
            if (d < D) {
                sp = fmaf(pr.tau[d], had[d], sp);                 // l22:107,125-128 (chromatic)
                grad[d] = 2.f * pr.tau[d] * lap[d];               // l22:137-139
            }
        obj += sp;
    }
    if (spatial_out) *spatial_out = sp;
    if (pr.has_indicator) {
        float gi[BT_MAX_D];
        obj += indicator_terms(pr, D, x, gi);
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) grad[d] += gi[d];
    }
    return obj;
}

// likelihood of one row: nll and (optionally) gradient wrt the sampled dims
template <bool WITH_GRAD>
__device__ __forceinline__ float lik_row(const FmapDev& fm, int L, const float* __restrict__ lf_row,
                                         const float* __restrict__ jac_row, const ObsConst* __restrict__ obs_n,
                                         float (&grad)[BT_MAX_D]) {
    float nll = 0.f;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) grad[d] = 0.f;
    for (int l = 0; l < L; ++l) {
        const float4* op = reinterpret_cast<const float4*>(obs_n + l);
        union { float4 v[3]; ObsConst o; } u;
        u.v[0] = __ldg(op); u.v[1] = __ldg(op + 1); u.v[2] = __ldg(op + 2);
        float v, G;
        mixing_line<WITH_GRAD>(lf_row[l], u.o, v, G);
        nll += v;                                                    // :568
        if (WITH_GRAD) {
#pragma unroll
            for (int d = 0; d < BT_MAX_D; ++d)
                if (d < fm.D) {
                    const float J = (fm.tslot[d] < 0) ? 1.0f : jac_row[fm.tslot[d] * L + l];
                    grad[d] += nan_to_num(G * J);                    // :673-674
                }
        }
    }
    return nll;
}

// ---- compute_all: likelihood terms of every pixel from (lf, jac) --------------------------
__global__ void lik_all_kernel(FmapDev fm, int N, int L, const float* __restrict__ lf_rows,
                               const float* __restrict__ jac_rows, const ObsConst* __restrict__ obs,
                               float* __restrict__ nll_lik, float* __restrict__ grad_lik, float* __restrict__ lf_cache) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float g[BT_MAX_D];
    const float nll = lik_row<true>(fm, L, lf_rows + (size_t)n * L, jac_rows + (size_t)n * fm.T * L,
                                    obs + (size_t)n * L, g);
    nll_lik[n] = nll;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < fm.D) grad_lik[(size_t)n * fm.D + d] = g[d];
    if (lf_cache != lf_rows)
        for (int l = 0; l < L; ++l) lf_cache[(size_t)n * L + l] = lf_rows[(size_t)n * L + l];
}

// objective / gradient assembly for bt_get_current and bt_init_v_from_grad (posterior.py:402-446, 212-241)
__global__ void assemble_kernel(PriorDev pr, int N, int D, int max_degree, const float* __restrict__ theta,
                                const int* __restrict__ nbr, const float* __restrict__ nll_lik,
                                const float* __restrict__ grad_lik, float* __restrict__ obj_chrom,
                                float* __restrict__ obj_glob, float* __restrict__ grad, float* __restrict__ v_init,
                                int* __restrict__ j_init) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float x[BT_MAX_D], gp[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? theta[(size_t)n * D + d] : 0.f;
    float sp;
    const float pobj = prior_eval(pr, theta, nbr + (size_t)n * max_degree, max_degree, D, x, gp, &sp);
    if (obj_chrom) obj_chrom[n] = nll_lik[n] + pobj;
    if (obj_glob) obj_glob[n] = nll_lik[n] + pobj - 0.5f * sp;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            const float g = nan_to_num(grad_lik[(size_t)n * D + d] + gp[d]);
            if (grad) grad[(size_t)n * D + d] = g;
            if (v_init) v_init[(size_t)n * D + d] = g * g;
            if (j_init) j_init[(size_t)n * D + d] = 0;
        }
}

// ---- PMALA ---------------------------------------------------------------------------------
// propose (my_sampler.py:756-808): one thread per pixel of the colour class
__global__ void pmala_propose_kernel(PriorDev pr, int n_pix, const int* __restrict__ site_pix, int D, int max_degree,
                                     const float* __restrict__ theta, const int* __restrict__ nbr,
                                     const float* __restrict__ nll_lik, const float* __restrict__ grad_lik,
                                     const float* __restrict__ v, float eps, float eta, int chromatic,
                                     uint64_t seed, int64_t t, int site, const int64_t* __restrict__ gid,
                                     const float* __restrict__ z_inject,
                                     float* __restrict__ cand_rows, float* __restrict__ logq_fwd,
                                     float* __restrict__ obj_cur) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int n = site_pix[i];
    float x[BT_MAX_D], gp[BT_MAX_D], z[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? theta[(size_t)n * D + d] : 0.f;
    float sp;
    const float pobj = prior_eval(pr, theta, nbr + (size_t)n * max_degree, max_degree, D, x, gp, &sp);
    obj_cur[i] = nll_lik[n] + pobj - (chromatic ? 0.f : 0.5f * sp);       // :745-750, :857
    if (z_inject) {
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) z[d] = (d < D) ? z_inject[(size_t)n * D + d] : 0.f;
    } else {
        bt_normals(z, D, seed, t, BT_PURPOSE_PMALA_Z, site, gid ? gid[n] : (int64_t)n, 0u);
    }
    float lq = 0.f;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            const float g = nan_to_num(grad_lik[(size_t)n * D + d] + gp[d]);     // posterior.py:240
            const float G = 1.0f / (eta + sqrtf(v[(size_t)n * D + d]));          // :761
            const float zs = z[d] * sqrtf(eps * G);                             // :768
            const float mu = x[d] - 0.5f * eps * G * g;                          // :796-800
            const float c = mu + zs;                                             // :802
            const float dc = c - mu;
            lq += -0.5f * logf(G) - dc * dc / (2.0f * eps * G);                  // :804-808
            cand_rows[(size_t)i * D + d] = c;
        }
    logq_fwd[i] = lq;
}

// accept / reject and state update (my_sampler.py:822-896)
__global__ void pmala_accept_kernel(PriorDev pr, FmapDev fm, int n_pix, const int* __restrict__ site_pix, int D, int L,
                                    int max_degree, float* __restrict__ theta, const int* __restrict__ nbr,
                                    const ObsConst* __restrict__ obs, float* __restrict__ nll_lik,
                                    float* __restrict__ grad_lik, float* __restrict__ lf_cache,
                                    float* __restrict__ v, int* __restrict__ jt,
                                    const float* __restrict__ cand_rows, const float* __restrict__ lf_rows,
                                    const float* __restrict__ jac_rows, const float* __restrict__ logq_fwd,
                                    const float* __restrict__ obj_cur, float eps, float eta, float alpha,
                                    int chromatic, uint64_t seed, int64_t t, int site,
                                    const int64_t* __restrict__ gid, const float* __restrict__ u_inject,
                                    float* __restrict__ accept, float* __restrict__ logpa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int n = site_pix[i];
    float c[BT_MAX_D], gp[BT_MAX_D], gl[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) c[d] = (d < D) ? cand_rows[(size_t)i * D + d] : 0.f;
    const float nll_c = lik_row<true>(fm, L, lf_rows + (size_t)i * L, jac_rows + (size_t)i * fm.T * L,
                                      obs + (size_t)n * L, gl);
    float sp;
    const float pobj = prior_eval(pr, theta, nbr + (size_t)n * max_degree, max_degree, D, c, gp, &sp);
    const float obj_c = nll_c + pobj - (chromatic ? 0.f : 0.5f * sp);
    float lq = 0.f;
    float vnew[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            const float g = nan_to_num(gl[d] + gp[d]);
            const float vc = alpha * v[(size_t)n * D + d] + (1.0f - alpha) * g * g;   // :823-825
            const float G = 1.0f / (eta + sqrtf(vc));                                 // :826
            const float mu = c[d] - 0.5f * eps * G * g;                               // :841-845
            const float dc = theta[(size_t)n * D + d] - mu;
            lq += -0.5f * logf(G) - dc * dc / (2.0f * eps * G);                       // :847-851
            vnew[d] = vc;
        }
    const float lpa = -obj_c + obj_cur[i] + lq - logq_fwd[i];                         // :865-870
    const float u = u_inject ? u_inject[n] : bt_uniform(seed, t, BT_PURPOSE_PMALA_U, site, gid ? gid[n] : (int64_t)n);
    const bool acc = logf(u) < lpa;                                                   // :873-874
    accept[n] = acc ? 1.f : 0.f;
    logpa[n] = lpa;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            v[(size_t)n * D + d] = vnew[d];                                           // :886-888 (always)
            jt[(size_t)n * D + d] = acc ? 0 : jt[(size_t)n * D + d] + 1;              // :890-896
            if (acc) {
                theta[(size_t)n * D + d] = c[d];                                      // :876-880
                grad_lik[(size_t)n * D + d] = gl[d];
            }
        }
    if (acc) {
        nll_lik[n] = nll_c;
        for (int l = 0; l < L; ++l) lf_cache[(size_t)n * L + l] = lf_rows[(size_t)i * L + l];
    }
}

// ---- MTM -----------------------------------------------------------------------------------
// One draw from the smooth indicator prior of dimension d (sampler/utils/utils.py:50-97, Appendix B of the paper):
// with probability w = Z_u / (1 + Z_u), Z_u = 2 (hi - lo) / (delta Gamma(1/4)), uniform in [lo, hi]; otherwise a
// generalised-Gaussian tail exp(-(x/delta)^4) glued to the nearer bound (utils.py:12-47: x = +-Z^(1/4) with
// Z ~ Gamma(1/4, delta^4)).  Gamma(1/4): G_a = G_{a+1} U^{1/a} with Marsaglia-Tsang for shape 5/4, so
// Z^(1/4) = delta G_{5/4}^{1/4} U.  All randomness from Philox blocks (purpose MTM_IND, sub-index 16 (k D + d) + j).
__device__ __forceinline__ float smooth_indicator_draw(const PriorDev& pr, int d, uint64_t seed, int64_t t, int site,
                                                       int64_t gid, uint32_t sub) {
    const float lo = pr.lo[d], hi = pr.hi[d];
    const float zu = 2.f * (hi - lo) * pr.inv_delta * (1.0f / 3.6256099082219083f);
    uint32_t r[4];
    bt_philox(r, seed, t, BT_PURPOSE_MTM_IND, site, gid, sub * 16u);
    if (u01(r[0]) * (1.f + zu) < zu) return lo + (hi - lo) * u01(r[1]);                // :83-85, :96
    const float dd = 1.25f - 1.0f / 3.0f, cc = rsqrtf(9.f * dd);
    float gam = dd;                                                                   // (fallback after 15 rejections: p < 1e-20)
    for (uint32_t j = 1; j < 16u; ++j) {
        uint32_t q[4];
        bt_philox(q, seed, t, BT_PURPOSE_MTM_IND, site, gid, sub * 16u + j);
        float x, x_unused;
        box_muller(q[0], q[1], x, x_unused);
        const float v1 = 1.f + cc * x;
        if (v1 <= 0.f) continue;
        const float v = v1 * v1 * v1;
        if (logf(u01(q[2])) < 0.5f * x * x + dd - dd * v + dd * logf(v)) { gam = dd * v; break; }
    }
    const float mag = sqrtf(sqrtf(gam)) * u01(r[3]) / pr.inv_delta;                    // delta G^(1/4)
    return (r[2] & 1u) ? hi + mag : lo - mag;                                         // :87-95
}

// Row layout of the candidate buffers (cand_rows, lf_rows) of one colour class: the n_pix x K proposals first, [i][k], then
// the n_pix current values.  Only the proposals go through the forward map: ln f of a current value is already cached
// per pixel (lf_cache) and is copied behind them (mtm_current_lf_kernel) -- one forward-map row in K + 1 saved.
__device__ __forceinline__ size_t mtm_row(int i, int k, int K, int n_pix) {
    return k < K ? (size_t)i * K + k : (size_t)n_pix * K + i;
}
__global__ void mtm_current_lf_kernel(int n_pix, const int* __restrict__ site_pix, int L, int K,
                                      const float* __restrict__ lf_cache, float* __restrict__ lf_rows) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_pix * L) return;
    const int i = (int)(idx / L), l = (int)(idx - (long long)i * L);
    lf_rows[((size_t)n_pix * K + i) * L + l] = lf_cache[(size_t)site_pix[i] * L + l];
}

// candidate generation (sampler/utils/utils.py:274-349): thread per (pixel, k); slot K = current value
__global__ void mtm_gen_kernel(PriorDev pr, int n_pix, const int* __restrict__ site_pix, int D, int K, int max_degree,
                               const float* __restrict__ theta, const int* __restrict__ nbr, uint64_t seed,
                               int64_t t, int site, const int64_t* __restrict__ gid,
                               const float* __restrict__ cand_inject, float* __restrict__ cand_rows) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_pix * (K + 1)) return;
    const int i = (int)(idx / (K + 1)), k = (int)(idx - (long long)i * (K + 1));
    const int n = site_pix[i];
    float* out = cand_rows + mtm_row(i, k, K, n_pix) * D;
    if (k == K) {                                                                     // my_sampler.py:1011-1015
        for (int d = 0; d < D; ++d) out[d] = theta[(size_t)n * D + d];
        return;
    }
    if (cand_inject) {
        for (int d = 0; d < D; ++d) out[d] = cand_inject[((size_t)n * K + k) * D + d];
        return;
    }
    if (!pr.has_spatial) {                                                            // my_sampler.py:135-143
        const int64_t g0 = gid ? gid[n] : (int64_t)n;
        for (int d = 0; d < D; ++d) out[d] = smooth_indicator_draw(pr, d, seed, t, site, g0, (uint32_t)(k * BT_MAX_D + d));
        return;
    }
    const int* nb = nbr + (size_t)n * max_degree;
    int deg = 0;
    for (int j = 0; j < max_degree; ++j) deg += (nb[j] >= 0);
    if (deg == 0) {                                                                   // utils.py:321-322: zeros
        for (int d = 0; d < D; ++d) out[d] = 0.f;
        return;
    }
    const int64_t g = gid ? gid[n] : (int64_t)n;
    uint32_t r[4];
    bt_philox(r, seed, t, BT_PURPOSE_MTM_SUBSET, site, g, (uint32_t)k);
    // uniform over the 2^deg - 1 non-empty subsets (= Bernoulli(1/2) per neighbour, redrawn while empty, :326-328)
    const uint32_t mask = 1u + (uint32_t)(((uint64_t)r[0] * (uint64_t)((1u << deg) - 1u)) >> 32);
    float mean[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) mean[d] = 0.f;
    int m = 0, slot = 0;
    for (int j = 0; j < max_degree; ++j) {
        if (nb[j] < 0) continue;
        if ((mask >> slot) & 1u) {
            ++m;
#pragma unroll
            for (int d = 0; d < BT_MAX_D; ++d)
                if (d < D) mean[d] += theta[(size_t)nb[j] * D + d];
        }
        ++slot;
    }
    float z[BT_MAX_D];
    bt_normals(z, D, seed, t, BT_PURPOSE_MTM_Z, site, g, (uint32_t)k);
    const float inv_m = 1.0f / (float)m;
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) out[d] = mean[d] * inv_m + z[d] * inv_m * rsqrtf(pr.tau_prop[d]);  // :336-345
}

// warp reductions
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Negative log weight of every candidate row (my_sampler.py:1017-1105) for graphs of degree <= 4: ONE THREAD PER ROW (all
// lanes busy, where the warp-per-pixel kernel below idles 20 % of them at K + 1 = 51), everything in registers.  The
// proposal log-density (utils.py:419-495) needs, per dimension, log sum_S sqrt(tau) m_S exp(-(x - mu_S)^2 tau m_S^2) over
// the non-empty neighbour subsets S; since (x - mu_S) m_S = sum_{j in S} (x - theta_j), the exponent is
// -tau (sum_{j in S} a_j)^2 with a_j = x - theta_j: 11 additions give all 15 subset sums, no means, no divisions.
__global__ void __launch_bounds__(128)
mtm_nl_kernel(PriorDev pr, FmapDev fm, int n_pix, const int* __restrict__ site_pix, int D, int L, int K, int max_degree,
              const float* __restrict__ theta, const int* __restrict__ nbr, const ObsConst* __restrict__ obs,
              const float* __restrict__ cand_rows, const float* __restrict__ lf_rows, float* __restrict__ nl) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)n_pix * (K + 1)) return;
    const int i = (int)(row / (K + 1));
    const int n = site_pix[i];
    const size_t r = mtm_row(i, (int)(row - (long long)i * (K + 1)), K, n_pix);
    float x[BT_MAX_D], dummy[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? cand_rows[r * D + d] : 0.f;
    float val = lik_row<false>(fm, L, lf_rows + r * L, nullptr, obs + (size_t)n * L, dummy);                // :1021-1025
    if (pr.has_indicator) val += indicator_terms(pr, D, x, dummy);                                         // posterior.py:105-110
    if (pr.has_spatial) {
        // valid neighbours compacted to the front, in edge-list order
        int nb[4], deg = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) nb[j] = -1;
        for (int j = 0; j < max_degree && j < 4; ++j) {
            const int m = nbr[(size_t)n * max_degree + j];
            if (m >= 0) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) if (jj == deg) nb[jj] = m;
                ++deg;
            }
        }
        if (deg == 0) { nl[row] = -INFINITY; return; }               // isolated pixel: empty logsumexp
        const uint32_t nsub = (1u << deg) - 1u;
        float sp = 0.f, total = 0.f;
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) {
            if (d >= D) continue;
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = (j < deg) ? x[d] - theta[(size_t)nb[j] * D + d] : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) sp = fmaf(pr.tau[d] * a[j], a[j], sp);                              // posterior.py:96-103
            // subset sums, index = mask - 1
            float ssum[15];
            ssum[0] = a[0]; ssum[1] = a[1]; ssum[2] = a[0] + a[1]; ssum[3] = a[2];
            ssum[4] = a[0] + a[2]; ssum[5] = a[1] + a[2]; ssum[6] = ssum[2] + a[2]; ssum[7] = a[3];
            ssum[8] = a[0] + a[3]; ssum[9] = a[1] + a[3]; ssum[10] = ssum[2] + a[3]; ssum[11] = a[2] + a[3];
            ssum[12] = ssum[4] + a[3]; ssum[13] = ssum[5] + a[3]; ssum[14] = ssum[6] + a[3];
            const float tau = pr.tau_prop[d], st = sqrtf(tau);
            float e[15], mx = -INFINITY;                             // numba_logsumexp_stable, utils.py:390-416
#pragma unroll
            for (int q = 0; q < 15; ++q) {
                e[q] = -tau * ssum[q] * ssum[q];
                if ((uint32_t)q < nsub) mx = fmaxf(mx, e[q]);
            }
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 15; ++q) {
                const float m = (float)__popc((uint32_t)q + 1u);
                if ((uint32_t)q < nsub) acc = fmaf(st * m, expf(e[q] - mx), acc);
            }
            total += logf(acc) + mx;
        }
        val += sp + total;           // chromatic factor 1 (l22:103-105) + proposal density (my_sampler.py:1093-1105)
    }
    nl[row] = val;
}

// weights, selection, accept (my_sampler.py:1017-1178): one warp per pixel, lanes stride over candidates.
// Shared memory per warp: the means mu_S (and sizes m_S) of all 2^deg - 1 non-empty neighbour subsets, computed once
// per pixel and reused by every candidate for the proposal log-density
//   log q(x) = sum_d log sum_S sqrt(tau_d) m_S exp(-(x_d - mu_Sd)^2 tau_d m_S^2)          (utils.py:419-495)
#define MTM_WARPS 4
// NL_READY: nl_scratch already holds the negative log weights (mtm_nl_kernel below): selection / accept only
template <bool NL_READY>
__global__ void __launch_bounds__(MTM_WARPS * 32)
mtm_select_kernel(PriorDev pr, FmapDev fm, int n_pix, const int* __restrict__ site_pix, int D, int L,
                  int K, int max_degree, int chromatic, float* __restrict__ theta,
                  const int* __restrict__ nbr, const ObsConst* __restrict__ obs,
                  int* __restrict__ jt, const float* __restrict__ cand_rows,
                  const float* __restrict__ lf_rows, float* __restrict__ nl_scratch,
                  uint64_t seed, int64_t t, int site, const int64_t* __restrict__ gid,
                  const float* __restrict__ usel_inject, const float* __restrict__ uacc_inject,
                  float* __restrict__ accept, float* __restrict__ logpa, uint8_t* __restrict__ accepted_any) {
    extern __shared__ float mtm_smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * MTM_WARPS + wib;
    if (i >= n_pix) return;
    const int n = site_pix[i];
    const int K1 = K + 1;
    const int nsub_max = (1 << max_degree) - 1;
    float* nl = nl_scratch + (size_t)i * K1;
    float mn = INFINITY;
    if (NL_READY) {
        for (int k = lane; k < K1; k += 32) mn = fminf(mn, nl[k]);
    } else {
    float* mu = mtm_smem + (size_t)wib * nsub_max * (BT_MAX_D + 1);      // [nsub][D] then m_S in slot D
    // ---- neighbours (registers; loops are unrolled over BT_MAX_DEGREE with guards)
    const int* nb = nbr + (size_t)n * max_degree;
    float nbv[BT_MAX_DEGREE][BT_MAX_D];
    int deg = 0;
#pragma unroll
    for (int j = 0; j < BT_MAX_DEGREE; ++j) {
        const int m = j < max_degree ? nb[j] : -1;
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) nbv[j][d] = 0.f;
        if (m >= 0) {
            // compact valid neighbours to the front (neighbour order = edge-list order)
#pragma unroll
            for (int jj = 0; jj < BT_MAX_DEGREE; ++jj)
                if (jj == deg) {
#pragma unroll
                    for (int d = 0; d < BT_MAX_D; ++d) nbv[jj][d] = (d < D) ? theta[(size_t)m * D + d] : 0.f;
                }
            ++deg;
        }
    }
    const int nsub = (1 << deg) - 1;
    for (int sidx = lane; sidx < nsub; sidx += 32) {
        const uint32_t mask = (uint32_t)sidx + 1u;
        float sum[BT_MAX_D];
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) sum[d] = 0.f;
#pragma unroll
        for (int j = 0; j < BT_MAX_DEGREE; ++j)
            if ((mask >> j) & 1u) {
#pragma unroll
                for (int d = 0; d < BT_MAX_D; ++d) sum[d] += nbv[j][d];
            }
        const float m = (float)__popc(mask);
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) mu[sidx * (BT_MAX_D + 1) + d] = sum[d] / m;
        mu[sidx * (BT_MAX_D + 1) + BT_MAX_D] = m;
    }
    __syncwarp();
    // ---- negative log weight of every candidate
    for (int k = lane; k < K1; k += 32) {
        const size_t row = mtm_row(i, k, K, n_pix);
        float x[BT_MAX_D], dummy[BT_MAX_D];
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? cand_rows[row * D + d] : 0.f;
        float val = lik_row<false>(fm, L, lf_rows + row * L, nullptr, obs + (size_t)n * L, dummy);  // :1021-1025
        if (pr.has_spatial) {                                                                      // posterior.py:96-103
            float sp = 0.f;
#pragma unroll
            for (int j = 0; j < BT_MAX_DEGREE; ++j)
                if (j < deg) {
#pragma unroll
                    for (int d = 0; d < BT_MAX_D; ++d)
                        if (d < D) { const float df = x[d] - nbv[j][d]; sp = fmaf(pr.tau[d] * df, df, sp); }
                }
            val += sp;   // a colour class never holds all N pixels -> chromatic factor 1 (l22:103-105)
        }
        if (pr.has_indicator) val += indicator_terms(pr, D, x, dummy);                             // posterior.py:105-110
        if (pr.has_spatial && nsub > 0) {                                                          // my_sampler.py:1093-1105
            float total = 0.f;
#pragma unroll
            for (int d = 0; d < BT_MAX_D; ++d) {
                if (d >= D) continue;
                const float tau = pr.tau_prop[d], st = sqrtf(tau);
                float mx = -INFINITY;                                   // numba_logsumexp_stable, utils.py:390-416
                for (int sidx = 0; sidx < nsub; ++sidx) {
                    const float m = mu[sidx * (BT_MAX_D + 1) + BT_MAX_D];
                    const float diff = x[d] - mu[sidx * (BT_MAX_D + 1) + d];
                    mx = fmaxf(mx, -diff * diff * tau * m * m);
                }
                float acc = 0.f;
                for (int sidx = 0; sidx < nsub; ++sidx) {
                    const float m = mu[sidx * (BT_MAX_D + 1) + BT_MAX_D];
                    const float diff = x[d] - mu[sidx * (BT_MAX_D + 1) + d];
                    acc += st * m * expf(-diff * diff * tau * m * m - mx);
                }
                total += logf(acc) + mx;
            }
            val += total;
        } else if (pr.has_spatial) {
            val = -INFINITY;            // isolated pixel: empty logsumexp
        }
        nl[k] = val;
        mn = fminf(mn, val);
    }
    }   // !NL_READY
    mn = warp_min(mn);
    __syncwarp();
    // ---- selection, ratio, accept.  The reference works with w_k = exp(-(nl_k - min nl)) in float64 (:1107-1114);
    // float32 would flush w_k to 0 beyond 87 nats, so everything is done in the log domain, relative to the best
    // proposal (numerator, selection) and to the best remaining candidate (denominator).  A weight is dropped when
    // float64 would have underflowed (nl_k - min > 745.13), which reproduces the reference's non-finite -> accepted
    // quirk (:1156-1159).  All sums are warp reductions over candidates in a fixed order (deterministic).
    const float F64_UNDERFLOW = 745.13f;
    float dp = INFINITY;
    for (int k = lane; k < K; k += 32) dp = fminf(dp, nl[k] - mn);
    dp = warp_min(dp);
    float s_prop = 0.f, s_all = 0.f, s_tot = 0.f;
    for (int k = lane; k < K1; k += 32) {
        const float dk = nl[k] - mn;
        s_tot += expf(-dk);
        if (k < K) {
            const float w = expf(-(dk - dp));
            s_all += w;
            if (dk <= F64_UNDERFLOW) s_prop += w;
        }
    }
    s_prop = warp_sum(s_prop); s_all = warp_sum(s_all); s_tot = warp_sum(s_tot);
    const int64_t g = gid ? gid[n] : (int64_t)n;
    // inverse-CDF selection with one uniform: Generator.choice == searchsorted(cdf, u, 'right') (:1123-1131)
    const float us = usel_inject ? usel_inject[n] : bt_uniform(seed, t, BT_PURPOSE_MTM_USEL, site, g);
    const float target = us * s_all;
    int sel = K - 1;
    float base = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        float w = (k < K) ? expf(-((nl[k] - mn) - dp)) : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {                          // inclusive scan in candidate order
            const float up = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += up;
        }
        const float cum = base + w;
        const unsigned hit = __ballot_sync(0xffffffffu, (k < K) && (cum > target));
        if (hit) { sel = k0 + __ffs(hit) - 1; break; }
        base += __shfl_sync(0xffffffffu, w, 31);
    }
    // denominator: sum over all candidates but the selected one (:1142-1144)
    float d_o = INFINITY;
    for (int k = lane; k < K1; k += 32)
        if (k != sel) d_o = fminf(d_o, nl[k] - mn);
    d_o = warp_min(d_o);
    float s_den = 0.f;
    for (int k = lane; k < K1; k += 32) {
        const float dk = nl[k] - mn;
        if (k != sel && dk <= F64_UNDERFLOW) s_den += expf(-(dk - d_o));
    }
    s_den = warp_sum(s_den);
    if (lane == 0) {
        const float log_num = (s_prop > 0.f) ? logf(s_prop) - dp : -INFINITY;          // :1114
        float log_den = (s_den > 0.f) ? logf(s_den) - d_o : -INFINITY;
        // the reference forms the denominator as (sum of all weights) - w_sel in float64 (:1142-1144): when the other
        // weights are below 2^-53 of the total the difference is exactly 0 -> log_rg = +inf -> replaced by 1e-15
        if (log_den < logf(s_tot) - 36.7368f) log_den = -INFINITY;
        float log_rg = log_num - log_den;
        if (!isfinite(log_rg)) log_rg = 1e-15f;                                        // :1156-1159
        const float ua = uacc_inject ? uacc_inject[n] : bt_uniform(seed, t, BT_PURPOSE_MTM_UACC, site, g);
        const bool acc = logf(ua) < log_rg;
        accept[n] = acc ? 1.f : 0.f;
        logpa[n] = log_rg;
        if (acc) {
            accepted_any[n] = 1;
            const size_t row = mtm_row(i, sel, K, n_pix);
            for (int d = 0; d < D; ++d) {
                theta[(size_t)n * D + d] = cand_rows[row * D + d];                 // :1164-1168
                jt[(size_t)n * D + d] = 0;                                         // :1174-1178
            }
        }
    }
}

// end of the MTM iteration (my_sampler.py:1182-1197): likelihood refresh + v update on accepted pixels
__global__ void mtm_finish_kernel(PriorDev pr, FmapDev fm, int N, int D, int L, int max_degree,
                                  const float* __restrict__ theta, const int* __restrict__ nbr,
                                  const ObsConst* __restrict__ obs, const float* __restrict__ lf_rows,
                                  const float* __restrict__ jac_rows, float* __restrict__ nll_lik,
                                  float* __restrict__ grad_lik, float* __restrict__ lf_cache, float* __restrict__ v,
                                  float alpha, uint8_t* __restrict__ accepted_any) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float gl[BT_MAX_D];
    const float nll = lik_row<true>(fm, L, lf_rows + (size_t)n * L, jac_rows + (size_t)n * fm.T * L, obs + (size_t)n * L, gl);
    nll_lik[n] = nll;
    for (int l = 0; l < L; ++l) lf_cache[(size_t)n * L + l] = lf_rows[(size_t)n * L + l];
    float x[BT_MAX_D], gp[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? theta[(size_t)n * D + d] : 0.f;
    const bool upd = accepted_any[n] != 0;
    if (upd) prior_eval(pr, theta, nbr + (size_t)n * max_degree, max_degree, D, x, gp);
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            grad_lik[(size_t)n * D + d] = gl[d];
            if (upd) {
                const float g = nan_to_num(gl[d] + gp[d]);
                v[(size_t)n * D + d] = alpha * v[(size_t)n * D + d] + (1.0f - alpha) * g * g;   // :1189-1195
            }
        }
    accepted_any[n] = 0;
}

// same on the compacted list of accepted pixels (row i of lf_rows / jac_rows belongs to pixel list[i]): pixels that kept
// their value keep their cached likelihood terms, so only the accepted ones go through the forward map + Jacobian
__global__ void mtm_finish_compact_kernel(PriorDev pr, FmapDev fm, int n_acc, const int* __restrict__ list, int D, int L,
                                          int max_degree, const float* __restrict__ theta, const int* __restrict__ nbr,
                                          const ObsConst* __restrict__ obs, const float* __restrict__ lf_rows,
                                          const float* __restrict__ jac_rows, float* __restrict__ nll_lik,
                                          float* __restrict__ grad_lik, float* __restrict__ lf_cache, float* __restrict__ v,
                                          float alpha, uint8_t* __restrict__ accepted_any) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_acc) return;
    const int n = list[i];
    float gl[BT_MAX_D];
    const float nll = lik_row<true>(fm, L, lf_rows + (size_t)i * L, jac_rows + (size_t)i * fm.T * L, obs + (size_t)n * L, gl);
    nll_lik[n] = nll;
    for (int l = 0; l < L; ++l) lf_cache[(size_t)n * L + l] = lf_rows[(size_t)i * L + l];
    float x[BT_MAX_D], gp[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? theta[(size_t)n * D + d] : 0.f;
    prior_eval(pr, theta, nbr + (size_t)n * max_degree, max_degree, D, x, gp);
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            grad_lik[(size_t)n * D + d] = gl[d];
            const float g = nan_to_num(gl[d] + gp[d]);
            v[(size_t)n * D + d] = alpha * v[(size_t)n * D + d] + (1.0f - alpha) * g * g;           // :1189-1195
        }
    accepted_any[n] = 0;
}

// ---- online model checking (my_sampler.py:158-214, sampling branch) ---------------------------------------------
// Per saved iterate: draw a replicated observation y_rep = max(omega, eps_m f + eps_a)
// (approx_censored_add_mult.py:268-280), evaluate its negative log-likelihood, and update the running means
//   clppd[n,l]  <- mean of exp(-nll[n,l])                      (:169-170)
//   p_y[n,l]    <- mean of [y_rep <= y]                        (:194-195)
//   p_llh[n]    <- mean of [sum_l nll_rep >= sum_l nll]        (:199-203)
// Everything is expressed in units of sigma_a like the rest of the likelihood (eps_a / sigma_a ~ N(0,1)).
#define BT_PURPOSE_MCHK_A 7u
#define BT_PURPOSE_MCHK_M 8u
__global__ void model_check_kernel(int N, int L, const float* __restrict__ lf_cache, const ObsConst* __restrict__ obs,
                                   float count, uint64_t seed, int64_t t, const int64_t* __restrict__ gid,
                                   const float* __restrict__ za_inject, const float* __restrict__ zm_inject,
                                   float* __restrict__ clppd, float* __restrict__ pvy, float* __restrict__ pvl, int mode) {
    // mode 0: sampling branch (everything); 1: p-values only; 2: clppd <- exp(-nll) (optimisation branch, :208-212)
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t g = gid ? gid[n] : (int64_t)n;
    const float wgt_old = count / (count + 1.0f), wgt_new = 1.0f / (count + 1.0f);
    float nll_sum = 0.f, nll_rep_sum = 0.f;
    for (int l = 0; l < L; ++l) {
        const float4* op = reinterpret_cast<const float4*>(obs + (size_t)n * L + l);
        union { float4 v[3]; ObsConst o; } u;
        u.v[0] = __ldg(op); u.v[1] = __ldg(op + 1); u.v[2] = __ldg(op + 2);
        const float lf = lf_cache[(size_t)n * L + l];
        float nll, dummy;
        mixing_line<false>(lf, u.o, nll, dummy);
        if (mode == 2) { clppd[(size_t)n * L + l] = expf(-nll); continue; }
        float za, zm;
        if (za_inject) {
            za = za_inject[(size_t)n * L + l];
            zm = zm_inject[(size_t)n * L + l];
        } else {
            uint32_t r[4];
            bt_philox(r, seed, t, BT_PURPOSE_MCHK_A, 0, g, (uint32_t)l);
            box_muller(r[0], r[1], za, zm);
        }
        // y_rep / sigma_a = max(omega / sigma_a, eps_m f / sigma_a + eps_a / sigma_a)
        const float r_f = expf(lf - u.o.lsa);
        const float sm = sqrtf(u.o.sm2);
        const float eps_m = expf(-0.5f * u.o.sm2 + sm * zm);                      // lognormal(-sigma_m^2/2, sigma_m)
        const float yr_rep = fmaxf(u.o.wr, eps_m * r_f + za);
        ObsConst o2 = u.o;
        o2.yr = yr_rep;
        // NB ln(y) is NOT refreshed: the reference replaces likelihood_rep.y only (my_sampler.py:176-177), so the
        // multiplicative branch of the replicated likelihood keeps ln(y) of the original observation
        o2.flags = (yr_rep <= u.o.wr ? 1u : 0u) | (yr_rep == u.o.wr ? 2u : 0u);
        float nll_rep;
        mixing_line<false>(lf, o2, nll_rep, dummy);
        nll_sum += nll;
        nll_rep_sum += nll_rep;
        const size_t idx = (size_t)n * L + l;
        if (mode == 0) clppd[idx] = clppd[idx] * wgt_old + expf(-nll) * wgt_new;
        pvy[idx] = pvy[idx] * wgt_old + ((yr_rep <= u.o.yr) ? wgt_new : 0.f);
    }
    if (mode != 2) pvl[n] = pvl[n] * wgt_old + ((nll_rep_sum >= nll_sum) ? wgt_new : 0.f);
}

// ---- reductions -----------------------------------------------------------------------------
// out[0] = sum nll, out[1..D] spatial per dim (global convention, factor 1/2), out[1+D..2D] indicator per dim
__global__ void objective_terms_kernel(PriorDev pr, int N, int D, int max_degree, const float* __restrict__ theta,
                                       const int* __restrict__ nbr, const float* __restrict__ nll_lik,
                                       const uint8_t* __restrict__ owned, double* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double part[1 + 2 * BT_MAX_D];
#pragma unroll
    for (int q = 0; q < 1 + 2 * BT_MAX_D; ++q) part[q] = 0.0;
    if (n < N && (!owned || owned[n])) {
        float x[BT_MAX_D], had[BT_MAX_D], lap[BT_MAX_D];
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? theta[(size_t)n * D + d] : 0.f;
        part[0] = nll_lik[n];
        if (pr.has_spatial) {
            neighbour_sums(theta, nbr + (size_t)n * max_degree, max_degree, D, x, had, lap);
#pragma unroll
            for (int d = 0; d < BT_MAX_D; ++d)
                if (d < D) part[1 + d] = 0.5 * (double)pr.tau[d] * (double)had[d];
        }
        if (pr.has_indicator) {
#pragma unroll
            for (int d = 0; d < BT_MAX_D; ++d)
                if (d < D) {
                    float dist = 0.f;
                    if (x[d] > pr.hi[d]) dist = x[d] - pr.hi[d];
                    else if (x[d] < pr.lo[d]) dist = pr.lo[d] - x[d];
                    const double sc = (double)dist * (double)pr.inv_delta;
                    part[1 + D + d] = sc * sc * sc * sc;
                }
        }
    }
    for (int q = 0; q < 1 + 2 * D; ++q) {
        double val = part[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if ((threadIdx.x & 31) == 0 && val != 0.0) atomicAdd(out + q, val);
    }
}

__global__ void reduce_stats_kernel(int N, const float* __restrict__ accept, const float* __restrict__ logpa,
                                    const uint8_t* __restrict__ owned, double* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0.0, b = 0.0;
    if (n < N && (!owned || owned[n])) { a = accept[n]; b = logpa[n]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
}

__global__ void gather_rows_kernel(const float* __restrict__ theta, const int* __restrict__ pix, int n, int D, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * D) out[i] = theta[(size_t)pix[i / D] * D + (i % D)];
}
__global__ void scatter_rows_kernel(float* __restrict__ theta, const int* __restrict__ pix, int n, int D, const float* __restrict__ in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * D) theta[(size_t)pix[i / D] * D + (i % D)] = in[i];
}

// debug: the Philox draws the non-injected path would use
__global__ void debug_pmala_noise_kernel(int n_pix, const int* site_pix, int D, uint64_t seed, int64_t t, int site,
                                         const int64_t* gid, float* z_out, float* u_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int n = site_pix[i];
    const int64_t g = gid ? gid[n] : (int64_t)n;
    float z[BT_MAX_D];
    bt_normals(z, D, seed, t, BT_PURPOSE_PMALA_Z, site, g, 0u);
    for (int d = 0; d < D; ++d) z_out[(size_t)n * D + d] = z[d];
    u_out[n] = bt_uniform(seed, t, BT_PURPOSE_PMALA_U, site, g);
}
__global__ void debug_mtm_u_kernel(int n_pix, const int* site_pix, uint64_t seed, int64_t t, int site, const int64_t* gid,
                                   float* usel, float* uacc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int n = site_pix[i];
    const int64_t g = gid ? gid[n] : (int64_t)n;
    usel[n] = bt_uniform(seed, t, BT_PURPOSE_MTM_USEL, site, g);
    uacc[n] = bt_uniform(seed, t, BT_PURPOSE_MTM_UACC, site, g);
}
__global__ void debug_scatter_cand_kernel(int n_pix, const int* site_pix, int D, int K, const float* cand_rows, float* cand_full) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_pix * K * D) return;
    const int d = (int)(idx % D);
    const long long ik = idx / D;
    const int k = (int)(ik % K), i = (int)(ik / K);
    cand_full[((size_t)site_pix[i] * K + k) * D + d] = cand_rows[((size_t)i * K + k) * D + d];
}
