// Optimisation mode of the sampler (MySamplerParams.is_stochastic = False):
//   PMALA branch  sampler/my_sampler.py:908-964   -- whole-map preconditioned gradient step, kept only if the GLOBAL objective
//                                                     decreases; a noisy retry otherwise; v follows the last evaluated candidate
//   MTM branch    sampler/my_sampler.py:1050-1086 -- one challenger index for the whole colour class (argmin over k of the
//                                                     class-wide sum of candidate objectives), kept per pixel where it improves
// The global comparisons are reductions (one all-reduce when the map is sharded); everything per pixel stays on the device.
#pragma once
#include "sampler_kernels.cuh"

// theta <- theta_snap - eps/2 G grad(theta_snap) [+ sqrt(eps G) z],  G = 1 / (eta + sqrt(v))          (:913-930, :940)
__global__ void optim_pmala_propose_kernel(PriorDev pr, int N, int D, int max_degree, const float* __restrict__ theta_snap,
                                           const int* __restrict__ nbr, const float* __restrict__ grad_lik_snap,
                                           const float* __restrict__ v, float eps, float eta, int with_noise, uint64_t seed,
                                           int64_t t, const int64_t* __restrict__ gid, const float* __restrict__ z_inject,
                                           float* __restrict__ theta_out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float x[BT_MAX_D], gp[BT_MAX_D], z[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) { x[d] = (d < D) ? theta_snap[(size_t)n * D + d] : 0.f; z[d] = 0.f; }
    prior_eval(pr, theta_snap, nbr + (size_t)n * max_degree, max_degree, D, x, gp);
    if (with_noise) {
        if (z_inject) {
#pragma unroll
            for (int d = 0; d < BT_MAX_D; ++d) z[d] = (d < D) ? z_inject[(size_t)n * D + d] : 0.f;
        } else {
            bt_normals(z, D, seed, t, BT_PURPOSE_PMALA_Z, 0, gid ? gid[n] : (int64_t)n, 0u);
        }
    }
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            const float g = nan_to_num(grad_lik_snap[(size_t)n * D + d] + gp[d]);     // posterior.py:240
            const float G = 1.0f / (eta + sqrtf(v[(size_t)n * D + d]));
            theta_out[(size_t)n * D + d] = x[d] - 0.5f * eps * G * g + z[d] * sqrtf(eps * G);
        }
}

// v <- alpha v + (1 - alpha) grad(theta)^2 at the resident (candidate) state                               (:952-953)
__global__ void optim_update_v_kernel(PriorDev pr, int N, int D, int max_degree, const float* __restrict__ theta,
                                      const int* __restrict__ nbr, const float* __restrict__ grad_lik, float alpha,
                                      float* __restrict__ v) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float x[BT_MAX_D], gp[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? theta[(size_t)n * D + d] : 0.f;
    prior_eval(pr, theta, nbr + (size_t)n * max_degree, max_degree, D, x, gp);
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d)
        if (d < D) {
            const float g = nan_to_num(grad_lik[(size_t)n * D + d] + gp[d]);
            v[(size_t)n * D + d] = alpha * v[(size_t)n * D + d] + (1.f - alpha) * g * g;
        }
}

// negative log conditional posterior of every candidate row: priors (posterior.py:59-111) + likelihood
// (abstract_likelihood.py:73-111), no proposal density (my_sampler.py:1044-1046).  One thread per candidate row.
__global__ void mtm_optim_nl_kernel(PriorDev pr, FmapDev fm, int n_pix, const int* __restrict__ site_pix, int D, int L, int K,
                                    int max_degree, const float* __restrict__ theta, const int* __restrict__ nbr,
                                    const ObsConst* __restrict__ obs, const float* __restrict__ cand_rows,
                                    const float* __restrict__ lf_rows, float* __restrict__ nl) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)n_pix * (K + 1)) return;
    const int i = (int)(row / (K + 1));
    const int n = site_pix[i];
    const size_t r = mtm_row(i, (int)(row - (long long)i * (K + 1)), K, n_pix);
    float x[BT_MAX_D], dummy[BT_MAX_D];
#pragma unroll
    for (int d = 0; d < BT_MAX_D; ++d) x[d] = (d < D) ? cand_rows[r * D + d] : 0.f;
    float val = lik_row<false>(fm, L, lf_rows + r * L, nullptr, obs + (size_t)n * L, dummy);
    if (pr.has_spatial) {
        float had[BT_MAX_D], lap[BT_MAX_D];
        neighbour_sums(theta, nbr + (size_t)n * max_degree, max_degree, D, x, had, lap);
        float sp = 0.f;
#pragma unroll
        for (int d = 0; d < BT_MAX_D; ++d)
            if (d < D) sp = fmaf(pr.tau[d], had[d], sp);
        val += sp;                                                   // chromatic factor 1 (l22:103-107)
    }
    if (pr.has_indicator) val += indicator_terms(pr, D, x, dummy);
    nl[row] = val;
}

// colsum[k] = sum over the pixels of the class (flagged in owned, or all) of nl[i, k], k < K                   (:1051-1053)
// one block per k, fixed-order tree reduction in float64 (deterministic)
__global__ void mtm_optim_colsum_kernel(int n_pix, const int* __restrict__ site_pix, int K, const float* __restrict__ nl,
                                        const uint8_t* __restrict__ owned, double* __restrict__ colsum) {
    __shared__ double red[256];
    const int k = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_pix; i += blockDim.x)
        if (!owned || owned[site_pix[i]]) acc += (double)nl[(size_t)i * (K + 1) + k];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) colsum[k] = red[0];
}

// keep the challenger where it improves on the current value and both are finite                               (:1073-1086)
__global__ void mtm_optim_select_kernel(int n_pix, const int* __restrict__ site_pix, int D, int K, int k_star,
                                        const float* __restrict__ cand_rows, const float* __restrict__ nl,
                                        float* __restrict__ theta, float* __restrict__ accept,
                                        uint8_t* __restrict__ accepted_any) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int n = site_pix[i];
    const float nc = nl[(size_t)i * (K + 1) + k_star], n0 = nl[(size_t)i * (K + 1) + K];
    const bool acc = (nc < n0) && isfinite(nc) && isfinite(n0);
    accept[n] = acc ? 1.f : 0.f;
    if (acc) {
        accepted_any[n] = 1;
        for (int d = 0; d < D; ++d) theta[(size_t)n * D + d] = cand_rows[((size_t)i * K + k_star) * D + d];
    }
}
