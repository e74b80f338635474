/*
 * beetroots_b200 -- C ABI of the B200-native sampler step (PMALA + MTM) of beetroots.
 *
 * The reference is pure Python, so "the FFI a maintainer would bind" is ctypes
 * (INTEGRATION.md shows the stub).  Every entry point below replaces a method of the reference
 * on the per-iteration hot path; the reference interface it stands for is cited as file:line
 * relative to /root/reference/beetroots/.
 *
 * Conventions
 *   - plain pointers and sizes only; "host" pointers are ordinary process memory (double /
 *     int32, C-contiguous, the dtypes the reference uses), "device" pointers are CUDA device
 *     addresses on the context's GPU;
 *   - every function returns 0 on success, a negative code otherwise; bt_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - one context = one GPU = one (band of a) map; all work is enqueued on the context's
 *     stream (bt_set_stream); functions taking or returning HOST data synchronise that stream
 *     before returning, functions working on device-resident state do not;
 *   - no hidden global state besides the per-thread error string.
 */
#ifndef BEETROOTS_B200_H
#define BEETROOTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bt_ctx bt_ctx;

#define BT_MAX_D 5        /* kappa, P, radm, Avmax, angle */
#define BT_MAX_L 64
#define BT_MAX_DEGREE 8

/* precision of the forward-map kernel */
#define BT_MLP_FP32 0      /* CUDA-core fp32: the parity path (rtol 1e-5 vs fp64 reference) */
#define BT_MLP_BF16_TC 1   /* tcgen05 bf16 operands, fp32 accumulation in TMEM */

const char* bt_last_error(void);
int bt_version(void);

/* N pixels held by this context (owned + ghost), D sampled parameters, L lines. */
int bt_ctx_create(bt_ctx** out, int device, int N, int D, int L, int max_degree);
int bt_ctx_destroy(bt_ctx* ctx);
/* cudaStream_t as an integer (0 = legacy default stream). */
int bt_set_stream(bt_ctx* ctx, uint64_t cuda_stream);
int bt_set_mlp_mode(bt_ctx* ctx, int mode);
int bt_synchronize(bt_ctx* ctx);

/* ---- model upload (once per run) -------------------------------------------------------- */

/* nnbma PolynomialNetwork -> DenselyConnected weights; replaces NeuralNetwork.load +
 * restrict_to_output_subset, modelling/forward_maps/neural_network_approx.py:38,87-105.
 * exponents: (n_feat0, order) int32 (-1 padded); weights[l]: (block_new[l], in_l) row-major
 * with in_l = n_feat0 + sum_{l'<l} block_new[l'];  w_out: (n_out, n_feat_total). */
int bt_upload_network(bt_ctx* ctx, int n_in, int order, int n_feat0, const int32_t* exponents,
                      const double* poly_means, const double* poly_stds, int n_blocks,
                      const int32_t* block_new, const double* const* weights,
                      const double* const* biases, const double* w_out, const double* b_out,
                      int n_out);

/* Fixed / sampled entries of theta; replaces ForwardMap.set_sampled_and_fixed_entries,
 * modelling/forward_maps/abstract_base.py:44-73.  d_full = n_in + 1 (kappa first). */
int bt_upload_forward_map(bt_ctx* ctx, int d_full, const double* arr_fixed_values,
                          int d_sampling, const int32_t* list_indices_to_sample);

/* (N, L) host arrays; replaces the attributes read by MixingModelsLikelihood
 * (modelling/likelihoods/approx_censored_add_mult.py:122-266). */
int bt_upload_observations(bt_ctx* ctx, const double* y, const double* sigma_a,
                           const double* sigma_m, const double* omega, const double* log_fm1,
                           const double* log_fp1);

/* Spatial prior weights (tau = weights, tau_init = initial_weights) and smooth-indicator box;
 * modelling/priors/abstract_spatial_prior.py:100-121, smooth_indicator_prior.py:133-166. */
int bt_upload_priors(bt_ctx* ctx, int has_spatial, const double* tau, const double* tau_init,
                     int has_indicator, const double* lower, const double* upper, double delta);

/* Pixel graph: neighbour table (N, max_degree) int32, -1 padded, neighbour order = order in
 * SpatialPrior.list_edges (abstract_spatial_prior.py:13-55); colour classes = dict_sites
 * (:125-153) as concatenated sorted pixel lists; global_id (N) or NULL (= arange) keys the
 * counter-based RNG so that chains do not depend on how the map is sharded. */
int bt_upload_graph(bt_ctx* ctx, const int32_t* nbr, int n_sites, const int32_t* site_sizes,
                    const int32_t* site_pixels, const int64_t* global_id);

/* ---- state ------------------------------------------------------------------------------ */

/* Theta (N, D) scaled space; v (N, D) RMSProp accumulator; j (N, D) counters.  NULL = keep. */
int bt_set_state(bt_ctx* ctx, const double* theta, const double* v, const int32_t* j);
int bt_get_state(bt_ctx* ctx, double* theta, double* v, int32_t* j);
/* float32 variants for the end-to-end bench / halo plumbing (host pointers). */
int bt_set_theta_f32(bt_ctx* ctx, const float* theta);
int bt_get_theta_f32(bt_ctx* ctx, float* theta);
/* device-side gather / scatter of theta rows (halo exchange buffers live in torch tensors). */
int bt_gather_theta_rows(bt_ctx* ctx, const int32_t* pix_dev, int n, float* out_dev);
int bt_scatter_theta_rows(bt_ctx* ctx, const int32_t* pix_dev, int n, const float* in_dev);

/* Posterior.compute_all (modelling/posterior.py:332-446) at the resident Theta: forward map +
 * Jacobian, likelihood value and gradient, cached per pixel. */
int bt_compute_all(bt_ctx* ctx);
/* v <- grad^2, j <- 0  (sampler/my_sampler.py:337,354). */
int bt_init_v_from_grad(bt_ctx* ctx);
/* Any pointer may be NULL.  log_f (N, L), nll_pix (N), grad_lik (N, D), obj_chromatic (N),
 * obj_global (N), grad (N, D): the entries of the ``current`` dict (posterior.py:421-446). */
int bt_get_current(bt_ctx* ctx, double* log_f, double* nll_pix, double* grad_lik,
                   double* obj_chromatic, double* obj_global, double* grad);
/* dict_objective of Posterior.compute_all_for_saver (posterior.py:270-330):
 * out[0] = nll, out[1..D] = nl_prior_spatial per dim, out[1+D..2D] = nl_prior_indicator per dim,
 * out[1+2D] = objective.  Sums run over pixels flagged in owned (N uint8) or all if NULL. */
int bt_objective_terms(bt_ctx* ctx, const uint8_t* owned, double* out);

/* SpatialPrior.neglog_pdf(Theta, idx_pix = all pixels, with_weights=False)
 * (modelling/priors/l22_discrete_grad_prior.py:95-128): out[d] = 1/2 sum_n sum_{i in V_n} (Theta_nd - Theta_id)^2, d < D --
 * the potential g(Theta) of the regularisation-weight update, Sampler.sample_regu_hyperparams
 * (sampler/abstract_sampler.py:172-210).  Pixels flagged in owned (N uint8), or all if NULL. */
int bt_spatial_prior_unweighted(bt_ctx* ctx, const uint8_t* owned, double* out);

/* ForwardMap.compute_all on arbitrary rows (neural_network_approx.py:152-305): theta (M, D)
 * host -> log_f (M, L), grad_log_f (M, D, L) or NULL. */
int bt_forward_map(bt_ctx* ctx, int M, const double* theta, double* log_f, double* grad_log_f);

/* ---- transition kernels, one colour class (site) at a time -------------------------------- */

/* MySampler.generate_new_sample_pmala_rmsprop, sampling branch, body of the site loop
 * (sampler/my_sampler.py:752-903).  z_inject (N, D) standard normals / u_inject (N) uniforms
 * are HOST float arrays indexed by local pixel, or NULL to draw from Philox(seed, t). */
int bt_pmala_site(bt_ctx* ctx, int site, double eps, double eta, double alpha, uint64_t seed,
                  int64_t t, const float* z_inject, const float* u_inject);

/* MySampler.generate_new_sample_mtm, sampling branch, body of the site loop
 * (sampler/my_sampler.py:1006-1178) with proposals of sampler/utils/utils.py:274-349 and the
 * proposal density of :419-495.  cand_inject (N, K, D), usel_inject (N), uacc_inject (N) host
 * float arrays or NULL (Philox). */
int bt_mtm_site(bt_ctx* ctx, int site, int K, uint64_t seed, int64_t t, const float* cand_inject,
                const float* usel_inject, const float* uacc_inject);
/* End of generate_new_sample_mtm (:1182-1197): refresh likelihood terms and update v on the
 * pixels accepted during this iteration.  Call after all sites (and after the halo exchange). */
int bt_mtm_finish(bt_ctx* ctx, double alpha);
/* Clear the per-iteration accept / log-ratio maps (start of either kernel, :736-737, :989-990). */
int bt_begin_iteration(bt_ctx* ctx);

/* accept (N) in {0,1}, log_proba (N): accept_total / log_proba_accept_total (:882-883,
 * :1170-1171).  Host pointers, may be NULL. */
int bt_get_step_stats(bt_ctx* ctx, double* accept, double* log_proba);
/* out[0] = sum accept, out[1] = sum log_proba over pixels flagged in owned (or all). */
int bt_reduce_step_stats(bt_ctx* ctx, const uint8_t* owned, double* out);

/* Fill out (n) with the standard normals / uniforms the Philox path would use for
 * (purpose, site, t): lets tests replay the device noise through the oracle. */
int bt_debug_philox_pmala(bt_ctx* ctx, int site, uint64_t seed, int64_t t, float* z_out,
                          float* u_out);
int bt_debug_philox_mtm(bt_ctx* ctx, int site, int K, uint64_t seed, int64_t t, float* cand_out,
                        float* usel_out, float* uacc_out);

/* ---- online model checking (sampler/my_sampler.py:158-267, sampling branch) ---------------------
 * One update per saved iterate: replicated observation y_rep = max(omega, eps_m f + eps_a)
 * (likelihoods/approx_censored_add_mult.py:268-280), running means of exp(-nll) (clppd), of
 * [y_rep <= y] and of [nll(y_rep) >= nll(y)] per pixel.  za / zm: HOST float (N, L) standard normals for
 * the additive / multiplicative noise, or NULL (Philox).  bt_model_check_get with finalize != 0 folds the
 * p-values-y onto [0, 0.5] (my_sampler.py:260-265).  Any output pointer may be NULL. */
int bt_model_check_reset(bt_ctx* ctx);
int bt_model_check_update(bt_ctx* ctx, uint64_t seed, int64_t t, const float* za_inject, const float* zm_inject);
int bt_model_check_get(bt_ctx* ctx, double* clppd, double* pval_y, double* pval_llh, int finalize);

/* optimisation mode of the model check (sampler/my_sampler.py:208-212, 222-256): clppd <- exp(-nll) of the resident
 * iterate (call when the objective improved); p-values only (ESS_OPTIM replications at the final iterate). */
int bt_model_check_snapshot_clppd(bt_ctx* ctx);
int bt_model_check_update_pvalues(bt_ctx* ctx, uint64_t seed, int64_t t, const float* za_inject, const float* zm_inject);

/* ---- optimisation mode (MySamplerParams.is_stochastic = False) -----------------------------------------------
 * PMALA branch, sampler/my_sampler.py:908-964 -- a whole-map candidate is evaluated in place and kept only if the
 * global objective decreases:
 *   bt_optim_begin            snapshot Theta and the cached likelihood terms
 *   bt_optim_pmala_propose    Theta <- Theta_snap - eps/2 G grad(Theta_snap) [+ sqrt(eps G) z if with_noise]  (:913-940);
 *                             z_inject (N, D) HOST float or NULL (Philox).  Then: (halo exchange,) bt_compute_all,
 *                             bt_objective_terms -> compare on the host (all-reduce when sharded)
 *   bt_optim_update_v         v <- alpha v + (1 - alpha) grad(Theta)^2 at the last evaluated candidate (:952-953)
 *   bt_optim_end              accept != 0 keeps the candidate, 0 restores the snapshot
 * MTM branch, sampler/my_sampler.py:1006-1086, one colour class:
 *   bt_mtm_optim_weights      candidates + their negative log conditional posterior (no proposal density);
 *                             colsum_out[k] (K, HOST) = sum over the class's pixels (flagged in owned, or all) of nl[i, k]
 *                             -> the caller takes the argmin (after an all-reduce when sharded) (:1051-1053)
 *   bt_mtm_optim_select       keep candidate k_star where it improves on the current value (:1073-1086)
 * followed by bt_mtm_finish after all classes, as in sampling mode (:1182-1197). */
int bt_optim_begin(bt_ctx* ctx);
int bt_optim_pmala_propose(bt_ctx* ctx, double eps, double eta, int with_noise, uint64_t seed, int64_t t,
                           const float* z_inject);
int bt_optim_update_v(bt_ctx* ctx, double alpha);
int bt_optim_end(bt_ctx* ctx, int accept);
int bt_mtm_optim_weights(bt_ctx* ctx, int site, int K, uint64_t seed, int64_t t, const float* cand_inject,
                         const uint8_t* owned, double* colsum_out);
int bt_mtm_optim_select(bt_ctx* ctx, int site, int K, int k_star);

/* ---- on-device chain store and posterior summaries ------------------------------------------------------
 * Replaces the round trip saver memory -> HDF5 `list_Theta` (sampler/saver/my_saver.py:46-83, abstract_saver.py:190-213)
 * -> per-pixel re-read + np.mean / np.percentile (inversion/results/utils/mc.py:158-215): the (thinned) chain stays in
 * HBM as [T][N][D] float32 and is summarised there.
 *   bt_chain_reserve   allocate room for `capacity` iterates (drops a previous chain; 0 frees it)
 *   bt_chain_push      append the resident Theta (device-to-device, asynchronous on the context's stream)
 *   bt_chain_get       iterates [first, first + count) as HOST double (count, N, D) -- what update_memory stores
 *   bt_chain_stats     over iterates [first, first + count), per (pixel, parameter): mean (N, D) and population variance
 *                      (N, D) in scaled space (either may be NULL), and for each of the n_q <= 16 percentiles p (in
 *                      [0, 100]) the two order statistics that bracket numpy's 'linear' virtual index p/100 (count-1):
 *                      q_lo, q_hi (n_q, N, D) and the interpolation weight q_frac (n_q), so that the caller can
 *                      interpolate after mapping to linear space exactly as the reference does
 *                      (percentile = lo + frac (hi - lo)). */
int bt_chain_reserve(bt_ctx* ctx, int64_t capacity);
int bt_chain_push(bt_ctx* ctx);
int64_t bt_chain_length(bt_ctx* ctx);
int bt_chain_get(bt_ctx* ctx, int64_t first, int64_t count, double* out);
int bt_chain_stats(bt_ctx* ctx, int64_t first, int64_t count, int n_q, const double* percentiles, double* mean,
                   double* var, double* q_lo, double* q_hi, double* q_frac);

/* which tensor-core forward-map kernel BT_MLP_BF16_TC runs for the uploaded network: 4 = pair kernel fed by the pre-pass
 * kernel for the leading small blocks (default), 3 = pair kernel alone,
 * 1 = single-CTA kernel (fallback for shapes the pair kernel does not take), 2 = first cta_group::2 kernel, 0 = none. */
int bt_mlp_kernel_variant(bt_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t bt_kernel_launches(bt_ctx* ctx);
/* device time (ms) spent in the forward-map kernel since the last call with reset != 0 and the
 * number of rows it processed; measured with CUDA events on the context's stream. */
int bt_mlp_profile(bt_ctx* ctx, int enable, int reset, double* ms, int64_t* rows, int64_t* launches);

/* device time and ALGORITHMIC HBM bytes of the memory-bound kernels of the step (the roofline entries of bench.py next to the
 * forward map's): arrays of BT_KP_COUNT entries, accumulated since the last call with reset != 0.  Bytes per launch are
 * counted from the launch shape (per colour pixel / per candidate row figures), not measured. */
#define BT_KP_PMALA_PROPOSE 0  /* my_sampler.py:752-838: stencil + indicator gradient + preconditioned proposal */
#define BT_KP_PMALA_ACCEPT 1   /* my_sampler.py:840-903 + likelihood of the candidate (approx_censored_add_mult.py:982-1246) */
#define BT_KP_MTM_GEN 2        /* utils.py:274-349 */
#define BT_KP_MTM_NL 3         /* likelihood + priors + proposal density of every candidate row (my_sampler.py:1017-1105) */
#define BT_KP_MTM_SELECT 4     /* my_sampler.py:1107-1178 */
#define BT_KP_MTM_FINISH 5     /* my_sampler.py:1182-1197 on the accepted pixels */
#define BT_KP_COUNT 6
int bt_kernel_profile(bt_ctx* ctx, int enable, int reset, double* ms, double* bytes, int64_t* launches);
/* allocate every scratch buffer a step over colour classes of up to max_rows candidate rows needs (candidate rows, their
 * ln f / Jacobian, the pre-pass hand-over image of the tensor-core forward map), so that no allocation happens inside a step. */
int bt_reserve_scratch(bt_ctx* ctx, int64_t max_rows, int64_t max_rows_with_jacobian);

#ifdef __cplusplus
}
#endif
#endif /* BEETROOTS_B200_H */
