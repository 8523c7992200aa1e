/* gpb200 -- C ABI of the B200-native Gaussian-process regression engine (libgpb200.so).
 *
 * This is the drop-in boundary for the GpRegressor hot path of C-bowman/inference-tools.  The reference
 * has no FFI: its boundary is Python (inference/gp/regression.py, covariance.py, mean.py,
 * acquisition.py) calling numpy.linalg.cholesky / scipy.linalg.solve_triangular / BLAS.  Every entry
 * point below names the reference interface it replaces (paths relative to inference/gp/).  The
 * Python shim in inference_tools_b200/gp binds these with ctypes (INTEGRATION.md shows the binding a
 * reference maintainer would add).
 *
 * Conventions
 *   - all array arguments are caller-owned C-order float64 HOST buffers unless the name ends in _dev
 *     (then they are device pointers on the context's GPU);
 *   - return value: 0 = ok; < 0 = CUDA / argument error, text in gpb_last_error();
 *   - `info` out-parameters follow LAPACK dpotrf: 0 = positive definite, k > 0 = the leading minor of
 *     order k is not positive definite (the shim raises numpy.linalg.LinAlgError exactly where the
 *     reference would, regression.py:241, 537, 555);
 *   - hyper-parameter vectors use the reference layout theta = [mean params | cov params]
 *     (regression.py:150-155), cov params concatenated in component order (covariance.py:61, 692-697);
 *   - a context owns all device memory, is bound to one GPU and is not re-entrant; calls release the
 *     GIL when made through ctypes.  There is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef GPB200_H
#define GPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpb_ctx gpb_ctx;

/* covariance kinds: covariance.py:181 SquaredExponential, :282 RationalQuadratic, :108 WhiteNoise,
 * :608 HeteroscedasticNoise; sums of them = CompositeCovariance (:47-105). */
enum { GPB_COV_SE = 0, GPB_COV_RQ = 1, GPB_COV_WHITE = 2, GPB_COV_HETERO = 3 };
/* mean kinds: mean.py:31 ConstantMean, :54 LinearMean, :86 QuadraticMean */
enum { GPB_MEAN_CONST = 0, GPB_MEAN_LINEAR = 1, GPB_MEAN_QUADRATIC = 2 };
/* gpb_get selectors: attributes GpRegressor exposes after set_hyperparameters (regression.py:239-244) */
enum { GPB_GET_K_XX = 0, GPB_GET_L = 1, GPB_GET_ALPHA = 2, GPB_GET_MU = 3 };
/* gpb_expected_improvement modes: ExpectedImprovement.__call__ (acquisition.py:76-86), opt_func
 * (:88-97), opt_func_gradient (:99-125) */
enum { GPB_EI_VALUE = 0, GPB_EI_NEG_LOG = 1, GPB_EI_NEG_LOG_GRAD = 2 };
/* gpb_acquisition kinds: ExpectedImprovement (acquisition.py:44-140), UpperConfidenceBound (:143-192), MaxVariance
 * (:195-232).  The modes above read, for every kind: the acquisition value (__call__), opt_func (the minimiser's
 * objective: -ln EI, -UCB, -sigma^2), opt_func and its gradient (opt_func_gradient). */
enum { GPB_ACQ_EI = 0, GPB_ACQ_UCB = 1, GPB_ACQ_MAXVAR = 2 };

const char* gpb_last_error(void);
int gpb_device_count(int* count);
int64_t gpb_launch_count(void);           /* kernels launched by this library so far (process-wide) */
double gpb_gemm_flops(void);              /* algorithmic FP64 flops issued through the GEMM kernels so far */
double gpb_gemm_flops_int8(void);         /* ... of which on the INT8 tensor-core path (each costs 28 int8 products) */

/* Runtime options, process-wide (initial values come from the GPB200_* environment variables):
 *   "gemm_i8"        0 = every FP64 GEMM on the FP64 tensor pipe (DMMA), 1 = INT8 tensor-core digit-split GEMM where
 *                    it pays (k >= gemm_i8_min_k and >= 148 tiles; default), 2 = wherever it applies
 *   "gemm_i8_min_k"  shortest k extent sent to the INT8 path (512)
 *   "gemm_i8_pair"   INT8 kernel tiles: 0 = single CTA, 1 = CTA pairs with uniform shared-memory slots (default), 2 = CTA
 *                    pairs with 3 A slots + 2 half-size B slots (measured equal)
 *   "gemm_tile", "gemm_tma", "gemm_i8_debug"  kernel-selection / probe switches of the GEMM dispatcher
 *   "graphs"         CUDA-graph replay of launch sequences (1)
 *   "i8_fallback"    repeat a factorisation on DMMA when the INT8 path reports a non-PD pivot (1)
 *   "predict_block"  block width of the left-looking predict solve against cached digit planes (0 = auto, -1 = the
 *                    full-width recursion)
 *   "predict_diag"   its diagonal blocks: 1 = recursion with FP64 leaf inverses (default), 2 = one INT8 product with the
 *                    block's explicit inverse (9 % faster solve, sigma up to 4x above the FP64 floor on dense data),
 *                    0 = the latter only for well-conditioned fits (amp / min L_ii <~ 30)
 *   "gemm_i8_max_k"  longest k extent of one INT8 launch (16384 = the int32 exactness limit); longer extents are chunked
 *   "grad_inverse"   K^-1 of gpb_lml_grad: 0 = recursive triangular inverse W, then W^T W (default); 1 = rows of L^-T by blocked
 *                    substitution, then Y Y^T (equal in time and accuracy)
 *   "gemm_i8_epi2"   INT8 kernel, second-sweep epilogue in two passes (TMEM released before global memory is touched):
 *                    0 = off, 1 = on, 2 = for k extents >= 2048 with 8 epilogue warps (default)
 *   "gemm_i8_prefetch"  INT8 kernel: k-blocks by which an L2 prefetch of the digit planes runs ahead of the loads (0 = off)
 *   "gemm_i8_epi"    epilogue warps of the INT8 kernel (0 = by k extent, 8, 16)      "i8_grad_phases"  diagnostic mask
 *   "i8_grad_guard"  a-posteriori error estimate of the INT8 inverse chain in gpb_lml_grad, DMMA repeat when it is too
 *                    large relative to the gradient (1)
 * Changing an option drops every context's captured graphs at its next call. */
int gpb_set_option(const char* name, int64_t value);
int gpb_get_option(const char* name, int64_t* value);

int gpb_ctx_create(int device, gpb_ctx** out);
void gpb_ctx_destroy(gpb_ctx* ctx);

/* GpRegressor.__init__ data intake (regression.py:94-133): x is n x d, y has n entries, noise_var =
 * y_err**2 (the diagonal of `sig`, regression.py:320) or NULL (zeros, :322), y_cov = dense n x n `sig`
 * (:293) or NULL. */
int gpb_set_data(gpb_ctx* ctx, const double* x, int64_t n, int d, const double* y, const double* noise_var,
                 const double* y_cov);
/* kernel= / mean= arguments (regression.py:136-143): component kinds in sum order. */
int gpb_set_model(gpb_ctx* ctx, const int* cov_kinds, int ncomp, int mean_kind);
/* Same with an explicit parameter layout, needed for ChangePoint kernels (covariance.py:371-605): the leaves (plain
 * kernels) in evaluation order, theta_offs[c] = offset of leaf c's first parameter inside theta_cov, regions[c] = index
 * of the change-point region the leaf belongs to or -1; n_regions = 0 when there is no ChangePoint; the change-point
 * parameters (location_a, width_a), a = 0 .. n_regions-2, start at cp_theta_off (covariance.py:484-489). */
int gpb_set_model_ex(gpb_ctx* ctx, const int* cov_kinds, const int* theta_offs, const int* regions, int ncomp,
                     int n_regions, int cp_axis, int cp_theta_off, int n_cov_params_total, int mean_kind);
int gpb_num_hyperpars(gpb_ctx* ctx, int* n_mean, int* n_cov);

/* CovarianceFunction.build_covariance(theta) (covariance.py:247-255, 343-348, 163-169, 674-680, 91-95);
 * add_sig != 0 adds the data-error term as regression.py:239 does.  K_out is n x n. */
int gpb_build_covariance(gpb_ctx* ctx, const double* theta_cov, int add_sig, double* K_out);
/* CovarianceFunction.covariance_and_gradients(theta) (covariance.py:268-276, 350-365, 171-175, 682-686,
 * 97-105): K_out n x n, dK_out n_cov x n x n. */
int gpb_covariance_and_gradients(gpb_ctx* ctx, const double* theta_cov, double* K_out, double* dK_out);
/* CovarianceFunction.__call__(u, v, theta) (covariance.py:240-245, 335-341, 160-161, 671-672, 86-89):
 * out is m x n. */
int gpb_cross_covariance(gpb_ctx* ctx, const double* u, int64_t m, const double* v, int64_t n,
                         const double* theta_cov, double* out);

/* GpRegressor.set_hyperparameters (regression.py:218-244): K_xx = K(theta)+sig, L = chol(K_xx),
 * alpha = K_xx^-1 (y - mu).  State stays on the device until gpb_get. */
int gpb_factor(gpb_ctx* ctx, const double* theta, int* info);
int gpb_get(gpb_ctx* ctx, int which, double* out);
/* GpRegressor.marginal_likelihood (regression.py:528-542); on info > 0 the shim returns -1e50 */
int gpb_lml(gpb_ctx* ctx, const double* theta, double* lml, int* info);
/* GpRegressor.marginal_likelihood_gradient (regression.py:544-567): grad has n_mean + n_cov entries */
int gpb_lml_grad(gpb_ctx* ctx, const double* theta, double* lml, double* grad, int* info);
/* The same for n_theta hyper-parameter vectors at once (thetas is n_theta x p row-major, p = n_mean + n_cov), sharded over
 * nctx contexts that hold the same data and model on different GPUs: vector r runs on context r % nctx, each context on
 * its own host thread.  Serves the multistart restarts (regression.py:585-605, the reference forks a process pool) and
 * population-batched differential evolution (:569-574).  grad_or_null = NULL evaluates values only (gpb_lml). */
int gpb_lml_grad_batch(gpb_ctx** ctxs, int nctx, const double* thetas, int n_theta, int p, double* lml,
                       double* grad_or_null, int* info);
/* GpRegressor.loo_likelihood / loo_likelihood_gradient / loo_predictions (regression.py:451-526) */
int gpb_loo(gpb_ctx* ctx, const double* theta, double* loo, double* grad_or_null, int* info);
int gpb_loo_predictions(gpb_ctx* ctx, double* mu, double* sigma);

/* GpRegressor.__call__ (regression.py:188-216): q is m x d; mu, sig have m entries */
int gpb_predict(gpb_ctx* ctx, const double* q, int64_t m, double* mu, double* sig);
/* same with device-resident buffers (inputs already in HBM) */
int gpb_predict_dev(gpb_ctx* ctx, const double* q_dev, int64_t m, double* mu_dev, double* sig_dev);
/* GpRegressor.gradient (regression.py:351-385), SquaredExponential only: mean m x d, cov m x d x d */
int gpb_gradient(gpb_ctx* ctx, const double* q, int64_t m, double* mean, double* cov);
/* GpRegressor.spatial_derivatives (regression.py:387-419), SquaredExponential only: m x d each */
int gpb_spatial_derivatives(gpb_ctx* ctx, const double* q, int64_t m, double* dmu, double* dvar);
/* GpRegressor.build_posterior (regression.py:421-449): mu m, sigma m x m (NULL = mean_only) */
int gpb_posterior(gpb_ctx* ctx, const double* q, int64_t m, double* mu, double* sigma_or_null);
/* ExpectedImprovement over a batch of candidates (acquisition.py:76-125); grad (m x d) only for
 * GPB_EI_NEG_LOG_GRAD (SquaredExponential only); argmax_or_null receives the index of the best value */
int gpb_expected_improvement(gpb_ctx* ctx, const double* q, int64_t m, double y_max, int mode, double* out,
                             double* grad_or_null, int64_t* argmax_or_null);

/* Adds one training point to the FITTED model with the hyper-parameters kept (extension for GpOptimiser.add_evaluation,
 * optimisation.py:136-190, which re-builds the regressor from scratch): appends a row to the Cholesky factor with one
 * forward substitution (O(n^2)) and recomputes alpha.  x_new has d entries; noise_var_new is ignored when the model was
 * set up without y_err.  info > 0: the enlarged matrix is not positive definite and the old state is kept.  Not
 * available with y_cov, HeteroscedasticNoise or ChangePoint models. */
int gpb_append_point(gpb_ctx* ctx, const double* x_new, double y_new, double noise_var_new, int* info);

/* Any acquisition function over a batch of candidates: kind = GPB_ACQ_*, param = y_max (EI, acquisition.py:41) or
 * kappa (UCB, :162); argbest_or_null receives the index of the best candidate (largest value = smallest opt_func; lowest
 * index on ties), found by a device reduction. */
int gpb_acquisition(gpb_ctx* ctx, int kind, double param, const double* q, int64_t m, int mode, double* out,
                    double* grad_or_null, int64_t* argbest_or_null);

/* Distributed regressor (one process per GPU) for N beyond one GPU (BASELINE.json config 5): K(theta) + sig lives as
 * block columns of width `block`, column j on rank j mod world; the only exchanges on the data path are NCCL broadcasts
 * of column panels (and of solved nbd-vectors in the backward solve).  Rank 0 creates the 128-byte NCCL unique id and
 * shares it (torch.distributed / any channel); every rank then calls gpb_dist_init.  All gpb_dist_* calls below are
 * COLLECTIVE: every rank of the communicator must make them in the same order.
 *   gpb_dist_factor   set_hyperparameters (regression.py:218-244): assemble own columns, right-looking blocked Cholesky
 *                     with one-panel look-ahead; L stays sharded, v = L^-1 (y - mu) ends up on every rank
 *   gpb_dist_lml      marginal_likelihood (regression.py:528-542) = gpb_dist_factor + -1/2 v.v - sum log L_ii
 *   gpb_dist_lml_grad marginal_likelihood_gradient (regression.py:544-567) = gpb_dist_factor + alpha + the rows of K^-1 this
 *                     rank owns (identity rows solved against the streamed panels, then Y_a Y_b^T products with the
 *                     streamed row blocks) + this rank's share of the traces + one all-reduce of the n_mean + n_cov
 *                     gradient entries; grad on every rank; seconds_out4 = assemble, factor sweep, their sum, gradient part.
 *                     No ChangePoint kernels, no dense y_cov.
 *   gpb_dist_alpha    alpha = L^-T v (regression.py:242-244), block back-substitution; alpha_out (n doubles, host) on
 *                     every rank
 *   gpb_dist_predict  __call__ (regression.py:188-216): each rank passes ITS OWN slab of query points (m may differ,
 *                     may be 0); per chunk of queries the panels of L are streamed through all ranks */
int gpb_dist_unique_id(char* out128);
int gpb_dist_init(gpb_ctx* ctx, int rank, int world, const char* id128);
int gpb_dist_factor(gpb_ctx* ctx, const double* theta, int block, int* info, double* seconds_out3);
int gpb_dist_lml(gpb_ctx* ctx, const double* theta, int block, double* lml, int* info, double* seconds_out3);
int gpb_dist_lml_grad(gpb_ctx* ctx, const double* theta, int block, double* lml, double* grad, int* info, double* seconds_out4);
int gpb_dist_alpha(gpb_ctx* ctx, double* alpha_out);
int gpb_dist_predict(gpb_ctx* ctx, const double* q, int64_t m, double* mu, double* sig);
int gpb_dist_finalize(gpb_ctx* ctx);
/* layout of the sweep for one rank (host-only, no GPU): block columns, owned ones, doubles of panel / staging storage */
int gpb_dist_plan(int64_t n, int block, int world, int rank, int* n_blocks, int* n_owned, int64_t* panel_doubles,
                  int64_t* staging_doubles, int* owners_or_null);

/* GpLinearInverter (reference inference/gp/inversion.py:11-249): posterior of y = A x + noise with a GP prior on x.
 * The context's training inputs are the parameter_spatial_positions (gpb_set_data with y = zeros, no noise) and its
 * model the prior covariance/mean (gpb_set_model*); theta = [mean, cov] as in inversion.py:126-133.
 *   gpb_linv_set_problem   inversion.py:85-118 (A is m x n row-major, y and y_err length m)
 *   gpb_linv_lml           inversion.py:170-187  marginal_likelihood
 *   gpb_linv_lml_grad      inversion.py:189-217  marginal_likelihood_gradient
 *   gpb_linv_posterior     inversion.py:138-168  calculate_posterior / calculate_posterior_mean (cov may be NULL)
 * info = 0, or the 1-based index of the first non-positive pivot of chol(A K A^T + Sigma). */
int gpb_linv_set_problem(gpb_ctx* ctx, const double* A, int64_t m, const double* y, const double* y_err);
int gpb_linv_lml(gpb_ctx* ctx, const double* theta, double* lml, int* info);
int gpb_linv_lml_grad(gpb_ctx* ctx, const double* theta, double* lml, double* grad, int* info);
int gpb_linv_posterior(gpb_ctx* ctx, const double* theta, double* mean, double* cov_or_null, int* info);

/* Counters of a context: "dmma_retries" (factorisations repeated on DMMA after the INT8 path reported a non-PD pivot),
 * "grad_guard_retries" (gradient evaluations whose inverse chain was repeated on DMMA by the error guard),
 * "grad_guard_est" (the guard's last estimate of the INT8 gradient error relative to max|grad|), "predict_block"
 * (block width of the cached-plane predict solve, 0 = recursion). */
int gpb_ctx_stat(gpb_ctx* ctx, const char* name, double* out);

/* CUDA-event phase timings (milliseconds) of the most recent call on this context:
 * names is a ';'-separated list written into name_buf, ms[i] the matching durations. */
int gpb_timers(gpb_ctx* ctx, char* name_buf, int name_buf_len, double* ms, int max_entries, int* n_entries);

/* Device memory helpers for callers that keep inputs resident in HBM (bench.py `value` leg). */
int gpb_dev_alloc(gpb_ctx* ctx, int64_t n_doubles, double** out_dev);
int gpb_dev_free(gpb_ctx* ctx, double* p_dev);
int gpb_dev_upload(gpb_ctx* ctx, double* dst_dev, const double* src_host, int64_t n_doubles);
int gpb_dev_download(gpb_ctx* ctx, double* dst_host, const double* src_dev, int64_t n_doubles);
int gpb_sync(gpb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GPB200_H */
