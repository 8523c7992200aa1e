// Launch wrappers of the non-GEMM kernels (kernels.cu, potrf.cu, solve.cu, lml.cu).
#pragma once
#include "common.cuh"

namespace gpb {

// ---- kernels.cu : covariance assembly, cross-covariance, row reductions, acquisition
// K(theta)+diag terms on the padded training grid, lower 128x128 tiles (mirror=1 also fills the upper
// triangle).  Padded rows/cols (>= n) hold the identity.
int launch_assemble_train(const CovParams& cp, const double* x, int n, int npad, const double* noise_var,
                          const double* y_cov, double* K, int64_t ld, int mirror, cudaStream_t s);
// rectangular block of the same matrix into a panel buffer (rows/cols multiples of 128, within npad)
int launch_assemble_block(const CovParams& cp, const double* x, int n, const double* noise_var, int row0, int nrows,
                          int col0, int ncols, double* out, int64_t ld, cudaStream_t s);
// dK/dtheta_p planes for the covariance_and_gradients API (small N; dense output, np x n x n)
int launch_assemble_grads(const CovParams& cp, const double* x, int n, double* K, double* dK, cudaStream_t s);
// generic cross covariance cov(u, v) (covariance.py:240-245, 335-341): out is m x n row-major
int launch_cross_cov(const CovParams& cp, const double* u, int m, const double* v, int n, double* out, int64_t ld,
                     cudaStream_t s);
// stacked cross-covariance rows for a chunk of queries: row q*nstack+0 = k(q, x_j); rows a=1..d (SE only)
// = ((x_ja - q_a)/l_a^2) k(q, x_j)  (covariance.py:257-266, regression.py:376, 414).  Columns >= n are 0.
int launch_cross_stack(const CovParams& cp, const double* q, int mq, int nstack, const double* x, int n, int npad,
                       double* S, int64_t ld, cudaStream_t s);
// out[r] = sum_j S[r][j] * vec[j]  (one warp per row, fixed-order reduction)
int launch_row_dot(const double* S, int64_t ld, int rows, int ncols, const double* vec, double* out, cudaStream_t s);
// out[c] = sum_r S[r][c] * vec[r]  (ncols % 128 == 0; ws holds col_dot_ws_bytes(nrows, ncols))
size_t col_dot_ws_bytes(int nrows, int ncols);
int launch_col_dot(const double* S, int64_t ld, int nrows, int ncols, const double* vec, double* out, double* ws,
                   cudaStream_t s);
// G[q][a][b] = sum_j X[q*ns+a][j] X[q*ns+b][j]
int launch_row_gram(const double* X, int64_t ld, int mq, int nstack, int ncols, double* G, cudaStream_t s);
// predictive mean/sigma (regression.py:212-216): mu = dot + mean(q); sig = sqrt|k(q,q) - G|
int launch_finalize_predict(const CovParams& cp, const MeanParams& mp, const double* q, int mq, int nstack,
                            const double* dots, const double* G, double* mu, double* sig, cudaStream_t s);
// gradient() outputs (regression.py:379-380): mean[q][a] = dots[q*ns+1+a]; cov[q][a][b] = R[b] - G[q][a+1][b+1]
int launch_finalize_gradient(const double* dots, const double* G, int mq, int d, const double* R_dev, double* mean,
                             double* cov, cudaStream_t s);
// spatial_derivatives() outputs (regression.py:413-414): dmu = dots rows a>=1; dvar[q][a] = -2 G[q][0][a+1]
int launch_finalize_spatial(const double* dots, const double* G, int mq, int d, double* dmu, double* dvar,
                            cudaStream_t s);
// Acquisition functions (acquisition.py:76-125 EI, :169-189 UCB, :213-229 MaxVariance): kind 0 EI (param = y_max),
// 1 UCB (param = kappa), 2 MaxVariance.  mode 0: value, 1: opt_func, 2: opt_func and its gradient
int launch_acquisition(const double* mu, const double* sig, const double* dmu, const double* dvar, int m, int d, int kind,
                       double param, int mode, double* out, double* grad, cudaStream_t s);
// index of the largest / smallest entry (lowest index on ties, NaNs ignored, -1 when there is none): result in
// ws_val[0] / ws_idx[0]; both workspaces hold 1 + ARGBEST_BLOCKS entries
constexpr int ARGBEST_BLOCKS = 592;
int launch_argbest(const double* v, int64_t m, int want_max, double* ws_val, int64_t* ws_idx, cudaStream_t s);
// resid = y - mean(x) (regression.py:243, 538); mu_out optional
int launch_residual(const MeanParams& mp, const double* x, const double* y, int n, int npad, double* resid,
                    double* mu_out, cudaStream_t s);
// small utility kernels
int launch_copy2d(const double* src, int64_t lds, double* dst, int64_t ldd, int rows, int cols, cudaStream_t s);

// ---- potrf.cu : blocked recursive Cholesky / triangular inverse / triangular solve drivers
struct LinalgWs {
    double* dinv;      // npad/NB blocks of NB x NB: inverses of the diagonal blocks of L
    double* tmp;       // scratch: rows_max x NB panel for the out-of-place leaf solves
    int64_t tmp_rows;  // capacity of tmp in rows
    int* info;         // device: 0 or 1-based index of the first non-positive pivot
};
int potrf_lower(double* A, int64_t ld, int n, const LinalgWs& ws, cudaStream_t s);
// X <- X * L^-T  (X: m x n row-major, L: n x n lower with inverted diagonal blocks in ws.dinv + blk0)
int trsm_right_lt(double* X, int64_t ldx, int m, const double* L, int64_t ldl, int n, int blk0, const LinalgWs& ws,
                  cudaStream_t s);
// W = L^-1 (lower, row-major; the strict upper triangle of W must be zero on entry); scratch n/2 x n/2
int trtri_lower(const double* L, int64_t ldl, double* W, int64_t ldw, int n, int blk0, const LinalgWs& ws,
                double* scratch, int64_t lds, cudaStream_t s);
// Kinv(lower tiles) = W^T W
int lauum_lower(const double* W, int64_t ldw, double* Kinv, int64_t ldk, int n, cudaStream_t s);
// the same inverse by substitution: Y = L^-T (upper, row-major; n x n overwritten), then Kinv(lower tiles) = Y Y^T
int trtri_rows_lower(const double* L, int64_t ldl, double* Y, int64_t ldy, int n, const LinalgWs& ws, cudaStream_t s);
int lauum_rows_lower(const double* Y, int64_t ldy, double* Kinv, int64_t ldk, int n, cudaStream_t s);

// ---- solve.cu : vector solves and reductions
// v = L^-1 r (fwd) ; a = L^-T v (bwd).  vec holds 2*npad doubles: [0,npad) = right-hand side (destroyed),
// [npad, 2*npad) = solution.
int trsv_lower_fwd(const double* L, int64_t ld, int npad, const double* dinv, double* vec, cudaStream_t s);
int trsv_lower_bwd(const double* L, int64_t ld, int npad, const double* dinv, double* vec, cudaStream_t s);
// out[0] = sum_i log L_ii (i < n); out[1] = a.b (i < n); out[2] = min_i L_ii
int launch_logdet_dot(const double* L, int64_t ld, const double* a, const double* b, int n, double* out2,
                      cudaStream_t s);

// ---- lml.cu : fused gradient traces (regression.py:563-566)
// grad layout: [mean params | cov params]; partial buffer sized by trace_partials_size().  fro2_dev (MAX_COMP doubles,
// optional): per smooth component sum_ij max_p dK_p,ij^2, an upper bound on every |dK_p|_F^2 (gradient error guard)
size_t trace_partials_size(int npad);
int launch_lml_grad(const CovParams& cp, const MeanParams& mp, int n_theta_mean, const double* x, int n, int npad, const double* alpha, const double* Kinv, int64_t ld,
                    double* partials, double* grad_dev, double* fro2_dev, cudaStream_t s);
// distributed layout (dist.cu): row blocks me + G idx (idx < na, width nbd) of Kinv in a row stack + diagonal blocks;
// every rank writes its share into grad_dev (zero on entry), the sum over ranks is the gradient
size_t trace_partials_size_stacked(int na, int nbd, int npad);
int launch_lml_grad_stacked(const CovParams& cp, const MeanParams& mp, int n_theta_mean, const double* x, int n, int npad,
                            const double* alpha, const double* Kst, int64_t ldst, const double* Kdiag, int nbd, int me, int G,
                            int na, bool with_mean, double* partials, double* grad_dev, cudaStream_t s);

// ---- loo.cu : leave-one-out objective (regression.py:451-526)
int launch_symmetrize(double* A, int64_t ld, int n, cudaStream_t s);
int launch_loo_predictions(const double* Kinv, int64_t ld, const double* alpha, const double* y, int n, double* mu,
                           double* sigma, cudaStream_t s);
int launch_loo(const CovParams& cp, const MeanParams& mp, int n_theta_mean, const double* x, int n, int npad,
               const double* alpha, double* Kinv, int64_t ld, double* ws, double* Dk, double* T, double* val_dev,
               double* grad_dev, const double* dK_all, int n_cov, cudaStream_t s);

}  // namespace gpb
