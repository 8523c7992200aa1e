// GpLinearInverter on the GPU (reference inference/gp/inversion.py:11-249): posterior of y = A x + noise with a
// Gaussian-process prior on x.  Everything is a composition of the engine's existing pieces:
//   K (full symmetric)         assemble_train_kernel (mirror)                       inversion.py:149, 180, 195
//   J = A K A^T + Sigma        two DMMA GEMMs (T = A K, J = T A^T) + diagonal add   :182, :196
//   L = chol(J), v, alpha      potrf_lower / trsv                                   :182-183, :203-209
//   gradient                   1/2 sum (alpha alpha^T - iJ) o (A dK A^T) = 1/2 sum (a a^T - M) o dK with a = A^T alpha and
//                              M = A^T iJ A = X X^T, X = A^T L^-T (one batched solve + one GEMM); then the SAME fused
//                              trace kernels as GpRegressor (lml.cu), so no dK_p or A dK_p A^T is ever materialised
//                              (the reference forms both per parameter, :196-197, :215)
//   posterior                  cov = K - (K A^T L^-T)(K A^T L^-T)^T, mean = mu + K A^T alpha  (Woodbury form of
//                              solve(I + K W, K), :150-155 -- same quantity, SPD factorisation instead of a general solve)
#include "ctx.cuh"

using namespace gpb;

struct gpb_linv {
    int64_t m = 0, mpad = 0;
    double *A = nullptr, *At = nullptr, *y = nullptr, *sig2 = nullptr;        // A: mpad x npad, At: npad x mpad
    double *J = nullptr, *dinv = nullptr, *T = nullptr, *X = nullptr, *tmp = nullptr;
    double *vec = nullptr, *resid = nullptr, *alpha = nullptr, *a = nullptr, *mu = nullptr, *f = nullptr;
    int* info = nullptr;
};

namespace {

__global__ void add_diag_kernel(double* __restrict__ J, int64_t ld, const double* __restrict__ d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) J[(int64_t)i * ld + i] += d[i];
}
// out = a - b  (n entries)
__global__ void sub_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}
__global__ void add_kernel(const double* a, const double* b, double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

void free_linv(gpb_linv* v) {
    if (!v) return;
    for (double* p : {v->A, v->At, v->y, v->sig2, v->J, v->dinv, v->T, v->X, v->tmp, v->vec, v->resid, v->alpha, v->a, v->mu, v->f})
        if (p) cudaFree(p);
    if (v->info) cudaFree(v->info);
    delete v;
}

// shared front part: K -> Kwork (full), mu, J = A K A^T + Sigma factored in place, r = y - A mu in v->vec[0..mpad)
int factor_j(gpb_ctx* c, const double* theta, CovParams& cp, MeanParams& mp, int* info_host) {
    gpb_linv* v = c->linv;
    const int npad = (int)c->npad, n = (int)c->n, mpad = (int)v->mpad;
    GPB_TRY(ctx_make_cov_params(c, theta + c->n_mean, cp));
    ctx_make_mean_params(c, theta, mp);
    GPB_TRY(ensure(c->Kwork, c->Kwork_cap, sizeof(double) * (size_t)npad * npad));
    GPB_TRY(launch_assemble_train(cp, c->x, n, npad, nullptr, nullptr, c->Kwork, npad, 1, c->s));
    GPB_TRY(launch_residual(mp, c->x, c->y, n, npad, v->resid /*scratch: -mu*/, v->mu, c->s));
    GemmArgs g1{mpad, npad, npad, v->A, npad, c->Kwork, npad, nullptr, 0, v->T, npad, nullptr, 0, 1.0, 0.0, GEMM_FULL};
    GPB_TRY(gemm_nt(g1, c->s));
    GemmArgs g2{mpad, mpad, npad, v->T, npad, v->A, npad, nullptr, 0, v->J, mpad, nullptr, 0, 1.0, 0.0, GEMM_FULL};
    GPB_TRY(gemm_nt(g2, c->s));
    add_diag_kernel<<<(mpad + 255) / 256, 256, 0, c->s>>>(v->J, mpad, v->sig2, mpad);
    GPB_CUDA(cudaGetLastError());
    GPB_TRY(launch_row_dot(v->A, npad, mpad, npad, v->mu, v->f, c->s));
    sub_kernel<<<(mpad + 255) / 256, 256, 0, c->s>>>(v->y, v->f, v->vec, mpad);
    GPB_CUDA(cudaGetLastError());
    GPB_CUDA(cudaMemcpyAsync(v->resid, v->vec, sizeof(double) * mpad, cudaMemcpyDeviceToDevice, c->s));
    count_launch(2);
    LinalgWs ws{v->dinv, v->tmp, std::max<int64_t>(mpad, npad), v->info};
    GPB_TRY(potrf_lower(v->J, mpad, mpad, ws, c->s));
    GPB_CUDA(cudaMemcpyAsync(info_host, v->info, sizeof(int), cudaMemcpyDeviceToHost, c->s));
    return 0;
}

int need_linv(gpb_ctx* c) {
    GPB_TRY(ctx_use(c));
    GPB_TRY(ctx_need_model(c));
    if (!c->linv) {
        set_error("gpb_linv_set_problem must be called first");
        return -2;
    }
    GPB_TRY(ensure(c->scal, c->scal_cap, sizeof(double) * 8));
    return 0;
}

}  // namespace

namespace gpb {
void linv_destroy(gpb_ctx* c) {
    free_linv(c->linv);
    c->linv = nullptr;
}
}  // namespace gpb

extern "C" {

int gpb_linv_set_problem(gpb_ctx* c, const double* A, int64_t m, const double* y, const double* y_err) {
    GPB_TRY(ctx_use(c));
    if (c->n == 0 || m <= 0) {
        set_error("gpb_linv_set_problem: call gpb_set_data (parameter positions) first and pass m > 0");
        return -2;
    }
    linv_destroy(c);
    gpb_linv* v = new gpb_linv();
    c->linv = v;
    const int64_t n = c->n, npad = c->npad, mpad = round_up(m, NB);
    v->m = m;
    v->mpad = mpad;
    std::vector<double> hA((size_t)mpad * npad, 0.0), hAt((size_t)npad * mpad, 0.0), hy(mpad, 0.0), hs(mpad, 1.0);
    for (int64_t i = 0; i < m; ++i) {
        for (int64_t j = 0; j < n; ++j) {
            hA[i * npad + j] = A[i * n + j];
            hAt[j * mpad + i] = A[i * n + j];
        }
        hy[i] = y[i];
        hs[i] = y_err[i] * y_err[i];  // padded rows keep 1: J is padded with the identity
    }
    auto up = [&](double*& dst, const std::vector<double>& src) -> int {
        GPB_CUDA(cudaMalloc(&dst, sizeof(double) * src.size()));
        GPB_CUDA(cudaMemcpy(dst, src.data(), sizeof(double) * src.size(), cudaMemcpyHostToDevice));
        return 0;
    };
    GPB_TRY(up(v->A, hA));
    GPB_TRY(up(v->At, hAt));
    GPB_TRY(up(v->y, hy));
    GPB_TRY(up(v->sig2, hs));
    const int64_t big = std::max(mpad, npad);
    GPB_CUDA(cudaMalloc(&v->J, sizeof(double) * mpad * mpad));
    GPB_CUDA(cudaMalloc(&v->dinv, sizeof(double) * mpad * NB));
    GPB_CUDA(cudaMalloc(&v->T, sizeof(double) * mpad * npad));
    GPB_CUDA(cudaMalloc(&v->X, sizeof(double) * npad * mpad));
    GPB_CUDA(cudaMalloc(&v->tmp, sizeof(double) * big * NB));
    GPB_CUDA(cudaMalloc(&v->vec, sizeof(double) * 2 * mpad));
    GPB_CUDA(cudaMalloc(&v->resid, sizeof(double) * big));
    GPB_CUDA(cudaMalloc(&v->alpha, sizeof(double) * mpad));
    GPB_CUDA(cudaMalloc(&v->a, sizeof(double) * npad));
    GPB_CUDA(cudaMalloc(&v->mu, sizeof(double) * npad));
    GPB_CUDA(cudaMalloc(&v->f, sizeof(double) * mpad));
    GPB_CUDA(cudaMalloc(&v->info, sizeof(int)));
    return 0;
}

int gpb_linv_lml(gpb_ctx* c, const double* theta, double* lml, int* info) {
    GPB_TRY(need_linv(c));
    gpb_linv* v = c->linv;
    CovParams cp;
    MeanParams mp;
    int info_h = 0;
    GPB_TRY(factor_j(c, theta, cp, mp, &info_h));
    const int mpad = (int)v->mpad;
    GPB_TRY(trsv_lower_fwd(v->J, mpad, mpad, v->dinv, v->vec, c->s));
    GPB_TRY(launch_logdet_dot(v->J, mpad, v->vec + mpad, v->vec + mpad, (int)v->m, c->scal, c->s));
    double sc[2];
    GPB_CUDA(cudaMemcpyAsync(sc, c->scal, sizeof(sc), cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    *lml = -0.5 * sc[1] - sc[0];
    return 0;
}

int gpb_linv_lml_grad(gpb_ctx* c, const double* theta, double* lml, double* grad, int* info) {
    GPB_TRY(need_linv(c));
    gpb_linv* v = c->linv;
    const int npad = (int)c->npad, n = (int)c->n, mpad = (int)v->mpad, nt = c->n_mean + c->n_cov;
    CovParams cp;
    MeanParams mp;
    int info_h = 0;
    GPB_TRY(factor_j(c, theta, cp, mp, &info_h));
    GPB_TRY(ensure(c->Kinv, c->Kinv_cap, sizeof(double) * (size_t)npad * npad));
    GPB_TRY(ensure(c->partials, c->partials_cap, std::max(trace_partials_size(npad), col_dot_ws_bytes(npad, npad))));
    GPB_TRY(ensure(c->grad_dev, c->grad_cap, sizeof(double) * (nt + 2)));
    // alpha = J^-1 r
    GPB_TRY(trsv_lower_fwd(v->J, mpad, mpad, v->dinv, v->vec, c->s));
    GPB_CUDA(cudaMemcpyAsync(v->vec, v->vec + mpad, sizeof(double) * mpad, cudaMemcpyDeviceToDevice, c->s));
    GPB_TRY(trsv_lower_bwd(v->J, mpad, mpad, v->dinv, v->vec, c->s));
    GPB_CUDA(cudaMemcpyAsync(v->alpha, v->vec + mpad, sizeof(double) * mpad, cudaMemcpyDeviceToDevice, c->s));
    GPB_TRY(launch_logdet_dot(v->J, mpad, v->resid, v->alpha, (int)v->m, c->scal, c->s));
    // X = A^T L^-T, M = X X^T (lower tiles), a = A^T alpha
    GPB_CUDA(cudaMemcpyAsync(v->X, v->At, sizeof(double) * (size_t)npad * mpad, cudaMemcpyDeviceToDevice, c->s));
    LinalgWs ws{v->dinv, v->tmp, std::max<int64_t>(mpad, npad), v->info};
    GPB_TRY(trsm_right_lt(v->X, mpad, npad, v->J, mpad, mpad, 0, ws, c->s));
    GemmArgs g{npad, npad, mpad, v->X, mpad, v->X, mpad, nullptr, 0, c->Kinv, npad, nullptr, 0, 1.0, 0.0, GEMM_LOWER};
    GPB_TRY(gemm_nt(g, c->s));
    GPB_TRY(launch_row_dot(v->At, mpad, npad, mpad, v->alpha, v->a, c->s));
    GPB_CUDA(cudaMemsetAsync(c->grad_dev, 0, sizeof(double) * (nt + 2), c->s));
    GPB_TRY(launch_lml_grad(cp, mp, c->n_mean, c->x, n, npad, v->a, c->Kinv, npad, c->partials, c->grad_dev, nullptr, c->s));
    double sc[2];
    GPB_CUDA(cudaMemcpyAsync(sc, c->scal, sizeof(sc), cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaMemcpyAsync(grad, c->grad_dev, sizeof(double) * nt, cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    *lml = -0.5 * sc[1] - sc[0];
    return 0;
}

int gpb_linv_posterior(gpb_ctx* c, const double* theta, double* mean, double* cov_or_null, int* info) {
    GPB_TRY(need_linv(c));
    gpb_linv* v = c->linv;
    const int npad = (int)c->npad, n = (int)c->n, mpad = (int)v->mpad;
    CovParams cp;
    MeanParams mp;
    int info_h = 0;
    GPB_TRY(factor_j(c, theta, cp, mp, &info_h));
    GPB_TRY(trsv_lower_fwd(v->J, mpad, mpad, v->dinv, v->vec, c->s));
    GPB_CUDA(cudaMemcpyAsync(v->vec, v->vec + mpad, sizeof(double) * mpad, cudaMemcpyDeviceToDevice, c->s));
    GPB_TRY(trsv_lower_bwd(v->J, mpad, mpad, v->dinv, v->vec, c->s));
    GPB_CUDA(cudaMemcpyAsync(v->alpha, v->vec + mpad, sizeof(double) * mpad, cudaMemcpyDeviceToDevice, c->s));
    // K A^T (npad x mpad); mean = mu + K A^T alpha
    GemmArgs g1{npad, mpad, npad, c->Kwork, npad, v->A, npad, nullptr, 0, v->X, mpad, nullptr, 0, 1.0, 0.0, GEMM_FULL};
    GPB_TRY(gemm_nt(g1, c->s));
    GPB_TRY(launch_row_dot(v->X, mpad, npad, mpad, v->alpha, v->a, c->s));
    add_kernel<<<(npad + 255) / 256, 256, 0, c->s>>>(v->a, v->mu, v->a, npad);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    GPB_CUDA(cudaMemcpyAsync(mean, v->a, sizeof(double) * n, cudaMemcpyDeviceToHost, c->s));
    if (cov_or_null) {
        LinalgWs ws{v->dinv, v->tmp, std::max<int64_t>(mpad, npad), v->info};
        GPB_TRY(trsm_right_lt(v->X, mpad, npad, v->J, mpad, mpad, 0, ws, c->s));
        GemmArgs g2{npad, npad, mpad, v->X, mpad, v->X, mpad, c->Kwork, npad, c->Kwork, npad, nullptr, 0, -1.0, 1.0, GEMM_FULL};
        GPB_TRY(gemm_nt(g2, c->s));
        GPB_CUDA(cudaMemcpy2DAsync(cov_or_null, sizeof(double) * n, c->Kwork, sizeof(double) * npad, sizeof(double) * n, n,
                                   cudaMemcpyDeviceToHost, c->s));
    }
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    return 0;
}

}  // extern "C"
