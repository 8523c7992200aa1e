// C-ABI of libgpb200 (declared in include/gpb200.h): context, device state and the host-side drivers
// that string the kernels together.  No CPU arithmetic on the path: the host only maps the log
// hyper-parameters (a handful of exp() calls), sequences launches and moves results.
#include "ctx.cuh"

#include <thread>

namespace gpb {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace gpb

using namespace gpb;

namespace {

int use(gpb_ctx* c) {
    if (!c) {
        set_error("null context");
        return -2;
    }
    GPB_CUDA(cudaSetDevice(c->device));
    return 0;
}

int need_model(gpb_ctx* c) {
    if (c->n == 0 || !c->model_set) {
        set_error("gpb_set_data and gpb_set_model must be called first");
        return -2;
    }
    return 0;
}

int n_cov_params(int kind, int64_t n, int d) {
    switch (kind) {
        case COV_SE: return d + 1;
        case COV_RQ: return d + 2;
        case COV_WHITE: return 1;
        default: return (int)n;
    }
}

// covariance.py:248-249 (SE), :344-346 (RQ), :167 (White), :678 (Hetero)
int make_cov_params(gpb_ctx* c, const double* tc, CovParams& cp) {
    std::memset(&cp, 0, sizeof(cp));
    cp.ncomp = c->ncomp;
    cp.d = c->d;
    cp.jitter = 1e-12;
    cp.hetero_log_sigma = nullptr;
    cp.n_regions = c->n_regions;
    cp.cp_axis = c->cp_axis;
    cp.cp_theta_off = c->cp_theta_off;
    for (int a = 0; a + 1 < c->n_regions; ++a) {  // covariance.py:592-595: theta = (location, width), not in log space
        cp.cp_loc[a] = tc[c->cp_theta_off + 2 * a];
        cp.cp_width[a] = tc[c->cp_theta_off + 2 * a + 1];
    }
    for (int i = 0; i < MAX_COMP; ++i) cp.region[i] = -1;
    for (int i = 0; i < c->ncomp; ++i) {
        const int off = c->theta_off[i];
        cp.kind[i] = c->kinds[i];
        cp.theta_off[i] = off;
        cp.region[i] = c->region[i];
        if (c->kinds[i] == COV_SE) {
            const double a = std::exp(tc[off]);
            cp.amp2[i] = a * a;
            for (int k = 0; k < c->d; ++k) {
                const double l = std::exp(tc[off + 1 + k]);
                cp.inv_l2[i][k] = 1.0 / (l * l);
            }
        } else if (c->kinds[i] == COV_RQ) {
            const double a = std::exp(tc[off]);
            cp.amp2[i] = a * a;
            cp.rq_alpha[i] = std::exp(tc[off + 1]);
            for (int k = 0; k < c->d; ++k) {
                const double l = std::exp(tc[off + 2 + k]);
                cp.inv_l2[i][k] = 1.0 / (l * l);
            }
        } else if (c->kinds[i] == COV_WHITE) {
            cp.amp2[i] = std::exp(2.0 * tc[off]);
        } else {
            GPB_TRY(ensure(c->theta_dev, c->theta_dev_cap, sizeof(double) * c->npad));
            GPB_CUDA(cudaMemsetAsync(c->theta_dev, 0, sizeof(double) * c->npad, c->s));
            GPB_CUDA(cudaMemcpyAsync(c->theta_dev, tc + off, sizeof(double) * c->n, cudaMemcpyHostToDevice, c->s));
            GPB_CUDA(cudaStreamSynchronize(c->s));  // tc is caller memory
            cp.hetero_log_sigma = c->theta_dev;
        }
    }
    return 0;
}

void make_mean_params(gpb_ctx* c, const double* tm, MeanParams& mp) {
    std::memset(&mp, 0, sizeof(mp));
    mp.kind = c->mean_kind;
    mp.d = c->d;
    mp.c0 = tm[0];
    for (int k = 0; k < c->d; ++k) {
        mp.xbar[k] = c->xbar[k];
        if (c->mean_kind >= MEAN_LINEAR) mp.lin[k] = tm[1 + k];
        if (c->mean_kind == MEAN_QUADRATIC) mp.quad[k] = tm[1 + c->d + k];
    }
}

int ensure_linalg_ws(gpb_ctx* c) {
    const size_t dinv_bytes = sizeof(double) * (size_t)c->npad * NB;
    GPB_TRY(ensure(c->dinv_work, c->dinv_work_cap, dinv_bytes));
    GPB_TRY(ensure(c->dinv_fit, c->dinv_fit_cap, dinv_bytes));
    const int64_t rows = std::max<int64_t>(c->npad, 148 * 128 * 4);
    GPB_TRY(ensure(c->tmp, c->tmp_cap, sizeof(double) * (size_t)rows * NB));
    c->tmp_rows = rows;
    if (!c->info_dev) GPB_CUDA(cudaMalloc(&c->info_dev, sizeof(int)));
    GPB_TRY(ensure(c->vec, c->vec_cap, sizeof(double) * 2 * c->npad));
    GPB_TRY(ensure(c->resid, c->resid_cap, sizeof(double) * c->npad));
    GPB_TRY(ensure(c->alpha_work, c->alpha_work_cap, sizeof(double) * c->npad));
    GPB_TRY(ensure(c->scal, c->scal_cap, sizeof(double) * 8));
    return 0;
}

}  // namespace
namespace gpb {
int ctx_use(gpb_ctx* c) { return use(c); }
int ctx_need_model(gpb_ctx* c) { return need_model(c); }
int ctx_make_cov_params(gpb_ctx* c, const double* t, CovParams& cp) { return make_cov_params(c, t, cp); }
void ctx_make_mean_params(gpb_ctx* c, const double* t, MeanParams& mp) { make_mean_params(c, t, mp); }
}  // namespace gpb
namespace {

LinalgWs ws_of(gpb_ctx* c, double* dinv) { return LinalgWs{dinv, c->tmp, c->tmp_rows, c->info_dev}; }

std::string pkey(const char* tag, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> dims) {
    std::string k(tag);
    for (const void* p : ptrs) k += ":" + std::to_string(reinterpret_cast<uintptr_t>(p));
    for (int64_t v : dims) k += "/" + std::to_string(v);
    return k;
}

void clear_graphs(gpb_ctx* c) {
    for (auto& kv : c->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    c->graphs.clear();
}

// Run a launch sequence whose arguments are only device pointers and shapes.  First use: plain launches (also
// performs the one-time kernel attribute setup).  Second use: the same sequence is captured into a CUDA graph;
// from then on every call is one cudaGraphLaunch -- the recursion issues hundreds of short kernels whose host
// launch cost otherwise dominates for N <= 8192.  GPB200_NO_GRAPHS=1 disables the cache.
template <class F>
int run_graphed(gpb_ctx* c, const std::string& key0, F&& body) {
    if (c->opt_epoch != option_epoch()) {  // a runtime option changed: recorded sequences may pick other kernels now
        clear_graphs(c);
        c->opt_epoch = option_epoch();
        c->use_graphs = option(OPT_GRAPHS) != 0;
    }
    if (!c->use_graphs) return body();
    // the per-thread DMMA override (retry after a failed INT8 factorisation) selects different kernels: separate graphs
    const std::string key = key0 + "#" + std::to_string(gemm_i8_override());
    auto& e = c->graphs[key];
    if (e.exec) {
        GPB_CUDA(cudaGraphLaunch(e.exec, c->s));
        count_launch((int)e.launches);
        credit_gemm_flops(e.flops, e.flops_i8);
        return 0;
    }
    if (e.uses++ == 0) return body();
    const int64_t l0 = thread_launch_count();
    const double f0 = thread_gemm_flops(), g0 = thread_gemm_flops_i8();
    GPB_CUDA(cudaStreamBeginCapture(c->s, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->s, &g);
    if (rc != 0 || ce != cudaSuccess || g == nullptr) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        c->use_graphs = false;  // fall back to plain launches for the rest of this context's life
        if (rc != 0) return rc;
        return body();
    }
    e.launches = thread_launch_count() - l0;
    e.flops = thread_gemm_flops() - f0;
    e.flops_i8 = thread_gemm_flops_i8() - g0;
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) {
        e.exec = nullptr;
        cudaGetLastError();
        c->use_graphs = false;
        return body();
    }
    GPB_CUDA(cudaGraphLaunch(e.exec, c->s));
    return 0;
}

// assemble K(theta)+sig into `K` (lower tiles), factor in place, solve for alpha.
//   resid_out: y - mu (optional copy), v_out: L^-1 r left in c->vec[npad..) when !want_alpha
enum SolveMode { SOLVE_NONE = 0, SOLVE_FWD = 1, SOLVE_BOTH = 2 };
int assemble_and_factor(gpb_ctx* c, const CovParams& cp, const MeanParams& mp, double* K, double* dinv, double* mu_out,
                        int solve, double* alpha_out, int* info_host) {
    const bool want_alpha = solve == SOLVE_BOTH;
    const int npad = (int)c->npad, n = (int)c->n;
    c->timer.mark("assemble");
    GPB_TRY(launch_assemble_train(cp, c->x, n, npad, c->has_noise ? c->noise : nullptr,
                                  c->has_ycov ? c->ycov : nullptr, K, npad, 0, c->s));
    c->timer.mark("potrf");
    GPB_TRY(run_graphed(c, pkey("potrf", {K, dinv, c->tmp, c->info_dev}, {npad}),
                        [&]() { return potrf_lower(K, npad, npad, ws_of(c, dinv), c->s); }));
    c->timer.mark("solve");
    GPB_TRY(launch_residual(mp, c->x, c->y, n, npad, c->vec, mu_out, c->s));
    if (solve == SOLVE_NONE) {  // the caller solves through the explicit inverse (marginal_likelihood_gradient path)
        GPB_CUDA(cudaMemcpyAsync(c->resid, c->vec, sizeof(double) * npad, cudaMemcpyDeviceToDevice, c->s));
        GPB_CUDA(cudaMemcpyAsync(info_host, c->info_dev, sizeof(int), cudaMemcpyDeviceToHost, c->s));
        return 0;
    }
    GPB_TRY(run_graphed(c, pkey("solve", {K, dinv, c->vec, c->resid, alpha_out}, {npad, want_alpha}), [&]() -> int {
        GPB_CUDA(cudaMemcpyAsync(c->resid, c->vec, sizeof(double) * npad, cudaMemcpyDeviceToDevice, c->s));
        GPB_TRY(trsv_lower_fwd(K, npad, npad, dinv, c->vec, c->s));
        if (want_alpha) {
            GPB_CUDA(cudaMemcpyAsync(c->vec, c->vec + npad, sizeof(double) * npad, cudaMemcpyDeviceToDevice, c->s));
            GPB_TRY(trsv_lower_bwd(K, npad, npad, dinv, c->vec, c->s));
            GPB_CUDA(cudaMemcpyAsync(alpha_out, c->vec + npad, sizeof(double) * npad, cudaMemcpyDeviceToDevice, c->s));
        }
        return 0;
    }));
    GPB_CUDA(cudaMemcpyAsync(info_host, c->info_dev, sizeof(int), cudaMemcpyDeviceToHost, c->s));
    return 0;
}

// A non-positive pivot reported while GEMMs ran on the INT8 path is re-checked on the FP64 DMMA kernels before it is
// believed: the digit splitting is accurate normwise (2^-55 of the row maxima), and a Schur complement that is positive
// only by the 1e-12 a^2 jitter (near-noise-free SE / RQ data) can lose its sign to that.  "i8_fallback" = 0 disables.
template <class F>
int with_dmma_retry(gpb_ctx* c, int* info, F&& body) {
    const double f0 = thread_gemm_flops_i8();
    int rc = body();
    if (rc == 0 && *info > 0 && option(OPT_I8_FALLBACK) && gemm_i8_override() < 0 && thread_gemm_flops_i8() > f0) {
        set_gemm_i8_override(0);
        rc = body();
        set_gemm_i8_override(-1);
        ++c->dmma_retries;
    }
    return rc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* gpb_last_error(void) { return g_last_error.c_str(); }

int gpb_device_count(int* count) {
    GPB_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int gpb_ctx_create(int device, gpb_ctx** out) {
    int count = 0;
    GPB_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) {
        set_error("gpb_ctx_create: no such CUDA device " + std::to_string(device) + " (visible: " +
                  std::to_string(count) + "); this library has no CPU fallback");
        return -2;
    }
    GPB_CUDA(cudaSetDevice(device));
    gpb_ctx* c = new gpb_ctx();
    c->device = device;
    GPB_CUDA(cudaStreamCreateWithFlags(&c->s, cudaStreamNonBlocking));
    c->timer.s = c->s;
    c->use_graphs = option(OPT_GRAPHS) != 0;
    c->opt_epoch = option_epoch();
    *out = c;
    return 0;
}

void gpb_ctx_destroy(gpb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->s);
    double* ptrs[] = {c->x, c->y, c->noise, c->ycov, c->theta_dev, c->Lfit, c->dinv_fit, c->alpha, c->mu, c->Kwork,
                      c->dinv_work, c->W, c->Kinv, c->partials, c->grad_dev, c->vec, c->resid, c->alpha_work, c->scal,
                      c->tmp, c->S, c->dots, c->G, c->qbuf, c->o1, c->o2, c->o3, c->R_dev,
                      reinterpret_cast<double*>(c->argws)};
    for (double* p : ptrs)
        if (p) cudaFree(p);
    for (void* p : {(void*)c->pp.lbuf, (void*)c->pp.wbuf, (void*)c->pp.xbuf, (void*)c->pp.tbuf, (void*)c->pp.lscale,
                    (void*)c->pp.wscale, (void*)c->pp.xscale, (void*)c->pp.tscale, (void*)c->pp.bound, (void*)c->pp.wtmp})
        if (p) cudaFree(p);
    dist_destroy(c);
    linv_destroy(c);
    clear_graphs(c);
    if (c->info_dev) cudaFree(c->info_dev);
    c->timer.reset();
    gemm_i8_release(c->s);
    cudaStreamDestroy(c->s);
    delete c;
}

int gpb_set_data(gpb_ctx* c, const double* x, int64_t n, int d, const double* y, const double* noise_var,
                 const double* y_cov) {
    GPB_TRY(use(c));
    if (n <= 0 || d <= 0 || d > MAX_DIM) {
        set_error("gpb_set_data: need n > 0 and 1 <= d <= " + std::to_string(MAX_DIM));
        return -2;
    }
    for (double** p : {&c->x, &c->y, &c->noise, &c->ycov}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    clear_graphs(c);
    linv_destroy(c);  // A is padded to this context's npad
    c->n = n;
    c->d = d;
    c->npad = round_up(n, NB);
    c->fitted = false;
    const size_t np = (size_t)c->npad;
    GPB_CUDA(cudaMalloc(&c->x, sizeof(double) * np * d));
    GPB_CUDA(cudaMalloc(&c->y, sizeof(double) * np));
    GPB_CUDA(cudaMemset(c->x, 0, sizeof(double) * np * d));
    GPB_CUDA(cudaMemset(c->y, 0, sizeof(double) * np));
    GPB_CUDA(cudaMemcpy(c->x, x, sizeof(double) * n * d, cudaMemcpyHostToDevice));
    GPB_CUDA(cudaMemcpy(c->y, y, sizeof(double) * n, cudaMemcpyHostToDevice));
    c->has_noise = noise_var != nullptr;
    if (noise_var) {
        GPB_CUDA(cudaMalloc(&c->noise, sizeof(double) * np));
        GPB_CUDA(cudaMemset(c->noise, 0, sizeof(double) * np));
        GPB_CUDA(cudaMemcpy(c->noise, noise_var, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    c->has_ycov = y_cov != nullptr;
    if (y_cov) {
        GPB_CUDA(cudaMalloc(&c->ycov, sizeof(double) * (size_t)n * n));
        GPB_CUDA(cudaMemcpy(c->ycov, y_cov, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
    }
    // column means of the training inputs (mean.py:59, 92), numpy pairwise order is not reproduced:
    // plain left-to-right accumulation in long double, rounded once
    for (int k = 0; k < d; ++k) {
        long double acc = 0.0L;
        for (int64_t i = 0; i < n; ++i) acc += x[i * d + k];
        c->xbar[k] = (double)(acc / (long double)n);
    }
    c->model_set = false;  // parameter layout depends on n (HeteroscedasticNoise): gpb_set_model must follow
    return 0;
}

int gpb_set_model_ex(gpb_ctx* c, const int* cov_kinds, const int* theta_offs, const int* regions, int ncomp,
                     int n_regions, int cp_axis, int cp_theta_off, int n_cov_params_total, int mean_kind) {
    GPB_TRY(use(c));
    if (ncomp < 1 || ncomp > MAX_COMP) {
        set_error("gpb_set_model: 1.." + std::to_string(MAX_COMP) + " covariance components supported");
        return -2;
    }
    if (c->n == 0) {
        set_error("gpb_set_model: call gpb_set_data first");
        return -2;
    }
    if (n_regions != 0 && (n_regions < 2 || n_regions > MAX_REG || cp_axis < 0 || cp_axis >= c->d)) {
        set_error("gpb_set_model: a ChangePoint needs 2.." + std::to_string(MAX_REG) + " regions and a valid axis");
        return -2;
    }
    int n_hetero = 0;
    for (int i = 0; i < ncomp; ++i) {
        if (cov_kinds[i] < COV_SE || cov_kinds[i] > COV_HETERO) {
            set_error("gpb_set_model: unknown covariance kind");
            return -2;
        }
        const int reg = regions ? regions[i] : -1;
        if (reg >= n_regions || reg < -1) {
            set_error("gpb_set_model: region index out of range");
            return -2;
        }
        n_hetero += cov_kinds[i] == COV_HETERO;
        c->kinds[i] = cov_kinds[i];
        c->theta_off[i] = theta_offs[i];
        c->region[i] = reg;
        if (theta_offs[i] < 0 || theta_offs[i] + n_cov_params(cov_kinds[i], c->n, c->d) > n_cov_params_total) {
            set_error("gpb_set_model: parameter offsets exceed the hyper-parameter vector");
            return -2;
        }
    }
    if (n_hetero > 1) {
        set_error("gpb_set_model: at most one HeteroscedasticNoise component");
        return -2;
    }
    if (mean_kind < MEAN_CONST || mean_kind > MEAN_QUADRATIC) {
        set_error("gpb_set_model: unknown mean kind");
        return -2;
    }
    c->ncomp = ncomp;
    c->n_cov = n_cov_params_total;
    c->n_regions = n_regions;
    c->cp_axis = cp_axis;
    c->cp_theta_off = cp_theta_off;
    c->mean_kind = mean_kind;
    c->n_mean = mean_kind == MEAN_CONST ? 1 : (mean_kind == MEAN_LINEAR ? 1 + c->d : 1 + 2 * c->d);
    c->model_set = true;
    c->fitted = false;
    return 0;
}

int gpb_set_model(gpb_ctx* c, const int* cov_kinds, int ncomp, int mean_kind) {
    GPB_TRY(use(c));
    if (ncomp < 1 || ncomp > MAX_COMP || c->n == 0) {
        set_error("gpb_set_model: call gpb_set_data first and pass 1.." + std::to_string(MAX_COMP) + " components");
        return -2;
    }
    int offs[MAX_COMP], regs[MAX_COMP], off = 0;
    for (int i = 0; i < ncomp; ++i) {
        offs[i] = off;
        regs[i] = -1;
        off += n_cov_params(cov_kinds[i], c->n, c->d);
    }
    return gpb_set_model_ex(c, cov_kinds, offs, regs, ncomp, 0, 0, 0, off, mean_kind);
}

int gpb_num_hyperpars(gpb_ctx* c, int* n_mean, int* n_cov) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    *n_mean = c->n_mean;
    *n_cov = c->n_cov;
    return 0;
}

int gpb_build_covariance(gpb_ctx* c, const double* theta_cov, int add_sig, double* K_out) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    CovParams cp;
    GPB_TRY(make_cov_params(c, theta_cov, cp));
    const size_t np = (size_t)c->npad;
    GPB_TRY(ensure(c->Kwork, c->Kwork_cap, sizeof(double) * np * np));
    GPB_TRY(launch_assemble_train(cp, c->x, (int)c->n, (int)c->npad, (add_sig && c->has_noise) ? c->noise : nullptr,
                                  (add_sig && c->has_ycov) ? c->ycov : nullptr, c->Kwork, c->npad, 1, c->s));
    GPB_CUDA(cudaMemcpy2DAsync(K_out, sizeof(double) * c->n, c->Kwork, sizeof(double) * np, sizeof(double) * c->n,
                               c->n, cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}

int gpb_covariance_and_gradients(gpb_ctx* c, const double* theta_cov, double* K_out, double* dK_out) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    CovParams cp;
    GPB_TRY(make_cov_params(c, theta_cov, cp));
    const size_t plane = (size_t)c->n * c->n;
    double *dK = nullptr, *K = nullptr;
    GPB_CUDA(cudaMalloc(&K, sizeof(double) * plane));
    GPB_CUDA(cudaMalloc(&dK, sizeof(double) * plane * c->n_cov));
    int r = launch_assemble_grads(cp, c->x, (int)c->n, K, dK, c->s);
    if (r == 0) {
        cudaMemcpyAsync(K_out, K, sizeof(double) * plane, cudaMemcpyDeviceToHost, c->s);
        cudaMemcpyAsync(dK_out, dK, sizeof(double) * plane * c->n_cov, cudaMemcpyDeviceToHost, c->s);
    }
    cudaError_t e = cudaStreamSynchronize(c->s);
    cudaFree(K);
    cudaFree(dK);
    if (r) return r;
    GPB_CUDA(e);
    return 0;
}

int gpb_cross_covariance(gpb_ctx* c, const double* u, int64_t m, const double* v, int64_t n, const double* theta_cov,
                         double* out) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    if (m == 0 || n == 0) return 0;
    CovParams cp;
    GPB_TRY(make_cov_params(c, theta_cov, cp));
    double *du = nullptr, *dv = nullptr, *dout = nullptr;
    GPB_CUDA(cudaMalloc(&du, sizeof(double) * m * c->d));
    GPB_CUDA(cudaMalloc(&dv, sizeof(double) * n * c->d));
    GPB_CUDA(cudaMalloc(&dout, sizeof(double) * m * n));
    cudaMemcpyAsync(du, u, sizeof(double) * m * c->d, cudaMemcpyHostToDevice, c->s);
    cudaMemcpyAsync(dv, v, sizeof(double) * n * c->d, cudaMemcpyHostToDevice, c->s);
    int r = launch_cross_cov(cp, du, (int)m, dv, (int)n, dout, n, c->s);
    if (r == 0) cudaMemcpyAsync(out, dout, sizeof(double) * m * n, cudaMemcpyDeviceToHost, c->s);
    cudaError_t e = cudaStreamSynchronize(c->s);
    cudaFree(du);
    cudaFree(dv);
    cudaFree(dout);
    if (r) return r;
    GPB_CUDA(e);
    return 0;
}

static int factor_impl(gpb_ctx* c, const double* theta, int* info) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    c->timer.reset();
    const size_t np = (size_t)c->npad;
    GPB_TRY(ensure_linalg_ws(c));
    GPB_TRY(ensure(c->Lfit, c->Lfit_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->alpha, c->alpha_cap, sizeof(double) * np));
    GPB_TRY(ensure(c->mu, c->mu_cap, sizeof(double) * np));
    c->fitted = false;
    c->pp.valid = false;  // cached digit planes of L belong to the previous fit
    c->pp.ns = 0;
    GPB_TRY(make_cov_params(c, theta + c->n_mean, c->cp_fit));
    make_mean_params(c, theta, c->mp_fit);
    int info_h = 0;
    GPB_TRY(assemble_and_factor(c, c->cp_fit, c->mp_fit, c->Lfit, c->dinv_fit, c->mu, SOLVE_BOTH, c->alpha, &info_h));
    // smallest pivot of the factor: the conditioning measure behind the predict solve's choice of diagonal-block variant
    GPB_TRY(launch_logdet_dot(c->Lfit, c->npad, c->alpha, c->alpha, (int)c->n, c->scal, c->s));
    double sc3[3] = {0.0, 0.0, 0.0};
    GPB_CUDA(cudaMemcpyAsync(sc3, c->scal, sizeof(sc3), cudaMemcpyDeviceToHost, c->s));
    c->timer.mark("end");
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    c->min_pivot_fit = sc3[2];
    if (info_h == 0) {
        c->theta_fit.assign(theta, theta + c->n_mean + c->n_cov);
        c->fitted = true;
    }
    return 0;
}

int gpb_get(gpb_ctx* c, int which, double* out) {
    GPB_TRY(use(c));
    if (!c->fitted) {
        set_error("gpb_get: no fitted state (call gpb_factor)");
        return -2;
    }
    const size_t np = (size_t)c->npad;
    const int64_t n = c->n;
    if (which == GPB_GET_ALPHA || which == GPB_GET_MU) {
        GPB_CUDA(cudaMemcpyAsync(out, which == GPB_GET_ALPHA ? c->alpha : c->mu, sizeof(double) * n,
                                 cudaMemcpyDeviceToHost, c->s));
        GPB_CUDA(cudaStreamSynchronize(c->s));
        return 0;
    }
    if (which == GPB_GET_L) {
        GPB_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * n, c->Lfit, sizeof(double) * np, sizeof(double) * n, n,
                                   cudaMemcpyDeviceToHost, c->s));
        GPB_CUDA(cudaStreamSynchronize(c->s));
        for (int64_t i = 0; i < n; ++i)  // numpy.linalg.cholesky returns zeros above the diagonal
            for (int64_t j = i + 1; j < n; ++j) out[i * n + j] = 0.0;
        return 0;
    }
    if (which == GPB_GET_K_XX) return gpb_build_covariance(c, c->theta_fit.data() + c->n_mean, 1, out);
    set_error("gpb_get: unknown selector");
    return -2;
}

static int lml_impl(gpb_ctx* c, const double* theta, double* lml, int* info) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    c->timer.reset();
    const size_t np = (size_t)c->npad;
    GPB_TRY(ensure_linalg_ws(c));
    GPB_TRY(ensure(c->Kwork, c->Kwork_cap, sizeof(double) * np * np));
    CovParams cp;
    MeanParams mp;
    GPB_TRY(make_cov_params(c, theta + c->n_mean, cp));
    make_mean_params(c, theta, mp);
    int info_h = 0;
    GPB_TRY(assemble_and_factor(c, cp, mp, c->Kwork, c->dinv_work, nullptr, SOLVE_FWD, nullptr, &info_h));
    // -0.5 v.v - sum log L_ii   (regression.py:538-539)
    GPB_TRY(launch_logdet_dot(c->Kwork, c->npad, c->vec + c->npad, c->vec + c->npad, (int)c->n, c->scal, c->s));
    double sc[2];
    GPB_CUDA(cudaMemcpyAsync(sc, c->scal, sizeof(sc), cudaMemcpyDeviceToHost, c->s));
    c->timer.mark("end");
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    *lml = -0.5 * sc[1] - sc[0];
    return 0;
}

// Error guard of the INT8 inverse chain.  The digit-split GEMM is accurate normwise: an entry of K^-1 = W^T W comes out
// with an absolute error of about 2^-56 sqrt(k) max|W|^2 (k = the chunk length of the product, 4096), however small the
// entry is, and the gradient trace 1/2 sum (alpha alpha^T - K^-1) o dK_p sums those errors against a smooth dK_p while the
// exact terms largely cancel.  The FP64 DMMA kernels are accurate componentwise and do not have this problem.  After the
// INT8 pass the estimate
//     est = GUARD_C 2^-56 sqrt(min(N, 4096)) (1 / min L_ii)^2 1/2 max_c sqrt(sum_ij max_p dK_p,ij^2)
// (the Frobenius sum comes out of the trace kernel for free) is compared with GUARD_TOL max(max|grad|, 1/2 |r.alpha|) --
// the second term is the size of the two quantities whose difference the gradient is, so the test does not degenerate
// where the gradient vanishes (the end of every L-BFGS run).  When est is larger, K^-1 and the traces are recomputed on
// DMMA from the same W.  GUARD_C = 20 and GUARD_TOL = 2e-10 are calibrated on profiles/conditioning_sweep_r2.md and
// profiles/grad_phase_sensitivity_r2.md (measured error / est between 4 and 94 with GUARD_C = 1 before the product was
// chunked; every case whose INT8 gradient error exceeded 1e-9 has est far above the threshold).
constexpr double GUARD_C = 20.0, GUARD_TOL = 2e-10;

static int lml_grad_impl(gpb_ctx* c, const double* theta, double* lml, double* grad, int* info) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    c->timer.reset();
    const size_t np = (size_t)c->npad;
    const int npad = (int)c->npad, n = (int)c->n, nt = c->n_mean + c->n_cov;
    GPB_TRY(ensure_linalg_ws(c));
    GPB_TRY(ensure(c->Kwork, c->Kwork_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->W, c->W_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->Kinv, c->Kinv_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->partials, c->partials_cap, std::max(trace_partials_size(npad), col_dot_ws_bytes(npad, npad))));
    GPB_TRY(ensure(c->grad_dev, c->grad_cap, sizeof(double) * (nt + 2 + MAX_COMP)));
    CovParams cp;
    MeanParams mp;
    GPB_TRY(make_cov_params(c, theta + c->n_mean, cp));
    make_mean_params(c, theta, mp);
    int info_h = 0;
    // diagnostic: "i8_grad_phases" bit mask (1 potrf, 2 trtri, 4 lauum) -- a cleared bit runs that phase on DMMA
    const int phases = (int)option(OPT_I8_GRAD_PHASES);
    struct PhaseMode {
        int prev;
        bool active;
        PhaseMode(int ph, int bit) : prev(gemm_i8_override()), active(prev < 0 && !(ph & bit)) {
            if (active) set_gemm_i8_override(0);
        }
        ~PhaseMode() {
            if (active) set_gemm_i8_override(prev);
        }
    };
    {
        PhaseMode pm(phases, 1);
        GPB_TRY(assemble_and_factor(c, cp, mp, c->Kwork, c->dinv_work, nullptr, SOLVE_NONE, nullptr, &info_h));
    }
    double sc[3], fro2[MAX_COMP];
    // The triangular inverse: c->W receives W = inv(L) (lower) from the recursion, or Y = W^T (upper) from the blocked
    // substitution -- option "grad_inverse"; the substitution form keeps every product at the scale of L (potrf.cu)
    const bool by_rows = option(OPT_GRAD_INVERSE) != 0;
    auto inverse_and_traces = [&](bool with_trtri) -> int {
        if (with_trtri) {
            c->timer.mark("trtri");
            {
                PhaseMode pm(phases, 2);
                if (by_rows) {
                    GPB_TRY(run_graphed(c, pkey("trtri_rows", {c->Kwork, c->W, c->dinv_work, c->tmp}, {npad}), [&]() -> int {
                        return trtri_rows_lower(c->Kwork, npad, c->W, npad, npad, ws_of(c, c->dinv_work), c->s);
                    }));
                } else {
                    GPB_TRY(run_graphed(c, pkey("trtri", {c->Kwork, c->W, c->Kinv, c->dinv_work, c->tmp}, {npad}), [&]() -> int {
                        GPB_CUDA(cudaMemsetAsync(c->W, 0, sizeof(double) * np * np, c->s));
                        return trtri_lower(c->Kwork, npad, c->W, npad, npad, 0, ws_of(c, c->dinv_work), c->Kinv, npad, c->s);
                    }));
                }
            }
            // alpha = K^-1 r through the explicit inverse, as the reference does here (regression.py:556-559):
            // v = W r, alpha = W^T v -- two fully parallel matrix-vector passes instead of 2 N/128 dependent block steps
            c->timer.mark("alpha");
            if (by_rows) {  // c->W holds W^T
                GPB_TRY(launch_col_dot(c->W, npad, npad, npad, c->resid, c->vec, c->partials, c->s));
                GPB_TRY(launch_row_dot(c->W, npad, npad, npad, c->vec, c->alpha_work, c->s));
            } else {
                GPB_TRY(launch_row_dot(c->W, npad, npad, npad, c->resid, c->vec, c->s));
                GPB_TRY(launch_col_dot(c->W, npad, npad, npad, c->vec, c->alpha_work, c->partials, c->s));
            }
            // LML = -0.5 r.alpha - sum log L_ii   (regression.py:559-560)
            GPB_TRY(launch_logdet_dot(c->Kwork, npad, c->resid, c->alpha_work, n, c->scal, c->s));
        }
        c->timer.mark("lauum");
        {
            PhaseMode pm(phases, 4);
            if (by_rows) GPB_TRY(lauum_rows_lower(c->W, npad, c->Kinv, npad, npad, c->s));
            else GPB_TRY(lauum_lower(c->W, npad, c->Kinv, npad, npad, c->s));
        }
        c->timer.mark("trace");
        GPB_CUDA(cudaMemsetAsync(c->grad_dev, 0, sizeof(double) * (nt + 2 + MAX_COMP), c->s));
        GPB_TRY(launch_lml_grad(cp, mp, c->n_mean, c->x, n, npad, c->alpha_work, c->Kinv, npad, c->partials, c->grad_dev,
                                c->grad_dev + nt + 2, c->s));
        GPB_CUDA(cudaMemcpyAsync(sc, c->scal, sizeof(sc), cudaMemcpyDeviceToHost, c->s));
        GPB_CUDA(cudaMemcpyAsync(grad, c->grad_dev, sizeof(double) * nt, cudaMemcpyDeviceToHost, c->s));
        GPB_CUDA(cudaMemcpyAsync(fro2, c->grad_dev + nt + 2, sizeof(fro2), cudaMemcpyDeviceToHost, c->s));
        c->timer.mark("end");
        GPB_CUDA(cudaStreamSynchronize(c->s));
        return 0;
    };
    const double i8_before = thread_gemm_flops_i8();
    GPB_TRY(inverse_and_traces(true));
    c->grad_guard_est = 0.0;
    if (info_h == 0 && option(OPT_I8_GRAD_GUARD) && gemm_i8_override() < 0 && thread_gemm_flops_i8() > i8_before) {
        double fmax2 = 0.0, gmax = 0.0;
        for (int i = 0; i < c->ncomp; ++i) fmax2 = std::max(fmax2, fro2[i]);
        for (int i = 0; i < nt; ++i) gmax = std::max(gmax, std::fabs(grad[i]));
        const double wmax = 1.0 / sc[2];
        const double est = GUARD_C * 1.3877787807814457e-17 * std::sqrt((double)std::min(n, 4096)) * wmax * wmax * 0.5 * std::sqrt(fmax2);
        const double gscale = std::max(gmax, 0.5 * std::fabs(sc[1]));
        c->grad_guard_est = gscale > 0.0 ? est / gscale : 0.0;
        if (!(est <= GUARD_TOL * gscale)) {  // also taken when anything is NaN
            c->timer.unmark_last();        // the "end" mark: the repeated phases extend this call's timeline
            // only K^-1 = W^T W is repeated: the phase-by-phase measurement (profiles/grad_phase_sensitivity_r2.md) puts
            // the error in that product; potrf and the triangular inverse on the INT8 path stay within 1e-10
            set_gemm_i8_override(0);
            const int rc = inverse_and_traces(false);
            set_gemm_i8_override(-1);
            ++c->grad_guard_retries;
            GPB_TRY(rc);
        }
    }
    *info = info_h;
    *lml = -0.5 * sc[1] - sc[0];
    return 0;
}

static int loo_impl(gpb_ctx* c, const double* theta, double* loo, double* grad, int* info) {
    GPB_TRY(use(c));
    GPB_TRY(need_model(c));
    c->timer.reset();
    const size_t np = (size_t)c->npad;
    const int npad = (int)c->npad, n = (int)c->n, nt = c->n_mean + c->n_cov;
    GPB_TRY(ensure_linalg_ws(c));
    GPB_TRY(ensure(c->Kwork, c->Kwork_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->W, c->W_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->Kinv, c->Kinv_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->grad_dev, c->grad_cap, sizeof(double) * (nt + 2)));
    GPB_TRY(ensure(c->partials, c->partials_cap, std::max(trace_partials_size(npad), sizeof(double) * 5 * np)));
    CovParams cp;
    MeanParams mp;
    GPB_TRY(make_cov_params(c, theta + c->n_mean, cp));
    make_mean_params(c, theta, mp);
    int info_h = 0;
    GPB_TRY(assemble_and_factor(c, cp, mp, c->Kwork, c->dinv_work, nullptr, SOLVE_BOTH, c->alpha_work, &info_h));
    c->timer.mark("trtri");
    GPB_TRY(run_graphed(c, pkey("trtri", {c->Kwork, c->W, c->Kinv, c->dinv_work, c->tmp}, {npad}), [&]() -> int {
        GPB_CUDA(cudaMemsetAsync(c->W, 0, sizeof(double) * np * np, c->s));
        return trtri_lower(c->Kwork, npad, c->W, npad, npad, 0, ws_of(c, c->dinv_work), c->Kinv, npad, c->s);
    }));
    c->timer.mark("lauum");
    GPB_TRY(lauum_lower(c->W, npad, c->Kinv, npad, npad, c->s));
    c->timer.mark("loo");
    // the factor (Kwork) and W are dead from here on: reuse them as the dK plane and the product buffer
    GPB_CUDA(cudaMemsetAsync(c->grad_dev, 0, sizeof(double) * (nt + 2), c->s));
    const double* dK_all = nullptr;
    if (c->n_regions && grad) {  // ChangePoint models: dense gradient planes from the generic kernel (small N by nature)
        const size_t plane = (size_t)n * n;
        GPB_TRY(ensure(c->S, c->S_cap, sizeof(double) * plane * (c->n_cov + 1)));
        GPB_TRY(launch_assemble_grads(cp, c->x, n, c->S, c->S + plane, c->s));
        dK_all = c->S + plane;
    }
    GPB_TRY(launch_loo(cp, mp, c->n_mean, c->x, n, npad, c->alpha_work, c->Kinv, npad, c->partials, c->Kwork, c->W,
                       c->scal, grad ? c->grad_dev : nullptr, dK_all, c->n_cov, c->s));
    double val = 0.0;
    GPB_CUDA(cudaMemcpyAsync(&val, c->scal, sizeof(double), cudaMemcpyDeviceToHost, c->s));
    if (grad) GPB_CUDA(cudaMemcpyAsync(grad, c->grad_dev, sizeof(double) * nt, cudaMemcpyDeviceToHost, c->s));
    c->timer.mark("end");
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    *loo = val;
    return 0;
}

int gpb_factor(gpb_ctx* c, const double* theta, int* info) {
    return with_dmma_retry(c, info, [&]() { return factor_impl(c, theta, info); });
}
int gpb_lml(gpb_ctx* c, const double* theta, double* lml, int* info) {
    return with_dmma_retry(c, info, [&]() { return lml_impl(c, theta, lml, info); });
}
int gpb_lml_grad(gpb_ctx* c, const double* theta, double* lml, double* grad, int* info) {
    return with_dmma_retry(c, info, [&]() { return lml_grad_impl(c, theta, lml, grad, info); });
}
int gpb_loo(gpb_ctx* c, const double* theta, double* loo, double* grad, int* info) {
    // The leave-one-out objective reads diag(K^-1) and K^-1 dK_p entry by entry: it needs the componentwise accuracy of
    // the FP64 DMMA kernels for the inverse chain (see the gradient error guard above), so the whole call runs on them.
    const int prev = gemm_i8_override();
    set_gemm_i8_override(0);
    const int rc = loo_impl(c, theta, loo, grad, info);
    set_gemm_i8_override(prev);
    return rc;
}

int gpb_ctx_stat(gpb_ctx* c, const char* name, double* out) {
    if (!c || !name || !out) {
        set_error("gpb_ctx_stat: bad arguments");
        return -2;
    }
    const std::string k(name);
    if (k == "dmma_retries") *out = (double)c->dmma_retries;
    else if (k == "grad_guard_retries") *out = (double)c->grad_guard_retries;
    else if (k == "grad_guard_est") *out = c->grad_guard_est;
    else if (k == "predict_block") *out = c->pp.valid ? (double)c->pp.nb : 0.0;
    else {
        set_error("gpb_ctx_stat: unknown statistic '" + k + "'");
        return -2;
    }
    return 0;
}

// Population / multistart evaluation: R hyper-parameter vectors, theta r on context r % nctx, one host thread per context
// (the contexts sit on different GPUs and hold the same data and model).  No collective: the units are independent.
int gpb_lml_grad_batch(gpb_ctx** ctxs, int nctx, const double* thetas, int n_theta, int p, double* lml,
                       double* grad_or_null, int* info) {
    if (!ctxs || nctx < 1 || n_theta < 0 || !thetas || !lml || !info) {
        set_error("gpb_lml_grad_batch: bad arguments");
        return -2;
    }
    for (int i = 0; i < nctx; ++i) {
        if (!ctxs[i] || !ctxs[i]->model_set || ctxs[i]->n_mean + ctxs[i]->n_cov != p) {
            set_error("gpb_lml_grad_batch: every context needs data, a model and p = n_mean + n_cov hyper-parameters");
            return -2;
        }
    }
    std::vector<int> rc(nctx, 0);
    std::vector<std::string> err(nctx);
    auto work = [&](int i) {
        for (int r = i; r < n_theta && rc[i] == 0; r += nctx) {
            const double* th = thetas + (size_t)r * p;
            rc[i] = grad_or_null ? gpb_lml_grad(ctxs[i], th, lml + r, grad_or_null + (size_t)r * p, info + r)
                                 : gpb_lml(ctxs[i], th, lml + r, info + r);
        }
        if (rc[i] != 0) err[i] = gpb_last_error();  // the message is thread-local: carry it to the caller
    };
    if (nctx == 1) {
        work(0);
    } else {
        std::vector<std::thread> threads;
        for (int i = 0; i < nctx; ++i) threads.emplace_back(work, i);
        for (auto& t : threads) t.join();
    }
    for (int i = 0; i < nctx; ++i)
        if (rc[i] != 0) {
            set_error(err[i]);
            return rc[i];
        }
    return 0;
}

int gpb_loo_predictions(gpb_ctx* c, double* mu, double* sigma) {
    GPB_TRY(use(c));
    if (!c->fitted) {
        set_error("gpb_loo_predictions: no fitted state (call gpb_factor)");
        return -2;
    }
    const size_t np = (size_t)c->npad;
    const int npad = (int)c->npad, n = (int)c->n;
    GPB_TRY(ensure(c->W, c->W_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->Kinv, c->Kinv_cap, sizeof(double) * np * np));
    GPB_TRY(ensure(c->partials, c->partials_cap, std::max(trace_partials_size(npad), sizeof(double) * 5 * np)));
    GPB_CUDA(cudaMemsetAsync(c->W, 0, sizeof(double) * np * np, c->s));
    GPB_TRY(trtri_lower(c->Lfit, npad, c->W, npad, npad, 0, ws_of(c, c->dinv_fit), c->Kinv, npad, c->s));
    GPB_TRY(lauum_lower(c->W, npad, c->Kinv, npad, npad, c->s));
    double* out = c->partials;
    GPB_TRY(launch_loo_predictions(c->Kinv, npad, c->alpha, c->y, n, out, out + np, c->s));
    GPB_CUDA(cudaMemcpyAsync(mu, out, sizeof(double) * n, cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaMemcpyAsync(sigma, out + np, sizeof(double) * n, cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ predict family
namespace {

enum PredMode { PM_PREDICT, PM_GRADIENT, PM_SPATIAL, PM_EI };

bool is_pure_se(const gpb_ctx* c) { return c->ncomp == 1 && c->kinds[0] == COV_SE && c->n_regions == 0; }

int64_t chunk_rows(const gpb_ctx* c) {
    // rows of the stacked cross-covariance per pass: one full wave of 3 CTAs/SM on the 128-wide leaf
    // solves (222 row tiles x 2 column tiles = 444 CTAs) when N is large, more when the rows are short
    const int64_t base = 222 * 128;
    if (c->npad >= 8192) return base;
    if (c->npad >= 2048) return base * 2;
    return base * 4;
}

// ---- blocked predict solve on cached digit planes --------------------------------------------------------------------------
// X = S L^-T for a chunk of stacked cross-covariance rows, left-looking over block columns of width nb:
//     T_j = S_j - X_{<j} L_{j,<j}^T        one INT8 GEMM, k = j nb, both operands already split
//     X_j = T_j L_jj^-T                     the recursive solve (trsm_right_lt) inside the nb x nb diagonal block
// L's off-diagonal planes are split ONCE per fit; each block of X is split once, right after it is solved.  A GEMM needs
// one scale per operand row over its whole k extent, while the blocks of a row of X appear one at a time -- so X uses an
// a-priori bound instead of the measured row maximum: row 0 of a query's stack is v = L^-1 k_q with |v|^2 <= k(q, q)
// (the posterior variance k(q,q) - |v|^2 is non-negative); rows a >= 1 (SquaredExponential gradient terms) are
// L^-1 (A_a o k_q) with squared norm <= a^2 / l_a^2, the prior variance of df/dx_a.  Values a rounding error beyond the
// bound are clamped by the split kernel.
// Two variants of the diagonal solve.  (fast) X_j = T_j inv(L_jj)^T as one INT8 product with the explicit inverse of the
// block, split once per fit.  The entries of inv(L_jj) grow like w = 1 / min L_ii, the digit-split product is accurate
// only normwise (relative to the row maxima), and sigma^2 = k(q,q) - |v|^2 cancels down to ~ k(q,q) / w^2 at the training
// points, so the relative error of sigma^2 is about 2^-56 sqrt(nb) w^3 (w in units of the prior sigma): measured on the
// conditioning sweep this variant lost a factor 5 (w = 10) to 1e5 (w = 1e4) against the FP64 path in sigma, and on a
// dense data set (SE 2-D, N = 32768, sigma^2 / k(q,q) down to 2e-6) a factor 4 even at w = 22
// (profiles/predict_variants_N32768_r2.json).  It is therefore NOT the default; option "predict_diag" = 2 selects it,
// = 0 selects it only while PREDICT_W_C w^3 <= PREDICT_W_TOL (w <~ 30).  (robust, default) the recursive solve inside the
// block, which multiplies by inverted 128-blocks on the FP64 tensor pipe and uses the INT8 path only for its benign
// X L21^T updates; it matches the FP64 path over the whole sweep (profiles/conditioning_sweep_r2.md) and costs ~9 % more
// time in the solve.
// Replaces the full-width recursion, which re-split L for every chunk and level (round-1 profile: 1108 split launches
// per step, 6.6 % of the GPU time) and ran 98 % of its flops at k <= N/2 GEMMs of decreasing size.
constexpr double X_BOUND_HEADROOM = 1.0 + 1e-6;
constexpr double PREDICT_W_C = 2.0 * 1.3877787807814457e-17 * 64.0, PREDICT_W_TOL = 5e-11;

bool predict_fast_diag(const gpb_ctx* c) {
    if (option(OPT_PREDICT_DIAG) == 1) return false;   // 0 = by conditioning, 1 = always robust, 2 = always fast
    if (option(OPT_PREDICT_DIAG) == 2) return true;
    double amp2 = 0.0;
    for (int i = 0; i < c->ncomp; ++i)
        if (c->kinds[i] <= COV_RQ) amp2 += c->cp_fit.amp2[i];
    if (!(c->min_pivot_fit > 0.0) || !(amp2 > 0.0)) return false;
    const double w = std::sqrt(amp2) / c->min_pivot_fit;
    return PREDICT_W_C * w * w * w <= PREDICT_W_TOL;
}

int predict_block_width(const gpb_ctx* c, int rows_pad) {
    const int mode = gemm_i8_override() >= 0 ? gemm_i8_override() : (int)option(OPT_GEMM_I8);
    // option "predict_block": -1 = off (full-width recursion), 0 = by variant (measured on the benchmark step: 4096 is the
    // fastest width with the INT8 diagonal solve, 1024 with the recursive one), else the width itself
    int want = (int)option(OPT_PREDICT_BLOCK);
    if (want == 0) want = predict_fast_diag(c) ? 4096 : 1024;
    if (mode == 0 || want < NB) return 0;
    const int npad = (int)c->npad;
    const int quarter = (npad / 4) / NB * NB;
    if (mode == 2) {  // forced INT8 (tests): any size the kernel's shape rules allow
        if (npad < 512) return 0;
        return std::min(want / NB * NB, std::max(256, quarter));
    }
    const int nb = std::min(want / NB * NB, std::max(512, quarter));
    if (npad < 2 * nb || nb < (int)option(OPT_GEMM_I8_MIN_K)) return 0;
    if ((int64_t)(rows_pad / NB) * (nb / NB) < 148) return 0;  // too few tiles to fill the machine: recursion on DMMA
    return nb;
}

int build_predict_planes(gpb_ctx* c, int nb, int ns, int64_t rows_cap) {
    PredictPlanes& pp = c->pp;
    const int npad = (int)c->npad;
    const int nblk = (npad + nb - 1) / nb;
    auto cols_of = [&](int j) { return std::min(nb, npad - j * nb); };
    const bool use_w = predict_fast_diag(c);
    if (!(pp.valid && pp.nb == nb && pp.use_w == use_w)) {
        pp.valid = false;
        size_t lbytes = 0, wbytes = 0;
        for (int j = 0; j < nblk; ++j) {
            lbytes += i8_plane_bytes(cols_of(j), (int64_t)j * nb);
            wbytes += i8_plane_bytes(cols_of(j), cols_of(j));
        }
        GPB_TRY(ensure(pp.lbuf, pp.lbuf_cap, std::max<size_t>(lbytes, 16)));
        GPB_TRY(ensure(pp.lscale, pp.lscale_cap, sizeof(double) * npad));
        pp.Lp.assign(nblk, I8Planes());
        pp.Wp.assign(nblk, I8Planes());
        if (use_w) {
            GPB_TRY(ensure(pp.wbuf, pp.wbuf_cap, wbytes));
            GPB_TRY(ensure(pp.wscale, pp.wscale_cap, sizeof(double) * npad));
            GPB_TRY(ensure(pp.wtmp, pp.wtmp_cap, sizeof(double) * 2 * (size_t)nb * nb));
        }
        size_t lo = 0, wo = 0;
        const LinalgWs ws = ws_of(c, c->dinv_fit);
        for (int j = 0; j < nblk; ++j) {
            const int k0 = j * nb, nbj = cols_of(j);
            I8Planes& L = pp.Lp[j];
            L.q = pp.lbuf + lo;
            L.scale = pp.lscale + k0;
            L.rows = nbj;
            L.ld = k0;
            L.plane = (int64_t)nbj * k0;
            lo += i8_plane_bytes(nbj, k0);
            if (j > 0) GPB_TRY(i8_split_rows(c->Lfit + (size_t)k0 * npad, npad, nbj, k0, false, nullptr, 1, L, 0, 0, c->s));
            if (!use_w) continue;
            I8Planes& W = pp.Wp[j];
            W.q = pp.wbuf + wo;
            W.scale = pp.wscale + k0;
            W.rows = nbj;
            W.ld = nbj;
            W.plane = (int64_t)nbj * nbj;
            wo += i8_plane_bytes(nbj, nbj);
            // inv(L_jj) in FP64 (recursion on the inverted 128-blocks), then its planes
            double* Wd = pp.wtmp;
            GPB_CUDA(cudaMemsetAsync(Wd, 0, sizeof(double) * (size_t)nbj * nbj, c->s));
            GPB_TRY(trtri_lower(c->Lfit + (size_t)k0 * npad + k0, npad, Wd, nbj, nbj, k0 / NB, ws, pp.wtmp + (size_t)nb * nb, nb,
                                c->s));
            GPB_TRY(i8_split_rows(Wd, nbj, nbj, nbj, false, nullptr, 1, W, 0, 0, c->s));
        }
        pp.nb = nb;
        pp.nblk = nblk;
        pp.use_w = use_w;
        pp.valid = true;
    }
    // per-chunk planes of the solved columns (and, fast variant, of the block before its diagonal solve)
    GPB_TRY(ensure(pp.xbuf, pp.xbuf_cap, i8_plane_bytes(rows_cap, npad)));
    GPB_TRY(ensure(pp.xscale, pp.xscale_cap, sizeof(double) * rows_cap));
    if (use_w) {
        GPB_TRY(ensure(pp.tbuf, pp.tbuf_cap, i8_plane_bytes(rows_cap, nb)));
        GPB_TRY(ensure(pp.tscale, pp.tscale_cap, sizeof(double) * rows_cap));
        pp.Tp.q = pp.tbuf;
        pp.Tp.scale = pp.tscale;
        pp.Tp.rows = rows_cap;
        pp.Tp.ld = nb;
        pp.Tp.plane = rows_cap * (int64_t)nb;
    }
    pp.Xp.q = pp.xbuf;
    pp.Xp.scale = pp.xscale;
    pp.Xp.rows = rows_cap;
    pp.Xp.ld = npad;
    pp.Xp.plane = rows_cap * (int64_t)npad;
    if (pp.ns != ns || !pp.bound) {  // a-priori bounds of the rows of a query's stack
        GPB_TRY(ensure(pp.bound, pp.bound_cap, sizeof(double) * (MAX_DIM + 1)));
        double b[MAX_DIM + 1] = {0};
        double kqq = 0.0;
        for (int i = 0; i < c->ncomp; ++i)
            if (c->kinds[i] <= COV_RQ) kqq += c->cp_fit.amp2[i];
        b[0] = X_BOUND_HEADROOM * std::sqrt(kqq);
        for (int a = 1; a < ns; ++a) b[a] = X_BOUND_HEADROOM * std::sqrt(c->cp_fit.amp2[0] * c->cp_fit.inv_l2[0][a - 1]);
        GPB_CUDA(cudaMemcpyAsync(pp.bound, b, sizeof(b), cudaMemcpyHostToDevice, c->s));
        GPB_CUDA(cudaStreamSynchronize(c->s));  // b is a stack array
        pp.ns = ns;
    }
    return 0;
}

int predict_solve_blocked(gpb_ctx* c, int rows_pad, int ns) {
    PredictPlanes& pp = c->pp;
    const int npad = (int)c->npad, nb = pp.nb;
    const LinalgWs ws = ws_of(c, c->dinv_fit);
    for (int j = 0; j < pp.nblk; ++j) {
        const int k0 = j * nb, nbj = std::min(nb, npad - k0);
        double* Sj = c->S + k0;
        if (j > 0)
            GPB_TRY(i8_gemm_planes(pp.Xp, 0, pp.Lp[j], 0, rows_pad, nbj, k0, Sj, npad, Sj, npad, -1.0, 1.0, GEMM_FULL, c->s));
        if (pp.use_w) {
            GPB_TRY(i8_split_rows(Sj, npad, rows_pad, nbj, true, nullptr, 1, pp.Tp, 0, 0, c->s));
            GPB_TRY(i8_gemm_planes(pp.Tp, 0, pp.Wp[j], 0, rows_pad, nbj, nbj, nullptr, 0, Sj, npad, 1.0, 0.0, GEMM_TRIL_B, c->s));
        } else {
            GPB_TRY(trsm_right_lt(Sj, npad, rows_pad, c->Lfit + (size_t)k0 * npad + k0, npad, nbj, k0 / NB, ws, c->s));
        }
        if (j + 1 < pp.nblk) GPB_TRY(i8_split_rows(Sj, npad, rows_pad, nbj, true, pp.bound, ns, pp.Xp, 0, k0, c->s));
    }
    return 0;
}

// Shared driver: out pointers are DEVICE pointers sized for all m queries.
//   PM_PREDICT : o_a = mu (m), o_b = sig (m)
//   PM_GRADIENT: o_a = mean (m x d), o_b = cov (m x d x d)
//   PM_SPATIAL : o_a = dmu (m x d), o_b = dvar (m x d)
//   PM_EI      : o_a = acquisition value (m), o_b = grad (m x d) or nullptr; ei_mode = GPB_EI_* mode, acq_kind / acq_param
int predict_driver(gpb_ctx* c, const double* q_dev, int64_t m, PredMode mode, double* o_a, double* o_b, int ei_mode,
                   double acq_param, int acq_kind = GPB_ACQ_EI) {
    const int npad = (int)c->npad, n = (int)c->n, d = c->d;
    const bool stacked = mode == PM_GRADIENT || mode == PM_SPATIAL || (mode == PM_EI && ei_mode == GPB_EI_NEG_LOG_GRAD);
    const int ns = stacked ? d + 1 : 1;
    const int64_t rc = chunk_rows(c);
    const int64_t qc = std::max<int64_t>(1, rc / ns);  // queries per chunk
    const int64_t qmax = std::min<int64_t>(qc, m);
    const int64_t rows_cap = round_up(qmax * ns, 256);
    GPB_TRY(ensure(c->S, c->S_cap, sizeof(double) * (size_t)rows_cap * npad));
    const int nb_solve = predict_block_width(c, (int)rows_cap);
    if (nb_solve) {
        c->timer.mark("planes");
        GPB_TRY(build_predict_planes(c, nb_solve, ns, rows_cap));
    }
    GPB_TRY(ensure(c->dots, c->dots_cap, sizeof(double) * (size_t)rows_cap));
    GPB_TRY(ensure(c->G, c->G_cap, sizeof(double) * (size_t)qmax * ns * ns));
    if (mode == PM_EI) {
        GPB_TRY(ensure(c->o1, c->o1_cap, sizeof(double) * (size_t)qmax));          // mu
        GPB_TRY(ensure(c->o2, c->o2_cap, sizeof(double) * (size_t)qmax));          // sig
        if (stacked) GPB_TRY(ensure(c->o3, c->o3_cap, sizeof(double) * (size_t)qmax * d * 2));  // dmu | dvar
    }
    if (mode == PM_GRADIENT) {  // R = (a / l)^2  (covariance.py:266)
        GPB_TRY(ensure(c->R_dev, c->R_cap, sizeof(double) * MAX_DIM));
        double R[MAX_DIM];
        for (int k = 0; k < d; ++k) R[k] = c->cp_fit.amp2[0] * c->cp_fit.inv_l2[0][k];
        GPB_CUDA(cudaMemcpyAsync(c->R_dev, R, sizeof(double) * d, cudaMemcpyHostToDevice, c->s));
        GPB_CUDA(cudaStreamSynchronize(c->s));
    }
    const LinalgWs ws = ws_of(c, c->dinv_fit);
    for (int64_t q0 = 0; q0 < m; q0 += qc) {
        const int mq = (int)std::min<int64_t>(qc, m - q0);
        const int rows = mq * ns;
        const bool blocked = nb_solve && predict_block_width(c, (int)round_up(rows, 256)) == nb_solve;
        const int rows_pad = (int)round_up(rows, blocked ? 256 : 128);
        const double* qp = q_dev + q0 * d;
        c->timer.mark("cross_cov");
        GPB_TRY(launch_cross_stack(c->cp_fit, qp, mq, ns, c->x, n, npad, c->S, npad, c->s));
        if (rows_pad > rows)
            GPB_CUDA(cudaMemsetAsync(c->S + (size_t)rows * npad, 0, sizeof(double) * (size_t)(rows_pad - rows) * npad,
                                     c->s));
        c->timer.mark("mean_dot");
        GPB_TRY(launch_row_dot(c->S, npad, rows, npad, c->alpha, c->dots, c->s));
        c->timer.mark("trsm");
        if (blocked) {
            GPB_TRY(run_graphed(c, pkey("pblk", {c->S, c->pp.lbuf, c->pp.xbuf, c->pp.wbuf, c->pp.tbuf, c->Lfit, c->dinv_fit, c->tmp},
                                     {npad, rows_pad, nb_solve, ns, c->pp.use_w}),
                                [&]() { return predict_solve_blocked(c, rows_pad, ns); }));
        } else {
            GPB_TRY(run_graphed(c, pkey("ptrsm", {c->S, c->Lfit, c->dinv_fit, c->tmp}, {npad, rows_pad}),
                                [&]() { return trsm_right_lt(c->S, npad, rows_pad, c->Lfit, npad, npad, 0, ws, c->s); }));
        }
        c->timer.mark("gram");
        GPB_TRY(launch_row_gram(c->S, npad, mq, ns, npad, c->G, c->s));
        c->timer.mark("finalize");
        switch (mode) {
            case PM_PREDICT:
                GPB_TRY(launch_finalize_predict(c->cp_fit, c->mp_fit, qp, mq, ns, c->dots, c->G, o_a + q0, o_b + q0, c->s));
                break;
            case PM_GRADIENT:
                GPB_TRY(launch_finalize_gradient(c->dots, c->G, mq, d, c->R_dev, o_a + q0 * d, o_b + q0 * d * d, c->s));
                break;
            case PM_SPATIAL:
                GPB_TRY(launch_finalize_spatial(c->dots, c->G, mq, d, o_a + q0 * d, o_b + q0 * d, c->s));
                break;
            case PM_EI:
                GPB_TRY(launch_finalize_predict(c->cp_fit, c->mp_fit, qp, mq, ns, c->dots, c->G, c->o1, c->o2, c->s));
                if (stacked)
                    GPB_TRY(launch_finalize_spatial(c->dots, c->G, mq, d, c->o3, c->o3 + (size_t)qmax * d, c->s));
                GPB_TRY(launch_acquisition(c->o1, c->o2, stacked ? c->o3 : nullptr,
                                           stacked ? c->o3 + (size_t)qmax * d : nullptr, mq, d, acq_kind, acq_param,
                                           ei_mode, o_a + q0, stacked ? o_b + q0 * d : nullptr, c->s));
                break;
        }
    }
    c->timer.mark("end");
    return 0;
}

int need_fit(gpb_ctx* c) {
    if (!c->fitted) {
        set_error("no fitted state: call gpb_factor first");
        return -2;
    }
    return 0;
}

// host-buffer wrapper: uploads q, runs the driver into device outputs, downloads na/nb doubles per query
int predict_host(gpb_ctx* c, const double* q, int64_t m, PredMode mode, double* a, int64_t na, double* b, int64_t nb,
                 int ei_mode = 0, double y_max = 0.0, int acq_kind = GPB_ACQ_EI, int64_t* argbest = nullptr) {
    if (m == 0) return 0;
    c->timer.reset();
    c->timer.mark("h2d");
    GPB_TRY(ensure(c->qbuf, c->qbuf_cap, sizeof(double) * (size_t)m * (c->d + na + nb)));
    double* qd = c->qbuf;
    double* ad = qd + (size_t)m * c->d;
    double* bd = ad + (size_t)m * na;
    GPB_CUDA(cudaMemcpyAsync(qd, q, sizeof(double) * m * c->d, cudaMemcpyHostToDevice, c->s));
    GPB_TRY(predict_driver(c, qd, m, mode, ad, nb ? bd : nullptr, ei_mode, y_max, acq_kind));
    if (argbest) {  // best candidate on the device: largest value, i.e. smallest opt_func
        c->timer.marks.back().first = "argbest";
        GPB_TRY(ensure(c->argws, c->argws_cap, (sizeof(double) + sizeof(int64_t)) * (ARGBEST_BLOCKS + 1)));
        double* wv = reinterpret_cast<double*>(c->argws);
        int64_t* wi = reinterpret_cast<int64_t*>(wv + ARGBEST_BLOCKS + 1);
        GPB_TRY(launch_argbest(ad, m, ei_mode == GPB_EI_VALUE, wv, wi, c->s));
        GPB_CUDA(cudaMemcpyAsync(argbest, wi, sizeof(int64_t), cudaMemcpyDeviceToHost, c->s));
        c->timer.mark("end");
    }
    c->timer.marks.back().first = "d2h";
    GPB_CUDA(cudaMemcpyAsync(a, ad, sizeof(double) * m * na, cudaMemcpyDeviceToHost, c->s));
    if (nb && b) GPB_CUDA(cudaMemcpyAsync(b, bd, sizeof(double) * m * nb, cudaMemcpyDeviceToHost, c->s));
    c->timer.mark("end");
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}

}  // namespace

extern "C" {

int gpb_predict(gpb_ctx* c, const double* q, int64_t m, double* mu, double* sig) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    return predict_host(c, q, m, PM_PREDICT, mu, 1, sig, 1);
}

int gpb_predict_dev(gpb_ctx* c, const double* q_dev, int64_t m, double* mu_dev, double* sig_dev) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    if (m == 0) return 0;
    c->timer.reset();
    GPB_TRY(predict_driver(c, q_dev, m, PM_PREDICT, mu_dev, sig_dev, 0, 0.0));
    return 0;  // asynchronous on the context stream; gpb_sync() to wait
}

int gpb_gradient(gpb_ctx* c, const double* q, int64_t m, double* mean, double* cov) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    if (!is_pure_se(c)) {
        set_error("gradient terms are only available for the SquaredExponential covariance (covariance.py:38-44)");
        return -3;
    }
    return predict_host(c, q, m, PM_GRADIENT, mean, c->d, cov, (int64_t)c->d * c->d);
}

int gpb_spatial_derivatives(gpb_ctx* c, const double* q, int64_t m, double* dmu, double* dvar) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    if (!is_pure_se(c)) {
        set_error("gradient terms are only available for the SquaredExponential covariance (covariance.py:38-44)");
        return -3;
    }
    return predict_host(c, q, m, PM_SPATIAL, dmu, c->d, dvar, c->d);
}

int gpb_acquisition(gpb_ctx* c, int kind, double param, const double* q, int64_t m, int mode, double* out,
                    double* grad_or_null, int64_t* argbest_or_null) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    if (kind < GPB_ACQ_EI || kind > GPB_ACQ_MAXVAR || mode < GPB_EI_VALUE || mode > GPB_EI_NEG_LOG_GRAD) {
        set_error("gpb_acquisition: unknown acquisition kind or mode");
        return -2;
    }
    if (mode == GPB_EI_NEG_LOG_GRAD && !is_pure_se(c)) {
        set_error("gradient terms are only available for the SquaredExponential covariance (covariance.py:38-44)");
        return -3;
    }
    const bool g = mode == GPB_EI_NEG_LOG_GRAD;
    if (argbest_or_null) *argbest_or_null = -1;
    GPB_TRY(predict_host(c, q, m, PM_EI, out, 1, g ? grad_or_null : nullptr, g ? c->d : 0, mode, param, kind,
                         (argbest_or_null && m > 0) ? argbest_or_null : nullptr));
    if (argbest_or_null && m > 0 && *argbest_or_null < 0) *argbest_or_null = 0;  // all NaN: the host scan's answer
    return 0;
}

int gpb_expected_improvement(gpb_ctx* c, const double* q, int64_t m, double y_max, int mode, double* out,
                             double* grad_or_null, int64_t* argmax_or_null) {
    return gpb_acquisition(c, GPB_ACQ_EI, y_max, q, m, mode, out, grad_or_null, argmax_or_null);
}

// ---- incremental append (GpOptimiser.add_evaluation with fixed hyper-parameters) -----------------------------------------
// The reference re-builds the regressor from scratch for every added evaluation (optimisation.py:177-186): O(N^3).  With
// the hyper-parameters kept, adding one training point only appends a row to the factor:
//     l = L^-1 k(x_new, X),   d = sqrt(k(x_new, x_new) + diagonal terms - l.l),   L <- [[L, 0], [l^T, d]]
// one forward substitution (O(N^2)); the inverted 128-block of the diagonal gets its new row from the rows above it
// (rows of a lower-triangular inverse depend only on earlier rows), and alpha is recomputed from the new factor because
// the mean (x-bar of Linear / Quadratic means) and the residual change for every point.
__global__ void __launch_bounds__(128) append_row_kernel(double* __restrict__ L, int64_t ld, double* __restrict__ dinv_blk,
                                                         const double* __restrict__ l, int n, double diag_base,
                                                         int* __restrict__ info) {
    __shared__ double red[4];
    __shared__ double dval;
    const int tid = threadIdx.x, r = n % NB, b0 = (n / NB) * NB;
    double acc = 0.0;
    for (int k = tid; k < n; k += 128) acc = fma(l[k], l[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        const double d2 = diag_base - ((red[0] + red[1]) + (red[2] + red[3]));
        if (!(d2 > 0.0)) {
            *info = n + 1;
            dval = 0.0;
        } else {
            dval = sqrt(d2);
        }
    }
    __syncthreads();
    const double d = dval;
    if (d == 0.0) return;
    double* row = L + (int64_t)n * ld;
    for (int k = tid; k < n; k += 128) row[k] = l[k];
    if (tid == 0) row[n] = d;
    // new row r of inv(L_bb): W[r][c] = -(1/d) sum_{k=c}^{r-1} L[n][b0+k] W[k][c], W[r][r] = 1/d
    if (tid < NB) {
        double w = 0.0;
        if (tid < r) {
            for (int k = tid; k < r; ++k) w = fma(l[b0 + k], dinv_blk[k * NB + tid], w);
            w = -w / d;
        } else if (tid == r) {
            w = 1.0 / d;
        }
        dinv_blk[r * NB + tid] = w;
    }
}

// grows every per-point buffer of the context to a new padded size (the factor keeps its contents; new rows / columns are
// the identity padding)
static int grow_padding(gpb_ctx* c, int64_t npad_new) {
    const int64_t np0 = c->npad, d = c->d;
    linv_destroy(c);  // a linear-inversion problem attached to this context is padded to the old size
    auto regrow = [&](double*& p, size_t new_doubles, size_t keep_doubles) -> int {
        double* q = nullptr;
        GPB_CUDA(cudaMalloc(&q, sizeof(double) * new_doubles));
        GPB_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * new_doubles, c->s));
        if (p) {
            GPB_CUDA(cudaMemcpyAsync(q, p, sizeof(double) * keep_doubles, cudaMemcpyDeviceToDevice, c->s));
            GPB_CUDA(cudaStreamSynchronize(c->s));
            GPB_CUDA(cudaFree(p));
        }
        p = q;
        return 0;
    };
    GPB_TRY(regrow(c->x, (size_t)npad_new * d, (size_t)np0 * d));
    GPB_TRY(regrow(c->y, (size_t)npad_new, (size_t)np0));
    if (c->has_noise) GPB_TRY(regrow(c->noise, (size_t)npad_new, (size_t)np0));
    double* Lnew = nullptr;
    GPB_CUDA(cudaMalloc(&Lnew, sizeof(double) * (size_t)npad_new * npad_new));
    GPB_CUDA(cudaMemsetAsync(Lnew, 0, sizeof(double) * (size_t)npad_new * npad_new, c->s));
    GPB_CUDA(cudaMemcpy2DAsync(Lnew, sizeof(double) * npad_new, c->Lfit, sizeof(double) * np0, sizeof(double) * np0, np0,
                               cudaMemcpyDeviceToDevice, c->s));
    std::vector<double> eye((size_t)NB * NB, 0.0);
    for (int i = 0; i < NB; ++i) eye[(size_t)i * NB + i] = 1.0;
    GPB_CUDA(cudaMemcpy2DAsync(Lnew + (size_t)np0 * npad_new + np0, sizeof(double) * npad_new, eye.data(), sizeof(double) * NB,
                               sizeof(double) * NB, NB, cudaMemcpyHostToDevice, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    GPB_CUDA(cudaFree(c->Lfit));
    c->Lfit = Lnew;
    c->Lfit_cap = sizeof(double) * (size_t)npad_new * npad_new;
    GPB_TRY(regrow(c->dinv_fit, (size_t)npad_new * NB, (size_t)np0 * NB));
    c->dinv_fit_cap = sizeof(double) * (size_t)npad_new * NB;
    GPB_CUDA(cudaMemcpyAsync(c->dinv_fit + (size_t)np0 * NB, eye.data(), sizeof(double) * NB * NB, cudaMemcpyHostToDevice, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    GPB_TRY(regrow(c->alpha, (size_t)npad_new, (size_t)np0));
    c->alpha_cap = sizeof(double) * (size_t)npad_new;
    GPB_TRY(regrow(c->mu, (size_t)npad_new, (size_t)np0));
    c->mu_cap = sizeof(double) * (size_t)npad_new;
    // objective / solve workspaces are re-created at their next use
    for (double** p : {&c->dinv_work, &c->vec, &c->resid, &c->alpha_work, &c->tmp, &c->Kwork, &c->W, &c->Kinv})
        if (*p) {
            GPB_CUDA(cudaFree(*p));
            *p = nullptr;
        }
    c->dinv_work_cap = c->vec_cap = c->resid_cap = c->alpha_work_cap = c->tmp_cap = c->Kwork_cap = c->W_cap = c->Kinv_cap = 0;
    c->npad = npad_new;
    return 0;
}

int gpb_append_point(gpb_ctx* c, const double* x_new, double y_new, double noise_var_new, int* info) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    if (c->has_ycov) {
        set_error("gpb_append_point: not available with a dense y_cov");
        return -2;
    }
    for (int i = 0; i < c->ncomp; ++i)
        if (c->kinds[i] == COV_HETERO) {
            set_error("gpb_append_point: HeteroscedasticNoise has one hyper-parameter per point; re-fit instead");
            return -2;
        }
    if (c->n_regions) {
        set_error("gpb_append_point: ChangePoint kernels are not supported; re-fit instead");
        return -2;
    }
    clear_graphs(c);            // captured sequences bake in pointers and sizes
    c->pp.valid = false;        // cached digit planes of L
    const int64_t n = c->n;
    if (n == c->npad) GPB_TRY(grow_padding(c, c->npad + NB));
    const int npad = (int)c->npad, d = c->d;
    GPB_TRY(ensure_linalg_ws(c));
    // the new point and its data
    GPB_CUDA(cudaMemcpyAsync(c->x + (size_t)n * d, x_new, sizeof(double) * d, cudaMemcpyHostToDevice, c->s));
    GPB_CUDA(cudaMemcpyAsync(c->y + n, &y_new, sizeof(double), cudaMemcpyHostToDevice, c->s));
    if (c->has_noise) GPB_CUDA(cudaMemcpyAsync(c->noise + n, &noise_var_new, sizeof(double), cudaMemcpyHostToDevice, c->s));
    for (int k = 0; k < d; ++k) c->xbar[k] = (c->xbar[k] * (double)n + x_new[k]) / (double)(n + 1);
    for (int k = 0; k < d; ++k) c->mp_fit.xbar[k] = c->xbar[k];
    // l = L^-1 k(x_new, X): cross-covariance row into the right-hand side half of `vec`, forward substitution
    GPB_TRY(launch_cross_stack(c->cp_fit, c->x + (size_t)n * d, 1, 1, c->x, (int)n, npad, c->vec, npad, c->s));
    GPB_TRY(trsv_lower_fwd(c->Lfit, npad, npad, c->dinv_fit, c->vec, c->s));
    double diag = c->has_noise ? noise_var_new : 0.0;
    for (int i = 0; i < c->ncomp; ++i) {
        if (c->kinds[i] <= COV_RQ) diag += c->cp_fit.amp2[i] * (1.0 + c->cp_fit.jitter);
        else if (c->kinds[i] == COV_WHITE) diag += c->cp_fit.amp2[i];
    }
    GPB_CUDA(cudaMemsetAsync(c->info_dev, 0, sizeof(int), c->s));
    append_row_kernel<<<1, 128, 0, c->s>>>(c->Lfit, npad, c->dinv_fit + (size_t)(n / NB) * NB * NB, c->vec + npad, (int)n, diag,
                                           c->info_dev);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    int info_h = 0;
    GPB_CUDA(cudaMemcpyAsync(&info_h, c->info_dev, sizeof(int), cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    *info = info_h;
    if (info_h != 0) {   // not positive definite with the new point: the fitted state of the old n points stays valid
        return 0;
    }
    c->n = n + 1;
    // alpha = K^-1 (y - mu) with the enlarged factor (every residual changes when x-bar moves)
    GPB_TRY(launch_residual(c->mp_fit, c->x, c->y, (int)c->n, npad, c->vec, c->mu, c->s));
    GPB_TRY(trsv_lower_fwd(c->Lfit, npad, npad, c->dinv_fit, c->vec, c->s));
    GPB_CUDA(cudaMemcpyAsync(c->vec, c->vec + npad, sizeof(double) * npad, cudaMemcpyDeviceToDevice, c->s));
    GPB_TRY(trsv_lower_bwd(c->Lfit, npad, npad, c->dinv_fit, c->vec, c->s));
    GPB_CUDA(cudaMemcpyAsync(c->alpha, c->vec + npad, sizeof(double) * npad, cudaMemcpyDeviceToDevice, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}

int gpb_posterior(gpb_ctx* c, const double* q, int64_t m, double* mu, double* sigma) {
    GPB_TRY(use(c));
    GPB_TRY(need_fit(c));
    if (m == 0) return 0;
    c->timer.reset();
    const int npad = (int)c->npad, n = (int)c->n, d = c->d;
    const int mp = (int)round_up(m, 128);
    GPB_TRY(ensure(c->S, c->S_cap, sizeof(double) * (size_t)mp * npad));
    GPB_TRY(ensure(c->dots, c->dots_cap, sizeof(double) * (size_t)mp));
    // + 1: the covariance block is moved up by one double when m (d + 1) is odd (16-byte alignment below)
    GPB_TRY(ensure(c->qbuf, c->qbuf_cap, sizeof(double) * ((size_t)m * (d + 1) + 1 + (size_t)mp * mp)));
    double* qd = c->qbuf;
    double* mud = qd + (size_t)m * d;
    double* sg = mud + m;
    if ((reinterpret_cast<uintptr_t>(sg) & 15) != 0) sg += 1;  // 16-byte alignment for the GEMM epilogue
    GPB_CUDA(cudaMemcpyAsync(qd, q, sizeof(double) * m * d, cudaMemcpyHostToDevice, c->s));
    GPB_TRY(launch_cross_stack(c->cp_fit, qd, (int)m, 1, c->x, n, npad, c->S, npad, c->s));
    if (mp > m) GPB_CUDA(cudaMemsetAsync(c->S + (size_t)m * npad, 0, sizeof(double) * (size_t)(mp - m) * npad, c->s));
    GPB_TRY(launch_row_dot(c->S, npad, (int)m, npad, c->alpha, c->dots, c->s));
    GPB_TRY(launch_finalize_predict(c->cp_fit, c->mp_fit, qd, (int)m, 1, c->dots, nullptr, mud, nullptr, c->s));
    GPB_CUDA(cudaMemcpyAsync(mu, mud, sizeof(double) * m, cudaMemcpyDeviceToHost, c->s));
    if (sigma) {
        GPB_TRY(trsm_right_lt(c->S, npad, mp, c->Lfit, npad, npad, 0, ws_of(c, c->dinv_fit), c->s));
        GPB_CUDA(cudaMemsetAsync(sg, 0, sizeof(double) * (size_t)mp * mp, c->s));
        GPB_TRY(launch_cross_cov(c->cp_fit, qd, (int)m, qd, (int)m, sg, mp, c->s));
        GemmArgs g{mp, mp, npad, c->S, npad, c->S, npad, sg, mp, sg, mp, nullptr, 0, -1.0, 1.0, GEMM_FULL};
        GPB_TRY(gemm_nt(g, c->s));
        GPB_CUDA(cudaMemcpy2DAsync(sigma, sizeof(double) * m, sg, sizeof(double) * mp, sizeof(double) * m, m,
                                   cudaMemcpyDeviceToHost, c->s));
    }
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}

int gpb_timers(gpb_ctx* c, char* name_buf, int name_buf_len, double* ms, int max_entries, int* n_entries) {
    GPB_TRY(use(c));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    std::vector<std::pair<std::string, double>> acc;
    auto& mk = c->timer.marks;
    for (size_t i = 0; i + 1 < mk.size(); ++i) {
        float t = 0.f;
        GPB_CUDA(cudaEventElapsedTime(&t, mk[i].second, mk[i + 1].second));
        bool found = false;
        for (auto& a : acc)
            if (a.first == mk[i].first) {
                a.second += t;
                found = true;
                break;
            }
        if (!found) acc.emplace_back(mk[i].first, (double)t);
    }
    std::string names;
    int k = 0;
    for (auto& a : acc) {
        if (k >= max_entries) break;
        if (k) names += ";";
        names += a.first;
        ms[k++] = a.second;
    }
    if ((int)names.size() + 1 > name_buf_len) {
        set_error("gpb_timers: name buffer too small");
        return -2;
    }
    std::memcpy(name_buf, names.c_str(), names.size() + 1);
    *n_entries = k;
    return 0;
}

int gpb_dev_alloc(gpb_ctx* c, int64_t n_doubles, double** out_dev) {
    GPB_TRY(use(c));
    GPB_CUDA(cudaMalloc(out_dev, sizeof(double) * (size_t)n_doubles));
    return 0;
}
int gpb_dev_free(gpb_ctx* c, double* p) {
    GPB_TRY(use(c));
    GPB_CUDA(cudaFree(p));
    return 0;
}
int gpb_dev_upload(gpb_ctx* c, double* dst, const double* src, int64_t n) {
    GPB_TRY(use(c));
    GPB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}
int gpb_dev_download(gpb_ctx* c, double* dst, const double* src, int64_t n) {
    GPB_TRY(use(c));
    GPB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->s));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}
int gpb_sync(gpb_ctx* c) {
    GPB_TRY(use(c));
    GPB_CUDA(cudaStreamSynchronize(c->s));
    return 0;
}

}  // extern "C"
