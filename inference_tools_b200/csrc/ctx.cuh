// Context object shared by api.cu, dist.cu and inverter.cu.
#pragma once
#include "../../include/gpb200.h"
#include "kernels.cuh"
#include "gemm_i8.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <utility>

namespace gpb {
struct PhaseTimer {
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    std::vector<cudaEvent_t> pool;
    cudaStream_t s = nullptr;
    void reset() {
        for (auto& m : marks) pool.push_back(m.second);
        marks.clear();
    }
    void mark(const char* name) {  // the interval that ENDS at the next mark is attributed to `name`
        cudaEvent_t e;
        if (!pool.empty()) {
            e = pool.back();
            pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        cudaEventRecord(e, s);
        marks.emplace_back(name, e);
    }
    void unmark_last() {  // drop the most recent mark (a call that extends its own timeline)
        if (marks.empty()) return;
        pool.push_back(marks.back().second);
        marks.pop_back();
    }
    ~PhaseTimer() {
        reset();
        for (auto e : pool) cudaEventDestroy(e);
    }
};

template <typename T>
int ensure(T*& p, size_t& cap, size_t bytes) {
    if (cap >= bytes && p) return 0;
    if (p) GPB_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
    GPB_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return 0;
}

}  // namespace gpb

// Digit planes of the fitted factor for the blocked predict solve (api.cu: predict_solve_blocked), built once per fit:
//   Lp[j]  rows of block row j of L left of its diagonal block (nb_j x j nb), the B operand of T_j = S_j - X_<j L_j,<j^T
//   Wp[j]  inv(L_jj) (nb_j x nb_j, lower), only for well-conditioned fits (use_w): the B operand of X_j = T_j inv(L_jj)^T
// and per query chunk: Xp, the solved columns (a-priori row scale), Tp, the current block before its diagonal solve.
struct PredictPlanes {
    bool valid = false, use_w = false;
    int nb = 0, nblk = 0, ns = 0;
    std::vector<gpb::I8Planes> Lp, Wp;
    gpb::I8Planes Xp, Tp;
    signed char *lbuf = nullptr, *wbuf = nullptr, *xbuf = nullptr, *tbuf = nullptr;
    double *lscale = nullptr, *wscale = nullptr, *xscale = nullptr, *tscale = nullptr, *bound = nullptr, *wtmp = nullptr;
    size_t lbuf_cap = 0, wbuf_cap = 0, xbuf_cap = 0, tbuf_cap = 0, lscale_cap = 0, wscale_cap = 0, xscale_cap = 0,
           tscale_cap = 0, bound_cap = 0, wtmp_cap = 0;
};

struct gpb_dist;  // dist.cu
struct gpb_linv;  // inverter.cu
using gpb::MAX_COMP;
using gpb::MAX_DIM;

struct gpb_ctx {
    int device = 0;
    cudaStream_t s = nullptr;
    int64_t n = 0, npad = 0;
    int d = 0;
    double *x = nullptr, *y = nullptr, *noise = nullptr, *ycov = nullptr;
    bool has_noise = false, has_ycov = false;
    double xbar[MAX_DIM] = {0};
    int ncomp = 0, kinds[MAX_COMP] = {0}, theta_off[MAX_COMP] = {0}, mean_kind = 0, n_mean = 0, n_cov = 0;
    int region[MAX_COMP] = {-1, -1, -1, -1}, n_regions = 0, cp_axis = 0, cp_theta_off = 0;  // ChangePoint layout
    bool model_set = false;
    double* theta_dev = nullptr;
    size_t theta_dev_cap = 0;
    // fitted state (set_hyperparameters)
    double *Lfit = nullptr, *dinv_fit = nullptr, *alpha = nullptr, *mu = nullptr;
    size_t Lfit_cap = 0, dinv_fit_cap = 0, alpha_cap = 0, mu_cap = 0;
    std::vector<double> theta_fit;
    bool fitted = false;
    double min_pivot_fit = 0.0;  // min_i L_ii of the fitted factor
    gpb::CovParams cp_fit;
    gpb::MeanParams mp_fit;
    // objective-evaluation workspace
    double *Kwork = nullptr, *dinv_work = nullptr, *W = nullptr, *Kinv = nullptr, *partials = nullptr,
           *grad_dev = nullptr, *vec = nullptr, *resid = nullptr, *alpha_work = nullptr, *scal = nullptr, *tmp = nullptr;
    size_t Kwork_cap = 0, dinv_work_cap = 0, W_cap = 0, Kinv_cap = 0, partials_cap = 0, grad_cap = 0, vec_cap = 0,
           resid_cap = 0, alpha_work_cap = 0, scal_cap = 0, tmp_cap = 0;
    int64_t tmp_rows = 0;
    int* info_dev = nullptr;
    // predict workspace
    double *S = nullptr, *dots = nullptr, *G = nullptr, *qbuf = nullptr, *o1 = nullptr, *o2 = nullptr, *o3 = nullptr,
           *R_dev = nullptr;
    size_t S_cap = 0, dots_cap = 0, G_cap = 0, qbuf_cap = 0, o1_cap = 0, o2_cap = 0, o3_cap = 0, R_cap = 0;
    char* argws = nullptr;  // arg-best reduction workspace (values, then indices)
    size_t argws_cap = 0;
    gpb::PhaseTimer timer;
    // CUDA-graph cache of the pointer/shape-only launch sequences (potrf, solves, inverse, predict solve)
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int uses = 0;
        int64_t launches = 0;
        double flops = 0.0, flops_i8 = 0.0;
    };
    std::map<std::string, GraphEntry> graphs;
    bool use_graphs = true;
    uint64_t opt_epoch = 0;  // option_epoch() the graphs were recorded under
    int64_t dmma_retries = 0;  // factorisations repeated on the DMMA kernels after the INT8 path reported info > 0
    int64_t grad_guard_retries = 0;  // gradient evaluations whose inverse chain was repeated on DMMA (error guard)
    double grad_guard_est = 0.0;     // last estimate of the INT8 gradient error relative to max|grad|
    PredictPlanes pp;
    gpb_dist* dist = nullptr;
    gpb_linv* linv = nullptr;
};


namespace gpb {
// helpers implemented in api.cu, used by dist.cu
int ctx_use(gpb_ctx* c);
int ctx_need_model(gpb_ctx* c);
int ctx_make_cov_params(gpb_ctx* c, const double* theta_cov, CovParams& cp);
void ctx_make_mean_params(gpb_ctx* c, const double* theta_mean, MeanParams& mp);
void dist_destroy(gpb_ctx* c);  // dist.cu
void linv_destroy(gpb_ctx* c);  // inverter.cu
}  // namespace gpb
