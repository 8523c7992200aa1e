// Shared declarations for the gpb200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace gpb {

// ---------------------------------------------------------------- error plumbing
void set_error(const std::string& msg);  // api.cu; message returned by gpb_last_error()

#define GPB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::gpb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

#define GPB_TRY(expr)                    \
    do {                                 \
        int _r = (expr);                 \
        if (_r != 0) return _r;          \
    } while (0)

// ---------------------------------------------------------------- blocking constants
constexpr int NB = 128;          // diagonal-block / padding granularity of every dense matrix
constexpr int MAX_DIM = 8;       // spatial dimensions held in constant-size kernel parameter structs
constexpr int MAX_COMP = 4;      // covariance components (leaves) in a composite kernel
constexpr int MAX_REG = 4;       // regions of a ChangePoint kernel (covariance.py:371-605)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------- GEMM (gemm_dmma.cu)
// D = alpha * A * B^T + beta * C on logical operands A(m,k) [M x K], B(n,k) [N x K]; C/D/D2 are
// row-major M x N.  Storage of the operands is selectable:
//   a_kmajor: A stored row-major M x K (k contiguous), else stored row-major K x M (m contiguous)
//   b_kmajor: B stored row-major N x K (k contiguous), else stored row-major K x N (n contiguous)
// M % 128 == 0, N % 64 == 0, K % 16 == 0.  Triangular operands are exploited through per-tile
// k-ranges; the skipped parts must still hold finite values only where they are read (they are
// not read at all for whole skipped k-tiles; inside a diagonal tile zeros must be stored).
enum GemmFlags : int {
    GEMM_FULL = 0,
    GEMM_LOWER = 1,    // only tiles that intersect the lower triangle (row >= col) are computed
    GEMM_TRIK_A = 2,   // A(m,k) == 0 for k < m  -> k starts at the tile's first row
    GEMM_TRIK_B = 4,   // B(n,k) == 0 for k < n  -> k starts at the tile's first column
    GEMM_TRIL_B = 8,   // B(n,k) == 0 for k > n  -> k ends after the tile's last column
    GEMM_TRIL_A = 16,  // A(m,k) == 0 for k > m  -> k ends after the tile's last row
    GEMM_A_MMAJOR = 32,  // A stored K x M
    GEMM_B_NMAJOR = 64   // B stored K x N
};
struct GemmArgs {
    int M, N, K;
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    const double* C; int64_t ldc;   // may be nullptr when beta == 0
    double* D; int64_t ldd;
    double* D2; int64_t ldd2;       // optional second copy of the output (nullptr = off)
    double alpha, beta;
    int flags;
    // INT8 path only: longest k extent per launch (0 = the "gemm_i8_max_k" option, 16384).  Every chunk has its own row
    // scales, so a short chunk keeps more bits of operand rows whose entries decay along k (columns of inv(L) in
    // K^-1 = W^T W: profiles/grad_phase_sensitivity_r2.md); the chunks are accumulated in FP64.
    int max_k = 0;
    // GEMM_TRIK_A with a block structure instead of the tile's own: A(m, k) == 0 for k < (m / trik_a_blk) * trik_a_step
    // (0 = the plain rule, k < m).  The stacked row blocks of the distributed inverse (dist.cu) are block-cyclic: row
    // block i of the stack starts its non-zeros world * nbd columns after row block i - 1.
    int trik_a_blk = 0, trik_a_step = 0;
};
// first k of a tile whose first row is row0 under GEMM_TRIK_A
__host__ __device__ inline int trik_a_begin(int row0, int blk, int step) { return blk > 0 ? (row0 / blk) * step : row0; }
int gemm_nt(const GemmArgs& a, cudaStream_t s);
void gemm_i8_release(cudaStream_t s);  // gemm_i8.cu: frees the digit-plane workspace tied to a stream

// ---------------------------------------------------------------- process-wide counters and options (options.cu)
// Counters are atomics (restarts run on worker threads, one context each); every thread also keeps private copies so
// that a CUDA-graph capture can attribute launches / flops to the sequence it recorded.
void count_launch(int n = 1);  // every kernel launch is counted through this
int64_t launch_count();
int64_t thread_launch_count();
void credit_gemm_flops(double f, double f_i8);  // algorithmic FP64 flops of a GEMM launch (f_i8: issued on the INT8 path)
double gemm_flops_issued();
double gemm_flops_issued_i8();
double thread_gemm_flops();
double thread_gemm_flops_i8();

// Runtime options (gpb_set_option / gpb_get_option; initial values from the GPB200_* environment variables).
enum Option : int {
    OPT_GEMM_I8 = 0,      // 0 = FP64 DMMA kernels only, 1 = INT8 tensor-core path where it pays (default), 2 = wherever it applies
    OPT_GEMM_I8_MIN_K,    // shortest k extent sent to the INT8 path (default 512)
    OPT_GEMM_I8_PAIR,     // INT8 path tiles: 0 single CTA, 1 CTA pairs (cta_group::2) with uniform slots, 2 pairs with the wide slot layout; default 1
    OPT_GEMM_I8_DEBUG,    // timing probes of the INT8 kernel (results meaningless)
    OPT_GEMM_TILE,        // DMMA tile override: 0 auto, 2 small, 3 wide, 6 previous default
    OPT_GEMM_TMA,         // TMA-staged DMMA kernel for k-major operands (default 1)
    OPT_GRAPHS,           // CUDA-graph replay of pointer/shape-only launch sequences (default 1)
    OPT_I8_FALLBACK,      // re-run a factorisation on the DMMA kernels when the INT8 path reports a non-PD pivot (default 1)
    OPT_PREDICT_BLOCK,    // block width of the left-looking predict solve against cached digit planes (0 = auto, -1 = recursion)
    OPT_PREDICT_DIAG,     // diagonal blocks of the blocked predict solve: 1 = recursion (robust, default), 2 = INT8 inverse, 0 = by conditioning
    OPT_I8_GRAD_GUARD,    // a-posteriori error estimate of the INT8 inverse chain in marginal_likelihood_gradient (default 1)
    OPT_GEMM_I8_MAX_K,    // longest k extent of one INT8 launch (<= 16384, the int32 exactness limit); longer ones are chunked,
                          // each chunk with its own row scales, and accumulated in FP64
    OPT_GEMM_I8_EPI,      // epilogue warps per CTA of the INT8 kernel: 16 (default) or 8
    OPT_I8_GRAD_PHASES,   // diagnostic bit mask: which phases of gpb_lml_grad may use the INT8 path (1 potrf, 2 trtri, 4 lauum; default 7)
    OPT_GRAD_INVERSE,     // K^-1 of marginal_likelihood_gradient: 0 = recursive triangular inverse, W^T W (default); 1 = rows of L^-T by
                          // blocked substitution, Y Y^T (measured equal in time and accuracy: profiles/dist_grad_parity_r2.json)
    OPT_GEMM_I8_EPI2,     // INT8 kernel, second-sweep epilogue in two passes (accumulators to registers, release, then global memory)
    OPT_GEMM_I8_PREFETCH, // INT8 kernel: k-blocks by which an L2 prefetch of the digit planes runs ahead of the loads (0 = off)
    OPT_COUNT
};
int64_t option(Option o);
uint64_t option_epoch();  // bumped by every gpb_set_option: contexts drop their captured graphs when it changes
// per-thread override of OPT_GEMM_I8 (-1 = none): the DMMA retry after a failed INT8 factorisation
int gemm_i8_override();
void set_gemm_i8_override(int v);

// ---------------------------------------------------------------- model description (kernels.cu)
enum CovKind : int { COV_SE = 0, COV_RQ = 1, COV_WHITE = 2, COV_HETERO = 3 };
enum MeanKind : int { MEAN_CONST = 0, MEAN_LINEAR = 1, MEAN_QUADRATIC = 2 };

// Hyper-parameters already mapped from log space on the host (covariance.py:248-249, 344-346).
struct CovParams {
    int ncomp, d;
    int kind[MAX_COMP];
    int theta_off[MAX_COMP];             // offset of the component's first parameter inside theta_cov
    double amp2[MAX_COMP];               // a^2 (SE/RQ) or sigma^2 (WHITE)
    double rq_alpha[MAX_COMP];           // RQ shape parameter
    double inv_l2[MAX_COMP][MAX_DIM];    // 1 / l_k^2
    const double* hetero_log_sigma;      // device pointer (N entries) for COV_HETERO, else nullptr
    double jitter;                       // 1e-12 (covariance.py:221, 318)
    // ChangePoint (covariance.py:371-605): leaf c belongs to region[c] (-1 = not under a change-point); the leaf's
    // covariance is weighted by g_r(u) g_r(v), g_0 = 1 - f_0, g_r = f_{r-1} (1 - f_r), g_last = f_last,
    // f_a(x) = 1 / (1 + exp(-(x[cp_axis] - cp_loc[a]) / cp_width[a]))
    int region[MAX_COMP];
    int n_regions, cp_axis, cp_theta_off;
    double cp_loc[MAX_REG - 1], cp_width[MAX_REG - 1];
};
struct MeanParams {
    int kind, d;
    double c0;
    double lin[MAX_DIM], quad[MAX_DIM], xbar[MAX_DIM];
};

}  // namespace gpb
