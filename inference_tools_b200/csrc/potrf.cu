// Blocked FP64 Cholesky, triangular solve and triangular inverse for sm_100a.
//
// Replaces numpy.linalg.cholesky (LAPACK dpotrf) and scipy.linalg.solve_triangular (dtrtrs) at the
// reference call sites regression.py:241, 213, 376, 410, 447, 460, 479-480, 501-502, 537-538, 555-556.
//
// Structure: cache-oblivious recursion on 128-column blocks.  All O(N^3) work is issued as large
// DMMA GEMMs (gemm_dmma.cu) whose K dimension is a whole half of the current sub-problem, so the
// tensor pipe sees long main loops instead of the K = nb rank updates of a right-looking sweep:
//   potrf(A)      = potrf(A11); A21 <- A21 L11^-T; A22 -= A21 A21^T (lower tiles); potrf(A22)
//   X L^-T        = [X1 L11^-T,  (X2 - X1' L21^T) L22^-T]
//   inv(L)        = [[W11, 0], [-W22 (L21 W11), W22]]
// Leaves are 128 x 128: one CTA factors the diagonal block in shared memory (warp-parallel
// right-looking sweep) and also leaves its explicit inverse, so every leaf solve is a GEMM too.
#include "kernels.cuh"

namespace gpb {
namespace {

constexpr int LDD = NB + 1;
constexpr int DIAG_SMEM = (NB * LDD + 3 * NB) * (int)sizeof(double);

// Factor one 128x128 diagonal block in place (lower triangle) and write inv(L_block) (full block,
// zero upper triangle) to dinv.  info: 1-based global index of the first non-positive pivot.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int64_t ld, double* __restrict__ dinv,
                                                         int* __restrict__ info, int row_offset) {
    extern __shared__ double sm[];
    double* S = sm;
    double* col = sm + NB * LDD;   // two column snapshots (double buffered)
    double* dg = col + 2 * NB;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        S[r * LDD + c] = A[(int64_t)r * ld + c];
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    bool failed = false;
    for (int j = 0; j < NB; ++j) {
        const double ajj = S[j * LDD + j];
        if (!(ajj > 0.0)) {  // also catches NaN; uniform across the CTA
            failed = true;
            if (tid == 0) atomicCAS(info, 0, row_offset + j + 1);
            break;
        }
        const double dj = sqrt(ajj);
        if (tid < NB) {
            if (tid > j) {
                const double v = S[tid * LDD + j] / dj;
                S[tid * LDD + j] = v;
                col[tid] = v;
            } else if (tid == j) {
                dg[j] = dj;
            }
        }
        __syncthreads();
        const int a0 = (j >= ty) ? (j - ty) / 16 + 1 : 0;
        const int b0 = (j >= tx) ? (j - tx) / 16 + 1 : 0;
        for (int a = a0; a < NB / 16; ++a) {
            const int i = ty + 16 * a;
            const double ci = col[i];
            for (int b = b0; tx + 16 * b <= i; ++b) {
                const int k = tx + 16 * b;
                S[i * LDD + k] = fma(-ci, col[k], S[i * LDD + k]);
            }
        }
        __syncthreads();
    }
    if (failed) {
        // leave a defined (NaN) factor behind so downstream kernels stay finite-state; info carries the error
        for (int idx = tid; idx < NB * NB; idx += 256) dinv[idx] = 0.0;
        return;
    }
    // write L back (lower triangle only)
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        if (c < r) A[(int64_t)r * ld + c] = S[r * LDD + c];
        else if (c == r) A[(int64_t)r * ld + c] = dg[r];
    }
    // in-place inverse of the lower-triangular factor (unblocked dtrti2 order: last column first)
    if (tid < NB) S[tid * LDD + tid] = dg[tid];
    __syncthreads();
    const int i = tid >> 1, h = tid & 1;
    for (int j = NB - 1; j >= 0; --j) {
        double* cj = col + (j & 1) * NB;
        if (tid < NB && tid > j) cj[tid] = S[tid * LDD + j];
        const double wjj = 1.0 / S[j * LDD + j];
        __syncthreads();
        double sum = 0.0;
        if (i > j) {
            for (int k = j + 1 + h; k <= i; k += 2) sum = fma(S[i * LDD + k], cj[k], sum);
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        if (h == 0) {
            if (i > j) S[i * LDD + j] = -wjj * sum;
            else if (i == j) S[j * LDD + j] = wjj;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 7, c = idx & (NB - 1);
        dinv[idx] = (c <= r) ? S[r * LDD + c] : 0.0;
    }
}

int launch_diag(double* A, int64_t ld, int blk, const LinalgWs& ws, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        GPB_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
        configured = true;
    }
    potrf_diag_kernel<<<1, 256, DIAG_SMEM, s>>>(A, ld, ws.dinv + (int64_t)blk * NB * NB, ws.info, blk * NB);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

inline int split(int n) { return ((n / NB + 1) / 2) * NB; }  // first part gets the larger half

int potrf_rec(double* A, int64_t ld, int n, int blk0, const LinalgWs& ws, cudaStream_t s) {
    if (n == NB) return launch_diag(A, ld, blk0, ws, s);
    const int n1 = split(n), n2 = n - n1;
    GPB_TRY(potrf_rec(A, ld, n1, blk0, ws, s));
    double* A21 = A + (int64_t)n1 * ld;
    double* A22 = A21 + n1;
    GPB_TRY(trsm_right_lt(A21, ld, n2, A, ld, n1, blk0, ws, s));
    GemmArgs g{n2, n2, n1, A21, ld, A21, ld, A22, ld, A22, ld, nullptr, 0, -1.0, 1.0, GEMM_LOWER};
    GPB_TRY(gemm_nt(g, s));
    return potrf_rec(A22, ld, n2, blk0 + n1 / NB, ws, s);
}

}  // namespace

int potrf_lower(double* A, int64_t ld, int n, const LinalgWs& ws, cudaStream_t s) {
    if (n % NB) {
        set_error("potrf_lower: n must be a multiple of 128");
        return -2;
    }
    GPB_CUDA(cudaMemsetAsync(ws.info, 0, sizeof(int), s));
    return potrf_rec(A, ld, n, 0, ws, s);
}

int trsm_right_lt(double* X, int64_t ldx, int m, const double* L, int64_t ldl, int n, int blk0, const LinalgWs& ws,
                  cudaStream_t s) {
    if (m == 0 || n == 0) return 0;
    if (n == NB) {
        const double* dinv = ws.dinv + (int64_t)blk0 * NB * NB;
        for (int64_t r0 = 0; r0 < m; r0 += ws.tmp_rows) {
            const int rows = (int)std::min<int64_t>(ws.tmp_rows, m - r0);
            double* Xr = X + r0 * ldx;
            GemmArgs g{rows, NB, NB, Xr, ldx, dinv, NB, nullptr, 0, ws.tmp, NB, nullptr, 0, 1.0, 0.0, GEMM_TRIL_B};
            GPB_TRY(gemm_nt(g, s));
            GPB_TRY(launch_copy2d(ws.tmp, NB, Xr, ldx, rows, NB, s));
        }
        return 0;
    }
    const int n1 = split(n), n2 = n - n1;
    GPB_TRY(trsm_right_lt(X, ldx, m, L, ldl, n1, blk0, ws, s));
    const double* L21 = L + (int64_t)n1 * ldl;
    GemmArgs g{m, n2, n1, X, ldx, L21, ldl, X + n1, ldx, X + n1, ldx, nullptr, 0, -1.0, 1.0, GEMM_FULL};
    GPB_TRY(gemm_nt(g, s));
    return trsm_right_lt(X + n1, ldx, m, L21 + n1, ldl, n2, blk0 + n1 / NB, ws, s);
}

int trtri_lower(const double* L, int64_t ldl, double* W, int64_t ldw, int n, int blk0, const LinalgWs& ws,
                double* scratch, int64_t lds, cudaStream_t s) {
    if (n == NB) return launch_copy2d(ws.dinv + (int64_t)blk0 * NB * NB, NB, W, ldw, NB, NB, s);
    const int n1 = split(n), n2 = n - n1;
    GPB_TRY(trtri_lower(L, ldl, W, ldw, n1, blk0, ws, scratch, lds, s));
    const double* L21 = L + (int64_t)n1 * ldl;
    double* W21 = W + (int64_t)n1 * ldw;
    double* W22 = W21 + n1;
    GPB_TRY(trtri_lower(L21 + n1, ldl, W22, ldw, n2, blk0 + n1 / NB, ws, scratch, lds, s));
    // T = L21 W11 : B(n,k) = W11[k][n] (stored K x N), zero for k < n
    GemmArgs g1{n2, n1, n1, L21, ldl, W, ldw, nullptr, 0, scratch, lds, nullptr, 0, 1.0, 0.0,
                GEMM_B_NMAJOR | GEMM_TRIK_B};
    GPB_TRY(gemm_nt(g1, s));
    // W21 = -W22 T : A = W22 lower triangular (zero for k > m); B(n,k) = T[k][n] (stored K x N)
    GemmArgs g2{n2, n1, n2, W22, ldw, scratch, lds, nullptr, 0, W21, ldw, nullptr, 0, -1.0, 0.0,
                GEMM_B_NMAJOR | GEMM_TRIL_A};
    return gemm_nt(g2, s);
}

int lauum_lower(const double* W, int64_t ldw, double* Kinv, int64_t ldk, int n, cudaStream_t s) {
    // Kinv_ij = sum_k W[k][i] W[k][j], k >= max(i, j)
    GemmArgs g{n, n, n, W, ldw, W, ldw, nullptr, 0, Kinv, ldk, nullptr, 0, 1.0, 0.0,
               GEMM_A_MMAJOR | GEMM_B_NMAJOR | GEMM_TRIK_A | GEMM_TRIK_B | GEMM_LOWER};
    return gemm_nt(g, s);
}

}  // namespace gpb
