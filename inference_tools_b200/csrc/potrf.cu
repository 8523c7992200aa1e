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

// Factor one 128x128 diagonal block in place (lower triangle) and write inv(L_block) (full block,
// zero upper triangle) to dinv.  info: 1-based global index of the first non-positive pivot.
//
// Register-resident right-looking sweep: 256 threads as a 16 x 16 grid, thread (ty,tx) owns the 64
// elements (ty + 16a, tx + 16b) of the block AND of the running inverse in registers.  Step j publishes
// column j of the factor and row j of the partially solved identity through 2 KB of shared memory (one
// __syncthreads per step, double buffered); every thread then applies the rank-1 update to its own
// registers:  A[i][k] -= l_ij l_kj (potrf)  and  B[i][c] -= l_ij W[j][c] (forward substitution of L W = I).
// The whole kernel needs 4 KB of shared memory, so it can be scheduled next to resident GEMM CTAs.
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, int64_t ld,
                                                            double* __restrict__ dinv, int* __restrict__ info,
                                                            int row_offset) {
    __shared__ double cbuf[2][NB];  // column j of A (unscaled)
    __shared__ double rbuf[2][NB];  // row j of B (unscaled)
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    double r[8][8], w[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b <= a) {
                r[a][b] = A[(int64_t)(ty + 16 * a) * ld + tx + 16 * b];
                w[a][b] = (a == b && tx == ty) ? 1.0 : 0.0;
            } else {
                r[a][b] = 0.0;
                w[a][b] = 0.0;
            }
        }
    bool failed = false;
#pragma unroll
    for (int bj = 0; bj < 8; ++bj) {
        for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * bj + jj;
            const int buf = jj & 1;
            if (tx == jj) {
#pragma unroll
                for (int a = bj; a < 8; ++a) cbuf[buf][ty + 16 * a] = r[a][bj];
            }
            if (ty == jj) {
#pragma unroll
                for (int b = 0; b <= bj; ++b) rbuf[buf][tx + 16 * b] = w[bj][b];
            }
            __syncthreads();
            const double ajj = cbuf[buf][j];
            if (!(ajj > 0.0)) {  // also catches NaN; uniform across the CTA
                failed = true;
                if (tid == 0) atomicCAS(info, 0, row_offset + j + 1);
                break;
            }
            const double dj = sqrt(ajj);
            const double rinv = 1.0 / dj;
            double ci[8], ck[8], wj[8];
#pragma unroll
            for (int a = bj; a < 8; ++a) ci[a] = cbuf[buf][ty + 16 * a] * rinv;  // l_ij for my rows
#pragma unroll
            for (int b = bj; b < 8; ++b) ck[b] = cbuf[buf][tx + 16 * b] * rinv;  // l_kj for my columns
#pragma unroll
            for (int b = 0; b <= bj; ++b) wj[b] = rbuf[buf][tx + 16 * b] * rinv;  // W[j][c] for my columns
#pragma unroll
            for (int a = bj; a < 8; ++a) {
                const bool row_below = (a > bj) || (ty > jj);  // i > j
                if (row_below) {
#pragma unroll
                    for (int b = bj; b <= a; ++b) {
                        const bool col_right = (b > bj) || (tx > jj);  // k > j
                        const bool lower = (a > b) || (tx <= ty);      // k <= i
                        if (col_right && lower) r[a][b] = fma(-ci[a], ck[b], r[a][b]);
                    }
#pragma unroll
                    for (int b = 0; b <= bj; ++b) {
                        const bool col_left = (b < bj) || (tx <= jj);  // c <= j
                        if (col_left) w[a][b] = fma(-ci[a], wj[b], w[a][b]);
                    }
                }
            }
            // the owners keep the finished column j of L and row j of W
            if (tx == jj) {
#pragma unroll
                for (int a = bj; a < 8; ++a) {
                    if ((a > bj) || (ty > jj)) r[a][bj] = ci[a];
                    else if (ty == jj) r[a][bj] = dj;
                }
            }
            if (ty == jj) {
#pragma unroll
                for (int b = 0; b <= bj; ++b) w[bj][b] = wj[b];
            }
        }
        if (failed) break;
    }
    if (failed) {
        for (int idx = tid; idx < NB * NB; idx += 256) dinv[idx] = 0.0;
        return;
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int i = ty + 16 * a, k = tx + 16 * b;
            if (b <= a && k <= i) A[(int64_t)i * ld + k] = r[a][b];
            dinv[i * NB + k] = (b <= a && k <= i) ? w[a][b] : 0.0;
        }
}

int launch_diag(double* A, int64_t ld, int blk, const LinalgWs& ws, cudaStream_t s) {
    potrf_diag_kernel<<<1, 256, 0, s>>>(A, ld, ws.dinv + (int64_t)blk * NB * NB, ws.info, blk * NB);
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
