// Pre-split operands of the INT8 tensor-core FP64 GEMM (gemm_i8.cu): digit planes that outlive one GEMM call.
//
// gemm_nt() splits both operands on every call.  Callers that multiply the SAME matrix many times (the predict solve:
// every query chunk against the fitted factor L) split it once and keep the seven int8 planes plus the row scales.
#pragma once
#include "common.cuh"

namespace gpb {

struct I8Planes {
    signed char* q = nullptr;  // planes[s][row][k], s = 0..6: row stride ld bytes, plane stride `plane` bytes
    double* scale = nullptr;   // per-row scale (2^e for a B operand, 2^(e-14) for an A operand)
    int64_t rows = 0, ld = 0, plane = 0;
};
inline size_t i8_plane_bytes(int64_t rows, int64_t ld) { return (size_t)7 * rows * ld; }

// Split `rows` k-contiguous rows X[r * ld + k], k in [0, K), into out's rows [row_off, row_off + rows), byte columns
// [k_off, k_off + K).  as_a: the operand will be the A (left) side of the product (its scale carries the 2^-14 of the
// digit weights).  bound (device, optional): a-priori bounds on |x|, row r uses bound[r % bound_period] instead of its
// measured maximum -- blocks of one logical row that are split at different times then share one scale; the scale is
// written only when k_off == 0 or no bound is given.
int i8_split_rows(const double* X, int64_t ld, int rows, int K, bool as_a, const double* bound, int bound_period,
                  const I8Planes& out, int row_off, int k_off, cudaStream_t s);

// D = alpha * A B^T + beta * C on pre-split operands: A rows [a_row_off, a_row_off + M), B rows [b_row_off, + N), k in
// [0, K) of both.  M % 128 == 0, N % 128 == 0, K % 64 == 0.  flags: GEMM_TRI*/GEMM_LOWER as gemm_nt (no *MAJOR flags).
int i8_gemm_planes(const I8Planes& A, int a_row_off, const I8Planes& B, int b_row_off, int M, int N, int K,
                   const double* C, int64_t ldc, double* D, int64_t ldd, double alpha, double beta, int flags,
                   cudaStream_t s);

}  // namespace gpb
