// Test-only C-ABI hooks that expose internal primitives on host buffers (used by tests/ and tools/).
#include "common.cuh"
#include <vector>

using namespace gpb;

extern "C" int gpb_test_gemm(int M, int N, int K, const double* A, const double* B, const double* C, double alpha,
                             double beta, int flags, double* D, int reps, double* ms_out) {
    double *dA, *dB, *dC = nullptr, *dD;
    GPB_CUDA(cudaMalloc(&dA, sizeof(double) * (size_t)M * K));
    GPB_CUDA(cudaMalloc(&dB, sizeof(double) * (size_t)N * K));
    GPB_CUDA(cudaMalloc(&dD, sizeof(double) * (size_t)M * N));
    GPB_CUDA(cudaMemcpy(dA, A, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));
    GPB_CUDA(cudaMemcpy(dB, B, sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice));
    if (C) {
        GPB_CUDA(cudaMalloc(&dC, sizeof(double) * (size_t)M * N));
        GPB_CUDA(cudaMemcpy(dC, C, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
    }
    GPB_CUDA(cudaMemset(dD, 0, sizeof(double) * (size_t)M * N));
    GemmArgs g{M, N, K, dA, K, dB, K, dC, N, dD, N, nullptr, 0, alpha, beta, flags};
    GPB_TRY(gemm_nt(g, 0));
    GPB_CUDA(cudaDeviceSynchronize());
    GPB_CUDA(cudaMemcpy(D, dD, sizeof(double) * (size_t)M * N, cudaMemcpyDeviceToHost));
    if (reps > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) GPB_TRY(gemm_nt(g, 0));
        cudaEventRecord(e1);
        GPB_CUDA(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_out) *ms_out = ms / reps;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(dD);
    return 0;
}

#include "kernels.cuh"
// potrf of a host n x n matrix (n % 128 == 0): returns L (lower; upper untouched) and the inverted diagonal blocks
extern "C" int gpb_test_potrf(int n, double* A, double* dinv_out, int* info, int reps, double* ms_out) {
    double *dA, *dA0, *dinv, *tmp;
    int* dinfo;
    const size_t bytes = sizeof(double) * (size_t)n * n;
    GPB_CUDA(cudaMalloc(&dA, bytes));
    GPB_CUDA(cudaMalloc(&dA0, bytes));
    GPB_CUDA(cudaMalloc(&dinv, sizeof(double) * (size_t)n * NB));
    GPB_CUDA(cudaMalloc(&tmp, sizeof(double) * (size_t)n * NB));
    GPB_CUDA(cudaMalloc(&dinfo, sizeof(int)));
    GPB_CUDA(cudaMemcpy(dA0, A, bytes, cudaMemcpyHostToDevice));
    LinalgWs ws{dinv, tmp, n, dinfo};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < std::max(1, reps); ++r) {
        GPB_CUDA(cudaMemcpy(dA, dA0, bytes, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e0);
        GPB_TRY(potrf_lower(dA, n, n, ws, 0));
        cudaEventRecord(e1);
        GPB_CUDA(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    if (ms_out) *ms_out = best;
    GPB_CUDA(cudaMemcpy(A, dA, bytes, cudaMemcpyDeviceToHost));
    if (dinv_out) GPB_CUDA(cudaMemcpy(dinv_out, dinv, sizeof(double) * (size_t)n * NB, cudaMemcpyDeviceToHost));
    GPB_CUDA(cudaMemcpy(info, dinfo, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dA0); cudaFree(dinv); cudaFree(tmp); cudaFree(dinfo);
    return 0;
}

// given a host SPD matrix: L (potrf), W = inv(L), Kinv (lower tiles) and X <- X L^-T for an m x n block
extern "C" int gpb_test_inverse(int n, const double* A, double* W_out, double* Kinv_out, int m, double* X_inout,
                                double* ms3_out) {
    double *dA, *dinv, *tmp, *dW, *dK, *dX = nullptr;
    int* dinfo;
    const size_t bytes = sizeof(double) * (size_t)n * n;
    GPB_CUDA(cudaMalloc(&dA, bytes));
    GPB_CUDA(cudaMalloc(&dW, bytes));
    GPB_CUDA(cudaMalloc(&dK, bytes));
    GPB_CUDA(cudaMalloc(&dinv, sizeof(double) * (size_t)n * NB));
    const int64_t tr = std::max(n, m);
    GPB_CUDA(cudaMalloc(&tmp, sizeof(double) * (size_t)tr * NB));
    GPB_CUDA(cudaMalloc(&dinfo, sizeof(int)));
    GPB_CUDA(cudaMemcpy(dA, A, bytes, cudaMemcpyHostToDevice));
    LinalgWs ws{dinv, tmp, tr, dinfo};
    cudaEvent_t e[4];
    for (auto& x : e) cudaEventCreate(&x);
    GPB_TRY(potrf_lower(dA, n, n, ws, 0));
    GPB_CUDA(cudaMemset(dW, 0, bytes));
    GPB_CUDA(cudaMemset(dK, 0, bytes));
    cudaEventRecord(e[0]);
    GPB_TRY(trtri_lower(dA, n, dW, n, n, 0, ws, dK, n, 0));
    cudaEventRecord(e[1]);
    GPB_TRY(lauum_lower(dW, n, dK, n, n, 0));
    cudaEventRecord(e[2]);
    if (m > 0) {
        GPB_CUDA(cudaMalloc(&dX, sizeof(double) * (size_t)m * n));
        GPB_CUDA(cudaMemcpy(dX, X_inout, sizeof(double) * (size_t)m * n, cudaMemcpyHostToDevice));
        cudaEventRecord(e[2]);
        GPB_TRY(trsm_right_lt(dX, n, m, dA, n, n, 0, ws, 0));
    }
    cudaEventRecord(e[3]);
    GPB_CUDA(cudaDeviceSynchronize());
    if (ms3_out) {
        float t;
        cudaEventElapsedTime(&t, e[0], e[1]); ms3_out[0] = t;
        cudaEventElapsedTime(&t, e[1], e[2]); ms3_out[1] = t;
        cudaEventElapsedTime(&t, e[2], e[3]); ms3_out[2] = t;
    }
    GPB_CUDA(cudaMemcpy(W_out, dW, bytes, cudaMemcpyDeviceToHost));
    GPB_CUDA(cudaMemcpy(Kinv_out, dK, bytes, cudaMemcpyDeviceToHost));
    if (m > 0) GPB_CUDA(cudaMemcpy(X_inout, dX, sizeof(double) * (size_t)m * n, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dW); cudaFree(dK); cudaFree(dinv); cudaFree(tmp); cudaFree(dinfo); cudaFree(dX);
    return 0;
}

// D = alpha A B^T + beta C through one chosen GEMM implementation (impl 0: dispatcher, 1: INT8 tensor-core path).
// A holds M*K doubles (K x M when flags has GEMM_A_MMAJOR), B holds N*K (K x N with GEMM_B_NMAJOR).
namespace gpb { int gemm_nt_i8(const GemmArgs& a, cudaStream_t s, double* flops_out); }
extern "C" int gpb_test_gemm_impl(int impl, int M, int N, int K, const double* A, const double* B, const double* C,
                                  double alpha, double beta, int flags, double* D, int reps, double* ms_out) {
    double *dA, *dB, *dC = nullptr, *dD;
    GPB_CUDA(cudaMalloc(&dA, sizeof(double) * (size_t)M * K));
    GPB_CUDA(cudaMalloc(&dB, sizeof(double) * (size_t)N * K));
    GPB_CUDA(cudaMalloc(&dD, sizeof(double) * (size_t)M * N));
    GPB_CUDA(cudaMemcpy(dA, A, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));
    GPB_CUDA(cudaMemcpy(dB, B, sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice));
    if (C) {
        GPB_CUDA(cudaMalloc(&dC, sizeof(double) * (size_t)M * N));
        GPB_CUDA(cudaMemcpy(dC, C, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
    }
    GPB_CUDA(cudaMemset(dD, 0, sizeof(double) * (size_t)M * N));
    GemmArgs g{M, N, K, dA, (flags & GEMM_A_MMAJOR) ? M : K, dB, (flags & GEMM_B_NMAJOR) ? N : K, dC, N, dD, N, nullptr, 0,
               alpha, beta, flags};
    auto run = [&]() -> int {
        if (impl == 1) {
            const int rc = gemm_nt_i8(g, 0, nullptr);
            if (rc == 1) set_error("gpb_test_gemm_impl: the INT8 path does not apply to this call");
            return rc == 0 ? 0 : -2;
        }
        return gemm_nt(g, 0);
    };
    GPB_TRY(run());
    GPB_CUDA(cudaDeviceSynchronize());
    GPB_CUDA(cudaMemcpy(D, dD, sizeof(double) * (size_t)M * N, cudaMemcpyDeviceToHost));
    if (reps > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) GPB_TRY(run());
        cudaEventRecord(e1);
        GPB_CUDA(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_out) *ms_out = ms / reps;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    gemm_i8_release(0);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(dD);
    return 0;
}
