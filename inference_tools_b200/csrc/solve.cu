// Single right-hand-side triangular solves and the scalar reductions of the log marginal likelihood.
//
// alpha = L^-T (L^-1 (y - mu))   regression.py:242-244;   -0.5 v.v - sum log L_ii   regression.py:538-539.
// Column-oriented block substitution: one launch per 128-block; every CTA first forms the solved
// block x_i = inv(L_ii) b_i from the explicit diagonal-block inverse (redundantly, 128x128 matvec from
// L2), then updates its own slice of the remaining right-hand side.  HBM-bound: each pass reads the
// lower triangle once (4 N^2 bytes); no atomics, fixed summation order.
#include "kernels.cuh"

namespace gpb {
namespace {

// forward: x_i = Dinv_i b_i ; b[rows below] -= L[rows, blk i] x_i.   grid.x = max(1, rows_below / 128)
__global__ void __launch_bounds__(256) trsv_fwd_step_kernel(const double* __restrict__ L, int64_t ld,
                                                            const double* __restrict__ dinv, double* __restrict__ b,
                                                            double* __restrict__ xout, int blk, int npad) {
    __shared__ double bi[NB];
    __shared__ double xi[NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blk * NB;
    if (tid < NB) bi[tid] = b[c0 + tid];
    __syncthreads();
    for (int r = warp; r < NB; r += 8) {  // x_i[r] = sum_{k<=r} Dinv[r][k] b_i[k]
        const double* row = dinv + (int64_t)r * NB;
        double a = 0.0;
#pragma unroll
        for (int k = lane; k < NB; k += 32) a = fma(row[k], bi[k], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) xi[r] = a;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < NB) xout[c0 + tid] = xi[tid];
    const int r0 = c0 + NB + blockIdx.x * NB;
    if (r0 >= npad) return;
    for (int r = warp; r < NB; r += 8) {
        const double* row = L + (int64_t)(r0 + r) * ld + c0;
        double a = 0.0;
#pragma unroll
        for (int k = lane; k < NB; k += 32) a = fma(row[k], xi[k], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) b[r0 + r] -= a;
    }
}

// backward: x_i = Dinv_i^T b_i ; b[cols before] -= L[blk i rows, cols]^T x_i.  grid.x = max(1, blk)
__global__ void __launch_bounds__(256) trsv_bwd_step_kernel(const double* __restrict__ L, int64_t ld,
                                                            const double* __restrict__ dinv, double* __restrict__ b,
                                                            double* __restrict__ xout, int blk) {
    __shared__ double bi[NB];
    __shared__ double xi[NB];
    __shared__ double part[2 * NB];
    const int tid = threadIdx.x, t = tid & (NB - 1), h = tid >> 7;
    const int c0 = blk * NB;
    if (tid < NB) bi[tid] = b[c0 + tid];
    __syncthreads();
    {   // x_i[t] = sum_{k>=t} Dinv[k][t] b_i[k]; two halves of the k range
        double a = 0.0;
        for (int k = h * (NB / 2); k < (h + 1) * (NB / 2); ++k) a = fma(dinv[(int64_t)k * NB + t], bi[k], a);
        part[h * NB + t] = a;
    }
    __syncthreads();
    if (tid < NB) xi[tid] = part[tid] + part[NB + tid];
    __syncthreads();
    if (blockIdx.x == 0 && tid < NB) xout[c0 + tid] = xi[tid];
    if (blk == 0) return;
    const int col = blockIdx.x * NB + t;  // columns before the block
    {
        double a = 0.0;
        const double* base = L + (int64_t)(c0 + h * (NB / 2)) * ld + col;
        for (int r = 0; r < NB / 2; ++r) a = fma(base[(int64_t)r * ld], xi[h * (NB / 2) + r], a);
        part[h * NB + t] = a;
    }
    __syncthreads();
    if (tid < NB) b[col] -= part[tid] + part[NB + tid];
}

__global__ void __launch_bounds__(1024) logdet_dot_kernel(const double* __restrict__ L, int64_t ld,
                                                          const double* __restrict__ a, const double* __restrict__ b,
                                                          int n, double* __restrict__ out2) {
    __shared__ double s0[1024];
    __shared__ double s1[1024];
    __shared__ double s2[1024];
    const int tid = threadIdx.x;
    double ld_sum = 0.0, dot = 0.0, dmin = 1e300;
    for (int i = tid; i < n; i += 1024) {
        const double lii = L[(int64_t)i * ld + i];
        ld_sum += log(lii);
        dmin = fmin(dmin, lii);
        dot = fma(a[i], b[i], dot);
    }
    s0[tid] = ld_sum;
    s1[tid] = dot;
    s2[tid] = dmin;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) {
            s0[tid] += s0[tid + o];
            s1[tid] += s1[tid + o];
            s2[tid] = fmin(s2[tid], s2[tid + o]);
        }
        __syncthreads();
    }
    if (tid == 0) {
        out2[0] = s0[0];
        out2[1] = s1[0];
        out2[2] = s2[0];  // smallest pivot: 1 / out2[2] bounds |inv(L)| (gradient error guard, api.cu)
    }
}

}  // namespace

int trsv_lower_fwd(const double* L, int64_t ld, int npad, const double* dinv, double* vec, cudaStream_t s) {
    // `vec` holds 2*npad doubles: [0,npad) is the right-hand side (consumed), [npad,2*npad) receives the
    // solution.  Separate halves keep the step race free: every CTA of step i reads block i of the RHS
    // while CTA 0 publishes the solved block.
    const int nblk = npad / NB;
    double* xout = vec + npad;
    for (int i = 0; i < nblk; ++i) {
        const int grid = std::max(1, nblk - 1 - i);
        trsv_fwd_step_kernel<<<grid, 256, 0, s>>>(L, ld, dinv + (int64_t)i * NB * NB, vec, xout, i, npad);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    return 0;
}

int trsv_lower_bwd(const double* L, int64_t ld, int npad, const double* dinv, double* vec, cudaStream_t s) {
    const int nblk = npad / NB;
    double* xout = vec + npad;
    for (int i = nblk - 1; i >= 0; --i) {
        const int grid = std::max(1, i);
        trsv_bwd_step_kernel<<<grid, 256, 0, s>>>(L, ld, dinv + (int64_t)i * NB * NB, vec, xout, i);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    return 0;
}

int launch_logdet_dot(const double* L, int64_t ld, const double* a, const double* b, int n, double* out2,
                      cudaStream_t s) {
    logdet_dot_kernel<<<1, 1024, 0, s>>>(L, ld, a, b, n, out2);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace gpb
