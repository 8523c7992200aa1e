// Covariance assembly, cross-covariance, row reductions and acquisition kernels (sm_100a).
//
// None of these materialise the reference's (N,N,d) difference arrays (covariance.py:218-219,
// 315-316): every kernel recomputes K_ij from the coordinates.  Thread mapping shared by the
// assembly kernels: a CTA owns a 128 x 128 output tile, thread = one output column (its point is
// held in registers), the CTA's row points sit in shared memory and are read as warp broadcasts,
// so a warp stores 32 consecutive doubles (256 B) per row.  The kernels are FP64-ALU bound
// (~3d + 30 DFMA per element), not HBM bound: 8 N^2 bytes at N = 32768 is 1.3 ms of HBM time.
#include "kernels.cuh"
#include "cov_device.cuh"

namespace gpb {
namespace {

constexpr int TILE = 128;

// sum over the smooth components of w_c a^2 f(z_c), z_c = sum_k 0.5 d2_k / l_ck^2; gi/gj = region weights of the two
// points (ignored unless the model has a ChangePoint)
template <bool CP = true>
__device__ __forceinline__ double cov_from_d2(const CovParams& cp, const double (&d2)[MAX_DIM],
                                              const double (&gi)[MAX_REG], const double (&gj)[MAX_REG]) {
    double kv = 0.0;
    for (int c = 0; c < cp.ncomp; ++c) {
        const int kind = cp.kind[c];
        if (kind > COV_RQ) continue;
        const double w = (CP && cp.n_regions) ? leaf_weight(cp, c, gi, gj) : 1.0;
        double z = 0.0;
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k)
            if (k < cp.d) z += (0.5 * d2[k]) * cp.inv_l2[c][k];
        if (kind == COV_SE) {
            kv += w * cp.amp2[c] * exp(-z);
        } else {
            const double q = cp.rq_alpha[c];
            // (1 + Z/q)^-q evaluated as exp(-q ln F), the form the reference itself uses at covariance.py:356-358;
            // a few ulp from the ** of :341/:348 and less than half the cost of a double-precision pow()
            kv += w * cp.amp2[c] * exp(-q * log(1.0 + z / q));
        }
    }
    return kv;
}

// diagonal additions of the data covariance: a^2 * 1e-12 per smooth component (covariance.py:254-255,
// 348), sigma^2 (White :168-169), exp(2 theta_i) (Hetero :679-680), y_err^2 (regression.py:320)
template <bool CP = true>
__device__ __forceinline__ double diag_terms(const CovParams& cp, int gi, const double* noise_var,
                                             const double (&g)[MAX_REG]) {
    double v = 0.0;
    for (int c = 0; c < cp.ncomp; ++c) {
        const int kind = cp.kind[c];
        const double w = (CP && cp.n_regions) ? leaf_weight(cp, c, g, g) : 1.0;
        if (kind <= COV_RQ) v += w * cp.amp2[c] * cp.jitter;
        else if (kind == COV_WHITE) v += w * cp.amp2[c];
        else v += w * exp(2.0 * cp.hetero_log_sigma[gi]);
    }
    if (noise_var) v += noise_var[gi];
    return v;
}

__device__ __forceinline__ void lower_tile(int t, int& bi, int& bj) {
    int r = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while (r * (r + 1) / 2 > t) --r;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    bi = r;
    bj = t - r * (r + 1) / 2;
}

template <bool CP>
__global__ void __launch_bounds__(256) assemble_train_kernel(const CovParams cp, const double* __restrict__ x, int n,
                                                             const double* __restrict__ noise_var,
                                                             const double* __restrict__ y_cov, double* __restrict__ K,
                                                             int64_t ld, int mirror) {
    __shared__ double xs[TILE * MAX_DIM];
    __shared__ double gs[CP ? TILE : 1][MAX_REG];
    int bi, bj;
    lower_tile(blockIdx.x, bi, bj);
    const int row0 = bi * TILE, col0 = bj * TILE;
    const int tid = threadIdx.x, col = tid & (TILE - 1), half = tid >> 7;
    const int d = cp.d;
    for (int idx = tid; idx < TILE * d; idx += 256) xs[idx] = x[(int64_t)row0 * d + idx];
    const int gj = col0 + col;
    double xj[MAX_DIM], wj[MAX_REG] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) xj[k] = (k < d) ? x[(int64_t)gj * d + k] : 0.0;
    if (CP && cp.n_regions) {
        region_weights(cp, x[(int64_t)gj * d + cp.cp_axis], wj);
        if (tid < TILE) {
            double g[MAX_REG];
            region_weights(cp, x[(int64_t)(row0 + tid) * d + cp.cp_axis], g);
#pragma unroll
            for (int q = 0; q < MAX_REG; ++q) gs[tid][q] = g[q];
        }
    }
    __syncthreads();
    for (int r = 0; r < TILE / 2; ++r) {
        const int i = half * (TILE / 2) + r;
        const int gi = row0 + i;
        double d2[MAX_DIM], wi[MAX_REG];
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) {
            const double df = (k < d) ? xs[i * d + k] - xj[k] : 0.0;
            d2[k] = df * df;
        }
#pragma unroll
        for (int q = 0; q < MAX_REG; ++q) wi[q] = (CP && cp.n_regions) ? gs[CP ? i : 0][q] : 0.0;
        double v = cov_from_d2<CP>(cp, d2, wi, wj);
        if (gi == gj) v += diag_terms<CP>(cp, gi, noise_var, wi);
        if (gi >= n || gj >= n) v = (gi == gj) ? 1.0 : 0.0;
        else if (y_cov) v += y_cov[(int64_t)gi * n + gj];
        K[(int64_t)gi * ld + gj] = v;
        if (mirror && bi != bj) K[(int64_t)gj * ld + gi] = v;
    }
}

// Rectangular block rows [row0, ..) x cols [col0, ..) of K(theta)+diag terms into a panel buffer
// (distributed Cholesky: every rank assembles only the block columns it owns).  grid = (ncols/128, nrows/128)
template <bool CP>
__global__ void __launch_bounds__(256) assemble_block_kernel(const CovParams cp, const double* __restrict__ x, int n,
                                                             const double* __restrict__ noise_var, int row0, int col0,
                                                             double* __restrict__ out, int64_t ld) {
    __shared__ double xs[TILE * MAX_DIM];
    __shared__ double gs[CP ? TILE : 1][MAX_REG];
    const int r0 = row0 + blockIdx.y * TILE, c0 = col0 + blockIdx.x * TILE;
    const int tid = threadIdx.x, col = tid & (TILE - 1), half = tid >> 7;
    const int d = cp.d;
    for (int idx = tid; idx < TILE * d; idx += 256) xs[idx] = x[(int64_t)r0 * d + idx];
    const int gj = c0 + col;
    double xj[MAX_DIM], wj[MAX_REG] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) xj[k] = (k < d) ? x[(int64_t)gj * d + k] : 0.0;
    if (CP && cp.n_regions) {
        region_weights(cp, x[(int64_t)gj * d + cp.cp_axis], wj);
        if (tid < TILE) {
            double g[MAX_REG];
            region_weights(cp, x[(int64_t)(r0 + tid) * d + cp.cp_axis], g);
#pragma unroll
            for (int q = 0; q < MAX_REG; ++q) gs[tid][q] = g[q];
        }
    }
    __syncthreads();
    for (int r = 0; r < TILE / 2; ++r) {
        const int i = half * (TILE / 2) + r;
        const int gi = r0 + i;
        double d2[MAX_DIM], wi[MAX_REG];
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) {
            const double df = (k < d) ? xs[i * d + k] - xj[k] : 0.0;
            d2[k] = df * df;
        }
#pragma unroll
        for (int q = 0; q < MAX_REG; ++q) wi[q] = (CP && cp.n_regions) ? gs[CP ? i : 0][q] : 0.0;
        double v = cov_from_d2<CP>(cp, d2, wi, wj);
        if (gi == gj) v += diag_terms<CP>(cp, gi, noise_var, wi);
        if (gi >= n || gj >= n) v = (gi == gj) ? 1.0 : 0.0;
        out[(int64_t)(gi - row0) * ld + (gj - col0)] = v;
    }
}

// K (without sig) and every dK/dtheta plane, dense n x n each (covariance.py:268-276, 350-365,
// 171-175, 682-686).  API-parity path for small N; one thread per element.
__global__ void assemble_grads_kernel(const CovParams cp, const double* __restrict__ x, int n, double* __restrict__ K,
                                      double* __restrict__ dK) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    const int d = cp.d;
    const int64_t plane = (int64_t)n * n;
    double d2[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) {
        const double df = (k < d) ? x[(int64_t)i * d + k] - x[(int64_t)j * d + k] : 0.0;
        d2[k] = df * df;
    }
    double gi[MAX_REG] = {0.0, 0.0, 0.0, 0.0}, gj[MAX_REG] = {0.0, 0.0, 0.0, 0.0};
    if (cp.n_regions) {
        region_weights(cp, x[(int64_t)i * d + cp.cp_axis], gi);
        region_weights(cp, x[(int64_t)j * d + cp.cp_axis], gj);
    }
    double ktot = 0.0;
    double kreg[MAX_REG] = {0.0, 0.0, 0.0, 0.0};  // unweighted covariance of each ChangePoint region
    for (int c = 0; c < cp.ncomp; ++c) {
        const int kind = cp.kind[c];
        const double w = cp.n_regions ? leaf_weight(cp, c, gi, gj) : 1.0;
        int p = cp.theta_off[c];
        double kv = 0.0;
        if (kind == COV_SE) {
            double z = 0.0;
            for (int k = 0; k < d; ++k) z += (0.5 * d2[k]) * cp.inv_l2[c][k];
            kv = cp.amp2[c] * (exp(-z) + (i == j ? cp.jitter : 0.0));
            dK[plane * p++ + e] = w * 2.0 * kv;
            for (int k = 0; k < d; ++k) dK[plane * p++ + e] = w * (d2[k] * cp.inv_l2[c][k]) * kv;
        } else if (kind == COV_RQ) {
            double z = 0.0;
            for (int k = 0; k < d; ++k) z += (0.5 * d2[k]) * cp.inv_l2[c][k];
            const double q = cp.rq_alpha[c];
            const double F = 1.0 + z / q, lnF = log(F);
            kv = cp.amp2[c] * (exp(-q * lnF) + (i == j ? cp.jitter : 0.0));
            dK[plane * p++ + e] = w * 2.0 * kv;
            dK[plane * p++ + e] = w * -kv * (lnF * q - z / F);
            const double G = 2.0 * kv / F;
            for (int k = 0; k < d; ++k) dK[plane * p++ + e] = w * G * ((0.5 * d2[k]) * cp.inv_l2[c][k]);
        } else if (kind == COV_WHITE) {
            kv = (i == j) ? cp.amp2[c] : 0.0;
            dK[plane * p++ + e] = w * 2.0 * kv;
        } else {
            kv = (i == j) ? exp(2.0 * cp.hetero_log_sigma[i]) : 0.0;
            for (int m = 0; m < n; ++m) dK[plane * p++ + e] = (m == i) ? w * 2.0 * kv : 0.0;
        }
        ktot += w * kv;
        const int r = cp.region[c];
#pragma unroll
        for (int q2 = 0; q2 < MAX_REG; ++q2)
            if (q2 == r) kreg[q2] += kv;
    }
    // change-point parameters (covariance.py:574-583): dK = K_a (A + A^T) + K_{a+1} (B + B^T),
    // A = -dw (1 - w)^T, B = dw w^T, dw = df/dc or df/dwidth (logistic_and_gradient :597-602)
    for (int a = 0; a + 1 < cp.n_regions; ++a) {
        const double zi = (x[(int64_t)i * d + cp.cp_axis] - cp.cp_loc[a]) / cp.cp_width[a];
        const double zj = (x[(int64_t)j * d + cp.cp_axis] - cp.cp_loc[a]) / cp.cp_width[a];
        const double fi = 1.0 / (1.0 + exp(-zi)), fj = 1.0 / (1.0 + exp(-zj));
        const double dci = -fi * (1.0 - fi) / cp.cp_width[a], dcj = -fj * (1.0 - fj) / cp.cp_width[a];
        double ka = 0.0, kb = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < MAX_REG; ++q2) {
            if (q2 == a) ka = kreg[q2];
            if (q2 == a + 1) kb = kreg[q2];
        }
        for (int v = 0; v < 2; ++v) {
            const double dwi = v ? dci * zi : dci, dwj = v ? dcj * zj : dcj;
            const double AAt = -(dwi * (1.0 - fj) + dwj * (1.0 - fi));
            const double BBt = dwi * fj + dwj * fi;
            dK[plane * (cp.cp_theta_off + 2 * a + v) + e] = ka * AAt + kb * BBt;
        }
    }
    K[e] = ktot;
}

__global__ void cross_cov_kernel(const CovParams cp, const double* __restrict__ u, int m, const double* __restrict__ v,
                                 int n, double* __restrict__ out, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)m * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    const int d = cp.d;
    double d2[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) {
        const double df = (k < d) ? u[(int64_t)i * d + k] - v[(int64_t)j * d + k] : 0.0;
        d2[k] = df * df;
    }
    double wi[MAX_REG] = {0.0, 0.0, 0.0, 0.0}, wj[MAX_REG] = {0.0, 0.0, 0.0, 0.0};
    if (cp.n_regions) {
        region_weights(cp, u[(int64_t)i * d + cp.cp_axis], wi);
        region_weights(cp, v[(int64_t)j * d + cp.cp_axis], wj);
    }
    out[(int64_t)i * ld + j] = cov_from_d2(cp, d2, wi, wj);
}

// Stacked cross-covariance rows of one query chunk.  grid = (npad/128, ceil(mq/128)).
template <bool CP>
__global__ void __launch_bounds__(256) cross_stack_kernel(const CovParams cp, const double* __restrict__ q, int mq,
                                                          int nstack, const double* __restrict__ x, int n,
                                                          double* __restrict__ S, int64_t ld) {
    __shared__ double qs[TILE * MAX_DIM];
    __shared__ double gs[CP ? TILE : 1][MAX_REG];
    const int col0 = blockIdx.x * TILE, q0 = blockIdx.y * TILE;
    const int tid = threadIdx.x, col = tid & (TILE - 1), half = tid >> 7;
    const int d = cp.d;
    const int nq = min(TILE, mq - q0);
    for (int idx = tid; idx < nq * d; idx += 256) qs[idx] = q[(int64_t)q0 * d + idx];
    const int gj = col0 + col;
    double xj[MAX_DIM], wj[MAX_REG] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) xj[k] = (k < d && gj < n) ? x[(int64_t)gj * d + k] : 0.0;
    if (CP && cp.n_regions) {
        if (gj < n) region_weights(cp, x[(int64_t)gj * d + cp.cp_axis], wj);
        if (tid < nq) {
            double g[MAX_REG];
            region_weights(cp, q[(int64_t)(q0 + tid) * d + cp.cp_axis], g);
#pragma unroll
            for (int r = 0; r < MAX_REG; ++r) gs[tid][r] = g[r];
        }
    }
    __syncthreads();
    for (int r = 0; r < TILE / 2; ++r) {
        const int i = half * (TILE / 2) + r;
        if (i >= nq) break;
        double df[MAX_DIM], d2[MAX_DIM];
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) {
            df[k] = (k < d) ? xj[k] - qs[i * d + k] : 0.0;
            d2[k] = df[k] * df[k];
        }
        double wi[MAX_REG];
#pragma unroll
        for (int r2 = 0; r2 < MAX_REG; ++r2) wi[r2] = (CP && cp.n_regions) ? gs[CP ? i : 0][r2] : 0.0;
        const double kv = (gj < n) ? cov_from_d2<CP>(cp, d2, wi, wj) : 0.0;
        double* dst = S + (int64_t)(q0 + i) * nstack * ld + gj;
        dst[0] = kv;
        if (nstack > 1) {
#pragma unroll
            for (int k = 0; k < MAX_DIM; ++k)
                if (k < d) dst[(int64_t)(k + 1) * ld] = (df[k] * cp.inv_l2[0][k]) * kv;
        }
    }
}

// one warp per row: out[r] = S[r, 0:ncols] . vec
__global__ void __launch_bounds__(256) row_dot_kernel(const double* __restrict__ S, int64_t ld, int rows, int ncols,
                                                      const double* __restrict__ vec, double* __restrict__ out) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double2* row = reinterpret_cast<const double2*>(S + (int64_t)r * ld);
    const double2* v2 = reinterpret_cast<const double2*>(vec);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const int n2 = ncols >> 1;  // ncols is a multiple of 128
    for (int j = lane; j < n2; j += 64) {
        const double2 s0 = row[j], w0 = v2[j];
        const double2 s1 = row[j + 32], w1 = v2[j + 32];
        a0 = fma(s0.x, w0.x, a0);
        a1 = fma(s0.y, w0.y, a1);
        a2 = fma(s1.x, w1.x, a2);
        a3 = fma(s1.y, w1.y, a3);
    }
    double a = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[r] = a;
}

// out[c] = sum_r S[r][c] * vec[r]: stage 1, CTA (bx, by) reduces rows [1024 by, 1024 by + 1024) for columns 128 bx ..
// into part[by][c]; stage 2 adds the row chunks in order (fixed summation order, no atomics)
__global__ void __launch_bounds__(256) col_dot_partial_kernel(const double* __restrict__ S, int64_t ld, int nrows,
                                                              const double* __restrict__ vec, double* __restrict__ part,
                                                              int ncols) {
    __shared__ double sm[2][TILE];
    const int col = blockIdx.x * TILE + (threadIdx.x & (TILE - 1)), half = threadIdx.x >> 7;
    const int r0 = blockIdx.y * 1024 + half * 512;
    const int r1 = min(nrows, r0 + 512);
    double a0 = 0.0, a1 = 0.0;
    const double* p = S + (int64_t)r0 * ld + col;
    int r = r0;
    for (; r + 1 < r1; r += 2, p += 2 * ld) {
        a0 = fma(p[0], vec[r], a0);
        a1 = fma(p[ld], vec[r + 1], a1);
    }
    if (r < r1) a0 = fma(p[0], vec[r], a0);
    sm[half][threadIdx.x & (TILE - 1)] = a0 + a1;
    __syncthreads();
    if (half == 0) part[(int64_t)blockIdx.y * ncols + col] = sm[0][threadIdx.x] + sm[1][threadIdx.x];
}
__global__ void col_dot_final_kernel(const double* __restrict__ part, int nchunks, int ncols, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double a = 0.0;
    for (int k = 0; k < nchunks; ++k) a += part[(int64_t)k * ncols + c];
    out[c] = a;
}

// one warp per query: Gram matrix of its nstack solved rows
template <int NS>
__global__ void __launch_bounds__(256) row_gram_kernel(const double* __restrict__ X, int64_t ld, int mq, int nstack,
                                                       int ncols, double* __restrict__ G) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= mq) return;
    const double* base = X + (int64_t)qi * nstack * ld;
    double acc[NS * (NS + 1) / 2];
#pragma unroll
    for (int p = 0; p < NS * (NS + 1) / 2; ++p) acc[p] = 0.0;
    for (int j = lane; j < ncols; j += 32) {
        double v[NS];
#pragma unroll
        for (int a = 0; a < NS; ++a) v[a] = (a < nstack) ? base[(int64_t)a * ld + j] : 0.0;
        int p = 0;
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = a; b < NS; ++b) acc[p] = fma(v[a], v[b], acc[p]), ++p;
    }
    int p = 0;
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = a; b < NS; ++b) {
            double s = acc[p++];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && a < nstack && b < nstack) {
                G[((int64_t)qi * nstack + a) * nstack + b] = s;
                G[((int64_t)qi * nstack + b) * nstack + a] = s;
            }
        }
}

__device__ __forceinline__ double mean_at(const MeanParams& mp, const double* pt) {
    double m = mp.c0;
    if (mp.kind >= MEAN_LINEAR) {
        double lin = 0.0, quad = 0.0;
        for (int k = 0; k < mp.d; ++k) {
            const double dx = pt[k] - mp.xbar[k];
            lin += dx * mp.lin[k];
            if (mp.kind == MEAN_QUADRATIC) quad += (dx * dx) * mp.quad[k];
        }
        m += lin;
        if (mp.kind == MEAN_QUADRATIC) m += quad;
    }
    return m;
}

__global__ void finalize_predict_kernel(const CovParams cp, const MeanParams mp, const double* __restrict__ q, int mq,
                                        int nstack, const double* __restrict__ dots, const double* __restrict__ G,
                                        double* __restrict__ mu, double* __restrict__ sig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mq) return;
    mu[i] = dots[(int64_t)i * nstack] + mean_at(mp, q + (int64_t)i * mp.d);
    if (sig) {
        // k(q, q): a^2 of every smooth component (noise kernels contribute 0, covariance.py:160-161, 671-672),
        // weighted by the query point's squared region weight under a ChangePoint
        double g[MAX_REG] = {0.0, 0.0, 0.0, 0.0};
        if (cp.n_regions) region_weights(cp, q[(int64_t)i * cp.d + cp.cp_axis], g);
        double kqq = 0.0;
        for (int c = 0; c < cp.ncomp; ++c)
            if (cp.kind[c] <= COV_RQ) kqq += (cp.n_regions ? leaf_weight(cp, c, g, g) : 1.0) * cp.amp2[c];
        sig[i] = sqrt(fabs(kqq - G[(int64_t)i * nstack * nstack]));
    }
}

__global__ void finalize_gradient_kernel(const double* __restrict__ dots, const double* __restrict__ G, int mq, int d,
                                         const double* __restrict__ R, double* __restrict__ mean,
                                         double* __restrict__ cov) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mq) return;
    const int ns = d + 1;
    for (int a = 0; a < d; ++a) {
        mean[(int64_t)i * d + a] = dots[(int64_t)i * ns + 1 + a];
        for (int b = 0; b < d; ++b)
            cov[((int64_t)i * d + a) * d + b] = R[b] - G[((int64_t)i * ns + a + 1) * ns + b + 1];
    }
}

__global__ void finalize_spatial_kernel(const double* __restrict__ dots, const double* __restrict__ G, int mq, int d,
                                        double* __restrict__ dmu, double* __restrict__ dvar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mq) return;
    const int ns = d + 1;
    for (int a = 0; a < d; ++a) {
        dmu[(int64_t)i * d + a] = dots[(int64_t)i * ns + 1 + a];
        dvar[(int64_t)i * d + a] = -2.0 * G[(int64_t)i * ns * ns + a + 1];
    }
}

// Acquisition functions over a batch of candidates (acquisition.py:76-125 ExpectedImprovement, :169-189
// UpperConfidenceBound, :213-229 MaxVariance).  mode 0: value, 1: opt_func (the minimiser's objective), 2: opt_func and its
// gradient.  EI branches at Z < -3 exactly as the reference.
__global__ void acquisition_kernel(const double* __restrict__ mu, const double* __restrict__ sig,
                                   const double* __restrict__ dmu, const double* __restrict__ dvar, int m, int d, int kind,
                                   double param, int mode, double* __restrict__ out, double* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double s = sig[i];
    if (kind == 1) {  // UCB: param = kappa
        const double ucb = mu[i] + param * s;
        out[i] = mode == 0 ? ucb : -ucb;
        if (mode == 2)
            for (int a = 0; a < d; ++a)
                grad[(int64_t)i * d + a] = -(dmu[(int64_t)i * d + a] + 0.5 * param * dvar[(int64_t)i * d + a] / s);
        return;
    }
    if (kind == 2) {  // MaxVariance
        const double v = s * s;
        out[i] = mode == 0 ? v : -v;
        if (mode == 2)
            for (int a = 0; a < d; ++a) grad[(int64_t)i * d + a] = -dvar[(int64_t)i * d + a];
        return;
    }
    const double y_max = param;
    const double ir2pi = 0.3989422804014327, ir2 = 0.7071067811865476, rpi2 = 1.2533141373155003,
                 ln2pi = 1.8378770664093453;
    const double Z = (mu[i] - y_max) / s;
    if (Z < -3.0) {
        const double R = rpi2 * erfcx(-Z * ir2);
        const double H = 1.0 + Z * R;
        const double ln_ei = log(H) - 0.5 * (Z * Z + ln2pi) + log(s);
        if (mode == 0) out[i] = exp(ln_ei);
        else out[i] = -ln_ei;
        if (mode == 2)
            for (int a = 0; a < d; ++a)
                grad[(int64_t)i * d + a] = -((0.5 * dvar[(int64_t)i * d + a] / s + R * dmu[(int64_t)i * d + a]) / (H * s));
    } else {
        const double pdf = exp(-0.5 * Z * Z) * ir2pi;
        const double cdf = 0.5 * (1.0 + erf(Z * ir2));
        const double ei = s * (Z * cdf + pdf);
        if (mode == 0) out[i] = ei;
        else out[i] = -log(ei);
        if (mode == 2)
            for (int a = 0; a < d; ++a)
                grad[(int64_t)i * d + a] =
                    -((0.5 * pdf * dvar[(int64_t)i * d + a] / s + dmu[(int64_t)i * d + a] * cdf) / ei);
    }
}

// Index of the best entry of v[0..m): the largest (want_max) or the smallest; ties go to the lowest index and NaNs never
// win, exactly like a left-to-right host scan with a strict comparison.  Stage 1: one (value, index) pair per CTA;
// stage 2 (one CTA) reduces the pairs.  No atomics: fixed order, bit-reproducible.
__device__ __forceinline__ bool better(double a, int64_t ia, double b, int64_t ib, bool want_max) {
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    const bool gt = want_max ? a > b : a < b;
    return gt || (a == b && ia < ib);
}
__global__ void __launch_bounds__(256) argbest_kernel(const double* __restrict__ v, const int64_t* __restrict__ idx_in,
                                                      int64_t m, int want_max, double* __restrict__ val_out,
                                                      int64_t* __restrict__ idx_out) {
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    double bv = 0.0;
    int64_t bi = -1;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < m; i += (int64_t)gridDim.x * 256) {
        const double x = v[i];
        const int64_t ix = idx_in ? idx_in[i] : i;
        if (x != x || ix < 0) continue;
        if (better(x, ix, bv, bi, want_max != 0)) {
            bv = x;
            bi = ix;
        }
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o && better(sv[threadIdx.x + o], si[threadIdx.x + o], sv[threadIdx.x], si[threadIdx.x], want_max != 0)) {
            sv[threadIdx.x] = sv[threadIdx.x + o];
            si[threadIdx.x] = si[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        val_out[blockIdx.x] = sv[0];
        idx_out[blockIdx.x] = si[0];
    }
}

__global__ void residual_kernel(const MeanParams mp, const double* __restrict__ x, const double* __restrict__ y, int n,
                                int npad, double* __restrict__ resid, double* __restrict__ mu_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    if (i < n) {
        const double m = mean_at(mp, x + (int64_t)i * mp.d);
        resid[i] = y[i] - m;
        if (mu_out) mu_out[i] = m;
    } else {
        resid[i] = 0.0;
        if (mu_out) mu_out[i] = 0.0;
    }
}

__global__ void copy2d_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                              int rows, int cols2) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * cols2) return;
    const int r = (int)(e / cols2), c = (int)(e % cols2);
    reinterpret_cast<double2*>(dst + (int64_t)r * ldd)[c] = reinterpret_cast<const double2*>(src + (int64_t)r * lds)[c];
}

}  // namespace

#define GPB_LAUNCH_CHECK()          \
    GPB_CUDA(cudaGetLastError());   \
    count_launch();                 \
    return 0

int launch_assemble_train(const CovParams& cp, const double* x, int n, int npad, const double* noise_var,
                          const double* y_cov, double* K, int64_t ld, int mirror, cudaStream_t s) {
    const int nb = npad / TILE;
    if (cp.n_regions) assemble_train_kernel<true><<<nb * (nb + 1) / 2, 256, 0, s>>>(cp, x, n, noise_var, y_cov, K, ld, mirror);
    else assemble_train_kernel<false><<<nb * (nb + 1) / 2, 256, 0, s>>>(cp, x, n, noise_var, y_cov, K, ld, mirror);
    GPB_LAUNCH_CHECK();
}

int launch_assemble_block(const CovParams& cp, const double* x, int n, const double* noise_var, int row0, int nrows,
                          int col0, int ncols, double* out, int64_t ld, cudaStream_t s) {
    dim3 grid(ncols / TILE, nrows / TILE);
    if (cp.n_regions) assemble_block_kernel<true><<<grid, 256, 0, s>>>(cp, x, n, noise_var, row0, col0, out, ld);
    else assemble_block_kernel<false><<<grid, 256, 0, s>>>(cp, x, n, noise_var, row0, col0, out, ld);
    GPB_LAUNCH_CHECK();
}

int launch_assemble_grads(const CovParams& cp, const double* x, int n, double* K, double* dK, cudaStream_t s) {
    const int64_t e = (int64_t)n * n;
    assemble_grads_kernel<<<(unsigned)((e + 255) / 256), 256, 0, s>>>(cp, x, n, K, dK);
    GPB_LAUNCH_CHECK();
}

int launch_cross_cov(const CovParams& cp, const double* u, int m, const double* v, int n, double* out, int64_t ld,
                     cudaStream_t s) {
    const int64_t e = (int64_t)m * n;
    if (e == 0) return 0;
    cross_cov_kernel<<<(unsigned)((e + 255) / 256), 256, 0, s>>>(cp, u, m, v, n, out, ld);
    GPB_LAUNCH_CHECK();
}

int launch_cross_stack(const CovParams& cp, const double* q, int mq, int nstack, const double* x, int n, int npad,
                       double* S, int64_t ld, cudaStream_t s) {
    dim3 grid(npad / TILE, (mq + TILE - 1) / TILE);
    if (cp.n_regions) cross_stack_kernel<true><<<grid, 256, 0, s>>>(cp, q, mq, nstack, x, n, S, ld);
    else cross_stack_kernel<false><<<grid, 256, 0, s>>>(cp, q, mq, nstack, x, n, S, ld);
    GPB_LAUNCH_CHECK();
}

int launch_row_dot(const double* S, int64_t ld, int rows, int ncols, const double* vec, double* out, cudaStream_t s) {
    row_dot_kernel<<<(rows + 7) / 8, 256, 0, s>>>(S, ld, rows, ncols, vec, out);
    GPB_LAUNCH_CHECK();
}

size_t col_dot_ws_bytes(int nrows, int ncols) { return sizeof(double) * (size_t)((nrows + 1023) / 1024) * ncols; }

int launch_col_dot(const double* S, int64_t ld, int nrows, int ncols, const double* vec, double* out, double* ws,
                   cudaStream_t s) {
    const int nchunks = (nrows + 1023) / 1024;
    dim3 grid(ncols / TILE, nchunks);
    col_dot_partial_kernel<<<grid, 256, 0, s>>>(S, ld, nrows, vec, ws, ncols);
    GPB_CUDA(cudaGetLastError());
    col_dot_final_kernel<<<(ncols + 255) / 256, 256, 0, s>>>(ws, nchunks, ncols, out);
    GPB_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}

int launch_row_gram(const double* X, int64_t ld, int mq, int nstack, int ncols, double* G, cudaStream_t s) {
    const int grid = (mq + 7) / 8;
    if (nstack == 1) row_gram_kernel<1><<<grid, 256, 0, s>>>(X, ld, mq, nstack, ncols, G);
    else if (nstack <= 3) row_gram_kernel<3><<<grid, 256, 0, s>>>(X, ld, mq, nstack, ncols, G);
    else if (nstack <= 6) row_gram_kernel<6><<<grid, 256, 0, s>>>(X, ld, mq, nstack, ncols, G);
    else row_gram_kernel<MAX_DIM + 1><<<grid, 256, 0, s>>>(X, ld, mq, nstack, ncols, G);
    GPB_LAUNCH_CHECK();
}

int launch_finalize_predict(const CovParams& cp, const MeanParams& mp, const double* q, int mq, int nstack,
                            const double* dots, const double* G, double* mu, double* sig, cudaStream_t s) {
    finalize_predict_kernel<<<(mq + 255) / 256, 256, 0, s>>>(cp, mp, q, mq, nstack, dots, G, mu, sig);
    GPB_LAUNCH_CHECK();
}

int launch_finalize_gradient(const double* dots, const double* G, int mq, int d, const double* R_dev, double* mean,
                             double* cov, cudaStream_t s) {
    finalize_gradient_kernel<<<(mq + 255) / 256, 256, 0, s>>>(dots, G, mq, d, R_dev, mean, cov);
    GPB_LAUNCH_CHECK();
}

int launch_finalize_spatial(const double* dots, const double* G, int mq, int d, double* dmu, double* dvar,
                            cudaStream_t s) {
    finalize_spatial_kernel<<<(mq + 255) / 256, 256, 0, s>>>(dots, G, mq, d, dmu, dvar);
    GPB_LAUNCH_CHECK();
}

int launch_acquisition(const double* mu, const double* sig, const double* dmu, const double* dvar, int m, int d, int kind,
                       double param, int mode, double* out, double* grad, cudaStream_t s) {
    if (m == 0) return 0;
    acquisition_kernel<<<(m + 255) / 256, 256, 0, s>>>(mu, sig, dmu, dvar, m, d, kind, param, mode, out, grad);
    GPB_LAUNCH_CHECK();
}

int launch_argbest(const double* v, int64_t m, int want_max, double* ws_val, int64_t* ws_idx, cudaStream_t s) {
    const int blocks = (int)std::min<int64_t>(ARGBEST_BLOCKS, (m + 255) / 256);
    argbest_kernel<<<blocks, 256, 0, s>>>(v, nullptr, m, want_max, ws_val + 1, ws_idx + 1);
    GPB_CUDA(cudaGetLastError());
    argbest_kernel<<<1, 256, 0, s>>>(ws_val + 1, ws_idx + 1, blocks, want_max, ws_val, ws_idx);
    GPB_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}

int launch_residual(const MeanParams& mp, const double* x, const double* y, int n, int npad, double* resid,
                    double* mu_out, cudaStream_t s) {
    residual_kernel<<<(npad + 255) / 256, 256, 0, s>>>(mp, x, y, n, npad, resid, mu_out);
    GPB_LAUNCH_CHECK();
}

int launch_copy2d(const double* src, int64_t lds, double* dst, int64_t ldd, int rows, int cols, cudaStream_t s) {
    const int64_t e = (int64_t)rows * (cols / 2);
    if (e == 0) return 0;
    copy2d_kernel<<<(unsigned)((e + 255) / 256), 256, 0, s>>>(src, lds, dst, ldd, rows, cols / 2);
    GPB_LAUNCH_CHECK();
}

}  // namespace gpb
