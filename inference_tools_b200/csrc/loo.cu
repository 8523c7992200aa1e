// Leave-one-out cross-validation objective (regression.py:451-526, Rasmussen & Williams eqs. 5.10-5.14):
//   var_i = 1 / Kinv_ii,  LOO = -1/2 sum_i (var_i alpha_i^2 + log var_i)
//   dLOO/dtheta_p = sum_i ( c1_i (Z_p alpha)_i - c2_i (Z_p Kinv)_ii ),  Z_p = Kinv dK_p,
//   c1 = alpha var, c2 = 1/2 var (1 + var alpha^2);  mean parameters: sum_i c1_i (Kinv dmu_p)_i.
// The reference forms Z_p = iK @ dK and Z_p @ iK as two dense N^3 products per parameter
// (regression.py:511-514).  Here: one DMMA GEMM T = Kinv dK_p per smooth-kernel parameter, after which both
// terms are row reductions of T (against alpha and against Kinv); noise-kernel parameters reduce to O(N^2)
// row reductions of Kinv and never touch a GEMM.
#include "kernels.cuh"

namespace gpb {
namespace {

constexpr int TILE = 128;

// mirror the lower triangle into the upper one (32 x 32 tiles through shared memory)
__global__ void symmetrize_kernel(double* __restrict__ A, int64_t ld, int n) {
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int r0 = bi * 32, c0 = bj * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = A[(int64_t)(r0 + r) * ld + c0 + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int gi = c0 + r, gj = r0 + threadIdx.x;  // transposed position
        if (gj > gi) A[(int64_t)gi * ld + gj] = t[threadIdx.x][r];
    }
}

// dense symmetric dK/dtheta plane of smooth component c (slot 0: ln a, 1: ln alpha (RQ), 2+k: ln l_k); padded entries 0
__global__ void __launch_bounds__(256) dk_plane_kernel(const CovParams cp, int c, int slot, const double* __restrict__ x,
                                                       int n, double* __restrict__ out, int64_t ld) {
    __shared__ double xs[TILE * MAX_DIM];
    const int row0 = blockIdx.y * TILE, col0 = blockIdx.x * TILE;
    const int tid = threadIdx.x, col = tid & (TILE - 1), half = tid >> 7;
    const int d = cp.d;
    for (int idx = tid; idx < TILE * d; idx += 256) xs[idx] = x[(int64_t)row0 * d + idx];
    const int gj = col0 + col;
    double xj[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) xj[k] = (k < d) ? x[(int64_t)gj * d + k] : 0.0;
    __syncthreads();
    const bool is_rq = cp.kind[c] == COV_RQ;
    const double amp2 = cp.amp2[c], q = cp.rq_alpha[c];
    for (int r = 0; r < TILE / 2; ++r) {
        const int i = half * (TILE / 2) + r;
        const int gi = row0 + i;
        double z = 0.0, sk = 0.0;
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) {
            const double df = (k < d) ? xs[i * d + k] - xj[k] : 0.0;
            const double s = (0.5 * df * df) * cp.inv_l2[c][k < d ? k : 0];
            if (k < d) z += s;
            if (k == slot - 2) sk = s;
        }
        double v;
        if (!is_rq) {
            const double kv = amp2 * (exp(-z) + (gi == gj ? cp.jitter : 0.0));
            v = (slot == 0) ? 2.0 * kv : 2.0 * sk * kv;                     // covariance.py:273-275
        } else {
            const double F = 1.0 + z / q, lnF = log(F);
            const double kv = amp2 * (exp(-q * lnF) + (gi == gj ? cp.jitter : 0.0));
            v = (slot == 0) ? 2.0 * kv : (slot == 1 ? -kv * (lnF * q - z / F) : (2.0 * kv / F) * sk);  // :360-364
        }
        if (gi >= n || gj >= n) v = 0.0;
        out[(int64_t)gi * ld + gj] = v;
    }
}

// out[i] = sum_k A[i][k] * B[i][k] * (w ? w[k] : 1)   (warp per row)
__global__ void __launch_bounds__(256) rowdot2_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                      int64_t ld, int rows, int ncols, const double* __restrict__ w,
                                                      double* __restrict__ out) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* a = A + (int64_t)r * ld;
    const double* b = B + (int64_t)r * ld;
    double s0 = 0.0, s1 = 0.0;
    for (int k = lane; k < ncols; k += 64) {
        s0 = fma(a[k] * b[k], w ? w[k] : 1.0, s0);
        if (k + 32 < ncols) s1 = fma(a[k + 32] * b[k + 32], w ? w[k + 32] : 1.0, s1);
    }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
}

// per-point LOO terms from diag(Kinv) and alpha; single CTA also reduces the objective into val[0]
__global__ void __launch_bounds__(1024) loo_terms_kernel(const double* __restrict__ Kinv, int64_t ld,
                                                         const double* __restrict__ alpha, int n, int npad,
                                                         double* __restrict__ var, double* __restrict__ c1,
                                                         double* __restrict__ c2, double* __restrict__ val) {
    __shared__ double sm[1024];
    const int tid = threadIdx.x;
    double acc = 0.0;
    for (int i = tid; i < npad; i += 1024) {
        if (i < n) {
            const double v = 1.0 / Kinv[(int64_t)i * ld + i], a = alpha[i];
            var[i] = v;
            c1[i] = a * v;
            c2[i] = 0.5 * v * (1.0 + v * a * a);
            acc += v * a * a + log(v);
        } else {
            var[i] = 0.0;
            c1[i] = 0.0;
            c2[i] = 0.0;
        }
    }
    sm[tid] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) sm[tid] += sm[tid + o];
        __syncthreads();
    }
    if (tid == 0) val[0] = -0.5 * sm[0];
}

// g = scale * sum_i (c1_i * za_i - c2_i * dz_i)   (dz may be null)
__global__ void __launch_bounds__(1024) loo_grad_reduce_kernel(const double* __restrict__ c1, const double* __restrict__ c2,
                                                               const double* __restrict__ za, const double* __restrict__ dz,
                                                               int n, double scale, double* __restrict__ g) {
    __shared__ double sm[1024];
    const int tid = threadIdx.x;
    double acc = 0.0;
    for (int i = tid; i < n; i += 1024) acc += c1[i] * za[i] - (dz ? c2[i] * dz[i] : 0.0);
    sm[tid] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) sm[tid] += sm[tid + o];
        __syncthreads();
    }
    if (tid == 0) g[0] = scale * sm[0];
}

// Hetero: g_i = 2 sigma_i^2 (alpha_i u_i - w_i), u = Kinv c1, w_i = sum_j Kinv_ij^2 c2_j
__global__ void loo_hetero_kernel(const double* __restrict__ log_sigma, const double* __restrict__ alpha,
                                  const double* __restrict__ u, const double* __restrict__ w, int n,
                                  double* __restrict__ g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = 2.0 * exp(2.0 * log_sigma[i]) * (alpha[i] * u[i] - w[i]);
}

// dmu_p/dtheta column p of the mean function (mean.py:50-51, 80-83, 122-126)
__global__ void mean_grad_vec_kernel(const MeanParams mp, int p, const double* __restrict__ x, int n, int npad,
                                     double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double g = 0.0;
    if (i < n) {
        g = 1.0;
        if (p >= 1) {
            const int k = (p - 1) % mp.d;
            const double dx = x[(int64_t)i * mp.d + k] - mp.xbar[k];
            g = (p <= mp.d) ? dx : dx * dx;
        }
    }
    out[i] = g;
}

__global__ void loo_predictions_kernel(const double* __restrict__ Kinv, int64_t ld, const double* __restrict__ alpha,
                                       const double* __restrict__ y, int n, double* __restrict__ mu,
                                       double* __restrict__ sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = 1.0 / Kinv[(int64_t)i * ld + i];
    mu[i] = y[i] - alpha[i] * v;   // regression.py:464
    sigma[i] = sqrt(v);
}

}  // namespace

#define GPB_LAUNCH_OK()           \
    GPB_CUDA(cudaGetLastError()); \
    count_launch()

int launch_symmetrize(double* A, int64_t ld, int n, cudaStream_t s) {
    dim3 grid(n / 32, n / 32), block(32, 8);
    symmetrize_kernel<<<grid, block, 0, s>>>(A, ld, n);
    GPB_LAUNCH_OK();
    return 0;
}

int launch_loo_predictions(const double* Kinv, int64_t ld, const double* alpha, const double* y, int n, double* mu,
                           double* sigma, cudaStream_t s) {
    loo_predictions_kernel<<<(n + 255) / 256, 256, 0, s>>>(Kinv, ld, alpha, y, n, mu, sigma);
    GPB_LAUNCH_OK();
    return 0;
}

// ws: 5 vectors of npad doubles (var, c1, c2, t1, t2); Dk, T: npad x npad scratch (only when grad_dev != nullptr).
// dK_all (n_cov dense n x n planes from launch_assemble_grads) is only used for models with a ChangePoint: there the
// per-parameter planes come from the generic gradient kernel instead of dk_plane_kernel.
int launch_loo(const CovParams& cp, const MeanParams& mp, int n_theta_mean, const double* x, int n, int npad,
               const double* alpha, double* Kinv, int64_t ld, double* ws, double* Dk, double* T, double* val_dev,
               double* grad_dev, const double* dK_all, int n_cov, cudaStream_t s) {
    double *var = ws, *c1 = ws + npad, *c2 = ws + 2 * (size_t)npad, *t1 = ws + 3 * (size_t)npad, *t2 = ws + 4 * (size_t)npad;
    loo_terms_kernel<<<1, 1024, 0, s>>>(Kinv, ld, alpha, n, npad, var, c1, c2, val_dev);
    GPB_LAUNCH_OK();
    if (!grad_dev) return 0;
    GPB_TRY(launch_symmetrize(Kinv, ld, npad, s));
    // mean parameters: g = sum c1 * (Kinv dmu)
    const int n_mean = n_theta_mean;
    for (int p = 0; p < n_mean; ++p) {
        mean_grad_vec_kernel<<<(npad + 255) / 256, 256, 0, s>>>(mp, p, x, n, npad, t1);
        GPB_LAUNCH_OK();
        GPB_TRY(launch_row_dot(Kinv, ld, n, npad, t1, t2, s));
        loo_grad_reduce_kernel<<<1, 1024, 0, s>>>(c1, c2, t2, nullptr, n, 1.0, grad_dev + p);
        GPB_LAUNCH_OK();
    }
    if (cp.n_regions) {  // regression.py:497-523 with dK_p = covariance_and_gradients planes, one GEMM Kinv dK_p each
        for (int p = 0; p < n_cov; ++p) {
            GPB_CUDA(cudaMemsetAsync(Dk, 0, sizeof(double) * (size_t)npad * ld, s));
            GPB_CUDA(cudaMemcpy2DAsync(Dk, sizeof(double) * ld, dK_all + (size_t)p * n * n, sizeof(double) * n,
                                       sizeof(double) * n, n, cudaMemcpyDeviceToDevice, s));
            GemmArgs gm{npad, npad, npad, Kinv, ld, Dk, ld, nullptr, 0, T, ld, nullptr, 0, 1.0, 0.0, GEMM_FULL};
            GPB_TRY(gemm_nt(gm, s));
            GPB_TRY(launch_row_dot(T, ld, n, npad, alpha, t1, s));
            rowdot2_kernel<<<(n + 7) / 8, 256, 0, s>>>(T, Kinv, ld, n, npad, nullptr, t2);
            GPB_LAUNCH_OK();
            loo_grad_reduce_kernel<<<1, 1024, 0, s>>>(c1, c2, t1, t2, n, 1.0, grad_dev + n_theta_mean + p);
            GPB_LAUNCH_OK();
        }
        return 0;
    }
    for (int c = 0; c < cp.ncomp; ++c) {
        double* g = grad_dev + n_theta_mean + cp.theta_off[c];
        if (cp.kind[c] <= COV_RQ) {
            const bool rq = cp.kind[c] == COV_RQ;
            const int nslots = 2 + cp.d;
            for (int slot = 0; slot < nslots; ++slot) {
                if (slot == 1 && !rq) continue;
                dim3 grid(npad / TILE, npad / TILE);
                dk_plane_kernel<<<grid, 256, 0, s>>>(cp, c, slot, x, n, Dk, ld);
                GPB_LAUNCH_OK();
                // T = Kinv dK (dK symmetric => NT form with B = dK)
                GemmArgs gm{npad, npad, npad, Kinv, ld, Dk, ld, nullptr, 0, T, ld, nullptr, 0, 1.0, 0.0, GEMM_FULL};
                GPB_TRY(gemm_nt(gm, s));
                GPB_TRY(launch_row_dot(T, ld, n, npad, alpha, t1, s));                         // (Z alpha)_i
                rowdot2_kernel<<<(n + 7) / 8, 256, 0, s>>>(T, Kinv, ld, n, npad, nullptr, t2);  // (Z Kinv)_ii
                GPB_LAUNCH_OK();
                const int idx = rq ? slot : (slot == 0 ? 0 : slot - 1);
                loo_grad_reduce_kernel<<<1, 1024, 0, s>>>(c1, c2, t1, t2, n, 1.0, g + idx);
                GPB_LAUNCH_OK();
            }
        } else if (cp.kind[c] == COV_WHITE) {
            // dK = 2 sigma^2 I: (Z alpha)_i = 2 s2 (Kinv alpha)_i, (Z Kinv)_ii = 2 s2 sum_k Kinv_ik^2
            GPB_TRY(launch_row_dot(Kinv, ld, n, npad, alpha, t1, s));
            rowdot2_kernel<<<(n + 7) / 8, 256, 0, s>>>(Kinv, Kinv, ld, n, npad, nullptr, t2);
            GPB_LAUNCH_OK();
            loo_grad_reduce_kernel<<<1, 1024, 0, s>>>(c1, c2, t1, t2, n, 2.0 * cp.amp2[c], g);
            GPB_LAUNCH_OK();
        } else {
            // dK_i = 2 sigma_i^2 e_i e_i^T: g_i = 2 sigma_i^2 (alpha_i (Kinv c1)_i - sum_j Kinv_ij^2 c2_j)
            GPB_TRY(launch_row_dot(Kinv, ld, n, npad, c1, t1, s));
            rowdot2_kernel<<<(n + 7) / 8, 256, 0, s>>>(Kinv, Kinv, ld, n, npad, c2, t2);
            GPB_LAUNCH_OK();
            loo_hetero_kernel<<<(n + 255) / 256, 256, 0, s>>>(cp.hetero_log_sigma, alpha, t1, t2, n, g);
            GPB_LAUNCH_OK();
        }
    }
    return 0;
}

}  // namespace gpb
