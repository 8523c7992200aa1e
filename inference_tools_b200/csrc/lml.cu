// Fused trace / quadratic terms of the log-marginal-likelihood gradient (regression.py:563-566):
//     dLML/dtheta_mean,i = sum_n alpha_n dmu_n/dtheta_i
//     dLML/dtheta_cov,p  = 1/2 sum_ij (alpha_i alpha_j - Kinv_ij) dK_p,ij
// The reference materialises every dK_p as a dense N x N array (covariance.py:273-276, 356-364) and
// makes one pass per parameter.  Here dK_p,ij is recomputed from the coordinates inside ONE pass over
// the lower tiles of Kinv per smooth component (HBM-read bound: 4 N^2 bytes), with per-tile partial sums
// reduced in a fixed order (bit-reproducible).  Noise kernels only need diag(alpha alpha^T - Kinv).
#include "kernels.cuh"
#include "cov_device.cuh"

namespace gpb {
namespace {

constexpr int TILE = 128;
constexpr int NACC = MAX_DIM + 3;  // [ln a, (ln alpha_rq), ln l_1..l_d], then the Frobenius bound of the gradient planes
constexpr int FRO = MAX_DIM + 2;   // sum_ij max_p dK_p,ij^2 (error guard of the INT8 inverse chain, api.cu)

__device__ __forceinline__ void lower_tile(int t, int& bi, int& bj) {
    int r = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while (r * (r + 1) / 2 > t) --r;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    bi = r;
    bj = t - r * (r + 1) / 2;
}

// One 128 x 128 tile (global rows row0.., columns col0..) of the trace of smooth component `c`; `kinv_at(i, gj)` returns
// Kinv at tile row i, global column gj.  out[NACC] receives the tile's partial sums.
template <class Fetch>
__device__ __forceinline__ void trace_smooth_tile(const CovParams& cp, int c, const double* __restrict__ x, int n,
                                                  const double* __restrict__ alpha, int row0, int col0, Fetch kinv_at,
                                                  double* __restrict__ out) {
    __shared__ double xs[TILE * MAX_DIM];
    __shared__ double as[TILE];
    __shared__ double ws[TILE];   // ChangePoint weight of this leaf's region at the row points
    __shared__ double red[8][NACC];
    const int tid = threadIdx.x, col = tid & (TILE - 1), half = tid >> 7;
    const int d = cp.d;
    for (int idx = tid; idx < TILE * d; idx += 256) xs[idx] = x[(int64_t)row0 * d + idx];
    if (tid < TILE) as[tid] = alpha[row0 + tid];
    const int gj = col0 + col;
    double wcol = 1.0;
    if (cp.n_regions && cp.region[c] >= 0) {
        double g[MAX_REG];
        region_weights(cp, x[(int64_t)gj * d + cp.cp_axis], g);
        wcol = pick_region(g, cp.region[c]);
        if (tid < TILE) {
            region_weights(cp, x[(int64_t)(row0 + tid) * d + cp.cp_axis], g);
            ws[tid] = pick_region(g, cp.region[c]);
        }
    } else if (tid < TILE) {
        ws[tid] = 1.0;
    }
    double xj[MAX_DIM], il2[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) {
        xj[k] = (k < d) ? x[(int64_t)gj * d + k] : 0.0;
        il2[k] = (k < d) ? cp.inv_l2[c][k] : 0.0;
    }
    const double aj = alpha[gj];
    const double amp2 = cp.amp2[c], q = cp.rq_alpha[c];
    const bool is_rq = cp.kind[c] == COV_RQ;
    double acc[NACC];
#pragma unroll
    for (int p = 0; p < NACC; ++p) acc[p] = 0.0;
    __syncthreads();
    for (int r = 0; r < TILE / 2; ++r) {
        const int i = half * (TILE / 2) + r;
        const int gi = row0 + i;
        if (gi < gj || gi >= n || gj >= n) continue;
        const double w = (gi == gj) ? 1.0 : 2.0;  // symmetry: strict lower counted twice
        // dK of a leaf under a ChangePoint carries the region coefficient g_r(x_i) g_r(x_j) (covariance.py:570-572)
        const double Q = w * (ws[i] * wcol) * (as[i] * aj - kinv_at(i, gj));
        double s[MAX_DIM];  // 0.5 dx^2 / l^2 per dimension
        double z = 0.0;
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) {
            const double df = (k < d) ? xs[i * d + k] - xj[k] : 0.0;
            s[k] = (0.5 * df * df) * il2[k];
            z += s[k];
        }
        if (!is_rq) {
            // covariance.py:271-275: grads = [2K, (dx_k^2 / l_k^2) K]
            const double kv = amp2 * (exp(-z) + (gi == gj ? cp.jitter : 0.0));
            const double qk = Q * kv;
            acc[0] += qk;  // 0.5 * Q * 2K
#pragma unroll
            for (int k = 0; k < MAX_DIM; ++k) acc[2 + k] = fma(qk, s[k], acc[2 + k]);  // 0.5 * Q * 2 s_k K
            if ((r & 3) == 0) {     // Frobenius bound of the gradient planes from every 4th row (an estimate is enough)
                double gmax = 1.0;  // max_p |dK_p| / (2 K)
#pragma unroll
                for (int k = 0; k < MAX_DIM; ++k) gmax = fmax(gmax, s[k]);
                const double dk = 2.0 * kv * gmax * (ws[i] * wcol);
                acc[FRO] = fma(4.0 * w * dk, dk, acc[FRO]);
            }
        } else {
            // covariance.py:356-364: F = 1 + Z/q; grads = [2K, -K (q ln F - Z/F), (2K/F) s_k]
            const double F = 1.0 + z / q, lnF = log(F);
            const double kv = amp2 * (exp(-q * lnF) + (gi == gj ? cp.jitter : 0.0));
            const double qk = Q * kv;
            acc[0] += qk;
            acc[1] = fma(-0.5 * qk, lnF * q - z / F, acc[1]);
            const double qkf = qk / F;
#pragma unroll
            for (int k = 0; k < MAX_DIM; ++k) acc[2 + k] = fma(qkf, s[k], acc[2 + k]);
            if ((r & 3) == 0) {
                double gmax = fmax(1.0, 0.5 * fabs(lnF * q - z / F));
#pragma unroll
                for (int k = 0; k < MAX_DIM; ++k) gmax = fmax(gmax, s[k] / F);
                const double dk = 2.0 * kv * gmax * (ws[i] * wcol);
                acc[FRO] = fma(4.0 * w * dk, dk, acc[FRO]);
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int p = 0; p < NACC; ++p) {
        double v = acc[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][p] = v;
    }
    __syncthreads();
    if (tid < NACC) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][tid];
        out[tid] = v;
    }
}

// one smooth component `c`, dense Kinv (lower tiles); partials[tile][NACC]
__global__ void __launch_bounds__(256) trace_smooth_kernel(const CovParams cp, int c, const double* __restrict__ x,
                                                           int n, const double* __restrict__ alpha,
                                                           const double* __restrict__ Kinv, int64_t ld,
                                                           double* __restrict__ partials) {
    int bi, bj;
    lower_tile(blockIdx.x, bi, bj);
    const int row0 = bi * TILE, col0 = bj * TILE;
    const double* base = Kinv + (int64_t)row0 * ld;
    trace_smooth_tile(cp, c, x, n, alpha, row0, col0,
                      [=](int i, int gj) { return base[(int64_t)i * ld + gj]; }, partials + (int64_t)blockIdx.x * NACC);
}

// Distributed layout (dist.cu): this rank holds the row blocks a = me + G * idx (width nbd) of Kinv.  Row i of block a is
// row idx * nbd + i of the stack `Kst` (leading dimension ldst, column = global column, valid LEFT of the diagonal block);
// the diagonal blocks are in Kdiag[idx] (nbd x nbd, lower tiles valid).  grid = stack row tiles x column tiles; tiles right
// of the diagonal block write zero partials.
__global__ void __launch_bounds__(256) trace_smooth_stacked_kernel(const CovParams cp, int c, const double* __restrict__ x,
                                                                   int n, const double* __restrict__ alpha,
                                                                   const double* __restrict__ Kst, int64_t ldst,
                                                                   const double* __restrict__ Kdiag, int nbd, int me, int G,
                                                                   int tiles_n, double* __restrict__ partials) {
    const int ti = blockIdx.x / tiles_n, tj = blockIdx.x - ti * tiles_n;
    const int per_blk = nbd / TILE, idx = ti / per_blk, a = me + G * idx;
    const int lrow0 = (ti - idx * per_blk) * TILE;             // first row of the tile inside its row block
    const int row0 = a * nbd + lrow0, col0 = tj * TILE;
    double* out = partials + (int64_t)blockIdx.x * NACC;
    if (col0 > row0) {                                          // uniform per CTA
        if (threadIdx.x < NACC) out[threadIdx.x] = 0.0;
        return;
    }
    const bool diag = col0 >= a * nbd;  // inside the diagonal block: columns are local to Kdiag[idx]
    const double* base = diag ? Kdiag + (int64_t)idx * nbd * nbd + (int64_t)lrow0 * nbd - (int64_t)a * nbd
                              : Kst + ((int64_t)idx * nbd + lrow0) * ldst;
    const int64_t ldk = diag ? nbd : ldst;
    trace_smooth_tile(cp, c, x, n, alpha, row0, col0, [=](int i, int gj) { return base[(int64_t)i * ldk + gj]; }, out);
}

// grad[off + map(p)] = sum over tiles of partials[tile][p]; one CTA per accumulator slot
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partials, int ntiles, int d,
                                                              int is_rq, int off, double* __restrict__ grad,
                                                              double* __restrict__ fro2_out) {
    __shared__ double sm[256];
    const int p = blockIdx.x;  // accumulator slot
    if (p == 1 && !is_rq) return;
    if (p >= 2 + d && p != FRO) return;
    if (p == FRO && fro2_out == nullptr) return;
    double v = 0.0;
    for (int t = threadIdx.x; t < ntiles; t += 256) v += partials[(int64_t)t * NACC + p];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (p == FRO) {
            *fro2_out = sm[0];
        } else {
            const int idx = is_rq ? p : (p == 0 ? 0 : p - 1);
            grad[off + idx] = sm[0];
        }
    }
}

// change-point parameters (location, width) of change-point `a` (covariance.py:574-583):
//   dK = K_a (A + A^T) + K_{a+1} (B + B^T),  A = -dw (1 - f)^T,  B = dw f^T,  dw = df/dc or df/dwidth,
// with K_a the UNWEIGHTED covariance of region a (its leaves summed, own diagonal terms included), as the reference
// computes it.  partials[tile][0..1] = 1/2 sum Q dK for (location, width).
__global__ void __launch_bounds__(256) trace_cp_kernel(const CovParams cp, int a, const double* __restrict__ x, int n,
                                                       const double* __restrict__ alpha,
                                                       const double* __restrict__ Kinv, int64_t ld,
                                                       double* __restrict__ partials) {
    __shared__ double xs[TILE * MAX_DIM];
    __shared__ double as[TILE];
    __shared__ double fs[TILE], zs[TILE];
    __shared__ double red[8][2];
    int bi, bj;
    lower_tile(blockIdx.x, bi, bj);
    const int row0 = bi * TILE, col0 = bj * TILE;
    const int tid = threadIdx.x, col = tid & (TILE - 1), half = tid >> 7;
    const int d = cp.d;
    for (int idx = tid; idx < TILE * d; idx += 256) xs[idx] = x[(int64_t)row0 * d + idx];
    const double loc = cp.cp_loc[a], width = cp.cp_width[a];
    if (tid < TILE) {
        as[tid] = alpha[row0 + tid];
        const double z = (x[(int64_t)(row0 + tid) * d + cp.cp_axis] - loc) / width;
        zs[tid] = z;
        fs[tid] = 1.0 / (1.0 + exp(-z));
    }
    const int gj = col0 + col;
    double xj[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) xj[k] = (k < d) ? x[(int64_t)gj * d + k] : 0.0;
    const double zj = (x[(int64_t)gj * d + cp.cp_axis] - loc) / width;
    const double fj = 1.0 / (1.0 + exp(-zj));
    const double dcj = -fj * (1.0 - fj) / width;
    const double aj = alpha[gj];
    double acc0 = 0.0, acc1 = 0.0;
    __syncthreads();
    for (int r = 0; r < TILE / 2; ++r) {
        const int i = half * (TILE / 2) + r;
        const int gi = row0 + i;
        if (gi < gj || gi >= n || gj >= n) continue;
        const double Q = ((gi == gj) ? 0.5 : 1.0) * (as[i] * aj - Kinv[(int64_t)gi * ld + gj]);  // 1/2 * symmetry weight
        double d2[MAX_DIM];
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) {
            const double df = (k < d) ? xs[i * d + k] - xj[k] : 0.0;
            d2[k] = df * df;
        }
        double ka = 0.0, kb = 0.0;
        for (int c = 0; c < cp.ncomp; ++c) {
            const int reg = cp.region[c];
            if (reg != a && reg != a + 1) continue;
            double kv = 0.0;
            if (cp.kind[c] <= COV_RQ) {
                double z = 0.0;
#pragma unroll
                for (int k = 0; k < MAX_DIM; ++k)
                    if (k < d) z += (0.5 * d2[k]) * cp.inv_l2[c][k];
                const double base = (cp.kind[c] == COV_SE) ? exp(-z) : exp(-cp.rq_alpha[c] * log(1.0 + z / cp.rq_alpha[c]));
                kv = cp.amp2[c] * (base + (gi == gj ? cp.jitter : 0.0));
            } else if (gi == gj) {
                kv = (cp.kind[c] == COV_WHITE) ? cp.amp2[c] : exp(2.0 * cp.hetero_log_sigma[gi]);
            }
            if (reg == a) ka += kv;
            else kb += kv;
        }
        const double fi = fs[i], zi = zs[i];
        const double dci = -fi * (1.0 - fi) / width;
        acc0 = fma(Q, ka * -(dci * (1.0 - fj) + dcj * (1.0 - fi)) + kb * (dci * fj + dcj * fi), acc0);
        const double dwi = dci * zi, dwj = dcj * zj;
        acc1 = fma(Q, ka * -(dwi * (1.0 - fj) + dwj * (1.0 - fi)) + kb * (dwi * fj + dwj * fi), acc1);
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    }
    if (lane == 0) {
        red[warp][0] = acc0;
        red[warp][1] = acc1;
    }
    __syncthreads();
    if (tid < 2) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][tid];
        partials[(int64_t)blockIdx.x * NACC + tid] = v;
    }
}

__global__ void __launch_bounds__(256) reduce_pair_kernel(const double* __restrict__ partials, int ntiles, int off,
                                                          double* __restrict__ grad) {
    __shared__ double sm[256];
    const int p = blockIdx.x;
    double v = 0.0;
    for (int t = threadIdx.x; t < ntiles; t += 256) v += partials[(int64_t)t * NACC + p];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) grad[off + p] = sm[0];
}

// single CTA: mean-parameter gradients, White / Hetero gradients from diag(alpha alpha^T - Kinv)
__global__ void __launch_bounds__(1024) diag_terms_kernel(const CovParams cp, const MeanParams mp, int n_theta_mean, const double* __restrict__ x, int n,
                                                          const double* __restrict__ alpha,
                                                          const double* __restrict__ Kinv, int64_t ld,
                                                          double* __restrict__ grad) {
    __shared__ double sm[1024];
    const int tid = threadIdx.x;
    auto block_sum = [&](double v) -> double {
        sm[tid] = v;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) {
            if (tid < o) sm[tid] += sm[tid + o];
            __syncthreads();
        }
        const double r = sm[0];
        __syncthreads();
        return r;
    };
    // mean gradients (mean.py:50-51, 80-83, 122-126)
    const int d = mp.d;
    const int n_mean = (mp.kind == MEAN_CONST) ? 1 : (mp.kind == MEAN_LINEAR ? 1 + d : 1 + 2 * d);
    for (int p = 0; p < n_mean; ++p) {
        double v = 0.0;
        for (int i = tid; i < n; i += 1024) {
            double g = 1.0;
            if (p >= 1) {
                const int k = (p - 1) % d;
                const double dx = x[(int64_t)i * d + k] - mp.xbar[k];
                g = (p <= d) ? dx : dx * dx;
            }
            v = fma(alpha[i], g, v);
        }
        v = block_sum(v);
        if (tid == 0) grad[p] = v;
    }
    // noise components
    for (int c = 0; c < cp.ncomp; ++c) {
        if (cp.kind[c] == COV_WHITE) {
            double v = 0.0;
            for (int i = tid; i < n; i += 1024) {
                double w2 = 1.0;
                if (cp.n_regions && cp.region[c] >= 0) {
                    double g[MAX_REG];
                    region_weights(cp, x[(int64_t)i * cp.d + cp.cp_axis], g);
                    w2 = pick_region(g, cp.region[c]);
                    w2 *= w2;
                }
                v += w2 * (alpha[i] * alpha[i] - Kinv[(int64_t)i * ld + i]);
            }
            v = block_sum(v);
            if (tid == 0) grad[n_theta_mean + cp.theta_off[c]] = cp.amp2[c] * v;  // 0.5 * sum Q_ii * 2 sigma^2
        } else if (cp.kind[c] == COV_HETERO) {
            for (int i = tid; i < n; i += 1024) {
                double w2 = 1.0;
                if (cp.n_regions && cp.region[c] >= 0) {
                    double g[MAX_REG];
                    region_weights(cp, x[(int64_t)i * cp.d + cp.cp_axis], g);
                    w2 = pick_region(g, cp.region[c]);
                    w2 *= w2;
                }
                const double s2 = w2 * exp(2.0 * cp.hetero_log_sigma[i]);
                grad[n_theta_mean + cp.theta_off[c] + i] = s2 * (alpha[i] * alpha[i] - Kinv[(int64_t)i * ld + i]);
            }
        }
    }
}

// Distributed variant of diag_terms_kernel: this rank adds the terms of the diagonal entries it owns (Kdiag blocks);
// the mean-parameter gradients (alpha is replicated) are written by the rank with with_mean != 0 only, so that the sum
// over ranks is the gradient.  grad must be zero on entry.
__global__ void __launch_bounds__(1024) diag_terms_stacked_kernel(const CovParams cp, const MeanParams mp, int n_theta_mean,
                                                                  const double* __restrict__ x, int n,
                                                                  const double* __restrict__ alpha,
                                                                  const double* __restrict__ Kdiag, int nbd, int me, int G,
                                                                  int na, int with_mean, double* __restrict__ grad) {
    __shared__ double sm[1024];
    const int tid = threadIdx.x;
    auto block_sum = [&](double v) -> double {
        sm[tid] = v;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) {
            if (tid < o) sm[tid] += sm[tid + o];
            __syncthreads();
        }
        const double r = sm[0];
        __syncthreads();
        return r;
    };
    const int d = mp.d;
    const int n_mean = (mp.kind == MEAN_CONST) ? 1 : (mp.kind == MEAN_LINEAR ? 1 + d : 1 + 2 * d);
    if (with_mean) {
        for (int p = 0; p < n_mean; ++p) {
            double v = 0.0;
            for (int i = tid; i < n; i += 1024) {
                double g = 1.0;
                if (p >= 1) {
                    const int k = (p - 1) % d;
                    const double dx = x[(int64_t)i * d + k] - mp.xbar[k];
                    g = (p <= d) ? dx : dx * dx;
                }
                v = fma(alpha[i], g, v);
            }
            v = block_sum(v);
            if (tid == 0) grad[p] = v;
        }
    }
    const int owned = na * nbd;  // stack rows; row r is global point (me + G (r / nbd)) nbd + r % nbd
    for (int c = 0; c < cp.ncomp; ++c) {
        if (cp.kind[c] != COV_WHITE && cp.kind[c] != COV_HETERO) continue;
        double v = 0.0;
        for (int r = tid; r < owned; r += 1024) {
            const int idx = r / nbd, l = r - idx * nbd;
            const int i = (me + G * idx) * nbd + l;
            if (i >= n) continue;
            const double qii = alpha[i] * alpha[i] - Kdiag[(int64_t)idx * nbd * nbd + (int64_t)l * nbd + l];
            if (cp.kind[c] == COV_WHITE) v += qii;
            else grad[n_theta_mean + cp.theta_off[c] + i] = exp(2.0 * cp.hetero_log_sigma[i]) * qii;
        }
        if (cp.kind[c] == COV_WHITE) {
            v = block_sum(v);
            if (tid == 0) grad[n_theta_mean + cp.theta_off[c]] = cp.amp2[c] * v;
        }
    }
}

}  // namespace

size_t trace_partials_size_stacked(int na, int nbd, int npad) {
    return (size_t)na * (nbd / TILE) * (npad / TILE) * NACC * sizeof(double);
}

int launch_lml_grad_stacked(const CovParams& cp, const MeanParams& mp, int n_theta_mean, const double* x, int n, int npad,
                            const double* alpha, const double* Kst, int64_t ldst, const double* Kdiag, int nbd, int me, int G,
                            int na, bool with_mean, double* partials, double* grad_dev, cudaStream_t s) {
    if (cp.n_regions > 0) {
        set_error("the distributed gradient does not support ChangePoint kernels");
        return -2;
    }
    const int tiles_n = npad / TILE;
    const int ntiles = na * (nbd / TILE) * tiles_n;
    for (int c = 0; c < cp.ncomp && ntiles > 0; ++c) {
        if (cp.kind[c] > COV_RQ) continue;
        trace_smooth_stacked_kernel<<<ntiles, 256, 0, s>>>(cp, c, x, n, alpha, Kst, ldst, Kdiag, nbd, me, G, tiles_n, partials);
        GPB_CUDA(cudaGetLastError());
        reduce_partials_kernel<<<NACC, 256, 0, s>>>(partials, ntiles, cp.d, cp.kind[c] == COV_RQ,
                                                    n_theta_mean + cp.theta_off[c], grad_dev, nullptr);
        GPB_CUDA(cudaGetLastError());
        count_launch(2);
    }
    diag_terms_stacked_kernel<<<1, 1024, 0, s>>>(cp, mp, n_theta_mean, x, n, alpha, Kdiag, nbd, me, G, na, with_mean ? 1 : 0,
                                                 grad_dev);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

size_t trace_partials_size(int npad) {
    const int64_t nb = npad / TILE;
    return (size_t)(nb * (nb + 1) / 2) * NACC * sizeof(double);
}

int launch_lml_grad(const CovParams& cp, const MeanParams& mp, int n_theta_mean, const double* x, int n, int npad, const double* alpha, const double* Kinv,
                    int64_t ld, double* partials, double* grad_dev, double* fro2_dev, cudaStream_t s) {
    const int nb = npad / TILE;
    const int ntiles = nb * (nb + 1) / 2;
    for (int c = 0; c < cp.ncomp; ++c) {
        if (cp.kind[c] > COV_RQ) continue;
        trace_smooth_kernel<<<ntiles, 256, 0, s>>>(cp, c, x, n, alpha, Kinv, ld, partials);
        GPB_CUDA(cudaGetLastError());
        reduce_partials_kernel<<<NACC, 256, 0, s>>>(partials, ntiles, cp.d, cp.kind[c] == COV_RQ,
                                                    n_theta_mean + cp.theta_off[c], grad_dev, fro2_dev ? fro2_dev + c : nullptr);
        GPB_CUDA(cudaGetLastError());
        count_launch(2);
    }
    for (int a = 0; a + 1 < cp.n_regions; ++a) {
        // off-diagonal symmetry: the strict lower triangle counts twice, folded with the 1/2 into Q's weight, but
        // dK is not symmetric in (i, j) term by term -- (A + A^T) and (B + B^T) are, so the fold is exact
        trace_cp_kernel<<<ntiles, 256, 0, s>>>(cp, a, x, n, alpha, Kinv, ld, partials);
        GPB_CUDA(cudaGetLastError());
        reduce_pair_kernel<<<2, 256, 0, s>>>(partials, ntiles, n_theta_mean + cp.cp_theta_off + 2 * a, grad_dev);
        GPB_CUDA(cudaGetLastError());
        count_launch(2);
    }
    diag_terms_kernel<<<1, 1024, 0, s>>>(cp, mp, n_theta_mean, x, n, alpha, Kinv, ld, grad_dev);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace gpb
