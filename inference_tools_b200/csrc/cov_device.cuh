// Device helpers shared by the covariance kernels (kernels.cu, lml.cu, loo.cu).
#pragma once
#include "common.cuh"

namespace gpb {
namespace {

// ChangePoint region weights g_r of one point from its coordinate along the change-point axis
__device__ __forceinline__ void region_weights(const CovParams& cp, double xa, double (&g)[MAX_REG]) {
    double f[MAX_REG - 1];
#pragma unroll
    for (int a = 0; a < MAX_REG - 1; ++a)
        f[a] = (a < cp.n_regions - 1) ? 1.0 / (1.0 + exp(-(xa - cp.cp_loc[a]) / cp.cp_width[a])) : 0.0;
#pragma unroll
    for (int r = 0; r < MAX_REG; ++r) {
        double w = 1.0;
        if (r > 0) w *= f[r - 1];
        if (r < cp.n_regions - 1) w *= 1.0 - f[r < MAX_REG - 1 ? r : 0];
        g[r] = (r < cp.n_regions) ? w : 0.0;
    }
}

__device__ __forceinline__ double leaf_weight(const CovParams& cp, int c, const double (&gi)[MAX_REG],
                                              const double (&gj)[MAX_REG]) {
    const int r = cp.region[c];
    if (r < 0) return 1.0;
    double w = 0.0;
#pragma unroll
    for (int q = 0; q < MAX_REG; ++q)
        if (q == r) w = gi[q] * gj[q];
    return w;
}

__device__ __forceinline__ double pick_region(const double (&g)[MAX_REG], int r) {
    double w = 0.0;
#pragma unroll
    for (int q = 0; q < MAX_REG; ++q)
        if (q == r) w = g[q];
    return w;
}

}  // namespace
}  // namespace gpb
