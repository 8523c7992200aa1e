// TMA-staged variant of the FP64 tensor-core GEMM (both operands k-major):  D = alpha A B^T + beta C.
//
// Same tile and warp layout as gemm_dmma.cu (CTA 128 x 64, 4 warps of 64 x 32, mma.sync m8n8k4 f64), but the operand
// tiles are brought into shared memory by the Tensor Memory Accelerator: one elected thread issues two
// cp.async.bulk.tensor.2d loads per k-tile (16 x 128 and 16 x 64 doubles) that complete on an mbarrier, instead of
// 12 cp.async per thread with their address arithmetic.  The tiles land dense with the 128-byte swizzle
// (16-byte chunk index XOR (row & 7)); fragment reads stay conflict free by permuting the k index inside a k-tile:
// lane t of k-step s takes element {2s, 2s+1, 8+2s, 9+2s}[t] of the 16 -- the same permutation for A and B, so the
// products are unchanged.  Dense tiles need 24 KB per stage instead of 30 KB, which buys a third pipeline stage at
// 3 CTAs per SM (72 KB each).
#include "common.cuh"

#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace gpb {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3, THREADS = 128;
constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;       // 16384, 8192
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;                      // 24576
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 64 /*barriers*/;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 3) dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const GemmArgs p, const int tiles_n) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned tile area (the 128-byte swizzle pattern repeats every 8 rows = 1024 bytes)
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tiles + STAGES * STAGE_BYTES);

    const int tid = threadIdx.x;
    int bi, bj;
    if (p.flags & GEMM_LOWER) {
        const int t = blockIdx.x;
        int r = (int)((sqrtf(4.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (r * (r + 1) > t) --r;
        while ((r + 1) * (r + 2) <= t) ++r;
        bi = r;
        bj = t - r * (r + 1);
    } else {
        bi = blockIdx.x / tiles_n;
        bj = blockIdx.x - bi * tiles_n;
    }
    const int row0 = bi * BM, col0 = bj * BN;
    int k_begin = 0, k_end = p.K;
    if (p.flags & GEMM_TRIK_A) k_begin = max(k_begin, trik_a_begin(row0, p.trik_a_blk, p.trik_a_step));
    if (p.flags & GEMM_TRIK_B) k_begin = max(k_begin, col0);
    if (p.flags & GEMM_TRIL_B) k_end = min(k_end, col0 + BN);
    if (p.flags & GEMM_TRIL_A) k_end = min(k_end, row0 + BM);
    const int KT = (k_end - k_begin) / BK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int kt) {  // thread 0 only
        const int stage = kt % STAGES;
        const unsigned bar = smem_u32(&bars[stage]);
        const unsigned dstA = smem_u32(tiles + stage * STAGE_BYTES);
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_2d(dstA, &tmA, k_begin + kt * BK, row0, bar);
        tma_load_2d(dstA + A_BYTES, &tmB, k_begin + kt * BK, col0, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (s < KT) issue(s);
    }

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, t = lane & 3;
    // byte offset inside a 128-byte row of the element this lane reads at k-step s: element e = {2s,2s+1,8+2s,9+2s}[t],
    // chunk (e >> 1) XOR (row & 7) with row & 7 == g for every fragment row of this lane
    int koff[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) koff[s] = ((((t >> 1) * 4 + s) ^ g) << 4) | ((t & 1) << 3);
    const int a_row = (wm * 64 + g) * 128, b_row = A_BYTES + (wn * 32 + g) * 128;

    for (int kt = 0; kt < KT; ++kt) {
        const int stage = kt % STAGES;
        mbar_wait(smem_u32(&bars[stage]), (kt / STAGES) & 1);
        __syncthreads();  // every warp is done with tile kt-1: its stage may be refilled
        if (tid == 0 && kt + STAGES - 1 < KT) issue(kt + STAGES - 1);
        const unsigned char* base = tiles + stage * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const double*>(base + a_row + i * 1024 + koff[ks]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double*>(base + b_row + j * 1024 + koff[ks]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }

    const double alpha = p.alpha, beta = p.beta;
    const int r_base = row0 + wm * 64 + g;
    const int c_base = col0 + wn * 32 + 2 * t;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = r_base + 8 * i;
        double2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j].x = alpha * acc[i][j][0];
            v[j].y = alpha * acc[i][j][1];
        }
        if (beta != 0.0) {
            double2 c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = *reinterpret_cast<const double2*>(p.C + r * p.ldc + c_base + 8 * j);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j].x = fma(beta, c[j].x, v[j].x);
                v[j].y = fma(beta, c[j].y, v[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(p.D + r * p.ldd + c_base + 8 * j) = v[j];
        if (p.D2 != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(p.D2 + r * p.ldd2 + c_base + 8 * j) = v[j];
        }
    }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    static const EncodeFn fn = []() -> EncodeFn {  // function-local static: initialised once, thread-safe
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeFn>(p);
        return nullptr;
    }();
    return fn;
}

int make_map(CUtensorMap* m, const double* base, int64_t rows, int64_t k_extent, int64_t ld, int box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) return 1;
    const cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace

// Returns 0 when launched, 1 when this variant is not applicable (caller falls back to the cp.async kernel), < 0 on error.
int gemm_nt_tma(const GemmArgs& a, cudaStream_t s, double* flops_out) {
    if (a.flags & (GEMM_A_MMAJOR | GEMM_B_NMAJOR)) return 1;
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) || (a.lda & 1) || (a.ldb & 1) ||
        a.K < BK)
        return 1;
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, a.A, a.M, a.K, a.lda, BM) || make_map(&tmB, a.B, a.N, a.K, a.ldb, BN)) return 1;
    static std::once_flag configured_dev[64];
    int dev = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    cudaError_t cfg_err = cudaSuccess;
    std::call_once(configured_dev[dev & 63], [&]() {
        cfg_err = cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (cfg_err == cudaSuccess)
            cfg_err = cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    });
    GPB_CUDA(cfg_err);
    const int tm = a.M / BM, tn = a.N / BN;
    const int64_t tiles = (a.flags & GEMM_LOWER) ? (int64_t)tm * (tm + 1) : (int64_t)tm * tn;
    dgemm_tma_kernel<<<(unsigned)tiles, THREADS, SMEM_BYTES, s>>>(tmA, tmB, a, tn);
    GPB_CUDA(cudaGetLastError());
    if (flops_out) {
        if (!(a.flags & (GEMM_TRIK_A | GEMM_TRIK_B | GEMM_TRIL_A | GEMM_TRIL_B))) {
            *flops_out = (double)tiles * 2.0 * BM * BN * a.K;
        } else {
            double kext = 0.0;
            for (int bi = 0; bi < tm; ++bi) {
                const int ntile = (a.flags & GEMM_LOWER) ? 2 * (bi + 1) : tn;
                for (int bj = 0; bj < ntile; ++bj) {
                    int kb = 0, ke = a.K;
                    if (a.flags & GEMM_TRIK_A) kb = std::max(kb, trik_a_begin(bi * BM, a.trik_a_blk, a.trik_a_step));
                    if (a.flags & GEMM_TRIK_B) kb = std::max(kb, bj * BN);
                    if (a.flags & GEMM_TRIL_B) ke = std::min(ke, bj * BN + BN);
                    if (a.flags & GEMM_TRIL_A) ke = std::min(ke, bi * BM + BM);
                    kext += std::max(0, ke - kb);
                }
            }
            *flops_out = kext * 2.0 * BM * BN;
        }
    }
    return 0;
}

}  // namespace gpb
