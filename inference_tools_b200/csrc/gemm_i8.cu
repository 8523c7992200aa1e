// FP64 GEMM on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM):  D = alpha A B^T + beta C.
//
// Blackwell's tcgen05 path has no FP64 kind; its FP64 tensor rate (DMMA, gemm_dmma.cu / gemm_tma.cu) is 37 TFLOP/s
// while the same SM does 4.5 POP/s of exact int8 x int8 -> int32.  This file trades one for the other with an
// error-free splitting of the operands (the Ozaki scheme):
//
//   1. split_rows_kernel: every row of A (and of B) is scaled by a power of two 2^-e so that |x| <= 127/128, rounded
//      to a 55-bit integer and written as S = 7 balanced base-256 digits d_0..d_6 (d_0 in [-127,127], the others in
//      [-128,127]):  x = 2^e sum_s d_s 2^-(7+8s)  up to 2^(e-55).  Seven int8 planes per operand.
//   2. gemm_i8_kernel: the 28 digit products with s + t <= 6 are exact int32 GEMMs.  A CTA owns a 128 x 64 tile of
//      D and keeps all seven anti-diagonal sums  P_g = sum_{s+t=g} A_s B_t^T  live in TMEM (7 x 64 columns of 512),
//      so every k-block of the 7 + 7 digit planes is brought in ONCE (one 3-D TMA box per operand) and feeds 28
//      MMAs -- 4x less shared-memory fill per MMA than a plain int8 GEMM, which is what keeps this off the L2 limit.
//      One thread issues the MMAs; a TMA thread runs two 84 KB stages ahead; four epilogue warps read the seven
//      accumulators back (tcgen05.ld), combine them smallest-first in FP64 (Horner in 2^-8) and apply the row /
//      column scales, alpha, beta.
//
// Dropped terms (s + t >= 7) are below 2^-53 of rowmax(A) * rowmax(B) per product -- the same normwise bound as an
// FP64 dot product; the int32 sums are exact for K <= 16384 (7 * K * 2^14 < 2^31).  Operands with the k index
// contiguous only; the caller (gemm_nt) falls back to the DMMA kernels otherwise.
#include "common.cuh"

#include <cuda.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace gpb {
namespace {

constexpr int S = 7;                          // digit planes per operand
constexpr int BM = 128, BN = 64, KB = 64;     // tile of D; k-block in int8 elements (= bytes, one 64-byte swizzle row)
constexpr int A_PLANE = BM * KB, B_PLANE = BN * KB;       // 8192, 4096
constexpr int A_BYTES = S * A_PLANE, B_BYTES = S * B_PLANE;  // 57344, 28672
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;            // 86016
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers, tmem slot*/;
constexpr int THREADS = 192;                  // warp 0: TMA, warp 1: TMEM alloc + MMA issue, warps 2-5: epilogue
constexpr int TMEM_COLS = 512;
constexpr int RASTER = 8;                     // row-blocks per rasterisation group (B planes stay in L2 across them)
constexpr int MAX_K = 16384;

struct I8Args {
    int M, N, K;
    const double *sa, *sb;  // row scales: sa[i] = 2^(ea_i - 14), sb[j] = 2^eb_j
    const double* C;
    int64_t ldc;
    double* D;
    int64_t ldd;
    double* D2;
    int64_t ldd2;
    double alpha, beta;
    int flags, tiles_m, tiles_n;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// shared-memory matrix descriptor: k-major tile of 64-byte rows, 64-byte swizzle (8-row atoms of 512 bytes)
__device__ __forceinline__ uint64_t smem_desc(unsigned addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// instruction descriptor: D = s32, A = B = signed 8-bit, both k-major, N = 64, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_i8(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, const I8Args p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tiles + STAGES * STAGE_BYTES);
    // bars[0..1] full, bars[2..3] empty, bars[4] accumulators ready; then the TMEM base address slot
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bi, bj;
    if (p.flags & GEMM_LOWER) {
        const int t = blockIdx.x;
        int r = (int)((sqrtf(4.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (r * (r + 1) > t) --r;
        while ((r + 1) * (r + 2) <= t) ++r;
        bi = r;
        bj = t - r * (r + 1);
    } else {  // groups of RASTER row-blocks sweep the columns together
        const int per_group = RASTER * p.tiles_n;
        const int grp = blockIdx.x / per_group, r = blockIdx.x - grp * per_group;
        const int rows_in = min(RASTER, p.tiles_m - grp * RASTER);
        bj = r / rows_in;
        bi = grp * RASTER + (r - bj * rows_in);
    }
    const int row0 = bi * BM, col0 = bj * BN;
    int k_begin = 0, k_end = p.K;
    if (p.flags & GEMM_TRIK_A) k_begin = max(k_begin, row0);
    if (p.flags & GEMM_TRIK_B) k_begin = max(k_begin, col0);
    if (p.flags & GEMM_TRIL_B) k_end = min(k_end, col0 + BN);
    if (p.flags & GEMM_TRIL_A) k_end = min(k_end, row0 + BM);
    const int NKB = max(0, k_end - k_begin) / KB;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < NKB; ++kb) {
                const int stage = kb & 1;
                if (kb >= STAGES) mbar_wait(smem_u32(&bars[2 + stage]), ((kb >> 1) - 1) & 1);
                const unsigned bar = smem_u32(&bars[stage]);
                const unsigned dst = smem_u32(tiles + stage * STAGE_BYTES);
                mbar_expect_tx(bar, STAGE_BYTES);
                tma_load_3d(dst, &tmA, k_begin + kb * KB, row0, 0, bar);
                tma_load_3d(dst + A_BYTES, &tmB, k_begin + kb * KB, col0, 0, bar);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            for (int kb = 0; kb < NKB; ++kb) {
                const int stage = kb & 1;
                mbar_wait(smem_u32(&bars[stage]), (kb >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned a_base = smem_u32(tiles + stage * STAGE_BYTES), b_base = a_base + A_BYTES;
#pragma unroll
                for (int ks = 0; ks < KB / 32; ++ks) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const uint64_t adesc = smem_desc(a_base + s * A_PLANE + ks * 32);
#pragma unroll
                        for (int t = 0; t < S - s; ++t) {
                            const uint64_t bdesc = smem_desc(b_base + t * B_PLANE + ks * 32);
                            mma_i8(tmem_base + (unsigned)((s + t) * BN), adesc, bdesc, (unsigned)((kb | ks | s) != 0));
                        }
                    }
                }
                mma_commit(smem_u32(&bars[2 + stage]));  // frees the stage when these MMAs have read it
            }
            if (NKB > 0) mma_commit(smem_u32(&bars[4]));
        }
    } else {  // ---- epilogue: warp w may touch TMEM lanes 32 (w % 4) .. +31
        const int q = warp & 3;
        const int row = row0 + q * 32 + lane;
        const double sa = p.sa[row] * p.alpha;
        const double beta = p.beta;
        if (NKB > 0) {
            mbar_wait(smem_u32(&bars[4]), 0);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
            double v[16];
            if (NKB > 0) {
                int r[S][16];
#pragma unroll
                for (int g = 0; g < S; ++g)
                    tmem_ld16(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(g * BN + c * 16), r[g]);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    double acc = (double)r[S - 1][j];
#pragma unroll
                    for (int g = S - 2; g >= 0; --g) acc = fma(acc, 0.00390625, (double)r[g][j]);
                    v[j] = acc * sa * __ldg(p.sb + col0 + c * 16 + j);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.0;
            }
            const int64_t cc = col0 + c * 16;
            if (beta != 0.0) {
                const double2* src = reinterpret_cast<const double2*>(p.C + (int64_t)row * p.ldc + cc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double2 w = src[j];
                    v[2 * j] = fma(beta, w.x, v[2 * j]);
                    v[2 * j + 1] = fma(beta, w.y, v[2 * j + 1]);
                }
            }
            double2* dst = reinterpret_cast<double2*>(p.D + (int64_t)row * p.ldd + cc);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_double2(v[2 * j], v[2 * j + 1]);
            if (p.D2 != nullptr) {
                double2* dst2 = reinterpret_cast<double2*>(p.D2 + (int64_t)row * p.ldd2 + cc);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst2[j] = make_double2(v[2 * j], v[2 * j + 1]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// One CTA per row: power-of-two scale from the row maximum over the row's valid k range, then the 7 balanced
// base-256 digits of round(x 2^(55-e)), 16 consecutive k per thread (one 16-byte store per plane).
// which = 0: rows of A (block height 128), 1: rows of B (block height 64) -- for the triangular k ranges.
__global__ void __launch_bounds__(256) split_rows_kernel(const double* __restrict__ X, int64_t ld, int K, int rows,
                                                         signed char* __restrict__ q, double* __restrict__ scale_out,
                                                         double extra, int flags, int which) {
    const int row = blockIdx.x, tid = threadIdx.x;
    int k_lo = 0, k_hi = K;
    if (which == 0) {
        if (flags & GEMM_TRIK_A) k_lo = (row / BM) * BM;
        if (flags & GEMM_TRIL_A) k_hi = min(K, (row / BM + 1) * BM);
    } else {
        if (flags & GEMM_TRIK_B) k_lo = (row / BN) * BN;
        if (flags & GEMM_TRIL_B) k_hi = min(K, (row / BN + 1) * BN);
    }
    const double* x = X + (int64_t)row * ld;
    double amax = 0.0;
    for (int k = k_lo + 2 * tid; k < k_hi; k += 512) {
        const double2 w = *reinterpret_cast<const double2*>(x + k);
        const double ax = fabs(w.x), ay = fabs(w.y);
        amax = (ax <= 1.7e308 && ay <= 1.7e308) ? fmax(amax, fmax(ax, ay)) : 1e300;  // NaN / Inf poison the row
    }
    __shared__ double red[8];
    __shared__ double mult_s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((tid & 31) == 0) red[tid >> 5] = amax;
    __syncthreads();
    if (tid == 0) {
        double m = red[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmax(m, red[i]);
        if (m > 1e-280 && m < 1e280) {
            const int e = ilogb(m * (128.0 / 127.0)) + 1;  // |x| 2^-e <= 127/128
            mult_s = scalbn(1.0, 55 - e);
            scale_out[row] = scalbn(extra, e);
        } else {  // all-zero (padding) row, or values this scheme does not cover
            mult_s = 0.0;
            scale_out[row] = (m == 0.0) ? 0.0 : nan("");  // NaN marks an unusable row (propagates to D)
        }
    }
    __syncthreads();
    const double mult = mult_s;
    signed char* qrow = q + (int64_t)row * K;
    const int64_t plane = (int64_t)rows * K;
    for (int k0 = tid * 16; k0 < K; k0 += 256 * 16) {
        unsigned int packed[S][4];
#pragma unroll
        for (int s = 0; s < S; ++s) packed[s][0] = packed[s][1] = packed[s][2] = packed[s][3] = 0u;
        if (k0 >= k_lo && k0 < k_hi) {
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const double2 w = *reinterpret_cast<const double2*>(x + k0 + 2 * h);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    long long m = __double2ll_rn((u ? w.y : w.x) * mult);
                    const int e = 2 * h + u;
#pragma unroll
                    for (int s = S - 1; s >= 1; --s) {
                        const int d = (int)(signed char)(m & 0xFF);
                        m = (m - d) >> 8;
                        packed[s][e >> 2] |= ((unsigned)d & 0xFFu) << ((e & 3) * 8);
                    }
                    packed[0][e >> 2] |= ((unsigned)(int)m & 0xFFu) << ((e & 3) * 8);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(qrow + s * plane + k0) = make_uint4(packed[s][0], packed[s][1], packed[s][2], packed[s][3]);
    }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

// planes[S][rows][K] int8: dims (k, row, plane), box (64, box_rows, S)
int make_plane_map(CUtensorMap* m, const signed char* base, int64_t rows, int64_t K, int box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) return 1;
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)rows * (cuuint64_t)K};
    const cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)box_rows, (cuuint32_t)S};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<signed char*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

// Digit-plane workspace per (device, stream).  Buffers only ever grow, and outgrown ones stay allocated until
// gemm_i8_release(): CUDA graphs captured by the callers keep the addresses they were recorded with.
struct Workspace {
    signed char *qa = nullptr, *qb = nullptr;
    double *sa = nullptr, *sb = nullptr;
    size_t qa_cap = 0, qb_cap = 0, sa_cap = 0, sb_cap = 0;
    std::vector<void*> retired;
};
std::mutex g_ws_mutex;
std::map<std::pair<int, cudaStream_t>, Workspace> g_ws;

template <typename T>
int grow(T*& p, size_t& cap, size_t bytes, std::vector<void*>& retired) {
    if (cap >= bytes) return 0;
    if (p) retired.push_back(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4;
    GPB_CUDA(cudaMalloc(&p, want));
    cap = want;
    return 0;
}

}  // namespace

void gemm_i8_release(cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    auto it = g_ws.find({dev, s});
    if (it == g_ws.end()) return;
    Workspace& w = it->second;
    for (void* p : {(void*)w.qa, (void*)w.qb, (void*)w.sa, (void*)w.sb})
        if (p) cudaFree(p);
    for (void* p : w.retired) cudaFree(p);
    g_ws.erase(it);
}

// Returns 0 when launched, 1 when this path does not apply (caller falls back to the DMMA kernels), < 0 on error.
int gemm_nt_i8(const GemmArgs& a, cudaStream_t s, double* flops_out) {
    if (a.flags & (GEMM_A_MMAJOR | GEMM_B_NMAJOR)) return 1;
    if (a.K % KB || a.K > MAX_K || a.K <= 0 || (a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
        (reinterpret_cast<uintptr_t>(a.B) & 15) || (a.ldd & 1) || (reinterpret_cast<uintptr_t>(a.D) & 15))
        return 1;
    if (a.beta != 0.0 && ((a.ldc & 1) || (reinterpret_cast<uintptr_t>(a.C) & 15))) return 1;
    if (a.D2 && ((a.ldd2 & 1) || (reinterpret_cast<uintptr_t>(a.D2) & 15))) return 1;
    if (!get_encode()) return 1;
    int dev = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    Workspace* w;
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        w = &g_ws[{dev, s}];
    }
    GPB_TRY(grow(w->qa, w->qa_cap, (size_t)S * a.M * a.K, w->retired));
    GPB_TRY(grow(w->qb, w->qb_cap, (size_t)S * a.N * a.K, w->retired));
    GPB_TRY(grow(w->sa, w->sa_cap, sizeof(double) * a.M, w->retired));
    GPB_TRY(grow(w->sb, w->sb_cap, sizeof(double) * a.N, w->retired));
    static bool configured_dev[64] = {};
    if (!configured_dev[dev & 63]) {
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured_dev[dev & 63] = true;
    }
    split_rows_kernel<<<a.M, 256, 0, s>>>(a.A, a.lda, a.K, a.M, w->qa, w->sa, 1.0 / 16384.0, a.flags, 0);
    GPB_CUDA(cudaGetLastError());
    split_rows_kernel<<<a.N, 256, 0, s>>>(a.B, a.ldb, a.K, a.N, w->qb, w->sb, 1.0, a.flags, 1);
    GPB_CUDA(cudaGetLastError());
    CUtensorMap tmA, tmB;
    if (make_plane_map(&tmA, w->qa, a.M, a.K, BM) || make_plane_map(&tmB, w->qb, a.N, a.K, BN)) {
        set_error("gemm_nt_i8: cuTensorMapEncodeTiled failed");
        return -3;
    }
    const int tm = a.M / BM, tn = a.N / BN;
    const int64_t tiles = (a.flags & GEMM_LOWER) ? (int64_t)tm * (tm + 1) : (int64_t)tm * tn;
    I8Args p{a.M, a.N, a.K, w->sa, w->sb, a.C, a.ldc, a.D, a.ldd, a.D2, a.ldd2, a.alpha, a.beta, a.flags, tm, tn};
    gemm_i8_kernel<<<(unsigned)tiles, THREADS, SMEM_BYTES, s>>>(tmA, tmB, p);
    GPB_CUDA(cudaGetLastError());
    count_launch(3);
    if (flops_out) {
        if (!(a.flags & (GEMM_TRIK_A | GEMM_TRIK_B | GEMM_TRIL_A | GEMM_TRIL_B))) {
            *flops_out = (double)tiles * 2.0 * BM * BN * a.K;
        } else {
            double kext = 0.0;
            for (int bi = 0; bi < tm; ++bi) {
                const int ntile = (a.flags & GEMM_LOWER) ? 2 * (bi + 1) : tn;
                for (int bj = 0; bj < ntile; ++bj) {
                    int kb = 0, ke = a.K;
                    if (a.flags & GEMM_TRIK_A) kb = std::max(kb, bi * BM);
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
