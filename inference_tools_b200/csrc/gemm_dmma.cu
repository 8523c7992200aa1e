// FP64 tensor-core GEMM for sm_100a:  D = alpha * A * B^T + beta * C  (row-major, both operands K-contiguous).
//
// Blackwell's tcgen05/TMEM path has no FP64 kind; FP64 tensor work is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), measured at 37.1 TFLOP/s on this pool's B200s
// (profiles/fp64_peak_r1.json) -- the same rate as the DFMA pipe but at 1/8 of the register-file and
// shared-memory traffic per flop, which is what lets a tile kernel sit on the pipe limit.
//
// Tiling: CTA tile 128 x 64, 4 warps (2 x 2), warp tile 64 x 32 = 8 x 4 DMMA tiles (64 accumulator
// doubles per thread), BK = 16, double-buffered cp.async pipeline, 3 CTAs resident per SM (166 registers,
// 61 KB of shared memory each) so that every warp scheduler always has a warp with DMMAs ready while the
// other two sit at a barrier, wait for fragments or run an epilogue (C read-modify-write).  Shared-memory
// rows are padded to 20 doubles: the fragment read (row g, column t) then hits bank pairs (4g + t) mod 16
// -- conflict free.
//
// Every blocked driver in this library (Cholesky trailing update, panel solve through the inverted
// diagonal block, triangular inverse, K^-1 = Z Z^T, batched predict solve) is expressed in this one
// NT form; triangular operands are handled by per-tile k-ranges (GEMM_TRIK_*), symmetric outputs by
// enumerating lower tiles only (GEMM_LOWER).
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace gpb {
namespace {

constexpr int KALIGN = 16;  // K and every k-range boundary are multiples of this
// Tile configurations (warp tile = 8*WMT x 8*WNT DMMA tiles, WMW x WNW warps per CTA):
//   wide  <8,4,2,2,16,2,3>: CTA 128 x 64, 4 warps of 64 x 32, BK = 16, 2-stage pipeline, 3 CTAs/SM -- throughput
//                    configuration (default)
//   small <4,2,2,2>: CTA  64 x 32, 4 warps, 4+ CTAs/SM -- latency configuration for launches that cannot fill the
//                    machine with big tiles (the leaves and low levels of the recursive drivers)
template <int WMT, int WNT, int WMW, int WNW, int BK_ = 16, int STAGES_ = 3, int CTAS_ = 0>  // DMMA tiles per warp, warps per CTA, k-tile, stages, CTAs/SM
struct Cfg {
    static constexpr int BK = BK_, STAGES = STAGES_;
    static constexpr int LDSK = BK + 4;  // padded k-major smem row (20 or 36 doubles: mod 16 == 4 -> conflict free)
    static constexpr int WMT_ = WMT, WNT_ = WNT, WNW_ = WNW;
    static constexpr int THREADS = 32 * WMW * WNW;
    static constexpr int BM = 8 * WMT * WMW, BN = 8 * WNT * WNW;
    static constexpr int LDSM_A = BM + 4;  // m-major A tile rows: (BM + 4) mod 16 == 4 -> conflict free
    static constexpr int LDSM_B = BN + 4;
    static constexpr int A_STAGE = BM * LDSK;  // >= BK * LDSM_A
    static constexpr int B_STAGE = BN * LDSK;  // >= BK * LDSM_B
    static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);  // 92160 / 46080
    static constexpr int MIN_CTAS = CTAS_ ? CTAS_ : (BM == 128 ? 2 : 4);
};

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

template <bool AKM, bool BKM, class C>  // operand k-major or not; tile configuration
__global__ void __launch_bounds__(C::THREADS, C::MIN_CTAS) dgemm_kernel(const GemmArgs p, const int tiles_n) {
    constexpr int BM = C::BM, BN = C::BN, LDSM_A = C::LDSM_A, LDSM_B = C::LDSM_B, A_STAGE = C::A_STAGE,
                  B_STAGE = C::B_STAGE, THREADS = C::THREADS, BK = C::BK, LDSK = C::LDSK, STAGES = C::STAGES,
                  WMT = C::WMT_, WNT = C::WNT_, WNW = C::WNW_;
    constexpr int CH = BK / 2;         // 16-byte chunks per k-major row
    constexpr int RPK = THREADS / CH;  // rows per pass of the k-major loaders
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    const int tid = threadIdx.x;
    int bi, bj;
    if (p.flags & GEMM_LOWER) {
        // lower tiles enumerated row by row: tile row bi owns bj = 0 .. 2*bi+1  (BM = 2*BN)
        const int t = blockIdx.x;
        int r = (int)((sqrtf(4.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (r * (r + 1) > t) --r;
        while ((r + 1) * (r + 2) <= t) ++r;
        bi = r;
        bj = t - r * (r + 1);
    } else {
        // supertile rasterisation: bands of RASTER_H tile rows, column-major inside a band, so the ~300 CTAs
        // resident at any time cover a near-square patch (16 x ~18 tiles) and share both operands through L2
        constexpr int RASTER_H = 16;
        const int tiles_m = gridDim.x / tiles_n;
        const int band = blockIdx.x / (RASTER_H * tiles_n);
        const int band_rows = min(RASTER_H, tiles_m - band * RASTER_H);
        const int in_band = blockIdx.x - band * RASTER_H * tiles_n;
        bj = in_band / band_rows;
        bi = band * RASTER_H + (in_band - bj * band_rows);
    }
    const int row0 = bi * BM, col0 = bj * BN;

    int k_begin = 0, k_end = p.K;
    if (p.flags & GEMM_TRIK_A) k_begin = max(k_begin, trik_a_begin(row0, p.trik_a_blk, p.trik_a_step));
    if (p.flags & GEMM_TRIK_B) k_begin = max(k_begin, col0);
    if (p.flags & GEMM_TRIL_B) k_end = min(k_end, col0 + BN);
    if (p.flags & GEMM_TRIL_A) k_end = min(k_end, row0 + BM);
    const int KT = (k_end - k_begin) / BK;

    // per-thread 16-byte chunk coordinates of the global -> shared copies
    const double* Ag;
    const double* Bg;
    double *as_w, *bs_w;
    if (AKM) {  // BM rows x CH chunks: thread owns chunk (tid % CH) of rows tid / CH + RPK i
        Ag = p.A + (int64_t)(row0 + tid / CH) * p.lda + k_begin + (tid % CH) * 2;
        as_w = As + (tid / CH) * LDSK + (tid % CH) * 2;
    } else {    // 16 k-rows x BM/2 chunks: thread owns chunk (tid % CPR) of k-rows tid / CPR + (128 / CPR) i
        constexpr int CPR = BM / 2;
        Ag = p.A + (int64_t)(k_begin + tid / CPR) * p.lda + row0 + (tid % CPR) * 2;
        as_w = As + (tid / CPR) * LDSM_A + (tid % CPR) * 2;
    }
    if (BKM) {
        Bg = p.B + (int64_t)(col0 + tid / CH) * p.ldb + k_begin + (tid % CH) * 2;
        bs_w = Bs + (tid / CH) * LDSK + (tid % CH) * 2;
    } else {    // 16 k-rows x BN/2 chunks
        constexpr int CPR = BN / 2;
        Bg = p.B + (int64_t)(k_begin + tid / CPR) * p.ldb + col0 + (tid % CPR) * 2;
        bs_w = Bs + (tid / CPR) * LDSM_B + (tid % CPR) * 2;
    }

    auto load_stage = [&](int stage, int kt) {
        double* as = as_w + stage * A_STAGE;
        if (AKM) {
            const double* a = Ag + kt * BK;
#pragma unroll
            for (int i = 0; i < BM / RPK; ++i) cp_async16(as + i * RPK * LDSK, a + (int64_t)i * RPK * p.lda);
        } else {
            constexpr int RPP = THREADS / (BM / 2);  // k-rows per pass
            const double* a = Ag + (int64_t)kt * BK * p.lda;
#pragma unroll
            for (int i = 0; i < BK / RPP; ++i) cp_async16(as + i * RPP * LDSM_A, a + (int64_t)i * RPP * p.lda);
        }
        double* bs = bs_w + stage * B_STAGE;
        if (BKM) {
            const double* b = Bg + kt * BK;
#pragma unroll
            for (int i = 0; i < BN / RPK; ++i) cp_async16(bs + i * RPK * LDSK, b + (int64_t)i * RPK * p.ldb);
        } else {
            constexpr int RPP = THREADS / (BN / 2);
            const double* b = Bg + (int64_t)kt * BK * p.ldb;
#pragma unroll
            for (int i = 0; i < BK / RPP; ++i) cp_async16(bs + i * RPP * LDSM_B, b + (int64_t)i * RPP * p.ldb);
        }
    };

    double acc[WMT][WNT][2];
#pragma unroll
    for (int i = 0; i < WMT; ++i)
#pragma unroll
        for (int j = 0; j < WNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WNW, wn = warp % WNW;
    const int g = lane >> 2, t = lane & 3;
    // fragment read origins: A(row g, k t) / B(col g, k t) of the warp tile
    const double* as_r = AKM ? As + (wm * 8 * WMT + g) * LDSK + t : As + t * LDSM_A + wm * 8 * WMT + g;
    const double* bs_r = BKM ? Bs + (wn * 8 * WNT + g) * LDSK + t : Bs + t * LDSM_B + wn * 8 * WNT + g;
    constexpr int A_I = AKM ? 8 * LDSK : 8, A_K = AKM ? 4 : 4 * LDSM_A;   // strides: next 8 rows / next k-step
    constexpr int B_J = BKM ? 8 * LDSK : 8, B_K = BKM ? 4 : 4 * LDSM_B;

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const int stage = kt % STAGES;
        const double* as = as_r + stage * A_STAGE;
        const double* bs = bs_r + stage * B_STAGE;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[WMT], b[WNT];
#pragma unroll
            for (int i = 0; i < WMT; ++i) a[i] = as[i * A_I + ks * A_K];
#pragma unroll
            for (int j = 0; j < WNT; ++j) b[j] = bs[j * B_J + ks * B_K];
#pragma unroll
            for (int i = 0; i < WMT; ++i)
#pragma unroll
                for (int j = 0; j < WNT; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue straight from the accumulator fragments: each lane owns 2 adjacent columns (16 B), a
    // quad covers 64 contiguous bytes of a row -> two fully used 32 B sectors per row
    const double alpha = p.alpha, beta = p.beta;
    const int r_base = row0 + wm * 8 * WMT + g;
    const int c_base = col0 + wn * 8 * WNT + 2 * t;
#pragma unroll
    for (int i = 0; i < WMT; ++i) {
        const int64_t r = r_base + 8 * i;
        double2 v[WNT];
#pragma unroll
        for (int j = 0; j < WNT; ++j) {
            v[j].x = alpha * acc[i][j][0];
            v[j].y = alpha * acc[i][j][1];
        }
        if (beta != 0.0) {
            double2 c[WNT];
#pragma unroll
            for (int j = 0; j < WNT; ++j) c[j] = *reinterpret_cast<const double2*>(p.C + r * p.ldc + c_base + 8 * j);
#pragma unroll
            for (int j = 0; j < WNT; ++j) {
                v[j].x = fma(beta, c[j].x, v[j].x);
                v[j].y = fma(beta, c[j].y, v[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < WNT; ++j) *reinterpret_cast<double2*>(p.D + r * p.ldd + c_base + 8 * j) = v[j];
        if (p.D2 != nullptr) {
#pragma unroll
            for (int j = 0; j < WNT; ++j) *reinterpret_cast<double2*>(p.D2 + r * p.ldd2 + c_base + 8 * j) = v[j];
        }
    }
}

}  // namespace

int gemm_nt_tma(const GemmArgs& a, cudaStream_t s, double* flops_out);  // gemm_tma.cu
int gemm_nt_i8(const GemmArgs& a, cudaStream_t s, double* flops_out);   // gemm_i8.cu

template <class C>
int launch_cfg(const GemmArgs& a, cudaStream_t s) {
    constexpr int THREADS = C::THREADS;
    constexpr int BM = C::BM, BN = C::BN;
    using kern_t = void (*)(const GemmArgs, const int);
    const bool akm = !(a.flags & GEMM_A_MMAJOR), bkm = !(a.flags & GEMM_B_NMAJOR);
    kern_t kern = akm ? (bkm ? dgemm_kernel<true, true, C> : dgemm_kernel<true, false, C>)
                      : (bkm ? dgemm_kernel<false, true, C> : dgemm_kernel<false, false, C>);
    static std::once_flag configured_dev[64][4];  // per device and operand layout; worker threads may race here
    int dev = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    const int ki = (akm ? 0 : 2) + (bkm ? 0 : 1);
    cudaError_t cfg_err = cudaSuccess;
    std::call_once(configured_dev[dev & 63][ki], [&]() {
        cfg_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (cfg_err == cudaSuccess) cfg_err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    });
    GPB_CUDA(cfg_err);
    const int tm = a.M / BM, tn = a.N / BN;
    const int64_t tiles = (a.flags & GEMM_LOWER) ? (int64_t)tm * (tm + 1) : (int64_t)tm * tn;
    kern<<<(unsigned)tiles, THREADS, C::SMEM_BYTES, s>>>(a, tn);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    // algorithmic flops of this launch (2 * BM * BN * k-extent per computed tile)
    if (!(a.flags & (GEMM_TRIK_A | GEMM_TRIK_B | GEMM_TRIL_A | GEMM_TRIL_B))) {
        credit_gemm_flops((double)tiles * 2.0 * BM * BN * a.K, 0.0);
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
        credit_gemm_flops(kext * 2.0 * BM * BN, 0.0);
    }
    return 0;
}

int gemm_nt(const GemmArgs& a, cudaStream_t s) {
    if (a.M <= 0 || a.N <= 0) return 0;
    if (a.M % 128 || a.N % 64 || a.K % KALIGN || a.K < 0) {
        set_error("gemm_nt: M % 128, N % 64, K % 16 must be 0 (got " + std::to_string(a.M) + "," +
                  std::to_string(a.N) + "," + std::to_string(a.K) + ")");
        return -2;
    }
    if ((a.flags & GEMM_LOWER) && a.N < a.M) {
        set_error("gemm_nt: GEMM_LOWER needs N >= M");
        return -2;
    }
    // big tiles unless they cannot fill one wave of 2 CTAs per SM (148 SMs)
    const int64_t tm = a.M / 128, tn = a.N / 64;
    const int64_t big_tiles = (a.flags & GEMM_LOWER) ? tm * (tm + 1) : tm * tn;
    // INT8 tensor-core path (gemm_i8.cu: exact digit splitting, 28 int8 GEMMs per FP64 GEMM) for long k extents
    // (gpb_set_option "gemm_i8": 0 off, 1 where it pays, 2 wherever it applies; a thread's DMMA retry overrides it)
    const int use_i8 = gemm_i8_override() >= 0 ? gemm_i8_override() : (int)option(OPT_GEMM_I8);
    const int i8_min_k = (int)option(OPT_GEMM_I8_MIN_K);  // measured break-even ~K = 256-384 (profiles/gemm_i8_r1_ksweep.json)
    if (use_i8 && a.K >= i8_min_k && (big_tiles >= 148 || use_i8 == 2)) {  // 2 = force (tests)
        double fl = 0.0;
        const int rc = gemm_nt_i8(a, s, &fl);
        if (rc < 0) return rc;
        if (rc == 0) {
            credit_gemm_flops(fl, fl);
            return 0;
        }
    }
    // Configurations measured on B200 at 8192^3: <8,4,2,2> BK=16, 2 stages, 3 CTAs/SM (166 registers, 61 KB) 35.8 TF/s
    // (default: a third resident warp per scheduler covers the other two's barrier / fragment-load gaps); the same
    // tile with 3 stages and 2 CTAs/SM 34.7; 8 warps of 32 x 32 <4,4,4,2> 33.7; BK=32 x 2 stages 34.1;
    // supertile rasterisation: no change.
    const int force = (int)option(OPT_GEMM_TILE);  // 2 = small, 3 = wide
    const int pick = force ? force : (big_tiles < 296 ? 2 : 3);
    if (pick == 2) return launch_cfg<Cfg<4, 2, 2, 2>>(a, s);
    if (pick == 6) return launch_cfg<Cfg<8, 4, 2, 2>>(a, s);  // previous default: 3 stages, 2 CTAs/SM (34.7 TF/s)
    const int use_tma = (int)option(OPT_GEMM_TMA);
    if (use_tma && pick == 3) {  // TMA-staged variant (gemm_tma.cu) for k-major operands
        double fl = 0.0;
        const int rc = gemm_nt_tma(a, s, &fl);
        if (rc < 0) return rc;
        if (rc == 0) {
            count_launch();
            credit_gemm_flops(fl, 0.0);
            return 0;
        }
    }
    return launch_cfg<Cfg<8, 4, 2, 2, 16, 2, 3>>(a, s);
}

}  // namespace gpb
