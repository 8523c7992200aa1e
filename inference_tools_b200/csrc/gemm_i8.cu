// FP64 GEMM on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM):  D = alpha A B^T + beta C.
//
// Blackwell's tcgen05 path has no FP64 kind; its FP64 tensor rate (DMMA, gemm_dmma.cu / gemm_tma.cu) is 37 TFLOP/s
// while the same SM does 4.5 POP/s of exact int8 x int8 -> int32.  This file trades one for the other with an
// error-free splitting of the operands (the Ozaki scheme):
//
//   1. split_rows_kernel: every row of A (and of B) is scaled by a power of two 2^-e so that |x| <= 127/128, rounded
//      to a 55-bit integer and written as S = 7 balanced base-256 digits d_0..d_6 (d_0 in [-127,127], the others in
//      [-128,127]):  x = 2^e sum_s d_s 2^-(7+8s)  up to 2^(e-55).  Seven int8 planes per operand.
//   2. gemm_i8_kernel: the 28 digit products with s + t <= 6 are exact int32 GEMMs, grouped by anti-diagonal
//      P_g = sum_{s+t=g} A_s B_t^T.  A persistent CTA owns 128 x 128 tiles of D; one tcgen05.mma is 128 x 128 x 32,
//      the smallest shape that runs at the full 8192 MAC/clk/SM from shared memory (N = 64 is limited to 2/3 of it by
//      shared-memory bandwidth: tools/umma_probe.cu).  TMEM holds four 128-column int32 accumulators, so a tile takes
//      two sweeps over k: first P_3..P_6 (22 products, all 7 + 7 planes per k-block), then P_0..P_2 (6 products,
//      planes 0..2).  Every k-block of planes is fetched by ONE 3-D TMA box per operand and feeds 22 (6) MMA pairs:
//      ~45 bytes of L2->shared traffic per MMA clock instead of ~190 for a plain int8 GEMM of this tile.
//      Warp 0 is the TMA producer (224 KB of operand slots in an A ring and a B ring), warp 1 issues the MMAs (one
//      elected lane), 8 or 16 epilogue warps drain the accumulators (tcgen05.ld), combine the diagonals of a sweep as
//      integers (the first sweep's sum waits in a per-CTA scratch row) and apply the row / column scales, alpha, beta.
//      What paces the kernel and what was tried against it: profiles/gemm_i8_load_path_r2.md.
//
// Error per product: the two splitting errors (2^-55 of the row maxima) plus the dropped terms (s + t >= 7), at most
// 2^-51 of rowmax(A) * rowmax(B) with worst-case digits and ~2^-56 with real ones (tests/test_int8_split_model.py) -- a
// normwise bound of the size of an FP64 dot product's; the int32 sums are exact for k extents <= 16384
// (7 * K * 2^14 < 2^31), longer ones are chunked.
// M and N must be multiples of 128; the caller (gemm_nt) falls back to the DMMA kernels otherwise.
#include "common.cuh"
#include "gemm_i8.cuh"

#include <cuda.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace gpb {
namespace {

#ifndef I8_KB
#define I8_KB 64
#endif
constexpr int S = 7;                          // digit planes per operand
constexpr int S_HI = 3;                       // second sweep: diagonals 0..2 need planes 0..2 only
constexpr int BM = 128, BN = 128, KB = I8_KB;  // tile of D; k-block in int8 elements (= bytes, one swizzle row)
constexpr int A_PLANE = BM * KB, B_PLANE = BN * KB;       // 8192, 8192
constexpr int A_BYTES = S * A_PLANE, B_BYTES = S * B_PLANE;  // 57344, 57344
// Shared memory is 224 KB of operand slots in two rings, each slot with a full / empty barrier: an A ring of 56 KB slots
// (7 planes of 128 rows) and a B ring.  A k-block of the first sweep takes one slot of each ring; a k-block of the
// second sweep takes one slot (3 A planes, 3 B planes at +28 KB).
//   uniform layout (single-CTA tiles; CTA pairs with option "gemm_i8_pair" = 1): 2 + 2 slots of 56 KB, the second sweep
//     alternates between the rings (four k-blocks deep);
//   wide layout (CTA pairs, option "gemm_i8_pair" = 2): a CTA of a pair stages only HALF of the B planes (28 KB), so the same 224 KB hold
//     3 A slots + 2 B slots of 28 KB.  Measured equal to the uniform layout within 1 % on every shape
//     (profiles/gemm_i8_load_path_r2.md: the B ring is still two k-blocks deep and sets the pace); kept as an option.
//     The second sweep (36 KB per k-block) runs in the A ring, three deep.
constexpr int SLOT_BYTES = A_BYTES;                        // 57344
constexpr int SLOT_B_OFF = SLOT_BYTES / 2;                 // 28672
constexpr int HI_BYTES = S_HI * A_PLANE;                   // 24576 per operand in the second sweep
constexpr int STAGES = 4;                                  // 56 KB slots' worth of shared memory
constexpr int STAGE_BYTES = SLOT_BYTES;
constexpr int RING_MAX = 4;                                // most slots per ring (I8_KB = 32 probe build: 4 + 4 slots of 28 KB)
constexpr int NBS = KB == 64 ? 2 : 4;                      // B ring slots
constexpr int BAR_FULL_A = 0, BAR_EMPTY_A = RING_MAX, BAR_FULL_B = 2 * RING_MAX, BAR_EMPTY_B = 3 * RING_MAX,
              BAR_READY = 4 * RING_MAX, BAR_DRAINED = 4 * RING_MAX + 1, BAR_COUNT = 4 * RING_MAX + 2;
constexpr int SMEM_BYTES = (KB == 64 ? STAGES : 2 * STAGES) * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers, tmem slot*/;
// warp 0: TMA, warp 1: TMEM alloc + MMA issue, then EPI epilogue warps (EPI / 4 per TMEM lane quarter, 128 * 4 / EPI
// columns each; template parameter: 16 by default, 8 = the round-1 configuration, option "gemm_i8_epi")
constexpr int threads_for(int epi) { return 64 + 32 * epi; }
constexpr int TMEM_COLS = 512;
constexpr int RASTER = 8;                     // row-blocks per rasterisation group (B planes stay in L2 across them)
constexpr int MAX_K = 16384;
constexpr int DEBUG_NO_MMA = 1 << 20, DEBUG_NO_LOAD = 1 << 21;  // timing probes (results are meaningless)
constexpr int EPI_TWO_PASS = 1 << 22;                           // second-sweep epilogue: release TMEM before touching global memory

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
    double* scratch;     // gridDim.x * 128 * 128 doubles: first-sweep partial sums, private to each thread
    int k_off, k_total;  // this launch covers [k_off, k_off + K) of the full k extent (int32 sums stay exact)
    int a_row_off, b_row_off;  // first row of the operands inside their digit-plane buffers (cached planes hold more rows)
    int prefetch;              // k-blocks by which an L2 prefetch of the planes runs ahead of the shared-memory loads (0 = off)
    int trik_a_blk, trik_a_step;  // GemmArgs: block structure of GEMM_TRIK_A (0 = the tile's own rows)
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
template <int CTAS>
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    if (CTAS == 1) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                "r"(dst),
            "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
            : "memory");
    } else {  // executed by both CTAs of the pair; the transaction bytes are credited to the leader's barrier
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                "r"(dst),
            "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar & 0xFEFFFFFFu)
            : "memory");
    }
}
// the box of `map` at these coordinates towards L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// shared-memory matrix descriptor: k-major tile of 64-byte rows, 64-byte swizzle (8-row atoms of 512 bytes)
__device__ __forceinline__ uint64_t smem_desc(unsigned addr) {
    // stride between 8-row atoms = 8 * KB bytes; layout type 4 = 64-byte swizzle, 6 = 32-byte swizzle
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((8 * KB) >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(KB == 64 ? 4 : 6) << 61);
}
// instruction descriptor: D = s32, A = B = signed 8-bit, both k-major, N = 128, M = 128 per CTA of the group
template <int CTAS>
__device__ __forceinline__ void mma_i8(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned accumulate) {
    constexpr uint32_t idesc =
        (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CTAS) >> 4) << 24);
    if (CTAS == 1) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
            : "memory");
    }
}
// one lane of a converged warp (the compiler keeps tcgen05 instructions under this predicate branch-free)
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// arrives on `bar` when every MMA issued so far has completed; with a CTA pair, on the same barrier of both CTAs
template <int CTAS>
__device__ __forceinline__ void mma_commit(unsigned bar) {
    if (CTAS == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .b16 m;\n"
            "mov.b16 m, 3;\n"
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
            "}\n" ::"r"(bar)
            : "memory");
    }
}
// plain arrival on a barrier of CTA `cta` of the cluster (same shared-memory offset)
__device__ __forceinline__ void mbar_arrive_cluster(unsigned bar, unsigned cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(bar),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

struct TileRange {
    int row0, col0, k_begin, nkb;
};
// tile t of the launch: CTAS * 128 rows x 128 columns
template <int CTAS>
__device__ __forceinline__ TileRange tile_range(const I8Args& p, int t) {
    constexpr int BMT = BM * CTAS;
    int bi, bj;
    if (p.flags & GEMM_LOWER) {  // tiles that touch the lower triangle, row-block major
        if (CTAS == 1) {         // bj <= bi
            int r = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
            while (r * (r + 1) / 2 > t) --r;
            while ((r + 1) * (r + 2) / 2 <= t) ++r;
            bi = r;
            bj = t - r * (r + 1) / 2;
        } else {                 // 256-row blocks: bj <= 2 bi + 1
            int r = (int)((sqrtf(4.f * (float)t + 1.f) - 1.f) * 0.5f);
            while (r * (r + 1) > t) --r;
            while ((r + 1) * (r + 2) <= t) ++r;
            bi = r;
            bj = t - r * (r + 1);
        }
    } else {  // groups of RASTER row-blocks sweep the columns together (their B planes stay in L2)
        const int per_group = RASTER * p.tiles_n;
        const int grp = t / per_group, r = t - grp * per_group;
        const int rows_in = min(RASTER, p.tiles_m - grp * RASTER);
        bj = r / rows_in;
        bi = grp * RASTER + (r - bj * rows_in);
    }
    TileRange tr;
    tr.row0 = bi * BMT;
    tr.col0 = bj * BN;
    int k_begin = 0, k_end = p.k_total;  // in the full k extent, then clipped to this launch's chunk
    if (p.flags & GEMM_TRIK_A) k_begin = max(k_begin, trik_a_begin(tr.row0, p.trik_a_blk, p.trik_a_step));
    if (p.flags & GEMM_TRIK_B) k_begin = max(k_begin, tr.col0);
    if (p.flags & GEMM_TRIL_B) k_end = min(k_end, tr.col0 + BN);
    if (p.flags & GEMM_TRIL_A) k_end = min(k_end, tr.row0 + BMT);
    k_begin = max(k_begin, p.k_off) - p.k_off;
    k_end = min(k_end, p.k_off + p.K) - p.k_off;
    tr.k_begin = k_begin;
    tr.nkb = max(0, k_end - k_begin) / KB;
    return tr;
}

// issue the MMAs of one k-block for the diagonals [G0, G1]: accumulator g - G0 lives at TMEM columns (g - G0) * BN.
// B planes are BN / CTAS rows each in this CTA's shared memory (a CTA pair splits the columns of the tile).
template <int CTAS, int G0, int G1>
__device__ __forceinline__ void issue_kblock(unsigned tmem_base, unsigned a_base, unsigned b_base, int kb) {
#pragma unroll
    for (int ks = 0; ks < KB / 32; ++ks) {
#pragma unroll
        for (int s = 0; s <= G1; ++s) {
            const uint64_t adesc = smem_desc(a_base + s * A_PLANE + ks * 32);
#pragma unroll
            for (int t = 0; t <= G1 - s; ++t) {
                if (s + t < G0) continue;
                const uint64_t bdesc = smem_desc(b_base + t * (B_PLANE / CTAS) + ks * 32);
                // the first product into each accumulator (s = 0 of the first k-step) overwrites it
                mma_i8<CTAS>(tmem_base + (unsigned)((s + t - G0) * BN), adesc, bdesc, (unsigned)((kb | ks | s) != 0));
            }
        }
    }
}

// CTAS = 1: one CTA per 128 x 128 tile.  CTAS = 2: a cluster of two CTAs (cta_group::2) per 256 x 128 tile -- each CTA
// stages its own 128 rows of A and HALF of the B planes, the leader issues M = 256 MMAs that read both halves, and each
// CTA's TMEM receives its 128 rows of the accumulators: per MMA a CTA reads 6 KB of shared memory instead of 8 KB and
// fills 25 % less of it, which takes the kernel off the shared-memory bandwidth limit.
template <int CTAS, int EPI_WARPS, bool WIDE>
__global__ void __launch_bounds__(threads_for(EPI_WARPS), 1) gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmA_hi,
                                                             const __grid_constant__ CUtensorMap tmB_hi, const I8Args p,
                                                             const int n_tiles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tiles + (KB == 64 ? STAGES : 2 * STAGES) * STAGE_BYTES);
    // bars[BAR_FULL_A + i] / [BAR_EMPTY_A + i]: A ring slot i; [BAR_FULL_B + j] / [BAR_EMPTY_B + j]: B ring slot j;
    // then accumulators ready, accumulators drained
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + BAR_COUNT);
    static_assert(!WIDE || CTAS == 2, "the wide layout needs the half-size B slots of a CTA pair");
    constexpr int NA = KB == 64 ? (WIDE ? 3 : 2) : 4;                  // A ring slots
    constexpr int B_SLOT = WIDE ? SLOT_BYTES / 2 : SLOT_BYTES;         // bytes per B ring slot
    constexpr bool ALT = !WIDE;  // second sweep alternates between the rings (a B slot holds 56 KB) or stays in the A ring
    auto a_slot = [&](int u) { return smem_u32(tiles + (u % NA) * SLOT_BYTES); };
    auto b_slot = [&](int u) { return smem_u32(tiles + NA * SLOT_BYTES + (u % NBS) * B_SLOT); };
    auto bar_of = [&](int first, int i) { return smem_u32(&bars[first + i]); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned rank = 0;  // CTA within the pair; the leader (0) owns the full / drained barriers and issues the MMAs
    if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    const int unit = blockIdx.x / CTAS, n_units = gridDim.x / CTAS;  // CTA (pair) index: the tile scheduler's lane
    constexpr int B_ROWS = BN / CTAS;                                 // B rows staged by this CTA
    constexpr int B_BYTES_CTA = B_BYTES / CTAS, HI_B_BYTES = HI_BYTES / CTAS;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < BAR_DRAINED; ++i) mbar_init(smem_u32(&bars[i]), 1);  // full: the leader's arrive.expect_tx; empty / ready: one commit
        mbar_init(smem_u32(&bars[BAR_DRAINED]), CTAS * EPI_WARPS);  // one arrival per epilogue warp of the pair
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (CTAS == 2) cluster_sync_all();  // both CTAs are resident before the paired TMEM allocation
    if (warp == 1) {
        if (CTAS == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {  // ---- TMA producer (both CTAs of a pair): converged warp, one elected lane issues
        int ia = 0, ib = 0;  // slots of the A ring / B ring handed out so far
        // wait until the MMAs that read the slot's previous contents have completed
        auto acquire_a = [&](int u) { if (u >= NA) mbar_wait(bar_of(BAR_EMPTY_A, u % NA), ((u / NA) - 1) & 1); };
        auto acquire_b = [&](int u) { if (u >= NBS) mbar_wait(bar_of(BAR_EMPTY_B, u % NBS), ((u / NBS) - 1) & 1); };
        // The leader announces the bytes of the whole pair on its barrier; the other CTA's TMA is credited to the same
        // barrier and needs no arrival of its own (a remote mbarrier.arrive costs more than the load it would announce).
        // Early bytes are harmless: the phase cannot complete before the leader's arrival, and the other CTA cannot be a
        // whole phase ahead because it waits for the MMAs that read the slot (its local `empty` barrier) first.
        auto announce = [&](unsigned bar, unsigned bytes) {
            if (rank == 0) mbar_expect_tx(bar, bytes * CTAS);
        };
        for (int t = unit; t < n_tiles; t += n_units) {
            const TileRange tr = tile_range<CTAS>(p, t);
            const int arow = p.a_row_off + tr.row0 + (int)rank * BM, brow = p.b_row_off + tr.col0 + (int)rank * B_ROWS;
            for (int kb = 0; kb < tr.nkb; ++kb, ++ia, ++ib) {  // first sweep: B planes -> B ring, A planes -> A ring
                acquire_b(ib);
                acquire_a(ia);
                if (elect_one()) {
                    const int k = tr.k_begin + kb * KB;
                    const unsigned bar_a = bar_of(BAR_FULL_A, ia % NA), bar_b = bar_of(BAR_FULL_B, ib % NBS);
                    if ((p.flags & DEBUG_NO_LOAD) && ia >= NA && ib >= NBS) {  // timing probe: MMA rate without the loads
                        if (rank == 0) { mbar_arrive(bar_a); mbar_arrive(bar_b); }
                    } else {  // the smaller box first: it is the one with the shorter deadline in the wide layout
                        announce(bar_b, B_BYTES_CTA);
                        tma_load_3d<CTAS>(b_slot(ib), &tmB, k, brow, 0, bar_b);
                        announce(bar_a, A_BYTES);
                        tma_load_3d<CTAS>(a_slot(ia), &tmA, k, arow, 0, bar_a);
                        // The loads complete at the latency of their slowest sector: the CTA that leads its row / column
                        // block through k takes every L2 miss.  An L2 prefetch a few k-blocks ahead (no shared memory
                        // needed) turns those into hits.
                        if (p.prefetch > 0 && kb + p.prefetch < tr.nkb) {
                            tma_prefetch_3d(&tmB, k + p.prefetch * KB, brow, 0);
                            tma_prefetch_3d(&tmA, k + p.prefetch * KB, arow, 0);
                        }
                    }
                }
                __syncwarp();
            }
            for (int kb = 0; kb < tr.nkb; ++kb) {  // second sweep: one slot per k-block
                const bool in_b = ALT && (kb & 1);
                if (in_b) acquire_b(ib); else acquire_a(ia);
                if (elect_one()) {
                    const int k = tr.k_begin + kb * KB;
                    const unsigned bar = in_b ? bar_of(BAR_FULL_B, ib % NBS) : bar_of(BAR_FULL_A, ia % NA);
                    if ((p.flags & DEBUG_NO_LOAD) && ia >= NA && ib >= NBS) {
                        if (rank == 0) mbar_arrive(bar);
                    } else {
                        const unsigned dst = in_b ? b_slot(ib) : a_slot(ia);
                        announce(bar, HI_BYTES + HI_B_BYTES);
                        tma_load_3d<CTAS>(dst, &tmA_hi, k, arow, 0, bar);
                        tma_load_3d<CTAS>(dst + SLOT_B_OFF, &tmB_hi, k, brow, 0, bar);
                        if (p.prefetch > 0 && kb + 2 * p.prefetch < tr.nkb) {  // these k-blocks are shorter: twice as far ahead
                            tma_prefetch_3d(&tmB_hi, k + 2 * p.prefetch * KB, brow, 0);
                            tma_prefetch_3d(&tmA_hi, k + 2 * p.prefetch * KB, arow, 0);
                        }
                    }
                }
                __syncwarp();
                if (in_b) ++ib; else ++ia;
            }
        }
    } else if (warp == 1) {  // ---- MMA issuer (leader CTA): the whole warp walks the pipeline, one elected lane issues
        if (rank == 0) {
            int ia = 0, ib = 0, uses = 0;
            auto filled_a = [&](int u) { mbar_wait(bar_of(BAR_FULL_A, u % NA), (u / NA) & 1); };
            auto filled_b = [&](int u) { mbar_wait(bar_of(BAR_FULL_B, u % NBS), (u / NBS) & 1); };
            auto drained = [&]() {  // the epilogue warps (of both CTAs) have read the previous sweep's accumulators
                if (uses >= 1) {
                    mbar_wait(smem_u32(&bars[BAR_DRAINED]), (uses - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                ++uses;
            };
            const bool no_mma = p.flags & DEBUG_NO_MMA;  // timing probe: load rate without the MMAs
            auto release = [&](unsigned bar) {           // no-MMA probe: free a barrier in every CTA of the pair
                mbar_arrive(bar);
                if (CTAS == 2) mbar_arrive_cluster(bar, 1);
            };
            const unsigned bar_ready = smem_u32(&bars[BAR_READY]);
            for (int t = unit; t < n_tiles; t += n_units) {
                const TileRange tr = tile_range<CTAS>(p, t);
                if (tr.nkb == 0) continue;
                drained();
                for (int kb = 0; kb < tr.nkb; ++kb, ++ia, ++ib) {  // first sweep: diagonals S_HI..6, accumulator g - S_HI
                    filled_b(ib);
                    filled_a(ia);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const unsigned bar_a = bar_of(BAR_EMPTY_A, ia % NA), bar_b = bar_of(BAR_EMPTY_B, ib % NBS);
                    if (elect_one()) {
                        if (no_mma) {
                            release(bar_a);
                            release(bar_b);
                            if (kb == tr.nkb - 1) release(bar_ready);
                        } else {
                            issue_kblock<CTAS, S_HI, S - 1>(tmem_base, a_slot(ia), b_slot(ib), kb);
                            mma_commit<CTAS>(bar_a);  // both slots are free once these MMAs have read them
                            mma_commit<CTAS>(bar_b);
                            if (kb == tr.nkb - 1) mma_commit<CTAS>(bar_ready);
                        }
                    }
                    __syncwarp();
                }
                drained();
                for (int kb = 0; kb < tr.nkb; ++kb) {  // second sweep: diagonals 0..S_HI-1
                    const bool in_b = ALT && (kb & 1);
                    if (in_b) filled_b(ib); else filled_a(ia);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const unsigned bar = in_b ? bar_of(BAR_EMPTY_B, ib % NBS) : bar_of(BAR_EMPTY_A, ia % NA);
                    const unsigned base = in_b ? b_slot(ib) : a_slot(ia);
                    if (elect_one()) {
                        if (no_mma) {
                            release(bar);
                            if (kb == tr.nkb - 1) release(bar_ready);
                        } else {
                            issue_kblock<CTAS, 0, S_HI - 1>(tmem_base, base, base + SLOT_B_OFF, kb);
                            mma_commit<CTAS>(bar);
                            if (kb == tr.nkb - 1) mma_commit<CTAS>(bar_ready);
                        }
                    }
                    __syncwarp();
                    if (in_b) ++ib; else ++ia;
                }
            }
        }
    } else {  // ---- epilogue: warp w may touch TMEM lanes 32 (w % 4) .. +31; EPI_WARPS / 4 warps per quarter split the columns
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int COLS = BN / (EPI_WARPS / 4);  // columns per warp
        const int cbase = half * COLS;
        const unsigned lane_base = (unsigned)(q * 32) << 16;
        double* scr = p.scratch + ((size_t)blockIdx.x * BM + q * 32 + lane) * BN + cbase;
        const double beta = p.beta;
        const unsigned bar_ready = smem_u32(&bars[BAR_READY]), bar_drained = smem_u32(&bars[BAR_DRAINED]);
        auto release_tmem = [&]() {  // this warp's accumulator columns are in registers
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (CTAS == 1) mbar_arrive(bar_drained);
                else mbar_arrive_cluster(bar_drained, 0);
            }
        };
        int uses = 0;
        for (int t = unit; t < n_tiles; t += n_units) {
            const TileRange tr = tile_range<CTAS>(p, t);
            const int row = tr.row0 + (int)rank * BM + q * 32 + lane;
            if (beta != 0.0) {  // pull this thread's part of the C row towards L2 while the MMAs run
                const double* crow = p.C + (int64_t)row * p.ldc + tr.col0 + cbase;
#pragma unroll
                for (int j = 0; j < COLS; j += 16) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(crow + j));
            }
            // The diagonal sums are combined as INTEGERS before the single conversion to FP64 per sweep (int32 -> FP64
            // conversions are a quarter-rate instruction and the accumulators cannot be released before they are read):
            //   first sweep   I1 = P_3 2^24 + P_4 2^16 + P_5 2^8 + P_6   (|I1| < 2^56, exact in int64) -> scratch as FP64
            //   second sweep  I2 = P_0 2^16 + P_1 2^8 + P_2              (exact in FP64)
            //   sum_g P_g 2^-8g = 2^-16 (I2 + 2^-32 I1)
            if (tr.nkb > 0) {
                mbar_wait(bar_ready, uses & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
                for (int c = 0; c < COLS / 16; ++c) {
                    int r[S - S_HI][16];
#pragma unroll
                    for (int g = 0; g < S - S_HI; ++g)
                        tmem_ld16(tmem_base + lane_base + (unsigned)(g * BN + cbase + c * 16), r[g]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    if (c == COLS / 16 - 1) release_tmem();
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        long long i0 = r[0][j], i1 = r[0][j + 1];
#pragma unroll
                        for (int g = 1; g < S - S_HI; ++g) {
                            i0 = i0 * 256 + r[g][j];
                            i1 = i1 * 256 + r[g][j + 1];
                        }
                        *reinterpret_cast<double2*>(scr + c * 16 + j) = make_double2((double)i0, (double)i1);
                    }
                }
                ++uses;
                mbar_wait(bar_ready, uses & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            const double sa = p.sa[row] * p.alpha * (1.0 / 65536.0);
            if (EPI_WARPS == 8 && (p.flags & EPI_TWO_PASS) && tr.nkb > 0) {  // 8 warps: 204 registers per thread to spend
                // Two passes: first the whole accumulator row goes to registers (I2 is exact in FP64) and the TMEM columns
                // are released -- the MMA warp starts the next tile's first sweep -- then the partial sums of the first
                // sweep and C are fetched from global memory, off the critical path (one L2 round trip per chunk there).
                double i2[COLS];
#pragma unroll
                for (int c = 0; c < COLS / 16; ++c) {
                    int r[S_HI][16];
#pragma unroll
                    for (int g = 0; g < S_HI; ++g)
                        tmem_ld16(tmem_base + lane_base + (unsigned)(g * BN + cbase + c * 16), r[g]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        long long v = r[0][j];
#pragma unroll
                        for (int g = 1; g < S_HI; ++g) v = v * 256 + r[g][j];
                        i2[c * 16 + j] = (double)v;
                    }
                }
                release_tmem();
                ++uses;
#pragma unroll
                for (int c = 0; c < COLS / 16; ++c) {
                    const int col = tr.col0 + cbase + c * 16;
                    double v[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const double2 h = *reinterpret_cast<const double2*>(scr + c * 16 + j);
                        v[j] = fma(h.x, 2.3283064365386963e-10, i2[c * 16 + j]) * sa * __ldg(p.sb + col + j);
                        v[j + 1] = fma(h.y, 2.3283064365386963e-10, i2[c * 16 + j + 1]) * sa * __ldg(p.sb + col + j + 1);
                    }
                    if (beta != 0.0) {
                        const double2* src = reinterpret_cast<const double2*>(p.C + (int64_t)row * p.ldc + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const double2 w = src[j];
                            v[2 * j] = fma(beta, w.x, v[2 * j]);
                            v[2 * j + 1] = fma(beta, w.y, v[2 * j + 1]);
                        }
                    }
                    double2* dst = reinterpret_cast<double2*>(p.D + (int64_t)row * p.ldd + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_double2(v[2 * j], v[2 * j + 1]);
                    if (p.D2 != nullptr) {
                        double2* dst2 = reinterpret_cast<double2*>(p.D2 + (int64_t)row * p.ldd2 + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) dst2[j] = make_double2(v[2 * j], v[2 * j + 1]);
                    }
                }
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < COLS / 16; ++c) {  // second sweep: add I2, scale, store
                double v[16];
                const int col = tr.col0 + cbase + c * 16;
                if (tr.nkb > 0) {
                    int r[S_HI][16];
#pragma unroll
                    for (int g = 0; g < S_HI; ++g)
                        tmem_ld16(tmem_base + lane_base + (unsigned)(g * BN + cbase + c * 16), r[g]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    if (c == COLS / 16 - 1) {
                        release_tmem();
                        ++uses;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const double2 h = *reinterpret_cast<const double2*>(scr + c * 16 + j);
                        long long i0 = r[0][j], i1 = r[0][j + 1];
#pragma unroll
                        for (int g = 1; g < S_HI; ++g) {
                            i0 = i0 * 256 + r[g][j];
                            i1 = i1 * 256 + r[g][j + 1];
                        }
                        const double h0 = fma(h.x, 2.3283064365386963e-10, (double)i0);  // 2^-32
                        const double h1 = fma(h.y, 2.3283064365386963e-10, (double)i1);
                        v[j] = h0 * sa * __ldg(p.sb + col + j);
                        v[j + 1] = h1 * sa * __ldg(p.sb + col + j + 1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.0;
                }
                if (beta != 0.0) {
                    const double2* src = reinterpret_cast<const double2*>(p.C + (int64_t)row * p.ldc + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double2 w = src[j];
                        v[2 * j] = fma(beta, w.x, v[2 * j]);
                        v[2 * j + 1] = fma(beta, w.y, v[2 * j + 1]);
                    }
                }
                double2* dst = reinterpret_cast<double2*>(p.D + (int64_t)row * p.ldd + col);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_double2(v[2 * j], v[2 * j + 1]);
                if (p.D2 != nullptr) {
                    double2* dst2 = reinterpret_cast<double2*>(p.D2 + (int64_t)row * p.ldd2 + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst2[j] = make_double2(v[2 * j], v[2 * j + 1]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();  // nobody leaves while its peer may still signal it
    if (warp == 1) {
        __syncwarp();
        if (CTAS == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// valid k range [k_lo, k_hi) of an operand row, local to the chunk [k_off, k_off + K) of the full k extent.
// which = 0: rows of B (tile width 128); otherwise rows of A and `which` is the tile height (128, or 256 for a CTA
// pair) -- the triangular flags are per tile.
__device__ __forceinline__ void row_k_range(int row, int flags, int which, int k_off, int K, int& k_lo, int& k_hi) {
    int lo = 0, hi = k_off + K;
    if (which != 0) {
        if (flags & GEMM_TRIK_A) lo = (row / which) * which;
        if (flags & GEMM_TRIL_A) hi = min(hi, (row / which + 1) * which);
    } else {
        if (flags & GEMM_TRIK_B) lo = (row / BN) * BN;
        if (flags & GEMM_TRIL_B) hi = min(hi, (row / BN + 1) * BN);
    }
    k_lo = max(lo, k_off) - k_off;
    k_hi = max(k_lo, hi - k_off);
}
// scale 2^-e with |x| 2^-e <= 127/128 for the row maximum m; mult = 2^(55-e); stored scale = extra * 2^e
__device__ __forceinline__ void row_scale(double m, double extra, double& mult, double& scale) {
    if (m > 1e-280 && m < 1e280) {
        const int e = ilogb(m * (128.0 / 127.0)) + 1;
        mult = scalbn(1.0, 55 - e);
        scale = scalbn(extra, e);
    } else {
        // all-zero (padding) rows and rows whose largest entry is below 1e-280 (underflowing covariance tails)
        // contribute nothing; non-finite or absurdly large input poisons its row of D with NaN
        mult = 0.0;
        scale = (m < 1e280) ? 0.0 : nan("");
    }
}
__device__ __forceinline__ double finite_abs_max(double amax, double v) {
    const double a = fabs(v);
    return (a <= 1.7e308) ? fmax(amax, a) : 1e300;  // NaN / Inf poison the row
}
// the 7 balanced base-256 digits of m = round(x 2^(55-e)); digit s goes to byte `lane` of word[s]
__device__ __forceinline__ void put_digits(long long m, int lane, unsigned (&word)[S]) {
#pragma unroll
    for (int s = S - 1; s >= 1; --s) {
        const int d = (int)(signed char)(m & 0xFF);
        m = (m - d) >> 8;
        word[s] |= ((unsigned)d & 0xFFu) << (lane * 8);
    }
    word[0] |= ((unsigned)(int)m & 0xFFu) << (lane * 8);
}

// k-contiguous operand (X[row * ld + k]).  One CTA per row: row maximum over the valid k range, then the digits,
// 16 consecutive k per thread (one 16-byte store per plane).
// bound != nullptr: the row's scale comes from an a-priori bound bound[row % bound_period] on |x| instead of the measured
// maximum (operands whose blocks are split at different times but must share one scale per row); values are clamped to
// the representable range.  qld / qplane: row and plane strides of the digit buffer (compact: K and rows * K).
__global__ void __launch_bounds__(256) split_rows_kernel(const double* __restrict__ X, int64_t ld, int K, int k_off,
                                                         int rows, signed char* __restrict__ q,
                                                         double* __restrict__ scale_out, double extra, int flags,
                                                         int which, const double* __restrict__ bound, int bound_period,
                                                         int64_t qld, int64_t qplane) {
    const int row = blockIdx.x, tid = threadIdx.x;
    int k_lo, k_hi;
    row_k_range(row, flags, which, k_off, K, k_lo, k_hi);
    const double* x = X + (int64_t)row * ld;
    double amax = 0.0;
    if (bound != nullptr) {
        amax = bound[row % bound_period];
    } else {
        for (int k = k_lo + 2 * tid; k < k_hi; k += 512) {
            const double2 w = *reinterpret_cast<const double2*>(x + k);
            amax = finite_abs_max(finite_abs_max(amax, w.x), w.y);
        }
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
        double mult, scale;
        row_scale(m, extra, mult, scale);
        mult_s = mult;
        if (scale_out) scale_out[row] = scale;
    }
    __syncthreads();
    const double mult = mult_s;
    signed char* qrow = q + (int64_t)row * qld;
    const int64_t plane = qplane;
    const bool clamp = bound != nullptr;
    constexpr double LIM = 127.0 * 281474976710656.0;  // 127 * 2^48: the largest |m| whose top digit fits
    for (int k0 = tid * 16; k0 < K; k0 += 256 * 16) {
        unsigned packed[4][S];
#pragma unroll
        for (int wd = 0; wd < 4; ++wd)
#pragma unroll
            for (int s = 0; s < S; ++s) packed[wd][s] = 0u;
        if (k0 >= k_lo && k0 < k_hi) {
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const double2 w = *reinterpret_cast<const double2*>(x + k0 + 2 * h);
                double m0 = w.x * mult, m1 = w.y * mult;
                if (clamp) {
                    m0 = fmin(fmax(m0, -LIM), LIM);
                    m1 = fmin(fmax(m1, -LIM), LIM);
                }
                put_digits(__double2ll_rn(m0), (2 * h) & 3, packed[h >> 1]);
                put_digits(__double2ll_rn(m1), (2 * h + 1) & 3, packed[h >> 1]);
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(qrow + s * plane + k0) = make_uint4(packed[0][s], packed[1][s], packed[2][s], packed[3][s]);
    }
}

// Operand stored with the row index contiguous (X[k * ld + row]: GEMM_A_MMAJOR / GEMM_B_NMAJOR).
// Pass 1: maxima of 64 rows per CTA, reads coalesced along the rows.
__global__ void __launch_bounds__(256) split_cols_max_kernel(const double* __restrict__ X, int64_t ld, int K, int k_off,
                                                             double* __restrict__ scale_out, double* __restrict__ mult_out,
                                                             double extra, int flags, int which) {
    const int r = blockIdx.x * 64 + (threadIdx.x & 63), kg = threadIdx.x >> 6;
    int k_lo, k_hi;
    row_k_range(blockIdx.x * 64, flags, which, k_off, K, k_lo, k_hi);
    double amax = 0.0;
    for (int k = k_lo + kg; k < k_hi; k += 4) amax = finite_abs_max(amax, X[(int64_t)k * ld + r]);
    __shared__ double red[4][64];
    red[kg][threadIdx.x & 63] = amax;
    __syncthreads();
    if (kg == 0) {
        const int c = threadIdx.x & 63;
        const double m = fmax(fmax(red[0][c], red[1][c]), fmax(red[2][c], red[3][c]));
        double mult, scale;
        row_scale(m, extra, mult, scale);
        mult_out[r] = mult;
        scale_out[r] = scale;
    }
}
// Pass 2: a 64 (rows) x 64 (k) tile per CTA, transposed through shared memory: global reads coalesced along the rows,
// digit planes written as 64-byte k-contiguous row segments.
__global__ void __launch_bounds__(256) split_cols_digits_kernel(const double* __restrict__ X, int64_t ld, int K,
                                                                int k_off, int rows, signed char* __restrict__ q,
                                                                const double* __restrict__ mult_in, int flags,
                                                                int which) {
    __shared__ unsigned dig[S][64][17];  // [plane][row][k / 4], padded against bank conflicts
    const int r0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int rl = threadIdx.x & 63, kq0 = threadIdx.x >> 6;
    int k_lo, k_hi;
    row_k_range(r0, flags, which, k_off, K, k_lo, k_hi);
    const bool live = k0 >= k_lo && k0 < k_hi;  // ranges are multiples of 64
    const double mult = mult_in[r0 + rl];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kq = kq0 + 4 * i;
        unsigned word[S];
#pragma unroll
        for (int s = 0; s < S; ++s) word[s] = 0u;
        if (live) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                put_digits(__double2ll_rn(X[(int64_t)(k0 + 4 * kq + u) * ld + r0 + rl] * mult), u, word);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) dig[s][rl][kq] = word[s];
    }
    __syncthreads();
    const int64_t plane = (int64_t)rows * K;
    for (int t = threadIdx.x; t < S * 64 * 4; t += 256) {
        const int s = t / 256, rr = (t >> 2) & 63, quarter = t & 3;
        const unsigned* src = &dig[s][rr][quarter * 4];
        *reinterpret_cast<uint4*>(q + s * plane + (int64_t)(r0 + rr) * K + k0 + quarter * 16) =
            make_uint4(src[0], src[1], src[2], src[3]);
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

// planes[S][rows][ldk] int8 (plane stride `plane`): dims (k, row, plane), box (64, box_rows, box_planes); the map covers k
// in [0, K) from `base`
int make_plane_map(CUtensorMap* m, const signed char* base, int64_t rows, int64_t K, int64_t ldk, int64_t plane, int box_rows,
                   int box_planes) {
    EncodeFn enc = get_encode();
    if (!enc) return 1;
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)ldk, (cuuint64_t)plane};
    const cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<signed char*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, KB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

// Digit-plane workspace per (device, stream).  Buffers only ever grow, and outgrown ones stay allocated until
// gemm_i8_release(): CUDA graphs captured by the callers keep the addresses they were recorded with.
struct Workspace {
    signed char *qa = nullptr, *qb = nullptr;
    double *sa = nullptr, *sb = nullptr, *scratch = nullptr;
    size_t qa_cap = 0, qb_cap = 0, sa_cap = 0, sb_cap = 0, scratch_cap = 0;
    std::vector<void*> retired;
    int sm_count = 0;  // 0 until the kernels have been configured on this workspace's device
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

int get_workspace(cudaStream_t s, Workspace*& w) {
    int dev = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    w = &g_ws[{dev, s}];
    if (w->sm_count == 0) {  // once per (device, stream), under the lock: worker threads share nothing else here
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<1, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<1, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<2, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<2, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<2, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        GPB_CUDA(cudaFuncSetAttribute(gemm_i8_kernel<2, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int sms = 0;
        GPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        w->sm_count = sms;
    }
    return 0;
}

// tile height: CTA pairs (256 x 128 tiles) whenever the rows allow it; option "gemm_i8_pair" = 0 forces single-CTA tiles,
// 1 (default) = pairs with the uniform shared-memory layout, 2 = pairs with the wide layout (see SLOT_BYTES above).
// Measured on B200 at 8192^3: 9.65 ms against 10.44 ms -- the pair reads 6 KB instead of 8 KB of shared memory per
// MMA; its loads have a longer round trip (credited to the leader's barrier) and now set the pace.
int ctas_for(int M) { return (option(OPT_GEMM_I8_PAIR) && M % (2 * BM) == 0) ? 2 : 1; }

// One launch of gemm_i8_kernel on digit planes: k in [k_off, k_off + Kc) of the full extent k_total; the tensor maps
// start map_k_off bytes into the plane rows (k_off for planes that hold the whole extent, 0 for a compact chunk).
// qa / qb: planes with row strides lda_q / ldb_q and plane strides pa / pb, holding rows_a / rows_b rows.
int launch_planes(Workspace* w, cudaStream_t s, const signed char* qa, int64_t lda_q, int64_t pa, int64_t rows_a,
                  int a_row_off, const double* sa, const signed char* qb, int64_t ldb_q, int64_t pb, int64_t rows_b,
                  int b_row_off, const double* sb, int M, int N, int Kc, int k_off, int map_k_off, int k_total, const double* C,
                  int64_t ldc, double* D, int64_t ldd, double* D2, int64_t ldd2, double alpha, double beta, int flags,
                  int ctas, int trik_blk = 0, int trik_step = 0) {
    const int bmt = BM * ctas;
    const int tm = M / bmt, tn = N / BN;
    const int64_t tiles = (flags & GEMM_LOWER) ? (ctas == 2 ? (int64_t)tm * (tm + 1) : (int64_t)tm * (tm + 1) / 2)
                                               : (int64_t)tm * tn;
    const int units = (int)std::min<int64_t>(tiles, w->sm_count / ctas);
    const int grid = units * ctas;
    GPB_TRY(grow(w->scratch, w->scratch_cap, sizeof(double) * (size_t)grid * BM * BN, w->retired));
    CUtensorMap tmA, tmB, tmA_hi, tmB_hi;
    if (make_plane_map(&tmA, qa + map_k_off, rows_a, Kc, lda_q, pa, BM, S) ||
        make_plane_map(&tmB, qb + map_k_off, rows_b, Kc, ldb_q, pb, BN / ctas, S) ||
        make_plane_map(&tmA_hi, qa + map_k_off, rows_a, Kc, lda_q, pa, BM, S_HI) ||
        make_plane_map(&tmB_hi, qb + map_k_off, rows_b, Kc, ldb_q, pb, BN / ctas, S_HI)) {
        set_error("gemm_nt_i8: cuTensorMapEncodeTiled failed");
        return -3;
    }
    const int debug = (int)option(OPT_GEMM_I8_DEBUG);
    // second-sweep epilogue in two passes (8 epilogue warps only): +2-3 % for 2048 <= k < 4096, -1 % at k = 1024
    // (profiles/i8_epi2_ab_r2.json).  Option "gemm_i8_epi2": 0 = off, 1 = on, 2 = by k extent (default).
    const int64_t epi2 = option(OPT_GEMM_I8_EPI2);
    const int two_pass = (epi2 == 1 || (epi2 == 2 && Kc >= 2048)) ? EPI_TWO_PASS : 0;
    I8Args p{M, N, Kc, sa, sb, C, ldc, D, ldd, D2, ldd2, alpha, beta, flags | ((debug & 3) << 20) | two_pass, tm, tn, w->scratch, k_off, k_total,
             a_row_off, b_row_off, (int)std::min<int64_t>(64, std::max<int64_t>(0, option(OPT_GEMM_I8_PREFETCH))), trik_blk, trik_step};
    // epilogue warps per CTA: 16 pay off on long k extents (+1.4 % on the predict shape), 8 on short ones (+3 % at
    // k = 1024): profiles/i8_epilogue_ab_r2.json.  Option "gemm_i8_epi": 0 = by k extent, 8 or 16 = fixed.
    const int epi_opt = (int)option(OPT_GEMM_I8_EPI);
    const int epi = epi_opt == 8 ? 8 : (epi_opt == 16 ? 16 : (Kc >= 4096 ? 16 : 8));
    if (ctas == 1) {
        if (epi == 8) gemm_i8_kernel<1, 8, false><<<grid, threads_for(8), SMEM_BYTES, s>>>(tmA, tmB, tmA_hi, tmB_hi, p, (int)tiles);
        else gemm_i8_kernel<1, 16, false><<<grid, threads_for(16), SMEM_BYTES, s>>>(tmA, tmB, tmA_hi, tmB_hi, p, (int)tiles);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads_for(epi));
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const int nt = (int)tiles;
        const bool wide = option(OPT_GEMM_I8_PAIR) == 2;  // 2: 3 A slots + 2 half-size B slots; 1 (default): uniform slots
        if (wide) {
            if (epi == 8) GPB_CUDA(cudaLaunchKernelEx(&cfg, gemm_i8_kernel<2, 8, true>, tmA, tmB, tmA_hi, tmB_hi, p, nt));
            else GPB_CUDA(cudaLaunchKernelEx(&cfg, gemm_i8_kernel<2, 16, true>, tmA, tmB, tmA_hi, tmB_hi, p, nt));
        } else {
            if (epi == 8) GPB_CUDA(cudaLaunchKernelEx(&cfg, gemm_i8_kernel<2, 8, false>, tmA, tmB, tmA_hi, tmB_hi, p, nt));
            else GPB_CUDA(cudaLaunchKernelEx(&cfg, gemm_i8_kernel<2, 16, false>, tmA, tmB, tmA_hi, tmB_hi, p, nt));
        }
    }
    GPB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

double algorithmic_flops(int M, int N, int K, int flags, int ctas, int trik_blk = 0, int trik_step = 0) {
    const int bmt = BM * ctas, tm = M / bmt, tn = N / BN;
    const int64_t tiles = (flags & GEMM_LOWER) ? (ctas == 2 ? (int64_t)tm * (tm + 1) : (int64_t)tm * (tm + 1) / 2)
                                               : (int64_t)tm * tn;
    if (!(flags & (GEMM_TRIK_A | GEMM_TRIK_B | GEMM_TRIL_A | GEMM_TRIL_B))) return (double)tiles * 2.0 * bmt * BN * K;
    double kext = 0.0;
    for (int bi = 0; bi < tm; ++bi) {
        const int ntile = (flags & GEMM_LOWER) ? (ctas == 2 ? 2 * bi + 2 : bi + 1) : tn;
        for (int bj = 0; bj < ntile; ++bj) {
            int kb = 0, ke = K;
            if (flags & GEMM_TRIK_A) kb = std::max(kb, trik_a_begin(bi * bmt, trik_blk, trik_step));
            if (flags & GEMM_TRIK_B) kb = std::max(kb, bj * BN);
            if (flags & GEMM_TRIL_B) ke = std::min(ke, bj * BN + BN);
            if (flags & GEMM_TRIL_A) ke = std::min(ke, bi * bmt + bmt);
            kext += std::max(0, ke - kb);
        }
    }
    return kext * 2.0 * bmt * BN;
}

}  // namespace

void gemm_i8_release(cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    auto it = g_ws.find({dev, s});
    if (it == g_ws.end()) return;
    Workspace& w = it->second;
    for (void* p : {(void*)w.qa, (void*)w.qb, (void*)w.sa, (void*)w.sb, (void*)w.scratch})
        if (p) cudaFree(p);
    for (void* p : w.retired) cudaFree(p);
    g_ws.erase(it);
}

// ---- pre-split operands (gemm_i8.cuh): planes that outlive one GEMM call ------------------------------------------------
int i8_split_rows(const double* X, int64_t ld, int rows, int K, bool as_a, const double* bound, int bound_period,
                  const I8Planes& out, int row_off, int k_off, cudaStream_t s) {
    if (rows <= 0 || K <= 0) return 0;
    if (K % 64 || (reinterpret_cast<uintptr_t>(X) & 15) || (ld & 1) || (k_off % 64) || row_off + rows > out.rows ||
        k_off + K > out.ld) {
        set_error("i8_split_rows: bad shape / alignment");
        return -2;
    }
    split_rows_kernel<<<rows, 256, 0, s>>>(X, ld, K, 0, rows, out.q + (int64_t)row_off * out.ld + k_off,
                                           k_off == 0 || bound == nullptr ? out.scale + row_off : nullptr,
                                           as_a ? 1.0 / 16384.0 : 1.0, 0, 0, bound, bound_period > 0 ? bound_period : 1,
                                           out.ld, out.plane);
    GPB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int i8_gemm_planes(const I8Planes& A, int a_row_off, const I8Planes& B, int b_row_off, int M, int N, int K,
                   const double* C, int64_t ldc, double* D, int64_t ldd, double alpha, double beta, int flags,
                   cudaStream_t s) {
    if (M <= 0 || N <= 0) return 0;
    if (M % BM || N % BN || K % 64 || K <= 0 || (reinterpret_cast<uintptr_t>(D) & 15) || (ldd & 1) ||
        (beta != 0.0 && ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc & 1))) || !get_encode()) {
        set_error("i8_gemm_planes: shape / alignment not supported by the INT8 kernel");
        return -2;
    }
    Workspace* w;
    GPB_TRY(get_workspace(s, w));
    const int ctas = ctas_for(M);
    const int max_k = (int)std::min<int64_t>(MAX_K, std::max<int64_t>(64, option(OPT_GEMM_I8_MAX_K) / 64 * 64));
    const int chunks = (K + max_k - 1) / max_k;
    const int Kc_max = ((K / 64 + chunks - 1) / chunks) * 64;
    for (int k0 = 0, c = 0; k0 < K; k0 += Kc_max, ++c) {
        const int Kc = std::min(Kc_max, K - k0);
        GPB_TRY(launch_planes(w, s, A.q, A.ld, A.plane, A.rows, a_row_off, A.scale + a_row_off, B.q, B.ld, B.plane, B.rows,
                              b_row_off, B.scale + b_row_off, M, N, Kc, k0, k0, K, c == 0 ? C : D, c == 0 ? ldc : ldd, D, ldd,
                              nullptr, 0, alpha, c == 0 ? beta : 1.0, flags, ctas));
    }
    const double fl = algorithmic_flops(M, N, K, flags, ctas);
    credit_gemm_flops(fl, fl);
    return 0;
}

// Returns 0 when launched, 1 when this path does not apply (caller falls back to the DMMA kernels), < 0 on error.
int gemm_nt_i8(const GemmArgs& a, cudaStream_t s, double* flops_out) {
    const bool a_t = a.flags & GEMM_A_MMAJOR, b_t = a.flags & GEMM_B_NMAJOR;
    auto misaligned = [](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) || (ld & 1); };
    if (a.M % BM || a.N % BN || a.K % 64 || a.K <= 0 || (!a_t && misaligned(a.A, a.lda)) || (!b_t && misaligned(a.B, a.ldb)) ||
        misaligned(a.D, a.ldd))
        return 1;
    if (a.beta != 0.0 && misaligned(a.C, a.ldc)) return 1;
    if (a.D2 && misaligned(a.D2, a.ldd2)) return 1;
    if (!get_encode()) return 1;
    Workspace* w;
    GPB_TRY(get_workspace(s, w));
    const int64_t want_k = a.max_k > 0 ? std::min<int64_t>(a.max_k, option(OPT_GEMM_I8_MAX_K)) : option(OPT_GEMM_I8_MAX_K);
    const int max_k = (int)std::min<int64_t>(MAX_K, std::max<int64_t>(64, want_k / 64 * 64));
    const int chunks = (a.K + max_k - 1) / max_k;
    const int Kc_max = ((a.K / 64 + chunks - 1) / chunks) * 64;  // balanced chunks, multiples of 64
    GPB_TRY(grow(w->qa, w->qa_cap, (size_t)S * a.M * Kc_max, w->retired));
    GPB_TRY(grow(w->qb, w->qb_cap, (size_t)S * a.N * Kc_max, w->retired));
    GPB_TRY(grow(w->sa, w->sa_cap, sizeof(double) * 2 * a.M, w->retired));  // scales, then the 2^(55-e) multipliers
    GPB_TRY(grow(w->sb, w->sb_cap, sizeof(double) * 2 * a.N, w->retired));
    const int ctas = ctas_for(a.M);
    const int bmt = BM * ctas;
    // block-structured GEMM_TRIK_A: the splitting pass treats A as full (the skipped part holds zeros, which split to zero
    // digits and do not move a row maximum); only the kernel's k ranges use the structure
    const int a_flags = a.trik_a_blk > 0 ? (a.flags & ~GEMM_TRIK_A) : a.flags;
    for (int k0 = 0, c = 0; k0 < a.K; k0 += Kc_max, ++c) {
        const int Kc = std::min(Kc_max, a.K - k0);
        if (a_t) {
            const double* X = a.A + (int64_t)k0 * a.lda;
            split_cols_max_kernel<<<a.M / 64, 256, 0, s>>>(X, a.lda, Kc, k0, w->sa, w->sa + a.M, 1.0 / 16384.0, a_flags, bmt);
            split_cols_digits_kernel<<<dim3(a.M / 64, Kc / 64), 256, 0, s>>>(X, a.lda, Kc, k0, a.M, w->qa, w->sa + a.M,
                                                                             a_flags, bmt);
        } else {
            split_rows_kernel<<<a.M, 256, 0, s>>>(a.A + k0, a.lda, Kc, k0, a.M, w->qa, w->sa, 1.0 / 16384.0, a_flags, bmt,
                                                  nullptr, 1, Kc, (int64_t)a.M * Kc);
        }
        GPB_CUDA(cudaGetLastError());
        if (b_t) {
            const double* X = a.B + (int64_t)k0 * a.ldb;
            split_cols_max_kernel<<<a.N / 64, 256, 0, s>>>(X, a.ldb, Kc, k0, w->sb, w->sb + a.N, 1.0, a.flags, 0);
            split_cols_digits_kernel<<<dim3(a.N / 64, Kc / 64), 256, 0, s>>>(X, a.ldb, Kc, k0, a.N, w->qb, w->sb + a.N,
                                                                             a.flags, 0);
        } else {
            split_rows_kernel<<<a.N, 256, 0, s>>>(a.B + k0, a.ldb, Kc, k0, a.N, w->qb, w->sb, 1.0, a.flags, 0, nullptr, 1, Kc,
                                                  (int64_t)a.N * Kc);
        }
        GPB_CUDA(cudaGetLastError());
        count_launch(2 + (a_t ? 1 : 0) + (b_t ? 1 : 0));
        // the chunk's planes are compact (k local to the chunk); k_off / k_total only steer the triangular k ranges
        GPB_TRY(launch_planes(w, s, w->qa, Kc, (int64_t)a.M * Kc, a.M, 0, w->sa, w->qb, Kc, (int64_t)a.N * Kc, a.N, 0, w->sb, a.M,
                              a.N, Kc, k0, 0, a.K, c == 0 ? a.C : a.D, c == 0 ? a.ldc : a.ldd, a.D, a.ldd, a.D2, a.ldd2, a.alpha,
                              c == 0 ? a.beta : 1.0, a.flags, ctas, a.trik_a_blk, a.trik_a_step));
    }
    if (flops_out) *flops_out = algorithmic_flops(a.M, a.N, a.K, a.flags, ctas, a.trik_a_blk, a.trik_a_step);
    return 0;
}

}  // namespace gpb
