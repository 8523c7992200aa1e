// Block-column-cyclic distributed FP64 Cholesky + log marginal likelihood over the GPUs of one box
// (BASELINE.json config 5: N = 131072 does not fit one GPU's HBM as a dense matrix).
//
// One process per GPU.  Block column j (width `nbd`) of the lower triangle lives on rank j mod G as a
// contiguous panel (rows j*nbd .. npad+128, ld = nbd); every rank assembles its own panels from the
// coordinates (no communication).  Right-looking sweep with one-panel look-ahead:
//   step k:  owner(k) has factored panel k (diag block potrf + panel solve, potrf.cu)
//            ncclBroadcast(panel k) -> Pbuf[k % 2] on every rank            [stream s_comm]
//            owner(k+1): update column k+1 with P_k, factor panel k+1        [stream s_panel, high priority]
//            every rank: C_j -= P_k[rows >= j] P_k[rows of block j]^T for its other columns j > k  [stream s_main]
// so the broadcast of panel k+1 and its factorisation hide under the step-k trailing update.  The only
// exchange on the data path is the panel broadcast over NVLink (N^2/2 * 8 bytes per receiver in total).
//
// The residual r = y - mu rides along as one extra matrix ROW (row npad of every panel): after the sweep
// that row holds v = L^-1 r, so LML = -1/2 v.v - sum log L_ii (regression.py:538-539) needs no distributed
// triangular solve, only one 2-double ncclAllReduce.
//
// NCCL is loaded with dlopen at gpb_dist_init (torch's bundled libnccl.so.2 when torch is imported first), so
// libgpb200.so itself has no link-time dependency on it.
#include "ctx.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace gpb {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* names[] = {getenv("GPB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm) continue;
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        set_error(std::string("cannot dlopen libnccl.so.2 (import torch first or set GPB200_NCCL_LIB): ") + dlerror());
        return -1;
    }
#define LOAD(field, sym)                                                        \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, sym)); \
    if (!g_nccl.field) {                                                        \
        set_error(std::string("NCCL symbol missing: ") + sym);                  \
        return -1;                                                              \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return 0;
}

#define GPB_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r));                       \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)

// The residual rides along as extra matrix rows below row npad of every panel.  AUG = 256 of them (row 0 = residual, the
// rest zero) keeps every row count of the sweep a multiple of 256 when npad is, so the INT8 GEMM can use its 256 x 128
// CTA-pair tiles.
constexpr int AUG = 256;
__global__ void fill_aug_rows_kernel(double* __restrict__ aug, int64_t ld, int ncols, const double* __restrict__ resid) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    aug[c] = resid[c];
    for (int a = 1; a < AUG; ++a) aug[(int64_t)a * ld + c] = 0.0;
}

// sa[i] = sb[i] / 2^14: the same digit planes used as the left operand of a product (gemm_i8.cu: the A scale carries the
// 2^-14 of the digit weights)
__global__ void a_scale_kernel(const double* __restrict__ sb, double* __restrict__ sa, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sa[i] = sb[i] * (1.0 / 16384.0);
}

// acc[0] += sum log diag(panel block), acc[1] += sum aug_row0^2  (single CTA, sequential launches => fixed order)
__global__ void __launch_bounds__(1024) panel_reduce_kernel(const double* __restrict__ panel, int64_t ld, int ncols,
                                                            const double* __restrict__ aug_row, int n_valid,
                                                            double* __restrict__ acc) {
    __shared__ double s0[1024];
    __shared__ double s1[1024];
    const int tid = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int i = tid; i < ncols; i += 1024) {
        if (i < n_valid) a += log(panel[(int64_t)i * ld + i]);
        b = fma(aug_row[i], aug_row[i], b);
    }
    s0[tid] = a;
    s1[tid] = b;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) {
            s0[tid] += s0[tid + o];
            s1[tid] += s1[tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        acc[0] += s0[0];
        acc[1] += s1[0];
    }
}

// out[c] += sum_r Lblk[r][c] * a[r] for the (rows x cols) block Lblk of an owned panel (backward solve: the term
// L_ij^T alpha_i that block row i contributes to the right-hand side of block column j).  grid = (cols / 128); one CTA per
// 128 columns, so nothing races and the summation order is fixed.
__global__ void __launch_bounds__(256) bwd_update_kernel(const double* __restrict__ Lblk, int64_t ld, int rows,
                                                         const double* __restrict__ a, double* __restrict__ out) {
    __shared__ double part[2][NB];
    const int col = blockIdx.x * NB + (threadIdx.x & (NB - 1)), half = threadIdx.x >> 7;
    const int r0 = half * (rows / 2), r1 = half ? rows : rows / 2;
    double s0 = 0.0, s1 = 0.0;
    const double* p = Lblk + (int64_t)r0 * ld + col;
    int r = r0;
    for (; r + 1 < r1; r += 2, p += 2 * ld) {
        s0 = fma(p[0], a[r], s0);
        s1 = fma(p[ld], a[r + 1], s1);
    }
    if (r < r1) s0 = fma(p[0], a[r], s0);
    part[half][threadIdx.x & (NB - 1)] = s0 + s1;
    __syncthreads();
    if (half == 0) out[col] += part[0][threadIdx.x] + part[1][threadIdx.x];
}

// Y[idx * nbd + i][a * nbd + i] = 1 for the owned row blocks a = me + G idx: the rows of the identity this rank
// carries through the streamed forward solve (inverse rows for the gradient)
__global__ void unit_rows_kernel(double* __restrict__ Y, int64_t ld, int nbd, int me, int G, int64_t npad) {
    const int idx = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t col = (int64_t)(me + G * idx) * nbd + i;
    if (i < nbd && col < npad) Y[((int64_t)idx * nbd + i) * ld + col] = 1.0;
}

// rhs[c] = v[c] - acc[c]
__global__ void sub_kernel(const double* __restrict__ v, const double* __restrict__ acc, double* __restrict__ rhs, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) rhs[c] = v[c] - acc[c];
}

}  // namespace
}  // namespace gpb

using namespace gpb;

struct gpb_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t s_panel = nullptr, s_comm = nullptr;
    // layout of the current factorisation: block column j (width nbd) on rank j % world; an owned panel is stored as
    // (rows_aug - j nbd) x nbd doubles (diagonal block first, the 128 augmented rows last) followed by the nbd x 128
    // inverted 128-blocks of its diagonal block (what every later solve against L_jj needs)
    int nbd = 0, nblk = 0;
    int64_t npad = 0;
    std::vector<size_t> off;
    bool factored = false;
    double lml = 0.0;
    CovParams cp;
    MeanParams mp;
    double *panels = nullptr, *pbuf = nullptr, *tmp = nullptr, *acc = nullptr, *resid = nullptr, *winv = nullptr,
           *v_full = nullptr, *alpha_full = nullptr, *bacc = nullptr, *bvec = nullptr, *pscale = nullptr;
    size_t panels_cap = 0, pbuf_cap = 0, tmp_cap = 0, acc_cap = 0, resid_cap = 0, winv_cap = 0, v_cap = 0, alpha_cap = 0,
           bacc_cap = 0, bvec_cap = 0, pscale_cap = 0, pplanes_cap = 0;
    // gradient workspace: rows of L^-T (then of Kinv, left of the diagonal block) for the owned row blocks, stacked;
    // the diagonal blocks of Kinv; per-tile partial sums; this rank's share of the gradient
    double *ystack = nullptr, *kdiag = nullptr, *gpart = nullptr, *ggrad = nullptr;
    size_t ystack_cap = 0, kdiag_cap = 0, gpart_cap = 0, ggrad_cap = 0;
    signed char* pplanes = nullptr;  // digit planes of the two broadcast buffers (split once per step, used by every update)
    int64_t tmp_rows = 0;
    bool have_alpha = false;
    int* info = nullptr;
    size_t info_cap = 0;
    std::vector<cudaEvent_t> events;
    cudaEvent_t ev(size_t i) {
        while (events.size() <= i) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            events.push_back(e);
        }
        return events[i];
    }
    int owner(int j) const { return j % world; }
    int cols_of(int j) const { return (int)std::min<int64_t>(nbd, npad - (int64_t)j * nbd); }
    int64_t row_end(int j) const { return (int64_t)j * nbd + cols_of(j); }
    int64_t rows_aug() const { return npad + AUG; }
    int64_t panel_rows(int j) const { return rows_aug() - (int64_t)j * nbd; }
    size_t panel_doubles(int j) const { return (size_t)panel_rows(j) * nbd + (size_t)nbd * NB; }
    double* panel(int j) { return panels + off[j]; }
    double* dinv_of(int j) { return panel(j) + (size_t)panel_rows(j) * nbd; }
};

namespace gpb {
void dist_destroy(gpb_ctx* c) {
    gpb_dist* d = c->dist;
    if (!d) return;
    for (double* p : {d->panels, d->pbuf, d->tmp, d->acc, d->resid, d->winv, d->v_full, d->alpha_full, d->bacc, d->bvec,
                      d->pscale, d->ystack, d->kdiag, d->gpart, d->ggrad, reinterpret_cast<double*>(d->pplanes)})
        if (p) cudaFree(p);
    if (d->info) cudaFree(d->info);
    for (auto e : d->events) cudaEventDestroy(e);
    // the INT8 GEMM keeps a digit-plane workspace per (device, stream): free it with the stream, or it leaks on every
    // init / finalize cycle and a later stream that reuses the handle would inherit stale buffers
    if (d->s_panel) {
        gemm_i8_release(d->s_panel);
        cudaStreamDestroy(d->s_panel);
    }
    if (d->s_comm) {
        gemm_i8_release(d->s_comm);
        cudaStreamDestroy(d->s_comm);
    }
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    delete d;
    c->dist = nullptr;
}
}  // namespace gpb

namespace {

struct EventGuard {  // timing events of one call: destroyed on every return path
    cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
    int create() {
        for (auto& x : e) GPB_CUDA(cudaEventCreate(&x));
        return 0;
    }
    ~EventGuard() {
        for (auto x : e)
            if (x) cudaEventDestroy(x);
    }
};

// A local failure (CUDA / NCCL error) leaves the peers blocked inside the next collective: abort the communicator so they
// return an error instead of hanging, and mark this context's distributed state unusable.
int dist_fail(gpb_dist* d, int rc) {
    if (rc != 0 && d->comm && d->world > 1) {
        using AbortFn = ncclResult_t (*)(ncclComm_t);
        if (auto f = reinterpret_cast<AbortFn>(dlsym(g_nccl.handle, "ncclCommAbort"))) f(d->comm);
        d->comm = nullptr;
        d->factored = false;
    }
    return rc;
}

int bcast(gpb_dist* d, const double* src, double* dst, size_t count, int root, cudaStream_t s) {
    if (d->world > 1) {
        if (!d->comm) {
            set_error("distributed communicator was aborted after an earlier failure");
            return -1;
        }
        GPB_NCCL(g_nccl.Broadcast(src, dst, count, ncclDouble, root, d->comm, s));
    } else if (src != dst) {
        GPB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

// Assemble the owned block columns, factor (right-looking, one-panel look-ahead), leave L in the panels, v = L^-1 r on
// every rank and the log marginal likelihood in d->lml.
int dist_factor_impl(gpb_ctx* c, const double* theta, int block, int* info_out, double* seconds_out) {
    gpb_dist* d = c->dist;
    const int G = d->world, me = d->rank;
    const int npad = (int)c->npad, n = (int)c->n, nbd = block;
    d->factored = false;
    d->have_alpha = false;
    d->nbd = nbd;
    d->npad = npad;
    d->nblk = (npad + nbd - 1) / nbd;
    const int nblk = d->nblk;
    const int64_t rows_aug = d->rows_aug();

    // ---- storage: owned panels back to back
    d->off.assign(nblk, 0);
    size_t total = 0;
    for (int j = 0; j < nblk; ++j)
        if (d->owner(j) == me) {
            d->off[j] = total;
            total += d->panel_doubles(j);
        }
    const size_t pb_doubles = (size_t)rows_aug * nbd + (size_t)nbd * NB;
    GPB_TRY(ensure(d->panels, d->panels_cap, sizeof(double) * std::max<size_t>(total, 1)));
    GPB_TRY(ensure(d->pbuf, d->pbuf_cap, sizeof(double) * 2 * pb_doubles));
    GPB_TRY(ensure(d->winv, d->winv_cap, sizeof(double) * 2 * (size_t)nbd * nbd));
    d->tmp_rows = std::max<int64_t>(d->tmp_rows, rows_aug);
    GPB_TRY(ensure(d->tmp, d->tmp_cap, sizeof(double) * (size_t)d->tmp_rows * NB));
    GPB_TRY(ensure(d->acc, d->acc_cap, sizeof(double) * 4));
    GPB_TRY(ensure(d->resid, d->resid_cap, sizeof(double) * (size_t)npad));
    GPB_TRY(ensure(d->v_full, d->v_cap, sizeof(double) * (size_t)npad));
    GPB_TRY(ensure(d->info, d->info_cap, sizeof(int) * (size_t)(nblk + 1)));
    double* pb[2] = {d->pbuf, d->pbuf + pb_doubles};
    // Every trailing update of step k multiplies rows of the same panel P_k: its digit planes are split ONCE per step,
    // right after the broadcast, and shared by all updates (the per-call splitting of gemm_nt re-split the panel for
    // every owned column: N^3 / (6 nbd) elements of traffic, more than the factorisation's own at nbd = 1024).
    const int i8_mode = gemm_i8_override() >= 0 ? gemm_i8_override() : (int)option(OPT_GEMM_I8);
    const bool use_planes = i8_mode >= 1 && nbd % 64 == 0 && (nbd >= (int)option(OPT_GEMM_I8_MIN_K) || i8_mode == 2);
    I8Planes Pa[2], Pb[2];
    if (use_planes) {
        GPB_TRY(ensure(d->pplanes, d->pplanes_cap, 2 * i8_plane_bytes(rows_aug, nbd)));
        GPB_TRY(ensure(d->pscale, d->pscale_cap, sizeof(double) * 4 * (size_t)rows_aug));
        for (int b = 0; b < 2; ++b) {
            Pb[b].q = d->pplanes + (size_t)b * i8_plane_bytes(rows_aug, nbd);
            Pb[b].scale = d->pscale + (size_t)(2 * b) * rows_aug;
            Pb[b].rows = rows_aug;
            Pb[b].ld = nbd;
            Pb[b].plane = rows_aug * (int64_t)nbd;
            Pa[b] = Pb[b];
            Pa[b].scale = d->pscale + (size_t)(2 * b + 1) * rows_aug;
        }
    }

    GPB_TRY(ctx_make_cov_params(c, theta + c->n_mean, d->cp));
    ctx_make_mean_params(c, theta, d->mp);

    cudaStream_t s_main = c->s, s_panel = d->s_panel, s_comm = d->s_comm;
    EventGuard t;
    GPB_TRY(t.create());
    const size_t EV_ASM = 0, EV_TAIL = 1;
    auto EV_PANEL = [&](int k) { return (size_t)2 + 4 * (size_t)k; };
    auto EV_BCAST = [&](int k) { return (size_t)3 + 4 * (size_t)k; };
    auto EV_UPD = [&](int k) { return (size_t)4 + 4 * (size_t)k; };
    auto EV_LA = [&](int k) { return (size_t)5 + 4 * (size_t)k; };
    std::vector<char> did_la(nblk, 0);

    // ---- assemble owned panels + residual row
    GPB_CUDA(cudaEventRecord(t.e[0], s_main));
    GPB_CUDA(cudaMemsetAsync(d->info, 0, sizeof(int) * (nblk + 1), s_main));
    GPB_CUDA(cudaMemsetAsync(d->acc, 0, sizeof(double) * 4, s_main));
    GPB_TRY(launch_residual(d->mp, c->x, c->y, n, npad, d->resid, nullptr, s_main));
    for (int j = 0; j < nblk; ++j) {
        if (d->owner(j) != me) continue;
        const int cj = d->cols_of(j);
        GPB_TRY(launch_assemble_block(d->cp, c->x, n, c->has_noise ? c->noise : nullptr, j * nbd, npad - j * nbd, j * nbd, cj,
                                      d->panel(j), nbd, s_main));
        fill_aug_rows_kernel<<<(cj + 255) / 256, 256, 0, s_main>>>(d->panel(j) + (size_t)(npad - j * nbd) * nbd, nbd, cj,
                                                                  d->resid + j * nbd);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    GPB_CUDA(cudaEventRecord(d->ev(EV_ASM), s_main));
    GPB_CUDA(cudaEventRecord(t.e[1], s_main));
    GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_ASM), 0));

    // Panel k on its owner (stream s_panel; column k fully updated by stream order / waits): Cholesky of the diagonal
    // block in place, then the rows below it (and the residual row) are solved against L_kk^T as ONE GEMM with the
    // explicit inverse W = inv(L_kk), k extent nbd -- at nbd = 1024 that runs on the INT8 tensor-core path like the
    // trailing updates (the recursive solve it replaces ran its k <= 512 pieces on DMMA and was the serial bottleneck
    // once the updates got fast).  The result goes straight into the broadcast buffer (NCCL broadcasts in place on the
    // root); the panel's own storage is brought up to date off the critical path, after the broadcast.
    auto factor_panel = [&](int k) -> int {
        const int ck = d->cols_of(k);
        LinalgWs ws{d->dinv_of(k), d->tmp, d->tmp_rows, d->info + nblk};  // potrf_lower clears ws.info at entry
        GPB_TRY(potrf_lower(d->panel(k), nbd, ck, ws, s_panel));
        // keep the first failure of this panel (index local to the panel; made global on the host after the sweep)
        GPB_CUDA(cudaMemcpyAsync(d->info + k, d->info + nblk, sizeof(int), cudaMemcpyDeviceToDevice, s_panel));
        const int64_t below = rows_aug - d->row_end(k);
        GPB_CUDA(cudaMemsetAsync(d->winv, 0, sizeof(double) * (size_t)ck * ck, s_panel));
        GPB_TRY(trtri_lower(d->panel(k), nbd, d->winv, ck, ck, 0, ws, d->winv + (size_t)nbd * nbd, nbd, s_panel));
        if (k >= 2) {  // pb[k & 1] was read by step k - 2's updates
            GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_UPD(k - 2)), 0));
            if (did_la[k - 2]) GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_LA(k - 2)), 0));
        }
        GemmArgs g{(int)below, ck, ck, d->panel(k) + (size_t)ck * nbd, nbd, d->winv, ck, nullptr, 0, pb[k & 1], nbd, nullptr, 0,
                   1.0, 0.0, GEMM_TRIL_B};
        GPB_TRY(gemm_nt(g, s_panel));
        GPB_CUDA(cudaEventRecord(d->ev(EV_PANEL(k)), s_panel));
        return 0;
    };
    auto update_col = [&](int j, int k, cudaStream_t s) -> int {  // C_j -= P_k[rows >= j*nbd] P_k[block j rows]^T
        const int M = (int)(rows_aug - (int64_t)j * nbd), Nc = d->cols_of(j), K = d->cols_of(k);
        const int row_off = (int)((int64_t)j * nbd - d->row_end(k));
        if (use_planes && K % 64 == 0 && ((int64_t)(M / NB) * (Nc / 64) >= 148 || i8_mode == 2))
            return i8_gemm_planes(Pa[k & 1], row_off, Pb[k & 1], row_off, M, Nc, K, d->panel(j), nbd, d->panel(j), nbd, -1.0, 1.0,
                                  GEMM_FULL, s);
        const double* P = pb[k & 1] + (size_t)row_off * nbd;
        GemmArgs g{M, Nc, K, P, nbd, P, nbd, d->panel(j), nbd, d->panel(j), nbd, nullptr, 0, -1.0, 1.0, GEMM_FULL};
        return gemm_nt(g, s);
    };

    if (d->owner(0) == me) GPB_TRY(factor_panel(0));
    for (int k = 0; k < nblk; ++k) {
        const int ok = d->owner(k);
        const int64_t below = rows_aug - d->row_end(k);
        const size_t count = (size_t)below * nbd;
        // (1) broadcast panel k (rows below the diagonal block) into Pbuf[k & 1]
        if (k >= 2) {
            GPB_CUDA(cudaStreamWaitEvent(s_comm, d->ev(EV_UPD(k - 2)), 0));
            if (did_la[k - 2]) GPB_CUDA(cudaStreamWaitEvent(s_comm, d->ev(EV_LA(k - 2)), 0));
        }
        if (ok == me) GPB_CUDA(cudaStreamWaitEvent(s_comm, d->ev(EV_PANEL(k)), 0));
        GPB_TRY(bcast(d, pb[k & 1], pb[k & 1], count, ok, s_comm));
        if (use_planes && d->cols_of(k) % 64 == 0) {  // digit planes of the received panel, once for all of step k's updates
            GPB_TRY(i8_split_rows(pb[k & 1], nbd, (int)below, d->cols_of(k), false, nullptr, 1, Pb[k & 1], 0, 0, s_comm));
            a_scale_kernel<<<(int)((below + 255) / 256), 256, 0, s_comm>>>(Pb[k & 1].scale, Pa[k & 1].scale, (int)below);
            GPB_CUDA(cudaGetLastError());
            count_launch();
        }
        GPB_CUDA(cudaEventRecord(d->ev(EV_BCAST(k)), s_comm));
        if (ok == me)  // the solved rows into the panel's own storage (needed by alpha / predict, not by the sweep)
            GPB_CUDA(cudaMemcpyAsync(d->panel(k) + (size_t)d->cols_of(k) * nbd, pb[k & 1], sizeof(double) * count,
                                     cudaMemcpyDeviceToDevice, s_comm));
        // (2) look-ahead: the owner of panel k+1 brings its column up to date and factors it
        if (k + 1 < nblk && d->owner(k + 1) == me) {
            GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_BCAST(k)), 0));
            if (k >= 1) GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_UPD(k - 1)), 0));
            GPB_TRY(update_col(k + 1, k, s_panel));
            GPB_CUDA(cudaEventRecord(d->ev(EV_LA(k)), s_panel));
            did_la[k] = 1;
            GPB_TRY(factor_panel(k + 1));
        }
        // (3) trailing update of the other owned columns; v_k = the residual row of the panel, kept on every rank.
        // On the owner of panel k+1 the trailing update waits until that panel is factored: the persistent INT8 GEMM
        // occupies every SM for the whole launch, so stream priority cannot get the ~60 short dependent kernels of the
        // panel chain in between trailing launches -- each would wait for a full trailing GEMM.  Ordering instead of
        // priority keeps the chain (the sweep's critical path) undisturbed; the owner catches up on its trailing work
        // while the panel travels.
        if (k + 1 < nblk && d->owner(k + 1) == me && G > 1) GPB_CUDA(cudaStreamWaitEvent(s_main, d->ev(EV_PANEL(k + 1)), 0));
        GPB_CUDA(cudaStreamWaitEvent(s_main, d->ev(EV_BCAST(k)), 0));
        GPB_CUDA(cudaMemcpyAsync(d->v_full + (size_t)k * nbd, pb[k & 1] + (size_t)(npad - d->row_end(k)) * nbd,
                                 sizeof(double) * d->cols_of(k), cudaMemcpyDeviceToDevice, s_main));
        for (int j = k + 2; j < nblk; ++j)
            if (d->owner(j) == me) GPB_TRY(update_col(j, k, s_main));
        GPB_CUDA(cudaEventRecord(d->ev(EV_UPD(k)), s_main));
    }
    // ---- reductions: log det from the owned diagonal blocks, v.v from the replicated v
    GPB_CUDA(cudaEventRecord(d->ev(EV_TAIL), s_comm));
    GPB_CUDA(cudaStreamWaitEvent(s_main, d->ev(EV_TAIL), 0));
    for (int j = 0; j < nblk; ++j) {
        if (d->owner(j) != me) continue;
        GPB_CUDA(cudaStreamWaitEvent(s_main, d->ev(EV_PANEL(j)), 0));
        const int cj = d->cols_of(j);
        panel_reduce_kernel<<<1, 1024, 0, s_main>>>(d->panel(j), nbd, cj, d->v_full + (size_t)j * nbd,
                                                    std::max(0, std::min(cj, n - j * nbd)), d->acc);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    if (G > 1) GPB_NCCL(g_nccl.AllReduce(d->acc, d->acc, 2, ncclDouble, ncclSum, d->comm, s_main));
    GPB_CUDA(cudaEventRecord(t.e[2], s_main));
    double acc[2];
    std::vector<int> info_h(nblk + 1, 0);
    GPB_CUDA(cudaMemcpyAsync(acc, d->acc, sizeof(acc), cudaMemcpyDeviceToHost, s_main));
    GPB_CUDA(cudaStreamSynchronize(s_main));
    GPB_CUDA(cudaStreamSynchronize(s_panel));
    GPB_CUDA(cudaStreamSynchronize(s_comm));
    GPB_CUDA(cudaMemcpy(info_h.data(), d->info, sizeof(int) * (nblk + 1), cudaMemcpyDeviceToHost));
    int first_bad = 0;
    for (int j = 0; j < nblk && !first_bad; ++j)
        if (d->owner(j) == me && info_h[j] > 0) first_bad = info_h[j] + j * nbd;
    if (G > 1) {  // smallest positive index over ranks: encode 0 as +inf
        int* dev_i = d->info + nblk;
        int enc = first_bad > 0 ? first_bad : 0x7fffffff;
        GPB_CUDA(cudaMemcpy(dev_i, &enc, sizeof(int), cudaMemcpyHostToDevice));
        GPB_NCCL(g_nccl.AllReduce(dev_i, dev_i, 1, ncclInt32, ncclMin, d->comm, s_main));
        GPB_CUDA(cudaMemcpyAsync(&enc, dev_i, sizeof(int), cudaMemcpyDeviceToHost, s_main));
        GPB_CUDA(cudaStreamSynchronize(s_main));
        first_bad = enc == 0x7fffffff ? 0 : enc;
    }
    *info_out = first_bad;
    d->lml = -0.5 * acc[1] - acc[0];
    d->factored = first_bad == 0;
    if (seconds_out) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, t.e[0], t.e[1]);
        cudaEventElapsedTime(&b, t.e[1], t.e[2]);
        seconds_out[0] = a * 1e-3;
        seconds_out[1] = b * 1e-3;
        seconds_out[2] = (a + b) * 1e-3;
    }
    return 0;
}

int check_dist(gpb_ctx* c, const char* who, bool need_factor) {
    GPB_TRY(ctx_use(c));
    GPB_TRY(ctx_need_model(c));
    if (!c->dist) {
        set_error(std::string(who) + ": call gpb_dist_init first");
        return -2;
    }
    if (need_factor && !c->dist->factored) {
        set_error(std::string(who) + ": no distributed factorisation (call gpb_dist_factor / gpb_dist_lml first)");
        return -2;
    }
    return 0;
}

// alpha = L^-T v on every rank (device: d->alpha_full, npad entries); computed once per factorisation
int dist_alpha_impl(gpb_ctx* c) {
    gpb_dist* d = c->dist;
    const int nbd = d->nbd, nblk = d->nblk, me = d->rank;
    const int64_t npad = d->npad;
    cudaStream_t s = c->s;
    GPB_TRY(ensure(d->alpha_full, d->alpha_cap, sizeof(double) * (size_t)npad));
    GPB_TRY(ensure(d->bacc, d->bacc_cap, sizeof(double) * (size_t)npad));
    GPB_TRY(ensure(d->bvec, d->bvec_cap, sizeof(double) * 2 * (size_t)nbd));
    if (d->have_alpha) return 0;
    GPB_CUDA(cudaMemsetAsync(d->bacc, 0, sizeof(double) * (size_t)npad, s));
    for (int i = nblk - 1; i >= 0; --i) {
        const int ci = d->cols_of(i);
        double* ai = d->alpha_full + (size_t)i * nbd;
        if (d->owner(i) == me) {
            sub_kernel<<<(ci + 255) / 256, 256, 0, s>>>(d->v_full + (size_t)i * nbd, d->bacc + (size_t)i * nbd, d->bvec, ci);
            GPB_CUDA(cudaGetLastError());
            count_launch();
            GPB_TRY(trsv_lower_bwd(d->panel(i), nbd, ci, d->dinv_of(i), d->bvec, s));
            GPB_CUDA(cudaMemcpyAsync(ai, d->bvec + ci, sizeof(double) * ci, cudaMemcpyDeviceToDevice, s));
        }
        GPB_TRY(bcast(d, ai, ai, ci, d->owner(i), s));
        for (int j = me; j < i; j += d->world) {  // owned block columns left of i
            const double* Lij = d->panel(j) + (size_t)((int64_t)(i - j) * nbd) * nbd;
            bwd_update_kernel<<<d->cols_of(j) / NB, 256, 0, s>>>(Lij, nbd, ci, ai, d->bacc + (size_t)j * nbd);
            GPB_CUDA(cudaGetLastError());
            count_launch();
        }
    }
    d->have_alpha = true;
    return 0;
}

// Rows of K^-1 for the gradient traces (the reference forms the whole inverse on one host: regression.py:556-557).
// Rank `me` owns the row blocks a = me + G idx (the same cyclic rule as the column panels), stacked in d->ystack.
//   Phase 1  Y = E L^-T for the owned rows E of the identity: the column panels of L are streamed through all ranks as in
//            gpb_dist_predict (one broadcast per panel, double-buffered); row block a of Y is (columns a of L^-1)^T and is
//            zero left of column a nbd, so step j only touches the row blocks a <= j.
//   Phase 2  K^-1[a, b] = Y_a Y_b^T for b <= a (k runs over the columns >= a nbd): the owner of row block b broadcasts it,
//            packed from its first non-zero column, and every rank multiplies its row blocks a >= b with it.  K^-1[a, b] for
//            b < a is written over the zero part of Y's row block a (column block b -- never read again as an operand),
//            the diagonal blocks go to d->kdiag.  The products run in k-chunks with their own row scales on the INT8 path
//            for the reason given at lauum_lower (potrf.cu).
// Flops: N^3 / 3 (phase 1) + N^3 / 3 (phase 2: the k range of every tile starts at its row block's first non-zero column)
// over all ranks;
// NVLink: 2 x N^2 / 2 x 8 bytes received per rank.
int dist_inverse_rows(gpb_ctx* c, int* na_out) {
    gpb_dist* d = c->dist;
    const int G = d->world, me = d->rank, nbd = d->nbd, nblk = d->nblk, npad = (int)d->npad;
    const int na = nblk > me ? (nblk - 1 - me) / G + 1 : 0;
    *na_out = na;
    const int64_t stack_rows = (int64_t)na * nbd;
    cudaStream_t s = c->s, s_comm = d->s_comm;
    GPB_TRY(ensure(d->ystack, d->ystack_cap, sizeof(double) * (size_t)std::max<int64_t>(stack_rows, 1) * npad));
    GPB_TRY(ensure(d->kdiag, d->kdiag_cap, sizeof(double) * (size_t)std::max(na, 1) * nbd * nbd));
    if (stack_rows > d->tmp_rows) {
        d->tmp_rows = stack_rows;
        GPB_TRY(ensure(d->tmp, d->tmp_cap, sizeof(double) * (size_t)d->tmp_rows * NB));
    }
    auto yrow = [&](int idx) { return d->ystack + (size_t)idx * nbd * npad; };
    if (na > 0) {
        GPB_CUDA(cudaMemsetAsync(d->ystack, 0, sizeof(double) * (size_t)stack_rows * npad, s));
        GPB_CUDA(cudaMemsetAsync(d->kdiag, 0, sizeof(double) * (size_t)na * nbd * nbd, s));
        unit_rows_kernel<<<dim3((nbd + 255) / 256, na), 256, 0, s>>>(d->ystack, npad, nbd, me, G, npad);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    const size_t pb_doubles = (size_t)d->rows_aug() * nbd + (size_t)nbd * NB;
    double* pb[2] = {d->pbuf, d->pbuf + pb_doubles};
    size_t ev = 0;
    std::vector<cudaEvent_t> used(2, nullptr);
    cudaEvent_t ev_ready = d->ev(ev++);
    GPB_CUDA(cudaEventRecord(ev_ready, s));  // earlier work on s (factor, alpha) is complete before s_comm reads the panels
    GPB_CUDA(cudaStreamWaitEvent(s_comm, ev_ready, 0));
    int64_t step = 0;
    // ---- phase 1: forward solve with the streamed panels
    for (int j = 0; j < nblk; ++j, ++step) {
        const int cj = d->cols_of(j), oj = d->owner(j);
        double* P = pb[step & 1];
        if (used[step & 1]) GPB_CUDA(cudaStreamWaitEvent(s_comm, used[step & 1], 0));
        const size_t count = (size_t)d->panel_rows(j) * nbd + (size_t)nbd * NB;  // panel + inverted diagonal blocks
        if (oj == me && G == 1) {
            GPB_CUDA(cudaMemcpyAsync(P, d->panel(j), sizeof(double) * count, cudaMemcpyDeviceToDevice, s_comm));
        } else {
            GPB_TRY(bcast(d, oj == me ? d->panel(j) : P, P, count, oj, s_comm));
        }
        cudaEvent_t eb = d->ev(ev++);
        GPB_CUDA(cudaEventRecord(eb, s_comm));
        GPB_CUDA(cudaStreamWaitEvent(s, eb, 0));
        const int cnt = j >= me ? (j - me) / G + 1 : 0;  // owned row blocks a <= j
        const int rows_act = cnt * nbd;
        if (rows_act > 0) {
            LinalgWs ws{P + (size_t)d->panel_rows(j) * nbd, d->tmp, d->tmp_rows, d->info + nblk};
            GPB_TRY(trsm_right_lt(d->ystack + (size_t)j * nbd, npad, rows_act, P, nbd, cj, 0, ws, s));
            const int rest = npad - (int)d->row_end(j);
            if (rest > 0) {
                GemmArgs g{rows_act, rest, cj, d->ystack + (size_t)j * nbd, npad, P + (size_t)cj * nbd, nbd,
                           d->ystack + d->row_end(j), npad, d->ystack + d->row_end(j), npad, nullptr, 0, -1.0, 1.0, GEMM_FULL};
                GPB_TRY(gemm_nt(g, s));
            }
        }
        cudaEvent_t eu = d->ev(ev++);
        GPB_CUDA(cudaEventRecord(eu, s));
        used[step & 1] = eu;
    }
    // ---- phase 2: K^-1 blocks from the streamed row blocks of Y
    cudaEvent_t ev_fwd = d->ev(ev++);
    GPB_CUDA(cudaEventRecord(ev_fwd, s));
    GPB_CUDA(cudaStreamWaitEvent(s_comm, ev_fwd, 0));
    const int chunk = std::min(4096, std::max(1024, (npad / 4) / 64 * 64));
    for (int b = 0; b < nblk; ++b, ++step) {
        const int cb = d->cols_of(b), ob = d->owner(b);
        const int64_t ldb = (int64_t)npad - (int64_t)b * nbd;  // packed row length: columns b nbd .. npad
        double* P = pb[step & 1];
        if (used[step & 1]) GPB_CUDA(cudaStreamWaitEvent(s_comm, used[step & 1], 0));
        if (ob == me)
            GPB_CUDA(cudaMemcpy2DAsync(P, sizeof(double) * ldb, yrow((b - me) / G) + (size_t)b * nbd, sizeof(double) * npad,
                                       sizeof(double) * ldb, cb, cudaMemcpyDeviceToDevice, s_comm));
        GPB_TRY(bcast(d, P, P, (size_t)cb * ldb, ob, s_comm));
        cudaEvent_t eb = d->ev(ev++);
        GPB_CUDA(cudaEventRecord(eb, s_comm));
        GPB_CUDA(cudaStreamWaitEvent(s, eb, 0));
        int idx = b <= me ? 0 : (b - me + G - 1) / G;  // first owned row block with a >= b
        if (idx < na && me + G * idx == b) {             // diagonal block (lower tiles): k >= b nbd
            GemmArgs g{cb, cb, (int)ldb, yrow(idx) + (size_t)b * nbd, npad, P, ldb, nullptr, 0,
                       d->kdiag + (size_t)idx * nbd * nbd, nbd, nullptr, 0, 1.0, 0.0, GEMM_LOWER};
            g.max_k = chunk;
            GPB_TRY(gemm_nt(g, s));
            ++idx;
        }
        if (idx < na) {  // row blocks a > b, stacked: both operands are zero left of column a1 nbd, and row block i of
                         // the stack is zero for another i G nbd columns (block-structured GEMM_TRIK_A: N^3/3 flops in all)
            const int64_t k0 = (int64_t)(me + G * idx) * nbd;
            GemmArgs g{(na - idx) * nbd, cb, (int)(npad - k0), yrow(idx) + k0, npad, P + (k0 - (int64_t)b * nbd), ldb, nullptr, 0,
                       yrow(idx) + (size_t)b * nbd, npad, nullptr, 0, 1.0, 0.0, GEMM_TRIK_A};
            g.trik_a_blk = nbd;
            g.trik_a_step = G * nbd;
            g.max_k = chunk;
            GPB_TRY(gemm_nt(g, s));
        }
        cudaEvent_t eu = d->ev(ev++);
        GPB_CUDA(cudaEventRecord(eu, s));
        used[step & 1] = eu;
    }
    return 0;
}

// marginal_likelihood_gradient (regression.py:544-567) on the distributed layout: factor, alpha, the owned rows of K^-1,
// then every rank's share of the traces 1/2 sum (alpha alpha^T - K^-1) o dK_p over its row blocks (lml.cu, stacked
// variant) and one all-reduce of the p gradient entries.
int dist_lml_grad_impl(gpb_ctx* c, const double* theta, int block, double* lml, double* grad, int* info_out,
                       double* seconds_out) {
    gpb_dist* d = c->dist;
    if (seconds_out) seconds_out[3] = 0.0;
    GPB_TRY(dist_factor_impl(c, theta, block, info_out, seconds_out));
    *lml = d->lml;
    if (*info_out != 0) return 0;
    cudaStream_t s = c->s;
    EventGuard t;
    GPB_TRY(t.create());
    GPB_CUDA(cudaEventRecord(t.e[0], s));
    GPB_TRY(dist_alpha_impl(c));
    int na = 0;
    GPB_TRY(dist_inverse_rows(c, &na));
    const int p = c->n_mean + c->n_cov, npad = (int)d->npad;
    GPB_TRY(ensure(d->ggrad, d->ggrad_cap, sizeof(double) * (size_t)std::max(p, 1)));
    GPB_TRY(ensure(d->gpart, d->gpart_cap, std::max<size_t>(trace_partials_size_stacked(na, d->nbd, npad), 16)));
    GPB_CUDA(cudaMemsetAsync(d->ggrad, 0, sizeof(double) * (size_t)p, s));
    GPB_TRY(launch_lml_grad_stacked(d->cp, d->mp, c->n_mean, c->x, (int)c->n, npad, d->alpha_full, d->ystack, npad, d->kdiag,
                                    d->nbd, d->rank, d->world, na, d->rank == 0, d->gpart, d->ggrad, s));
    if (d->world > 1) GPB_NCCL(g_nccl.AllReduce(d->ggrad, d->ggrad, (size_t)p, ncclDouble, ncclSum, d->comm, s));
    GPB_CUDA(cudaEventRecord(t.e[1], s));
    GPB_CUDA(cudaMemcpyAsync(grad, d->ggrad, sizeof(double) * (size_t)p, cudaMemcpyDeviceToHost, s));
    GPB_CUDA(cudaStreamSynchronize(s));
    GPB_CUDA(cudaStreamSynchronize(d->s_comm));
    if (seconds_out) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.e[0], t.e[1]);
        seconds_out[3] = ms * 1e-3;
    }
    return 0;
}

}  // namespace

extern "C" {

int gpb_dist_unique_id(char* out128) {
    GPB_TRY(load_nccl());
    ncclUniqueId id;
    GPB_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

int gpb_dist_init(gpb_ctx* c, int rank, int world, const char* id128) {
    GPB_TRY(ctx_use(c));
    if (world < 1 || rank < 0 || rank >= world) {
        set_error("gpb_dist_init: bad rank/world");
        return -2;
    }
    dist_destroy(c);
    gpb_dist* d = new gpb_dist();
    d->rank = rank;
    d->world = world;
    int lo = 0, hi = 0;
    GPB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = highest priority
    GPB_CUDA(cudaStreamCreateWithPriority(&d->s_panel, cudaStreamNonBlocking, hi));
    GPB_CUDA(cudaStreamCreateWithPriority(&d->s_comm, cudaStreamNonBlocking, hi));
    if (world > 1) {
        GPB_TRY(load_nccl());
        ncclUniqueId id;
        std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
        GPB_NCCL(g_nccl.CommInitRank(&d->comm, world, id, rank));
    }
    c->dist = d;
    return 0;
}

// Host-only description of the block-column-cyclic layout used by the distributed path (no GPU needed): for `rank` of
// `world`, the number of owned block columns, the doubles of panel storage and of broadcast staging, and (optionally) the
// owner of every block column.  Tests use it to check that the ranks tile the matrix exactly once.
int gpb_dist_plan(int64_t n, int block, int world, int rank, int* n_blocks, int* n_owned, int64_t* panel_doubles,
                  int64_t* staging_doubles, int* owners_or_null) {
    if (n <= 0 || block < NB || block % NB || world < 1 || rank < 0 || rank >= world) {
        set_error("gpb_dist_plan: bad arguments");
        return -2;
    }
    const int64_t npad = round_up(n, NB), rows_aug = npad + AUG;
    const int nblk = (int)((npad + block - 1) / block);
    int owned = 0;
    int64_t total = 0;
    for (int j = 0; j < nblk; ++j) {
        if (owners_or_null) owners_or_null[j] = j % world;
        if (j % world == rank) {
            ++owned;
            total += (rows_aug - (int64_t)j * block) * block + (int64_t)block * NB;
        }
    }
    *n_blocks = nblk;
    *n_owned = owned;
    *panel_doubles = total;
    *staging_doubles = 2 * (rows_aug * block + (int64_t)block * NB);
    return 0;
}

int gpb_dist_finalize(gpb_ctx* c) {
    GPB_TRY(ctx_use(c));
    dist_destroy(c);
    return 0;
}

// Distributed set_hyperparameters (regression.py:218-244) for N beyond one GPU: K(theta) + sig assembled and factored in
// the block-column-cyclic layout; L stays in the panels, v = L^-1 (y - mu) on every rank.  block = panel width (multiple
// of 128).  seconds_out[0..2] = assemble, factor sweep, total (device time on this rank, CUDA events).
int gpb_dist_factor(gpb_ctx* c, const double* theta, int block, int* info_out, double* seconds_out) {
    GPB_TRY(check_dist(c, "gpb_dist_factor", false));
    if (c->has_ycov) {
        set_error("gpb_dist_factor: dense y_cov is not supported on the distributed path");
        return -2;
    }
    if (block < NB || block % NB) {
        set_error("gpb_dist_factor: block must be a positive multiple of 128");
        return -2;
    }
    return dist_fail(c->dist, dist_factor_impl(c, theta, block, info_out, seconds_out));
}

// Distributed marginal_likelihood(theta) (regression.py:528-542): the factorisation above plus
// -1/2 v.v - sum log L_ii (one 2-double all-reduce).
int gpb_dist_lml(gpb_ctx* c, const double* theta, int block, double* lml, int* info_out, double* seconds_out) {
    GPB_TRY(gpb_dist_factor(c, theta, block, info_out, seconds_out));
    *lml = c->dist->lml;
    return 0;
}

// alpha = L^-T v (regression.py:242-244), block back-substitution over the column panels: for i = last .. 0 the owner of
// block column i solves alpha_i = L_ii^-T (v_i - acc_i), broadcasts the nbd values, and every rank adds L_ij^T alpha_i to
// the running sums acc_j of the block columns j < i it owns.  alpha_out (host, n doubles) is filled on every rank.
int gpb_dist_alpha(gpb_ctx* c, double* alpha_out) {
    GPB_TRY(check_dist(c, "gpb_dist_alpha", true));
    gpb_dist* d = c->dist;
    auto body = [&]() -> int {
        GPB_TRY(dist_alpha_impl(c));
        if (alpha_out)
            GPB_CUDA(cudaMemcpyAsync(alpha_out, d->alpha_full, sizeof(double) * c->n, cudaMemcpyDeviceToHost, c->s));
        GPB_CUDA(cudaStreamSynchronize(c->s));
        return 0;
    };
    return dist_fail(d, body());
}

// Distributed marginal_likelihood_gradient(theta) (regression.py:544-567).  COLLECTIVE.  grad receives the
// n_mean + n_cov entries on every rank; seconds_out[0..3] = assemble, factor sweep, their sum, alpha + inverse rows +
// traces (device time on this rank).  The factorisation stays valid for gpb_dist_alpha / gpb_dist_predict.
int gpb_dist_lml_grad(gpb_ctx* c, const double* theta, int block, double* lml, double* grad, int* info_out,
                      double* seconds_out) {
    GPB_TRY(check_dist(c, "gpb_dist_lml_grad", false));
    if (c->has_ycov) {
        set_error("gpb_dist_lml_grad: dense y_cov is not supported on the distributed path");
        return -2;
    }
    if (c->n_regions > 0) {
        set_error("gpb_dist_lml_grad: ChangePoint kernels are not supported on the distributed path");
        return -2;
    }
    if (block < NB || block % NB) {
        set_error("gpb_dist_lml_grad: block must be a positive multiple of 128");
        return -2;
    }
    return dist_fail(c->dist, dist_lml_grad_impl(c, theta, block, lml, grad, info_out, seconds_out));
}

// GpRegressor.__call__ (regression.py:188-216) against the distributed factor.  COLLECTIVE: every rank calls it with its
// own slab of query points (m may differ between ranks and may be 0); the query points are the sharded unit.  Per chunk
// of queries the column panels of L are streamed through every rank (one ncclBroadcast per panel, double-buffered, so the
// transfer of panel j+1 hides behind the solve with panel j):  X_j = S_j L_jj^-T;  S_{>j} -= X_j L_{>j,j}^T, then
// mu = m(q) + X v (v = L^-1 r is replicated, so no backward solve is needed) and sigma = sqrt|k(q,q) - |X row|^2|.
int gpb_dist_predict(gpb_ctx* c, const double* q, int64_t m, double* mu, double* sig) {
    GPB_TRY(check_dist(c, "gpb_dist_predict", true));
    gpb_dist* d = c->dist;
    auto body = [&]() -> int {
        const int nbd = d->nbd, nblk = d->nblk, me = d->rank, npad = (int)d->npad, n = (int)c->n, dd = c->d;
        cudaStream_t s = c->s, s_comm = d->s_comm;
        // rows per chunk: S (rows x npad doubles) within ~1/5 of the device memory
        size_t free_b = 0, total_b = 0;
        GPB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        int64_t qc = std::min<int64_t>(32768, std::max<int64_t>(256, (int64_t)(total_b / 5 / (sizeof(double) * (size_t)npad)) / 256 * 256));
        int64_t my_chunks = (m + qc - 1) / qc, chunks = my_chunks;
        if (d->world > 1) {  // agree on the number of passes over the panels
            long long* dev = reinterpret_cast<long long*>(d->acc + 2);
            long long v = my_chunks;
            GPB_CUDA(cudaMemcpyAsync(dev, &v, sizeof(v), cudaMemcpyHostToDevice, s));
            GPB_NCCL(g_nccl.AllReduce(dev, dev, 1, ncclInt64, ncclMax, d->comm, s));
            GPB_CUDA(cudaMemcpyAsync(&v, dev, sizeof(v), cudaMemcpyDeviceToHost, s));
            GPB_CUDA(cudaStreamSynchronize(s));
            chunks = v;
        }
        if (chunks == 0) return 0;
        const int64_t rows_cap = std::max<int64_t>(256, round_up(std::min<int64_t>(qc, std::max<int64_t>(m, 1)), 256));
        if (m > 0) {
            GPB_TRY(ensure(c->S, c->S_cap, sizeof(double) * (size_t)rows_cap * npad));
            GPB_TRY(ensure(c->dots, c->dots_cap, sizeof(double) * (size_t)rows_cap));
            GPB_TRY(ensure(c->G, c->G_cap, sizeof(double) * (size_t)rows_cap));
            GPB_TRY(ensure(c->qbuf, c->qbuf_cap, sizeof(double) * (size_t)m * (dd + 2)));
            d->tmp_rows = std::max<int64_t>(d->tmp_rows, rows_cap);
            GPB_TRY(ensure(d->tmp, d->tmp_cap, sizeof(double) * (size_t)d->tmp_rows * NB));
            GPB_CUDA(cudaMemcpyAsync(c->qbuf, q, sizeof(double) * m * dd, cudaMemcpyHostToDevice, s));
        }
        double* qd = c->qbuf;
        double* mud = qd + (size_t)m * dd;
        double* sgd = mud + m;
        const size_t pb_doubles = (size_t)d->rows_aug() * nbd + (size_t)nbd * NB;
        double* pb[2] = {d->pbuf, d->pbuf + pb_doubles};
        size_t ev = 0;  // events: bcast(step), used(step) alternate
        std::vector<cudaEvent_t> used(2, nullptr);
        cudaEvent_t ev_ready = d->ev(ev++);
        GPB_CUDA(cudaEventRecord(ev_ready, s));       // panels / previous work on s complete before s_comm reads them
        GPB_CUDA(cudaStreamWaitEvent(s_comm, ev_ready, 0));
        int64_t step = 0;
        for (int64_t ch = 0; ch < chunks; ++ch) {
            const int64_t q0 = ch * qc;
            const int mq = (int)std::max<int64_t>(0, std::min<int64_t>(qc, m - q0));
            const int rows_pad = (int)round_up(mq, 256);
            if (mq > 0) {
                GPB_TRY(launch_cross_stack(d->cp, qd + q0 * dd, mq, 1, c->x, n, npad, c->S, npad, s));
                if (rows_pad > mq)
                    GPB_CUDA(cudaMemsetAsync(c->S + (size_t)mq * npad, 0, sizeof(double) * (size_t)(rows_pad - mq) * npad, s));
            }
            for (int j = 0; j < nblk; ++j, ++step) {
                const int cj = d->cols_of(j), ok = d->owner(j);
                double* P = pb[step & 1];
                if (used[step & 1]) GPB_CUDA(cudaStreamWaitEvent(s_comm, used[step & 1], 0));
                const size_t count = (size_t)d->panel_rows(j) * nbd + (size_t)nbd * NB;  // panel + inverted diagonal blocks
                if (ok == me && d->world == 1) {
                    GPB_CUDA(cudaMemcpyAsync(P, d->panel(j), sizeof(double) * count, cudaMemcpyDeviceToDevice, s_comm));
                } else {
                    GPB_TRY(bcast(d, ok == me ? d->panel(j) : P, P, count, ok, s_comm));
                }
                cudaEvent_t eb = d->ev(ev++);
                GPB_CUDA(cudaEventRecord(eb, s_comm));
                GPB_CUDA(cudaStreamWaitEvent(s, eb, 0));
                if (mq > 0) {
                    LinalgWs ws{P + (size_t)d->panel_rows(j) * nbd, d->tmp, d->tmp_rows, d->info + nblk};
                    GPB_TRY(trsm_right_lt(c->S + (size_t)j * nbd, npad, rows_pad, P, nbd, cj, 0, ws, s));
                    const int rest = npad - (int)d->row_end(j);
                    if (rest > 0) {
                        GemmArgs g{rows_pad, rest, cj, c->S + (size_t)j * nbd, npad, P + (size_t)cj * nbd, nbd,
                                   c->S + d->row_end(j), npad, c->S + d->row_end(j), npad, nullptr, 0, -1.0, 1.0, GEMM_FULL};
                        GPB_TRY(gemm_nt(g, s));
                    }
                }
                cudaEvent_t eu = d->ev(ev++);
                GPB_CUDA(cudaEventRecord(eu, s));
                used[step & 1] = eu;
            }
            if (mq > 0) {
                GPB_TRY(launch_row_dot(c->S, npad, mq, npad, d->v_full, c->dots, s));
                GPB_TRY(launch_row_gram(c->S, npad, mq, 1, npad, c->G, s));
                GPB_TRY(launch_finalize_predict(d->cp, d->mp, qd + q0 * dd, mq, 1, c->dots, c->G, mud + q0, sgd + q0, s));
            }
        }
        if (m > 0) {
            GPB_CUDA(cudaMemcpyAsync(mu, mud, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
            GPB_CUDA(cudaMemcpyAsync(sig, sgd, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
        }
        GPB_CUDA(cudaStreamSynchronize(s));
        GPB_CUDA(cudaStreamSynchronize(s_comm));
        return 0;
    };
    return dist_fail(d, body());
}

}  // extern "C"
