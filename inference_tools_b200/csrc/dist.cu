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

// aug rows of one panel: row 0 = residual slice of the block's columns, rows 1..127 = 0
__global__ void fill_aug_rows_kernel(double* __restrict__ aug, int64_t ld, int ncols, const double* __restrict__ resid) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    aug[c] = resid[c];
    for (int a = 1; a < NB; ++a) aug[(int64_t)a * ld + c] = 0.0;
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

}  // namespace
}  // namespace gpb

using namespace gpb;

struct gpb_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t s_panel = nullptr, s_comm = nullptr;
    double *panels = nullptr, *pbuf = nullptr, *dinv = nullptr, *tmp = nullptr, *acc = nullptr, *resid = nullptr;
    size_t panels_cap = 0, pbuf_cap = 0, dinv_cap = 0, tmp_cap = 0, acc_cap = 0, resid_cap = 0;
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
};

namespace gpb {
void dist_destroy(gpb_ctx* c) {
    gpb_dist* d = c->dist;
    if (!d) return;
    for (double* p : {d->panels, d->pbuf, d->dinv, d->tmp, d->acc, d->resid})
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

// Host-only description of the block-column-cyclic layout used by gpb_dist_lml (no GPU needed): for `rank` of `world`,
// the number of owned block columns, the doubles of panel storage and of broadcast staging, and (optionally) the owner
// of every block column.  Tests use it to check that the ranks tile the matrix exactly once.
int gpb_dist_plan(int64_t n, int block, int world, int rank, int* n_blocks, int* n_owned, int64_t* panel_doubles,
                  int64_t* staging_doubles, int* owners_or_null) {
    if (n <= 0 || block < NB || block % NB || world < 1 || rank < 0 || rank >= world) {
        set_error("gpb_dist_plan: bad arguments");
        return -2;
    }
    const int64_t npad = round_up(n, NB), rows_aug = npad + NB;
    const int nblk = (int)((npad + block - 1) / block);
    int owned = 0;
    int64_t total = 0;
    for (int j = 0; j < nblk; ++j) {
        if (owners_or_null) owners_or_null[j] = j % world;
        if (j % world == rank) {
            ++owned;
            total += (rows_aug - (int64_t)j * block) * block;
        }
    }
    *n_blocks = nblk;
    *n_owned = owned;
    *panel_doubles = total;
    *staging_doubles = 2 * rows_aug * block;
    return 0;
}

int gpb_dist_finalize(gpb_ctx* c) {
    GPB_TRY(ctx_use(c));
    dist_destroy(c);
    return 0;
}

// Distributed marginal_likelihood(theta) (regression.py:528-542).  block = panel width (multiple of 128).
// seconds_out[0..2] = assemble, factor sweep, total (device time on this rank, CUDA events).
int gpb_dist_lml(gpb_ctx* c, const double* theta, int block, double* lml, int* info_out, double* seconds_out) {
    GPB_TRY(ctx_use(c));
    GPB_TRY(ctx_need_model(c));
    gpb_dist* d = c->dist;
    if (!d) {
        set_error("gpb_dist_lml: call gpb_dist_init first");
        return -2;
    }
    if (c->has_ycov) {
        set_error("gpb_dist_lml: dense y_cov is not supported on the distributed path");
        return -2;
    }
    if (block < NB || block % NB) {
        set_error("gpb_dist_lml: block must be a positive multiple of 128");
        return -2;
    }
    const int G = d->world, me = d->rank;
    const int npad = (int)c->npad, n = (int)c->n, nbd = block;
    const int nblk = (npad + nbd - 1) / nbd;
    const int64_t rows_aug = (int64_t)npad + NB;
    auto cols_of = [&](int j) { return std::min(nbd, npad - j * nbd); };
    auto row_end = [&](int j) { return j * nbd + cols_of(j); };
    auto owner = [&](int j) { return j % G; };

    // ---- storage: owned panels back to back
    std::vector<size_t> off(nblk, 0);
    size_t total = 0;
    for (int j = 0; j < nblk; ++j)
        if (owner(j) == me) {
            off[j] = total;
            total += (size_t)(rows_aug - (int64_t)j * nbd) * nbd;
        }
    GPB_TRY(ensure(d->panels, d->panels_cap, sizeof(double) * std::max<size_t>(total, 1)));
    GPB_TRY(ensure(d->pbuf, d->pbuf_cap, sizeof(double) * 2 * (size_t)rows_aug * nbd));
    GPB_TRY(ensure(d->dinv, d->dinv_cap, sizeof(double) * (size_t)nbd * NB));
    GPB_TRY(ensure(d->tmp, d->tmp_cap, sizeof(double) * (size_t)rows_aug * NB));
    GPB_TRY(ensure(d->acc, d->acc_cap, sizeof(double) * 4));
    GPB_TRY(ensure(d->resid, d->resid_cap, sizeof(double) * (size_t)npad));
    GPB_TRY(ensure(d->info, d->info_cap, sizeof(int) * (size_t)(nblk + 1)));
    double* pb[2] = {d->pbuf, d->pbuf + (size_t)rows_aug * nbd};
    auto panel = [&](int j) { return d->panels + off[j]; };

    CovParams cp;
    MeanParams mp;
    GPB_TRY(ctx_make_cov_params(c, theta + c->n_mean, cp));
    ctx_make_mean_params(c, theta, mp);

    cudaStream_t s_main = c->s, s_panel = d->s_panel, s_comm = d->s_comm;
    cudaEvent_t t0, t1, t2;
    GPB_CUDA(cudaEventCreate(&t0));
    GPB_CUDA(cudaEventCreate(&t1));
    GPB_CUDA(cudaEventCreate(&t2));
    // event slots
    const size_t EV_ASM = 0;
    auto EV_PANEL = [&](int k) { return (size_t)1 + 4 * (size_t)k; };
    auto EV_BCAST = [&](int k) { return (size_t)2 + 4 * (size_t)k; };
    auto EV_UPD = [&](int k) { return (size_t)3 + 4 * (size_t)k; };
    auto EV_LA = [&](int k) { return (size_t)4 + 4 * (size_t)k; };
    std::vector<char> did_la(nblk, 0);

    // ---- assemble owned panels + residual row
    GPB_CUDA(cudaEventRecord(t0, s_main));
    GPB_CUDA(cudaMemsetAsync(d->info, 0, sizeof(int) * (nblk + 1), s_main));
    GPB_CUDA(cudaMemsetAsync(d->acc, 0, sizeof(double) * 4, s_main));
    GPB_TRY(launch_residual(mp, c->x, c->y, n, npad, d->resid, nullptr, s_main));
    for (int j = 0; j < nblk; ++j) {
        if (owner(j) != me) continue;
        const int cj = cols_of(j);
        GPB_TRY(launch_assemble_block(cp, c->x, n, c->has_noise ? c->noise : nullptr, j * nbd, npad - j * nbd, j * nbd, cj,
                                      panel(j), nbd, s_main));
        fill_aug_rows_kernel<<<(cj + 255) / 256, 256, 0, s_main>>>(panel(j) + (size_t)(npad - j * nbd) * nbd, nbd, cj,
                                                                  d->resid + j * nbd);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    GPB_CUDA(cudaEventRecord(d->ev(EV_ASM), s_main));
    GPB_CUDA(cudaEventRecord(t1, s_main));
    GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_ASM), 0));

    auto factor_panel = [&](int k) -> int {  // on s_panel; column k must be fully updated (stream order / waits)
        const int ck = cols_of(k);
        LinalgWs ws{d->dinv, d->tmp, rows_aug, d->info + nblk};  // potrf_lower clears ws.info at entry
        GPB_TRY(potrf_lower(panel(k), nbd, ck, ws, s_panel));
        // keep the first failure of this panel as a global 1-based index in info[k]
        // (device-side copy: info[k] = info[nblk] ? info[nblk] + k*nbd : 0 is done on the host after the sweep)
        GPB_CUDA(cudaMemcpyAsync(d->info + k, d->info + nblk, sizeof(int), cudaMemcpyDeviceToDevice, s_panel));
        const int64_t below = rows_aug - row_end(k);
        GPB_TRY(trsm_right_lt(panel(k) + (size_t)ck * nbd, nbd, (int)below, panel(k), nbd, ck, 0, ws, s_panel));
        GPB_CUDA(cudaEventRecord(d->ev(EV_PANEL(k)), s_panel));
        return 0;
    };
    auto update_col = [&](int j, int k, cudaStream_t s) -> int {  // C_j -= P_k[rows >= j*nbd] P_k[block j rows]^T
        const double* P = pb[k & 1] + (size_t)(j * nbd - row_end(k)) * nbd;
        GemmArgs g{(int)(rows_aug - (int64_t)j * nbd), cols_of(j), cols_of(k), P, nbd, P, nbd, panel(j), nbd, panel(j), nbd,
                   nullptr, 0, -1.0, 1.0, GEMM_FULL};
        return gemm_nt(g, s);
    };

    if (owner(0) == me) GPB_TRY(factor_panel(0));
    for (int k = 0; k < nblk; ++k) {
        const int ok = owner(k);
        const int64_t below = rows_aug - row_end(k);
        const size_t count = (size_t)below * nbd;
        // (1) broadcast panel k (rows below the diagonal block) into Pbuf[k & 1]
        if (k >= 2) {
            GPB_CUDA(cudaStreamWaitEvent(s_comm, d->ev(EV_UPD(k - 2)), 0));
            if (did_la[k - 2]) GPB_CUDA(cudaStreamWaitEvent(s_comm, d->ev(EV_LA(k - 2)), 0));
        }
        if (ok == me) GPB_CUDA(cudaStreamWaitEvent(s_comm, d->ev(EV_PANEL(k)), 0));
        if (G > 1) {
            const double* src = (ok == me) ? panel(k) + (size_t)cols_of(k) * nbd : pb[k & 1];
            GPB_NCCL(g_nccl.Broadcast(src, pb[k & 1], count, ncclDouble, ok, d->comm, s_comm));
        } else {
            GPB_CUDA(cudaMemcpyAsync(pb[k & 1], panel(k) + (size_t)cols_of(k) * nbd, sizeof(double) * count,
                                     cudaMemcpyDeviceToDevice, s_comm));
        }
        GPB_CUDA(cudaEventRecord(d->ev(EV_BCAST(k)), s_comm));
        // (2) look-ahead: the owner of panel k+1 brings its column up to date and factors it
        if (k + 1 < nblk && owner(k + 1) == me) {
            GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_BCAST(k)), 0));
            if (k >= 1) GPB_CUDA(cudaStreamWaitEvent(s_panel, d->ev(EV_UPD(k - 1)), 0));
            GPB_TRY(update_col(k + 1, k, s_panel));
            GPB_CUDA(cudaEventRecord(d->ev(EV_LA(k)), s_panel));
            did_la[k] = 1;
            GPB_TRY(factor_panel(k + 1));
        }
        // (3) trailing update of the other owned columns
        GPB_CUDA(cudaStreamWaitEvent(s_main, d->ev(EV_BCAST(k)), 0));
        for (int j = k + 2; j < nblk; ++j)
            if (owner(j) == me) GPB_TRY(update_col(j, k, s_main));
        GPB_CUDA(cudaEventRecord(d->ev(EV_UPD(k)), s_main));
    }
    // ---- reductions: log det and v.v from the owned panels
    for (int j = 0; j < nblk; ++j) {
        if (owner(j) != me) continue;
        GPB_CUDA(cudaStreamWaitEvent(s_main, d->ev(EV_PANEL(j)), 0));
        const int cj = cols_of(j);
        panel_reduce_kernel<<<1, 1024, 0, s_main>>>(panel(j), nbd, cj, panel(j) + (size_t)(npad - j * nbd) * nbd,
                                                    std::max(0, std::min(cj, n - j * nbd)), d->acc);
        GPB_CUDA(cudaGetLastError());
        count_launch();
    }
    if (G > 1) GPB_NCCL(g_nccl.AllReduce(d->acc, d->acc, 2, ncclDouble, ncclSum, d->comm, s_main));
    GPB_CUDA(cudaEventRecord(t2, s_main));
    double acc[2];
    std::vector<int> info_h(nblk + 1, 0);
    GPB_CUDA(cudaMemcpyAsync(acc, d->acc, sizeof(acc), cudaMemcpyDeviceToHost, s_main));
    GPB_CUDA(cudaStreamSynchronize(s_main));
    GPB_CUDA(cudaStreamSynchronize(s_panel));
    GPB_CUDA(cudaStreamSynchronize(s_comm));
    GPB_CUDA(cudaMemcpy(info_h.data(), d->info, sizeof(int) * (nblk + 1), cudaMemcpyDeviceToHost));
    int first_bad = 0;
    for (int j = 0; j < nblk && !first_bad; ++j)
        if (owner(j) == me && info_h[j] > 0) first_bad = info_h[j] + j * nbd;
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
    *lml = -0.5 * acc[1] - acc[0];
    if (seconds_out) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, t0, t1);
        cudaEventElapsedTime(&b, t1, t2);
        seconds_out[0] = a * 1e-3;
        seconds_out[1] = b * 1e-3;
        seconds_out[2] = (a + b) * 1e-3;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaEventDestroy(t2);
    return 0;
}

}  // extern "C"
