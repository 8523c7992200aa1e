// Process-wide state of libgpb200: launch / flop counters and runtime options.
//
// GpRegressor(n_processes > 1) and gpb_lml_grad_batch drive several contexts from worker threads, so everything here
// is either atomic or thread-local.  Options replace the GPB200_* environment switches of round 1 (the variables still
// provide the initial values): tests flip the GEMM path per call instead of per process.
#include "common.cuh"
#include "../../include/gpb200.h"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace gpb {
namespace {

std::atomic<int64_t> g_launches{0};
std::atomic<double> g_flops{0.0}, g_flops_i8{0.0};
thread_local int64_t t_launches = 0;
thread_local double t_flops = 0.0, t_flops_i8 = 0.0;
thread_local int t_i8_override = -1;

void atomic_add(std::atomic<double>& a, double v) {
    double cur = a.load(std::memory_order_relaxed);
    while (!a.compare_exchange_weak(cur, cur + v, std::memory_order_relaxed)) {
    }
}

struct OptionTable {
    std::atomic<int64_t> v[OPT_COUNT];
    std::atomic<uint64_t> epoch{0};
    static int64_t env(const char* name, int64_t dflt) {
        const char* e = getenv(name);
        return e ? atoll(e) : dflt;
    }
    OptionTable() {
        v[OPT_GEMM_I8] = env("GPB200_GEMM_I8", 1);
        v[OPT_GEMM_I8_MIN_K] = env("GPB200_GEMM_I8_MINK", 512);
        v[OPT_GEMM_I8_PAIR] = env("GPB200_GEMM_I8_PAIR", 1);
        v[OPT_GEMM_I8_DEBUG] = env("GPB200_GEMM_I8_DEBUG", 0);
        v[OPT_GEMM_TILE] = env("GPB200_GEMM_TILE", 0);
        v[OPT_GEMM_TMA] = env("GPB200_GEMM_TMA", 1);
        v[OPT_GRAPHS] = getenv("GPB200_NO_GRAPHS") ? 0 : 1;
        v[OPT_I8_FALLBACK] = env("GPB200_I8_FALLBACK", 1);
        v[OPT_PREDICT_BLOCK] = env("GPB200_PREDICT_BLOCK", 0);
        v[OPT_PREDICT_DIAG] = env("GPB200_PREDICT_DIAG", 1);
        v[OPT_I8_GRAD_GUARD] = env("GPB200_I8_GRAD_GUARD", 1);
        v[OPT_GEMM_I8_MAX_K] = env("GPB200_GEMM_I8_MAX_K", 16384);
        v[OPT_GEMM_I8_EPI] = env("GPB200_GEMM_I8_EPI", 0);
        v[OPT_I8_GRAD_PHASES] = env("GPB200_I8_GRAD_PHASES", 7);
        v[OPT_GRAD_INVERSE] = env("GPB200_GRAD_INVERSE", 0);
        v[OPT_GEMM_I8_EPI2] = env("GPB200_GEMM_I8_EPI2", 2);
        v[OPT_GEMM_I8_PREFETCH] = env("GPB200_GEMM_I8_PREFETCH", 0);
    }
};
OptionTable& table() {
    static OptionTable t;  // thread-safe initialisation (C++11)
    return t;
}

const char* const kNames[OPT_COUNT] = {"gemm_i8",  "gemm_i8_min_k", "gemm_i8_pair", "gemm_i8_debug", "gemm_tile",
                                       "gemm_tma", "graphs",        "i8_fallback",  "predict_block", "predict_diag", "i8_grad_guard", "gemm_i8_max_k", "gemm_i8_epi", "i8_grad_phases", "grad_inverse", "gemm_i8_epi2", "gemm_i8_prefetch"};

int find_option(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < OPT_COUNT; ++i)
        if (std::strcmp(name, kNames[i]) == 0) return i;
    return -1;
}

}  // namespace

void count_launch(int n) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    t_launches += n;
}
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }
int64_t thread_launch_count() { return t_launches; }
void credit_gemm_flops(double f, double f_i8) {
    if (f != 0.0) atomic_add(g_flops, f);
    if (f_i8 != 0.0) atomic_add(g_flops_i8, f_i8);
    t_flops += f;
    t_flops_i8 += f_i8;
}
double gemm_flops_issued() { return g_flops.load(std::memory_order_relaxed); }
double gemm_flops_issued_i8() { return g_flops_i8.load(std::memory_order_relaxed); }
double thread_gemm_flops() { return t_flops; }
double thread_gemm_flops_i8() { return t_flops_i8; }

int64_t option(Option o) { return table().v[o].load(std::memory_order_relaxed); }
uint64_t option_epoch() { return table().epoch.load(std::memory_order_acquire); }
int gemm_i8_override() { return t_i8_override; }
void set_gemm_i8_override(int v) { t_i8_override = v; }

}  // namespace gpb

extern "C" {

int gpb_set_option(const char* name, int64_t value) {
    const int i = gpb::find_option(name);
    if (i < 0) {
        gpb::set_error(std::string("gpb_set_option: unknown option '") + (name ? name : "(null)") + "'");
        return -2;
    }
    gpb::table().v[i].store(value, std::memory_order_relaxed);
    gpb::table().epoch.fetch_add(1, std::memory_order_release);
    return 0;
}

int gpb_get_option(const char* name, int64_t* value) {
    const int i = gpb::find_option(name);
    if (i < 0 || !value) {
        gpb::set_error(std::string("gpb_get_option: unknown option '") + (name ? name : "(null)") + "'");
        return -2;
    }
    *value = gpb::table().v[i].load(std::memory_order_relaxed);
    return 0;
}

int64_t gpb_launch_count(void) { return gpb::launch_count(); }
double gpb_gemm_flops(void) { return gpb::gemm_flops_issued(); }
double gpb_gemm_flops_int8(void) { return gpb::gemm_flops_issued_i8(); }

}  // extern "C"
