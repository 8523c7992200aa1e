// Sustained full-chip rate of tcgen05.mma kind::i8 (the denominator of bench.py's roofline): every SM issues
// back-to-back 128 x 128 x 32 int8 MMAs (cta_group::1, both operands resident in shared memory with the 64-byte swizzle
// gemm_i8_kernel uses, 7 x 7 digit-plane pattern, four TMEM accumulators) for a chosen number of seconds; the rate is
// MMAs x 2 x 128 x 128 x 32 / CUDA-event time.  No global-memory traffic: this is the pipe's own ceiling at the clock the
// power limit allows, which is what a GEMM kernel can at best approach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int8_peak tools/int8_peak.cu;  ./tools/int8_peak [seconds]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(unsigned addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void mma_i8(unsigned d, uint64_t a, uint64_t b, unsigned idesc, unsigned acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.b32 %0, 1, 0, px;\n}\n" : "=r"(pred));
    return pred != 0;
}

constexpr int N = 128;
__global__ void __launch_bounds__(128, 1) peak_kernel(long long iters) {
    extern __shared__ unsigned char raw[];
    unsigned char* tiles = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ unsigned long long bar;
    __shared__ unsigned slot;
    // pseudo-random digits: constant operands toggle no data lines and stay far below the power limit a real GEMM hits
    for (int i = threadIdx.x; i < 112 * 1024 / 4; i += blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        ((unsigned*)tiles)[i] = h;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned tm = slot;
    constexpr unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | (8u << 24);
    if (threadIdx.x < 32) {
        const unsigned a_base = smem_u32(tiles), b_base = a_base + 7 * 8192;
        unsigned phase = 0;
        for (long long it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < 7; ++s)
#pragma unroll
                    for (int t = 0; t < 7 - s; ++t)
                        mma_i8(tm + (unsigned)(((s + t) & 3) * N), smem_desc(a_base + s * 8192), smem_desc(b_base + t * 8192), idesc, 1);
                // bound the in-flight queue like a real pipeline does: commit + wait every 64 steps
                if ((it & 63) == 63)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)));
            }
            __syncwarp();
            if ((it & 63) == 63) {
                unsigned done;
                do {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n"
                                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                } while (!done);
                phase ^= 1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tm));
}

int main(int argc, char** argv) {
    const double want_s = argc > 1 ? atof(argv[1]) : 3.0;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int smem = 120 * 1024;
    cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](long long iters) -> double {
        cudaEventRecord(e0);
        peak_kernel<<<sms, 128, smem>>>(iters);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms * 1e-3;
    };
    const long long probe_iters = 64 * 1024;
    run(probe_iters);                                     // warm-up
    const double t_probe = run(probe_iters);              // burst figure: a few tens of ms
    const long long iters = (long long)(probe_iters * (want_s / t_probe)) / 64 * 64;
    const double t = run(iters);                          // sustained figure
    const double ops_per_iter = 28.0 * 2.0 * 128 * N * 32 * sms;
    printf("{\"sms\": %d, \"burst_seconds\": %.4f, \"burst_int8_tops\": %.1f, \"sustained_seconds\": %.3f, \"sustained_int8_tops\": %.1f, "
           "\"mma_shape\": \"128x128x32 cta_group::1, SS, 7x7 digit-plane pattern\", \"fp64_equivalent_tflops_at_28_products\": %.1f}\n",
           sms, t_probe, ops_per_iter * probe_iters / t_probe / 1e12, t, ops_per_iter * iters / t / 1e12,
           ops_per_iter * iters / t / 1e12 / 28.0);
    return 0;
}
