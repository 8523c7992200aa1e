// Issue-rate probe for tcgen05.mma kind::i8 (M = 128) at several N, operands resident in shared memory (no TMA):
// clocks per MMA when one thread issues a long back-to-back stream.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
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
__device__ __forceinline__ void mma_i8_ts(unsigned d, unsigned a, uint64_t b, unsigned idesc, unsigned acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.b32 %0, 1, 0, px;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_cp(unsigned dst, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(dst), "l"(sdesc) : "memory");
}

// MODE 0: SS, 7 A planes x 7 B planes (28 MMAs per step, like gemm_i8_kernel); 1: TS with the 7 tcgen05.cp per step;
// 2: TS without the copies; 3: SS, one A tile against NB_T B tiles (A reused back to back)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* tiles = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ unsigned long long bar;
    __shared__ unsigned slot;
    for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) ((unsigned*)tiles)[i] = 0x01010101u;
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
    constexpr int NACC = 448 / N;
    if (threadIdx.x < 32) {
        const unsigned a_base = smem_u32(tiles), b_base = a_base + 7 * 8192;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                if (MODE == 1) {
#pragma unroll
                    for (int s = 0; s < 7; ++s) tmem_cp(tm + 448 + 8 * s, smem_desc(a_base + s * 8192));
                }
#pragma unroll
                for (int s = 0; s < 7; ++s)
#pragma unroll
                    for (int t = 0; t < 7 - s; ++t) {
                        const unsigned d = tm + (unsigned)(((s + t) % NACC) * N);
                        const uint64_t bd = smem_desc(b_base + t * (N * 64));
                        if (MODE == 1 || MODE == 2) mma_i8_ts(d, tm + 448 + 8 * s, bd, idesc, 1);
                        else mma_i8(d, smem_desc(a_base + (MODE == 3 ? 0 : s * 8192)), bd, idesc, 1);
                    }
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)));
        __syncwarp();
        unsigned done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        } while (!done);
        long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tm));
}

template <int N, int MODE>
void run(long long* out) {
    const int smem = 200 * 1024, iters = 500;
    cudaFuncSetAttribute(probe<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N, MODE><<<148, 128, smem>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long clk = 0;
    cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
    const double per = (double)clk / (iters * 28.0);
    printf("mode %d N %3d : %.1f clk/MMA, %.0f MAC/clk/SM (%s)\n", MODE, N, per, 128.0 * N * 32 / per, cudaGetErrorString(e));
}

int main() {
    long long* out;
    cudaMalloc(&out, 8);
    run<64, 0>(out); run<64, 1>(out); run<64, 2>(out); run<64, 3>(out);
    run<128, 0>(out); run<128, 1>(out); run<128, 2>(out); run<128, 3>(out);
    run<224, 0>(out); run<224, 2>(out); run<224, 3>(out);
    return 0;
}
