// FP64 issue-rate probe for B200 (sm_100a): DMMA.8x8x4 (mma.sync m8n8k4 f64) vs DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
// Output: one JSON line with achieved TFLOP/s for both pipes (the FP64 roofline denominators
// that MEASURED_PEAKS.json does not carry).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
    double c0[NACC], c1[NACC];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
    double c[NACC];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
        int grid = sms * ctas_per_sm, iters = 20000;
        dmma_loop<16><<<grid, 256>>>(out, 100, 1.0); CK(cudaDeviceSynchronize());
        double best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            dmma_loop<16><<<grid, 256>>>(out, iters, 1.0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = (double)grid * 8 /*warps*/ * iters * 16 * 512.0;
            double tf = flops / (ms * 1e-3) / 1e12; if (tf > best) best = tf;
        }
        printf(", \"dmma_tflops_%dcta\": %.2f", ctas_per_sm, best);
        dfma_loop<16><<<grid, 256>>>(out, 100, 1.0); CK(cudaDeviceSynchronize());
        best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            dfma_loop<16><<<grid, 256>>>(out, iters * 4, 1.0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = (double)grid * 256 * (iters * 4.0) * 16 * 2.0;
            double tf = flops / (ms * 1e-3) / 1e12; if (tf > best) best = tf;
        }
        printf(", \"dfma_tflops_%dcta\": %.2f", ctas_per_sm, best);
    }
    // sustained DMMA for ~3 s to see the power-capped rate
    {
        int grid = sms * 2, iters = 20000; double tot_ms = 0; int n = 0;
        while (tot_ms < 3000) {
            CK(cudaEventRecord(e0));
            dmma_loop<16><<<grid, 256>>>(out, iters, 1.0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            tot_ms += ms; ++n;
        }
        double flops = (double)n * grid * 8 * iters * 16 * 512.0;
        printf(", \"dmma_tflops_sustained\": %.2f", flops / (tot_ms * 1e-3) / 1e12);
    }
    printf("}\n");
    return 0;
}
