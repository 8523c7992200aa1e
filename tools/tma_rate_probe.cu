// Is gemm_i8_kernel's operand stream limited by TMA REQUESTS or by bytes?  Every SM streams 56 KB boxes of int8 digit
// planes (the GEMM's 3-D tensor map: k, row, plane) through a ring of 4 shared-memory slots with no consumer, walking k
// like the GEMM does (groups of 8 CTAs share a row block, so most boxes hit in L2).  Variants: 64-byte box rows
// (64 x 128 rows x 7 planes, SWIZZLE_64B -- what the kernel uses) against 128-byte rows (128 x 64 rows x 7 planes,
// SWIZZLE_128B): the same bytes in half the L2 requests.  Prints GB/s per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_rate_probe tools/tma_rate_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

constexpr int SLOT = 57344, SLOTS = 4;
// rows_per_box x row_bytes x 7 planes = 56 KB; k advances by row_bytes per box, CTA c works on row block c / 8
__global__ void __launch_bounds__(32, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int row_bytes,
                                                       int rows_per_box, int K, int row_blocks, int boxes) {
    extern __shared__ unsigned char raw[];
    unsigned char* tiles = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ unsigned long long bars[SLOTS];
    if (threadIdx.x == 0) {
        for (int i = 0; i < SLOTS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const int rb = (blockIdx.x / 8) % row_blocks;
    const int kblocks = K / row_bytes;
    int kb = (blockIdx.x * 7) % kblocks;  // de-phase the CTAs a little, as tiles do
    for (int i = 0; i < boxes; ++i) {
        const int s = i % SLOTS;
        const unsigned bar = smem_u32(&bars[s]);
        if (i >= SLOTS) mbar_wait(bar, ((i / SLOTS) - 1) & 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(SLOT) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                "r"(smem_u32(tiles + s * SLOT)), "l"(&map), "r"(kb * row_bytes), "r"(rb * rows_per_box), "r"(0), "r"(bar)
            : "memory");
        if (++kb == kblocks) kb = 0;
    }
    for (int i = boxes; i < boxes + SLOTS; ++i)
        if (i >= SLOTS) mbar_wait(smem_u32(&bars[i % SLOTS]), ((i / SLOTS) - 1) & 1);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 8192, K = argc > 2 ? atoi(argv[2]) : 8192, boxes = argc > 3 ? atoi(argv[3]) : 4096;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 1;
    EncodeFn enc = (EncodeFn)fn;
    signed char* buf;
    const size_t bytes = (size_t)7 * rows * K;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return 1;
    cudaMemset(buf, 1, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOTS * SLOT + 1024);
    printf("{\"rows\": %d, \"K\": %d, \"boxes_per_sm\": %d, \"sms\": %d", rows, K, boxes, sms);
    for (int variant = 0; variant < 3; ++variant) {
        const int row_bytes = variant == 0 ? 64 : 128;
        const int rpb = variant == 0 ? 128 : 64;
        // variant 2: 128-byte rows WITHOUT the L2 128-byte promotion hint
        CUtensorMap m;
        const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 7};
        const cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)rows * K};
        const cuuint32_t box[3] = {(cuuint32_t)row_bytes, (cuuint32_t)rpb, 7};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                variant == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 2;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            stream_kernel<<<sms, 32, SLOTS * SLOT + 1024>>>(m, row_bytes, rpb, K, rows / rpb, boxes);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) return 3;
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double gbs = (double)sms * boxes * SLOT / (best * 1e-3) / 1e9;
        printf(", \"%s\": {\"ms\": %.3f, \"GBps\": %.0f}",
               variant == 0 ? "rows64B" : (variant == 1 ? "rows128B" : "rows128B_nopromo"), best, gbs);
    }
    printf("}\n");
    return 0;
}
