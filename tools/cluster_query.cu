// How many thread-block clusters of size 2 / 4 / 8 with one 226 KB CTA per SM can be resident on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cluster_query tools/cluster_query.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    const int smem = 230656;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"sms\": %d", sms);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(sms / cs * cs);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf(", \"cluster%d\": %d", cs, e == cudaSuccess ? n : -(int)e);
    }
    printf("}\n");
    return 0;
}
