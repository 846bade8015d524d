// micro-benchmark: how many clusters of 8 / 4 / 2 CTAs (one CTA per SM: 220 KB of shared memory, 576 threads) fit on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *x) { extern __shared__ int s[]; if (x) s[threadIdx.x] = *x; }
int main() {
    const int smem = 223872, threads = 576;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cs > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: %d clusters co-resident = %d CTAs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
