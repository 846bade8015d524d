// micro-benchmark: cost of cooperative_groups grid.sync() on this GPU for persistent grids of several shapes
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k(int n, unsigned *sink) {
    cg::grid_group g = cg::this_grid();
    unsigned x = threadIdx.x;
    for (int i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; g.sync(); }
    if (x == 0xdeadbeef) *sink = x;
}
// hand-made barrier: one atomic counter per sync, thread 0 of each CTA arrives and spins on a generation word
__global__ void k2(int n, unsigned *bar, unsigned *sink) {
    unsigned x = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        x = x * 1664525u + 1013904223u;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned target = (unsigned)(i + 1) * gridDim.x;
            atomicAdd(bar, 1u);
            while (*((volatile unsigned *)bar) < target) {}
            __threadfence();
        }
        __syncthreads();
    }
    if (x == 0xdeadbeef) *sink = x;
}
int main() {
    unsigned *sink, *bar; cudaMalloc(&sink, 4); cudaMalloc(&bar, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int n = 200;
    for (int per_sm : {1, 2, 4, 8}) for (int threads : {128, 256}) {
        int ctas = 148 * per_sm;
        if ((long)ctas * threads > 148L * 2048) continue;
        void *args[] = {&n, &sink};
        cudaLaunchCooperativeKernel((void *)k, dim3(ctas), dim3(threads), args, 0, 0);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((void *)k, dim3(ctas), dim3(threads), args, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemset(bar, 0, 4);
        void *args2[] = {&n, &bar, &sink};
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((void *)k2, dim3(ctas), dim3(threads), args2, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms2; cudaEventElapsedTime(&ms2, e0, e1);
        printf("ctas %d x %d threads: grid.sync %.2f us, hand-made barrier %.2f us (err %s)\n", ctas, threads, 1e3 * ms / n, 1e3 * ms2 / n,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
