// nlmc_npmath.cu -- array entry points for the numpy-equivalent float64 tanh / arctanh of nlmc_npmath.h.
// They back the public helper atanh_saturated (NMC/nmc.py:230-255) and let the parity tests compare the device
// functions the LBP and replay kernels use against np.tanh / np.arctanh argument by argument.
#include "nlmc_common.cuh"
#include "nlmc_npmath.h"

namespace nlmc {

template <int WHICH>
__global__ void npmath_kernel(const double *__restrict__ x, double *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = WHICH == 0 ? nlmc_np_tanh(x[i]) : nlmc_np_arctanh(x[i]);
}

static int npmath_run(int which, const double *x, double *out, int64_t n, int device) {
    NLMC_REQUIRE(n >= 0 && (n == 0 || (x && out)), "nlmc_np_%s: bad arguments", which ? "arctanh" : "tanh");
    NLMC_CUDA(cudaSetDevice(device));
    if (n == 0) return NLMC_OK;
    double *d = nullptr;
    NLMC_CUDA(cudaMalloc(&d, sizeof(double) * 2 * (size_t)n));
    cudaError_t e = cudaMemcpy(d, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
        if (which == 0) npmath_kernel<0><<<grid, 256>>>(d, d + n, n);
        else npmath_kernel<1><<<grid, 256>>>(d, d + n, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error("nlmc_np_%s: %s", which ? "arctanh" : "tanh", cudaGetErrorString(e));
        return NLMC_ERR_CUDA;
    }
    return NLMC_OK;
}

}  // namespace nlmc

extern "C" {
int nlmc_np_tanh(const double *x, double *out, int64_t n, int device) { return nlmc::npmath_run(0, x, out, n, device); }
int nlmc_np_arctanh(const double *x, double *out, int64_t n, int device) { return nlmc::npmath_run(1, x, out, n, device); }
}
