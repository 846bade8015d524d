// nlmc_common.cuh -- shared host/device helpers and the handle layouts behind include/nlmc_b200.h.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "nlmc_b200.h"

namespace nlmc {

void set_error(const char *fmt, ...);

#define NLMC_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            ::nlmc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return NLMC_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

#define NLMC_REQUIRE(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            ::nlmc::set_error(__VA_ARGS__);    \
            return NLMC_ERR_ARG;               \
        }                                      \
    } while (0)

// Device scratch buffer that only grows (staging for host arrays passed through the C ABI).
struct Scratch {
    void *ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return NLMC_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        NLMC_CUDA(cudaMalloc(&ptr, bytes));
        cap = bytes;
        return NLMC_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return static_cast<T *>(ptr); }
};

// Large device buffers come from the stream-ordered allocator with the pool kept (no release back to the driver), so that
// creating and destroying handles call after call -- every NPT.run builds its own -- does not pay cudaMalloc / cudaFree
// (65 ms per 134 MB state at C5 size) each time.
inline cudaError_t pool_alloc(void **ptr, size_t bytes, int device, cudaStream_t st) {
    static bool configured[64] = {};
    if (device >= 0 && device < 64 && !configured[device]) {
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured[device] = true;
    }
    return cudaMallocAsync(ptr, bytes, st);
}
inline void pool_free(void *ptr, cudaStream_t st) {
    if (ptr) cudaFreeAsync(ptr, st);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace nlmc

// ------------------------------------------------------------------------------------------------
// handle layouts (opaque to callers)
// ------------------------------------------------------------------------------------------------
struct nlmc_instance {
    int device = 0;
    int n = 0;
    int nnz = 0;
    int max_deg = 0;
    bool integer_j = false;   // every stored value is an integer -> row sums exact in any order
    bool symmetric = false;   // pattern symmetric (rev index available)
    bool value_symmetric = false;  // J_ij == J_ji for every stored entry and no column appears twice in a row
    bool symmetry_known = false;   // ... evaluated lazily by nlmc::instance_value_symmetric()
    int32_t *row_ptr = nullptr;  // [n+1]
    int32_t *col = nullptr;      // [nnz]
    double *val = nullptr;       // [nnz]
    double *h = nullptr;         // [n]
    int32_t *rev = nullptr;      // [nnz] index of the transposed entry (built on demand for LBP)
    int8_t *int_val = nullptr;   // [nnz] J as int8 when every value is an integer in [-127,127] (K1-int), else NULL
    uint16_t *col16 = nullptr;   // [nnz] 16-bit column indices when n <= 65535
    cudaStream_t stream = nullptr;
    // host mirrors (colouring, validation, MSC packing)
    std::vector<int32_t> h_row_ptr, h_col;
    std::vector<double> h_val, h_h;
    // The device arrays above are uploaded on first use (nlmc::instance_device): the bit-packed and the dense engine build
    // their own layouts from the host mirrors and never read the CSR on the device (23 MB of pageable copies per NPT.run at
    // the size of config C5).
    bool device_ready = false;
    bool small_int = false;        // every value is an integer in [-127, 127]: int_val (and col16) exist on the device
    std::mutex device_mu;
};

namespace nlmc {
bool instance_value_symmetric(nlmc_instance *I);
int instance_device(nlmc_instance *I);   // NLMC_OK once row_ptr / col / val / h (and int_val / col16) are on the device
}

struct nlmc_replicas {
    nlmc_instance *inst = nullptr;
    int R = 0;
    int8_t *spins = nullptr;        // [R][n]
    double *h_eff = nullptr;        // [R][n], lazily allocated by nlmc_set_phase
    uint8_t *row_scaled = nullptr;  // [R][n], lazily allocated by nlmc_set_phase
    int32_t *flags = nullptr;       // [R] bit0: h_eff row valid, bit1: row_scaled row valid
    double *temp_x = nullptr;       // [R]
    std::vector<int32_t> h_flags;
    std::vector<double> h_temp_x;
    nlmc::Scratch s_perm, s_u, s_beta, s_lut, s_M, s_E;
};
