// nlmc_api.cu -- handle management of the C ABI (instances, replicas) and error reporting.
#include <algorithm>
#include <cmath>
#include <thread>
#include <utility>

#include "nlmc_common.cuh"
#include "nlmc_hostpar.h"

namespace nlmc {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace nlmc

namespace nlmc {
// J_ij == J_ji with no repeated column in a row: what the incremental-field kernels (K1-int, K2a, K3) rely on when a flip
// of site k pushes J_kj into the field of j.  The general replay kernel does not need it (the reference accepts any J
// through J.dot(m), NMC/nmc.py:86).  Evaluated on first use from the host mirror of the CSR (31 ms at C5 size, which the
// bit-packed path -- it checks its own six neighbours per site -- should not pay on every NPT.run call).
bool instance_value_symmetric(nlmc_instance *I) {
    if (I->symmetry_known) return I->value_symmetric;
    const int n = I->n;
    const int32_t *row_ptr = I->h_row_ptr.data(), *col = I->h_col.data();
    const double *val = I->h_val.data();
    bool value_symmetric = true;
    {
        // rows sorted by column (what scipy delivers for the lattices and dense matrices of the configs): binary search in
        // place; otherwise sorted copies of the rows
        bool sorted_rows = true;
        for (int i = 0; i < n && sorted_rows; ++i)
            for (int p = row_ptr[i] + 1; p < row_ptr[i + 1]; ++p)
                if (col[p] <= col[p - 1]) { sorted_rows = false; break; }
        if (sorted_rows) {
            for (int i = 0; i < n && value_symmetric; ++i)
                for (int p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
                    if (val[p] == 0.0) continue;
                    const int j = col[p];
                    const int32_t *b = col + row_ptr[j], *e = col + row_ptr[j + 1];
                    const int32_t *it = std::lower_bound(b, e, (int32_t)i);
                    if (it == e || *it != i || val[it - col] != val[p]) { value_symmetric = false; break; }
                }
        } else {
            std::vector<std::vector<std::pair<int32_t, double>>> rows((size_t)n);
            for (int i = 0; i < n; ++i) {
                auto &r = rows[(size_t)i];
                for (int p = row_ptr[i]; p < row_ptr[i + 1]; ++p) r.emplace_back(col[p], val[p]);
                std::sort(r.begin(), r.end());
                for (size_t q = 1; q < r.size() && value_symmetric; ++q)
                    if (r[q].first == r[q - 1].first) value_symmetric = false;
            }
            for (int i = 0; i < n && value_symmetric; ++i)
                for (const auto &e : rows[(size_t)i]) {
                    if (e.second == 0.0) continue;
                    const auto &rj = rows[(size_t)e.first];
                    auto it = std::lower_bound(rj.begin(), rj.end(), std::make_pair((int32_t)i, -1e300));
                    if (it == rj.end() || it->first != i || it->second != e.second) { value_symmetric = false; break; }
                }
        }
    }
    I->value_symmetric = value_symmetric;
    I->symmetry_known = true;
    return value_symmetric;
}
}  // namespace nlmc

extern "C" {

const char *nlmc_last_error(void) { return nlmc::g_err; }

int nlmc_version(void) { return 100; }

int nlmc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int nlmc_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor,
                     uint64_t *free_bytes, uint64_t *total_bytes) {
    cudaDeviceProp p;
    NLMC_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && name_len > 0) {
        strncpy(name, p.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (free_bytes || total_bytes) {
        size_t f = 0, t = 0;
        NLMC_CUDA(cudaSetDevice(device));
        NLMC_CUDA(cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = f;
        if (total_bytes) *total_bytes = t;
    }
    return NLMC_OK;
}

int nlmc_instance_create(int n, const int32_t *row_ptr, const int32_t *col, const double *val,
                         const double *h, int device, nlmc_instance **out) {
    NLMC_REQUIRE(out != nullptr, "nlmc_instance_create: out is NULL");
    *out = nullptr;
    NLMC_REQUIRE(n > 0 && row_ptr && h, "nlmc_instance_create: n must be > 0 and row_ptr/h non-NULL");
    NLMC_REQUIRE(row_ptr[0] == 0, "nlmc_instance_create: row_ptr[0] must be 0");
    const int nnz = row_ptr[n];
    NLMC_REQUIRE(nnz >= 0 && (nnz == 0 || (col && val)), "nlmc_instance_create: col/val missing");
    int max_deg = 0;
    for (int i = 0; i < n; ++i) {
        NLMC_REQUIRE(row_ptr[i + 1] >= row_ptr[i], "nlmc_instance_create: row_ptr not monotone at row %d", i);
        max_deg = std::max(max_deg, row_ptr[i + 1] - row_ptr[i]);
    }
    // one pass over the entries on the host workers: range check, integer / small-integer flags, and the host mirrors
    const int parts = nnz >= (1 << 18) ? nlmc::host_threads_shared() : 1;
    std::vector<int> bad((size_t)parts, -1);
    std::vector<char> not_int((size_t)parts, 0), not_small((size_t)parts, 0);
    auto *I = new nlmc_instance();
    I->h_col.resize((size_t)nnz);
    I->h_val.resize((size_t)nnz);
    nlmc::parallel_for(parts, [&](int t, int np) {
        const int per = (nnz + np - 1) / np, lo = std::min(nnz, per * t), hi = std::min(nnz, lo + per);
        for (int p = lo; p < hi; ++p) {
            I->h_col[(size_t)p] = col[p];
            I->h_val[(size_t)p] = val[p];
            if (col[p] < 0 || col[p] >= n) { if (bad[(size_t)t] < 0) bad[(size_t)t] = p; continue; }
            const double v = val[p];
            if (v != std::floor(v) || std::fabs(v) > 1e6) not_int[(size_t)t] = 1;
            if (!(std::fabs(v) <= 127.0)) not_small[(size_t)t] = 1;
        }
    });
    bool integer_j = true, small_int = nnz > 0;
    for (int t = 0; t < parts; ++t) {
        if (bad[(size_t)t] >= 0) {
            const int b = bad[(size_t)t];
            delete I;
            NLMC_REQUIRE(false, "nlmc_instance_create: column index out of range at entry %d", b);
        }
        if (not_int[(size_t)t]) integer_j = false;
        if (not_small[(size_t)t]) small_int = false;
    }
    small_int = small_int && integer_j;
    if (cudaSetDevice(device) != cudaSuccess) {
        nlmc::set_error("nlmc_instance_create: cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
        delete I;
        return NLMC_ERR_CUDA;
    }
    I->device = device;
    I->n = n;
    I->nnz = nnz;
    I->max_deg = max_deg;
    I->integer_j = integer_j;
    I->small_int = small_int;
    I->h_row_ptr.assign(row_ptr, row_ptr + n + 1);
    I->h_h.assign(h, h + n);
    if (cudaStreamCreateWithFlags(&I->stream, cudaStreamNonBlocking) != cudaSuccess) {
        nlmc::set_error("nlmc_instance_create: stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        nlmc_instance_destroy(I);
        return NLMC_ERR_CUDA;
    }
    *out = I;
    return NLMC_OK;
}

}  // extern "C"

namespace nlmc {

// The CSR on the device, uploaded on first use from the host mirrors (thread-safe; the engines that never read it -- the
// bit-packed and the dense one -- do not pay for it).  Device arrays come from the stream-ordered pool; the copies are queued
// on the instance's stream and waited for once.
int instance_device(nlmc_instance *I) {
    NLMC_REQUIRE(I != nullptr, "instance_device: NULL instance");
    std::lock_guard<std::mutex> lock(I->device_mu);
    if (I->device_ready) return NLMC_OK;
    NLMC_CUDA(cudaSetDevice(I->device));
    const int n = I->n, nnz = I->nnz;
    const size_t nz = (size_t)std::max(nnz, 1);
    cudaStream_t st = I->stream;
    // the compact int8 / uint16 copies the shared-memory replay kernel reads (small integer J only)
    std::vector<int8_t> v8(I->small_int ? (size_t)nnz : 0);
    std::vector<uint16_t> c16(I->small_int && n <= 65535 ? (size_t)nnz : 0);
    if (I->small_int) {
        const int parts = nnz >= (1 << 18) ? host_threads_shared() : 1;
        parallel_for(parts, [&](int t, int np) {
            const int per = (nnz + np - 1) / np, lo = std::min(nnz, per * t), hi = std::min(nnz, lo + per);
            for (int p = lo; p < hi; ++p) {
                v8[(size_t)p] = (int8_t)I->h_val[(size_t)p];
                if (!c16.empty()) c16[(size_t)p] = (uint16_t)I->h_col[(size_t)p];
            }
        });
    }
    auto up = [&](void **dst, const void *src, size_t alloc_bytes, size_t copy_bytes) {
        if (pool_alloc(dst, alloc_bytes, I->device, st) != cudaSuccess) return false;
        return copy_bytes == 0 || cudaMemcpyAsync(*dst, src, copy_bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
    };
    bool ok = up(reinterpret_cast<void **>(&I->row_ptr), I->h_row_ptr.data(), sizeof(int32_t) * (size_t)(n + 1), sizeof(int32_t) * (size_t)(n + 1)) &&
              up(reinterpret_cast<void **>(&I->col), I->h_col.data(), sizeof(int32_t) * nz, sizeof(int32_t) * (size_t)nnz) &&
              up(reinterpret_cast<void **>(&I->val), I->h_val.data(), sizeof(double) * nz, sizeof(double) * (size_t)nnz) &&
              up(reinterpret_cast<void **>(&I->h), I->h_h.data(), sizeof(double) * (size_t)n, sizeof(double) * (size_t)n);
    if (ok && I->small_int) {
        ok = up(reinterpret_cast<void **>(&I->int_val), v8.data(), (size_t)nnz, (size_t)nnz);
        if (ok && !c16.empty())
            ok = up(reinterpret_cast<void **>(&I->col16), c16.data(), sizeof(uint16_t) * (size_t)nnz, sizeof(uint16_t) * (size_t)nnz);
    }
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) {   // v8 / c16 go away; other streams read the arrays afterwards
        set_error("instance_device: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return NLMC_ERR_CUDA;
    }
    I->device_ready = true;
    return NLMC_OK;
}

}  // namespace nlmc

extern "C" {

/* Host-side format helper of the boundary: widen int8 spins to the float64 arrays the reference's API returns
 * (M is float64 +-1, NMC/nmc.py:52,89), on `threads` host threads (0 = hardware concurrency, at most 32). */
int nlmc_host_widen_i8_f64(const int8_t *in, double *out, uint64_t count, int threads) {
    NLMC_REQUIRE(count == 0 || (in && out), "nlmc_host_widen_i8_f64: NULL argument");
    nlmc::widen_i8_f64(in, out, count, threads);
    return NLMC_OK;
}

int nlmc_host_fetch_widen_blocks(const int8_t *dev_src, double *host_dst, int n_blocks, uint64_t block_elems,
                                 const int32_t *dst_block, int device, void *cuda_stream) {
    NLMC_REQUIRE(n_blocks >= 0 && (n_blocks == 0 || block_elems == 0 || (dev_src && host_dst)),
                 "nlmc_host_fetch_widen_blocks: NULL argument");
    if (n_blocks == 0 || block_elems == 0) return NLMC_OK;
    static thread_local int8_t *stage = nullptr;   // pinned, grow-only, one per host thread
    static thread_local size_t stage_cap = 0;
    constexpr int kMaxChunks = 8;
    static thread_local cudaEvent_t ev[kMaxChunks] = {};
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t bytes = (size_t)n_blocks * block_elems;
    NLMC_CUDA(cudaSetDevice(device));
    if (bytes > stage_cap) {
        if (stage) cudaFreeHost(stage);
        stage = nullptr; stage_cap = 0;
        NLMC_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&stage), bytes, cudaHostAllocDefault));
        stage_cap = bytes;
    }
    for (int k = 0; k < kMaxChunks; ++k)
        if (!ev[k]) NLMC_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    // chunks of whole blocks (a single block is cut into pieces of its own)
    if (n_blocks == 1) {
        const size_t per = ((bytes + kMaxChunks - 1) / kMaxChunks + 4095) & ~(size_t)4095;
        double *dst = host_dst + (dst_block ? (size_t)dst_block[0] * block_elems : 0);
        for (int k = 0; k < kMaxChunks; ++k) {
            const size_t lo = std::min(bytes, per * (size_t)k), hi = std::min(bytes, lo + per);
            if (lo < hi) NLMC_CUDA(cudaMemcpyAsync(stage + lo, dev_src + lo, hi - lo, cudaMemcpyDeviceToHost, st));
            NLMC_CUDA(cudaEventRecord(ev[k], st));
        }
        for (int k = 0; k < kMaxChunks; ++k) {
            const size_t lo = std::min(bytes, per * (size_t)k), hi = std::min(bytes, lo + per);
            NLMC_CUDA(cudaEventSynchronize(ev[k]));
            if (lo < hi) nlmc::widen_i8_f64(stage + lo, dst + lo, hi - lo, 0);
        }
        return NLMC_OK;
    }
    const int per_blocks = (n_blocks + kMaxChunks - 1) / kMaxChunks;
    for (int k = 0; k < kMaxChunks; ++k) {
        const int b0 = std::min(n_blocks, per_blocks * k), b1 = std::min(n_blocks, b0 + per_blocks);
        if (b0 < b1)
            NLMC_CUDA(cudaMemcpyAsync(stage + (size_t)b0 * block_elems, dev_src + (size_t)b0 * block_elems,
                                      (size_t)(b1 - b0) * block_elems, cudaMemcpyDeviceToHost, st));
        NLMC_CUDA(cudaEventRecord(ev[k], st));
    }
    for (int k = 0; k < kMaxChunks; ++k) {
        const int b0 = std::min(n_blocks, per_blocks * k), b1 = std::min(n_blocks, b0 + per_blocks);
        NLMC_CUDA(cudaEventSynchronize(ev[k]));
        for (int b = b0; b < b1; ++b)
            nlmc::widen_i8_f64(stage + (size_t)b * block_elems, host_dst + (size_t)(dst_block ? dst_block[b] : b) * block_elems,
                               block_elems, 0);
    }
    return NLMC_OK;
}

/* Touch every page of a freshly allocated host buffer on `threads` host threads (0 = all), so that the first-touch page
 * faults of a large result array (the 1 GB float64 M of config C5) are taken while the GPU is still sweeping instead of
 * inside the final widening. */
int nlmc_host_prefault(void *buf, uint64_t bytes, int threads) {
    NLMC_REQUIRE(bytes == 0 || buf, "nlmc_host_prefault: NULL argument");
    nlmc::prefault(buf, bytes, threads);
    return NLMC_OK;
}

int nlmc_instance_destroy(nlmc_instance *I) {
    if (!I) return NLMC_OK;
    cudaSetDevice(I->device);
    cudaDeviceSynchronize();  // handles built on this instance read its arrays from their own streams
    void *ptrs[] = {I->row_ptr, I->col, I->val, I->h, I->int_val, I->col16};
    if (I->stream) {  // back to the pool in stream order
        for (void *p : ptrs) nlmc::pool_free(p, I->stream);
    } else {
        for (void *p : ptrs) if (p) cudaFree(p);
    }
    if (I->rev) cudaFree(I->rev);
    if (I->stream) cudaStreamDestroy(I->stream);
    delete I;
    return NLMC_OK;
}

int nlmc_instance_n(const nlmc_instance *I) { return I ? I->n : NLMC_ERR_ARG; }
int nlmc_instance_is_integer(const nlmc_instance *I) { return I ? (I->integer_j ? 1 : 0) : NLMC_ERR_ARG; }
int nlmc_instance_is_symmetric(const nlmc_instance *I) {
    return I ? (nlmc::instance_value_symmetric(const_cast<nlmc_instance *>(I)) ? 1 : 0) : NLMC_ERR_ARG;
}

int nlmc_replicas_create(nlmc_instance *I, int R, const int8_t *init_spins, nlmc_replicas **out) {
    NLMC_REQUIRE(out != nullptr, "nlmc_replicas_create: out is NULL");
    *out = nullptr;
    NLMC_REQUIRE(I && R > 0, "nlmc_replicas_create: instance NULL or n_replicas <= 0");
    NLMC_CUDA(cudaSetDevice(I->device));
    auto *P = new nlmc_replicas();
    P->inst = I;
    P->R = R;
    P->h_flags.assign((size_t)R, 0);
    P->h_temp_x.assign((size_t)R, 1.0);
    const size_t bytes = (size_t)R * (size_t)I->n;
    if (cudaMalloc(&P->spins, bytes) != cudaSuccess || cudaMalloc(&P->flags, sizeof(int32_t) * (size_t)R) != cudaSuccess ||
        cudaMalloc(&P->temp_x, sizeof(double) * (size_t)R) != cudaSuccess ||
        cudaMemset(P->flags, 0, sizeof(int32_t) * (size_t)R) != cudaSuccess ||
        cudaMemcpy(P->temp_x, P->h_temp_x.data(), sizeof(double) * (size_t)R, cudaMemcpyHostToDevice) != cudaSuccess ||
        (init_spins ? cudaMemcpy(P->spins, init_spins, bytes, cudaMemcpyHostToDevice)
                    : cudaMemset(P->spins, 1, bytes)) != cudaSuccess) {
        nlmc::set_error("nlmc_replicas_create: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        nlmc_replicas_destroy(P);
        return NLMC_ERR_CUDA;
    }
    *out = P;
    return NLMC_OK;
}

int nlmc_replicas_destroy(nlmc_replicas *P) {
    if (!P) return NLMC_OK;
    cudaSetDevice(P->inst->device);
    if (P->spins) cudaFree(P->spins);
    if (P->h_eff) cudaFree(P->h_eff);
    if (P->row_scaled) cudaFree(P->row_scaled);
    if (P->flags) cudaFree(P->flags);
    if (P->temp_x) cudaFree(P->temp_x);
    P->s_perm.release();
    P->s_u.release();
    P->s_beta.release();
    P->s_lut.release();
    P->s_M.release();
    P->s_E.release();
    delete P;
    return NLMC_OK;
}

int nlmc_set_spins(nlmc_replicas *P, int first, int count, const int8_t *spins) {
    NLMC_REQUIRE(P && spins && first >= 0 && count >= 0 && first + count <= P->R, "nlmc_set_spins: bad range");
    NLMC_CUDA(cudaSetDevice(P->inst->device));
    const size_t n = (size_t)P->inst->n;
    NLMC_CUDA(cudaMemcpy(P->spins + (size_t)first * n, spins, (size_t)count * n, cudaMemcpyHostToDevice));
    return NLMC_OK;
}

int nlmc_get_spins(nlmc_replicas *P, int first, int count, int8_t *out) {
    NLMC_REQUIRE(P && out && first >= 0 && count >= 0 && first + count <= P->R, "nlmc_get_spins: bad range");
    NLMC_CUDA(cudaSetDevice(P->inst->device));
    const size_t n = (size_t)P->inst->n;
    NLMC_CUDA(cudaStreamSynchronize(P->inst->stream));
    NLMC_CUDA(cudaMemcpy(out, P->spins + (size_t)first * n, (size_t)count * n, cudaMemcpyDeviceToHost));
    return NLMC_OK;
}

int nlmc_set_phase(nlmc_replicas *P, int r, const double *h_eff, const uint8_t *row_scaled, double temp_x) {
    NLMC_REQUIRE(P && r >= 0 && r < P->R, "nlmc_set_phase: replica index out of range");
    NLMC_REQUIRE(!row_scaled || temp_x != 0.0, "nlmc_set_phase: temp_x must be non-zero when rows are scaled");
    NLMC_CUDA(cudaSetDevice(P->inst->device));
    const size_t n = (size_t)P->inst->n;
    int flags = 0;
    if (h_eff) {
        if (!P->h_eff) NLMC_CUDA(cudaMalloc(&P->h_eff, sizeof(double) * n * (size_t)P->R));
        NLMC_CUDA(cudaMemcpy(P->h_eff + (size_t)r * n, h_eff, sizeof(double) * n, cudaMemcpyHostToDevice));
        flags |= 1;
    }
    if (row_scaled) {
        if (!P->row_scaled) NLMC_CUDA(cudaMalloc(&P->row_scaled, n * (size_t)P->R));
        NLMC_CUDA(cudaMemcpy(P->row_scaled + (size_t)r * n, row_scaled, n, cudaMemcpyHostToDevice));
        flags |= 2;
    }
    P->h_flags[(size_t)r] = flags;
    P->h_temp_x[(size_t)r] = temp_x;
    NLMC_CUDA(cudaMemcpy(P->flags + r, &flags, sizeof(int32_t), cudaMemcpyHostToDevice));
    NLMC_CUDA(cudaMemcpy(P->temp_x + r, &temp_x, sizeof(double), cudaMemcpyHostToDevice));
    return NLMC_OK;
}

}  // extern "C"
