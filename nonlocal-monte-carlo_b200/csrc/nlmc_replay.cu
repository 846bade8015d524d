// nlmc_replay.cu -- K1 sweep_replay (exact-replay heat-bath sweeps) and K4 energy_csr.
//
// K1 replaces the reference's MCMC inner loop (NMC/nmc.py:62-89 and its three copies) in the mode
// where the random stream is injected: the visiting order of every sweep and the uniform of every
// attempt are the reference's own np.random draws, so the spin trajectory is reproduced bit for
// bit.  An exact replay is sequential inside a replica (attempt a sees the result of attempt
// a-1), so the parallelism is: one warp per replica, the 32 lanes over the stored entries of the
// visited row, and a 32-attempt look-ahead in which every lane fetches the row extent, field and
// scale flag of one upcoming attempt (those do not depend on the spins).  Spins live in shared
// memory when the replica fits (n <= kSmemSpinLimit), otherwise in global memory.
//
// Row sums follow scipy's csr_matvec, which the reference calls through J.dot(m) (nmc.py:86):
// sequential accumulation over the stored entries of the row, starting from 0, then "+ h[k]".
// When every J value is an integer and the row is not temperature-scaled the partial sums are
// exact integers and the lanes reduce them in parallel; otherwise the products are formed in
// parallel and accumulated in storage order by shuffles, which keeps the rounding identical.
#include <cmath>
#include <cstdlib>

#include "nlmc_common.cuh"
#include "nlmc_npmath.h"

namespace nlmc {

constexpr int kSmemSpinLimit = 200 * 1024;

struct ReplayArgs {
    int n, n_sweeps, lut_half, record_from, integer_j;
    const int32_t *rp, *ci;
    const double *val, *h_inst;
    int8_t *spins;
    const double *h_eff;
    const uint8_t *row_scaled;
    const int32_t *flags;
    const double *temp_x;
    const int32_t *perm;
    const double *u, *beta, *lut;
    int8_t *out_M;
    double *out_E;
};

template <bool kSmem>
__device__ __forceinline__ int ld_spin(const int8_t *m, int i) {
    if (kSmem) return m[i];
    return *reinterpret_cast<const volatile int8_t *>(m + i);
}
template <bool kSmem>
__device__ __forceinline__ void st_spin(int8_t *m, int i, int v) {
    if (kSmem) m[i] = (int8_t)v;
    else *reinterpret_cast<volatile int8_t *>(m + i) = (int8_t)v;
}

// np.tanh(y) in fp64, bit for bit (nlmc_npmath.h restates numpy's float64 routine).  For |y| >= 24 it is exactly
// +-1, which is the case of every spin frozen by h = +-1e4 in the NMC phases (nmc.py:381,400).
// The 288-entry table of the routine is staged in shared memory by the replay kernels: one warp per chain is latency
// bound, and 18 dependent-address loads from global memory per attempt cost more than the whole rest of the step.
__device__ __forceinline__ double tanh_sat(double y, const uint64_t *lut) { return nlmc_np_tanh_lut(y, lut); }
__device__ __forceinline__ void stage_tanh_lut(uint64_t *dst) {
    for (int i = threadIdx.x; i < 288; i += blockDim.x) dst[i] = nlmc_npm_tanh_lut[i];
    __syncthreads();
}

// Storage-order accumulation of 32 per-lane products: the lanes park them in shared memory, then every lane
// reads all 32 back (loads issued together, not interleaved with the adds) and runs the same dependent add chain
// -- the only part that is inherently sequential for a bit-exact emulation of scipy's csr_matvec row sum.
__device__ __forceinline__ double ordered_sum32(double acc, double prod, int count, double *scratch, int lane) {
    scratch[lane] = prod;
    __syncwarp();
    double2 v[16];
#pragma unroll
    for (int l = 0; l < 16; ++l) v[l] = reinterpret_cast<const double2 *>(scratch)[l];
#pragma unroll
    for (int l = 0; l < 16; ++l) {
        if (2 * l < count) acc = __dadd_rn(acc, v[l].x);
        if (2 * l + 1 < count) acc = __dadd_rn(acc, v[l].y);
    }
    __syncwarp();
    return acc;
}

template <bool kSmem>
__global__ void __launch_bounds__(32) sweep_replay_kernel(ReplayArgs a) {
    __shared__ __align__(16) double scratch[32];
    __shared__ uint64_t s_tlut[288];
    extern __shared__ __align__(16) int8_t smem_spins[];
    stage_tanh_lut(s_tlut);
    const int r = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.n;
    int8_t *g_spins = a.spins + (size_t)r * n;
    int8_t *m = kSmem ? smem_spins : g_spins;
    if (kSmem) {
        for (int i = lane; i < n; i += 32) m[i] = g_spins[i];
    }
    __syncwarp();

    const int flags = a.flags[r];
    const double *h_eff = (flags & 1) ? a.h_eff + (size_t)r * n : a.h_inst;
    const uint8_t *scaled = (flags & 2) ? a.row_scaled + (size_t)r * n : nullptr;
    const double temp_x = a.temp_x[r];
    const int lut_w = 2 * a.lut_half + 1;
    const int n_rec = a.n_sweeps - a.record_from;

    for (int s = 0; s < a.n_sweeps; ++s) {
        const size_t rs = (size_t)r * a.n_sweeps + s;
        const double beta = a.beta[rs];
        const int32_t *perm = a.perm + rs * n;
        const double *uu = a.u + rs * n;
        const double *lut = a.lut ? a.lut + rs * lut_w : nullptr;

        for (int base = 0; base < n; base += 32) {
            // look-ahead: lane l prepares attempt base+l
            const int idx = base + lane;
            int my_k = 0, my_rb = 0, my_re = 0, my_sc = 0;
            double my_u = 0.0, my_h = 0.0;
            if (idx < n) {
                my_k = perm[idx];
                my_u = uu[idx];
                my_rb = __ldg(a.rp + my_k);
                my_re = __ldg(a.rp + my_k + 1);
                my_h = h_eff[my_k];
                my_sc = scaled ? scaled[my_k] : 0;
            }
            const int cnt = min(32, n - base);
            for (int j = 0; j < cnt; ++j) {
                const int k = __shfl_sync(0xffffffffu, my_k, j);
                const int rb = __shfl_sync(0xffffffffu, my_rb, j);
                const int re = __shfl_sync(0xffffffffu, my_re, j);
                const int sc = __shfl_sync(0xffffffffu, my_sc, j);
                const double u_a = __shfl_sync(0xffffffffu, my_u, j);
                const double h_k = __shfl_sync(0xffffffffu, my_h, j);
                const bool exact = a.integer_j && !sc;
                double rowsum;
                if (exact) {
                    double part = 0.0;
                    for (int p = rb + lane; p < re; p += 32)
                        part += __ldg(a.val + p) * (double)ld_spin<kSmem>(m, __ldg(a.ci + p));
                    rowsum = warp_sum(part);
                } else {
                    rowsum = 0.0;
                    for (int pb = rb; pb < re; pb += 32) {
                        const int p = pb + lane;
                        double prod = 0.0;
                        if (p < re) {
                            double v = __ldg(a.val + p);
                            if (sc) v = __ddiv_rn(v, temp_x);  // J_c[all_clusters,:] / temp_x  (nmc.py:379)
                            prod = __dmul_rn(v, (double)ld_spin<kSmem>(m, __ldg(a.ci + p)));
                        }
                        rowsum = ordered_sum32(rowsum, prod, min(32, re - pb), scratch, lane);
                    }
                }
                const double x = __dadd_rn(rowsum, h_k);
                double t;
                if (lut != nullptr && exact && h_k == 0.0 && fabs(rowsum) <= (double)a.lut_half)
                    t = lut[(int)rowsum + a.lut_half];
                else
                    t = tanh_sat(__dmul_rn(beta, x), s_tlut);
                // np.sign(np.tanh(beta*x) - 2*rand() + 1)   (nmc.py:87)
                const double v = __dadd_rn(__dsub_rn(t, __dmul_rn(2.0, u_a)), 1.0);
                const int nw = (v > 0.0) - (v < 0.0);
                if (lane == 0) st_spin<kSmem>(m, k, nw);
                __syncwarp();
            }
        }

        if (a.out_M != nullptr && s >= a.record_from) {
            int8_t *dst = a.out_M + ((size_t)r * n_rec + (s - a.record_from)) * n;
            for (int i = lane; i < n; i += 32) dst[i] = (int8_t)ld_spin<kSmem>(m, i);
        }
        if (a.out_E != nullptr) {
            double quad = 0.0, lin = 0.0;
            for (int k = lane; k < n; k += 32) {
                double xs = 0.0;
                const int re = __ldg(a.rp + k + 1);
                for (int p = __ldg(a.rp + k); p < re; ++p)
                    xs += __ldg(a.val + p) * (double)ld_spin<kSmem>(m, __ldg(a.ci + p));
                const double mk = (double)ld_spin<kSmem>(m, k);
                quad += mk * xs;
                lin += mk * __ldg(a.h_inst + k);
            }
            quad = warp_sum(quad);
            lin = warp_sum(lin);
            if (lane == 0) a.out_E[rs] = -(quad / 2.0 + lin);
        }
    }
    if (kSmem) {
        __syncwarp();
        for (int i = lane; i < n; i += 32) g_spins[i] = m[i];
    }
}

// ------------------------------------------------------------------------------------------------
// K1-int: the same exact replay for integer-valued J (+-J instances) when the instance fits in shared
// memory.  Spins, CSR (int16/int32 columns, int8 values), the tanh LUT of the sweep and INCREMENTALLY
// maintained integer local fields f_k = sum_j J_kj m_j live in shared memory: an attempt reads f_k, looks
// tanh(beta*f_k) up, decides, and only when the spin changes walks its row to update the neighbours' fields
// (lanes over the row entries).  Integer sums are exact, so the decisions are bit-identical to the general
// kernel's.  Rows that an NMC phase rescales (J/temp_x is not an integer) still take the sequential
// storage-order sum, and sites with an effective field h_eff != 0 (frozen by +-1e4, or a real h) use
// tanh(beta*(f + h_eff)) directly.  Energies per sweep come from the fields: E = -(sum m_k f_k)/2 - sum h_k m_k.
template <typename ColT>
__global__ void __launch_bounds__(32) sweep_replay_int_kernel(ReplayArgs a, const ColT *__restrict__ g_col,
                                                              const int8_t *__restrict__ g_val, int nnz) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    __shared__ __align__(16) double scratch[32];
    __shared__ double div_tab[256];  // fl(v / temp_x) for every int8 coupling value v: no fp64 division per entry
    __shared__ uint64_t s_tlut[288];
    stage_tanh_lut(s_tlut);
    const int n = a.n;
    const int lut_w = 2 * a.lut_half + 1;
    // carve: lut (double) | fields (int32) | row_ptr (int32) | col | val (int8) | spins (int8) | mode (uint8)
    double *lut_s = reinterpret_cast<double *>(sm_raw);
    int32_t *fld = reinterpret_cast<int32_t *>(lut_s + (a.lut ? lut_w : 0));
    int32_t *rp_s = fld + n;
    ColT *col_s = reinterpret_cast<ColT *>(rp_s + n + 1);
    int8_t *val_s = reinterpret_cast<int8_t *>(col_s + nnz + (nnz & 1));
    int8_t *m = val_s + nnz;
    uint8_t *mode = reinterpret_cast<uint8_t *>(m + n);
    const int r = blockIdx.x, lane = threadIdx.x;
    int8_t *g_spins = a.spins + (size_t)r * n;
    const int flags = a.flags[r];
    const double *h_eff = (flags & 1) ? a.h_eff + (size_t)r * n : a.h_inst;
    const uint8_t *scaled = (flags & 2) ? a.row_scaled + (size_t)r * n : nullptr;
    const double temp_x = a.temp_x[r];
    for (int i = lane; i < 256; i += 32) div_tab[i] = __ddiv_rn((double)(i - 128), temp_x);  // J_c = J / temp_x (nmc.py:379)
    for (int i = lane; i <= n; i += 32) rp_s[i] = a.rp[i];
    for (int i = lane; i < nnz; i += 32) { col_s[i] = g_col[i]; val_s[i] = g_val[i]; }
    for (int i = lane; i < n; i += 32) {
        m[i] = g_spins[i];
        mode[i] = (uint8_t)((h_eff[i] != 0.0 ? 1 : 0) | ((scaled && scaled[i]) ? 2 : 0));
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {  // initial fields
        int f = 0;
        for (int p = rp_s[k]; p < rp_s[k + 1]; ++p) f += (int)val_s[p] * (int)m[col_s[p]];
        fld[k] = f;
    }
    __syncwarp();
    const int n_rec = a.n_sweeps - a.record_from;
    for (int s = 0; s < a.n_sweeps; ++s) {
        const size_t rs = (size_t)r * a.n_sweeps + s;
        const double beta = a.beta[rs];
        const int32_t *perm = a.perm + rs * n;
        const double *uu = a.u + rs * n;
        if (a.lut) {
            for (int i = lane; i < lut_w; i += 32) lut_s[i] = a.lut[rs * lut_w + i];
            __syncwarp();
        }
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane;
            int my_k = 0;
            double my_u = 0.0;
            if (idx < n) { my_k = perm[idx]; my_u = uu[idx]; }
            const int cnt = min(32, n - base);
            for (int j = 0; j < cnt; ++j) {
                const int k = __shfl_sync(0xffffffffu, my_k, j);
                const double u_a = __shfl_sync(0xffffffffu, my_u, j);
                const int md = mode[k];
                const int f = fld[k];
                const int old = m[k];
                double t;
                if (md == 0 && a.lut && abs(f) <= a.lut_half) {
                    t = lut_s[f + a.lut_half];
                } else if (!(md & 2)) {
                    const double h_k = (md & 1) ? h_eff[k] : 0.0;
                    t = tanh_sat(__dmul_rn(beta, __dadd_rn((double)f, h_k)), s_tlut);
                } else {
                    // Rescaled row (NMC backbone at beta/temp_x).  The reference's field is the storage-order sum of
                    // fl(J/temp_x)*m; it differs from f/temp_x (f = the exact integer field) by a few ulps only, so the
                    // decision sign(t - 2u + 1) is screened with f/temp_x first and the dependent fp64 add chain over
                    // the row is walked only when u lies within 1e-9 of the threshold (where those ulps could matter).
                    const double xa = __dadd_rn(__ddiv_rn((double)f, temp_x), h_eff[k]);
                    t = tanh_sat(__dmul_rn(beta, xa), s_tlut);
                    if (fabs(__dadd_rn(__dsub_rn(t, __dmul_rn(2.0, u_a)), 1.0)) <= 1e-9) {
                        double x = 0.0;
                        const int re = rp_s[k + 1];
                        for (int pb = rp_s[k]; pb < re; pb += 32) {
                            const int p = pb + lane;
                            double prod = 0.0;
                            if (p < re) prod = __dmul_rn(div_tab[(int)val_s[p] + 128], (double)m[col_s[p]]);
                            x = ordered_sum32(x, prod, min(32, re - pb), scratch, lane);
                        }
                        t = tanh_sat(__dmul_rn(beta, __dadd_rn(x, h_eff[k])), s_tlut);
                    }
                }
                const double v = __dadd_rn(__dsub_rn(t, __dmul_rn(2.0, u_a)), 1.0);  // nmc.py:87
                const int nw = (v > 0.0) - (v < 0.0);
                if (nw != old) {  // warp-uniform: walk the row, lanes over its entries
                    const int dm = nw - old;
                    const int re = rp_s[k + 1];
                    for (int p = rp_s[k] + lane; p < re; p += 32) fld[col_s[p]] += (int)val_s[p] * dm;
                    if (lane == 0) m[k] = (int8_t)nw;
                    __syncwarp();
                }
            }
        }
        if (a.out_M != nullptr && s >= a.record_from) {
            int8_t *dst = a.out_M + ((size_t)r * n_rec + (s - a.record_from)) * n;
            for (int i = lane; i < n; i += 32) dst[i] = m[i];
        }
        if (a.out_E != nullptr) {
            double quad = 0.0, lin = 0.0;
            for (int k = lane; k < n; k += 32) {
                quad += (double)((int)m[k] * fld[k]);
                lin += (double)m[k] * __ldg(a.h_inst + k);
            }
            quad = warp_sum(quad);
            lin = warp_sum(lin);
            if (lane == 0) a.out_E[rs] = -(quad / 2.0 + lin);
        }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) g_spins[i] = m[i];
}

// K4: one CTA per state; E = -(m^T J m / 2 + m^T h).  Integer J and h give exact integer sums.
__global__ void __launch_bounds__(256) energy_kernel(int n, const int32_t *__restrict__ rp,
                                                      const int32_t *__restrict__ ci,
                                                      const double *__restrict__ val,
                                                      const double *__restrict__ h,
                                                      const int8_t *__restrict__ states, double *__restrict__ out) {
    __shared__ double s_q[8], s_l[8];
    const int8_t *m = states + (size_t)blockIdx.x * n;
    double quad = 0.0, lin = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double xs = 0.0;
        const int re = rp[k + 1];
        for (int p = rp[k]; p < re; ++p) xs += val[p] * (double)m[ci[p]];
        const double mk = (double)m[k];
        quad += mk * xs;
        lin += mk * h[k];
    }
    quad = warp_sum(quad);
    lin = warp_sum(lin);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_q[w] = quad;
        s_l[w] = lin;
    }
    __syncthreads();
    if (w == 0) {
        quad = lane < (int)(blockDim.x >> 5) ? s_q[lane] : 0.0;
        lin = lane < (int)(blockDim.x >> 5) ? s_l[lane] : 0.0;
        quad = warp_sum(quad);
        lin = warp_sum(lin);
        if (lane == 0) out[blockIdx.x] = -(quad / 2.0 + lin);
    }
}

}  // namespace nlmc

extern "C" {

int nlmc_sweep_replay(nlmc_replicas *P, int n_sweeps, const int32_t *perm, const double *u,
                      const double *beta, const double *tanh_lut, int lut_half,
                      int8_t *out_M, int record_from, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(P != nullptr, "nlmc_sweep_replay: replicas handle is NULL");
    NLMC_REQUIRE(n_sweeps >= 0, "nlmc_sweep_replay: negative sweep count");
    if (n_sweeps == 0) return NLMC_OK;
    NLMC_REQUIRE(perm && u && beta, "nlmc_sweep_replay: perm, u and beta are required");
    NLMC_REQUIRE(record_from >= 0 && record_from <= n_sweeps, "nlmc_sweep_replay: record_from out of range");
    NLMC_REQUIRE(!tanh_lut || lut_half >= 0, "nlmc_sweep_replay: negative lut_half");
    { const int rc_dev = nlmc::instance_device(P->inst); if (rc_dev) return rc_dev; }   // the CSR on the device (uploaded on first use)
    nlmc_instance *I = P->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    const size_t R = (size_t)P->R, n = (size_t)I->n, S = (size_t)n_sweeps;
    const size_t n_rec = S - (size_t)record_from;
    const size_t lut_w = tanh_lut ? (size_t)(2 * lut_half + 1) : 0;
    int rc;
    if ((rc = P->s_perm.reserve(sizeof(int32_t) * R * S * n)) || (rc = P->s_u.reserve(sizeof(double) * R * S * n)) ||
        (rc = P->s_beta.reserve(sizeof(double) * R * S)))
        return rc;
    if (tanh_lut && (rc = P->s_lut.reserve(sizeof(double) * R * S * lut_w))) return rc;
    if (out_M && n_rec && (rc = P->s_M.reserve(R * n_rec * n))) return rc;
    if (out_E && (rc = P->s_E.reserve(sizeof(double) * R * S))) return rc;
    cudaStream_t st = I->stream;
    NLMC_CUDA(cudaMemcpyAsync(P->s_perm.ptr, perm, sizeof(int32_t) * R * S * n, cudaMemcpyHostToDevice, st));
    NLMC_CUDA(cudaMemcpyAsync(P->s_u.ptr, u, sizeof(double) * R * S * n, cudaMemcpyHostToDevice, st));
    NLMC_CUDA(cudaMemcpyAsync(P->s_beta.ptr, beta, sizeof(double) * R * S, cudaMemcpyHostToDevice, st));
    if (tanh_lut)
        NLMC_CUDA(cudaMemcpyAsync(P->s_lut.ptr, tanh_lut, sizeof(double) * R * S * lut_w, cudaMemcpyHostToDevice, st));

    ReplayArgs a;
    a.n = I->n;
    a.n_sweeps = n_sweeps;
    a.lut_half = lut_half;
    a.record_from = record_from;
    a.integer_j = I->integer_j ? 1 : 0;
    a.rp = I->row_ptr;
    a.ci = I->col;
    a.val = I->val;
    a.h_inst = I->h;
    a.spins = P->spins;
    a.h_eff = P->h_eff;
    a.row_scaled = P->row_scaled;
    a.flags = P->flags;
    a.temp_x = P->temp_x;
    a.perm = P->s_perm.as<int32_t>();
    a.u = P->s_u.as<double>();
    a.beta = P->s_beta.as<double>();
    a.lut = tanh_lut ? P->s_lut.as<double>() : nullptr;
    a.out_M = (out_M && n_rec) ? P->s_M.as<int8_t>() : nullptr;
    a.out_E = out_E ? P->s_E.as<double>() : nullptr;

    // integer instances that fit in shared memory take the incremental-field kernel
    const size_t lut_bytes = tanh_lut ? sizeof(double) * lut_w : 0;
    const bool small_cols = I->n <= 65535;
    const size_t nnz_s = (size_t)I->nnz;
    const size_t int_smem = lut_bytes + sizeof(int32_t) * (2 * n + 1) + (small_cols ? 2 : 4) * (nnz_s + (nnz_s & 1)) + nnz_s + 2 * n + 32;
    if (I->integer_j && I->int_val && nlmc::instance_value_symmetric(I) && int_smem <= 217 * 1024 && !getenv("NLMC_REPLAY_GENERAL")) {
        if (small_cols) {
            NLMC_CUDA(cudaFuncSetAttribute(sweep_replay_int_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)int_smem));
            sweep_replay_int_kernel<uint16_t><<<(unsigned)R, 32, int_smem, st>>>(a, I->col16, I->int_val, I->nnz);
        } else {
            NLMC_CUDA(cudaFuncSetAttribute(sweep_replay_int_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)int_smem));
            sweep_replay_int_kernel<int32_t><<<(unsigned)R, 32, int_smem, st>>>(a, I->col, I->int_val, I->nnz);
        }
    } else if (I->n <= kSmemSpinLimit) {
        const size_t smem = (n + 15) & ~(size_t)15;
        NLMC_CUDA(cudaFuncSetAttribute(sweep_replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sweep_replay_kernel<true><<<(unsigned)R, 32, smem, st>>>(a);
    } else {
        sweep_replay_kernel<false><<<(unsigned)R, 32, 0, st>>>(a);
    }
    NLMC_CUDA(cudaGetLastError());
    if (a.out_M) NLMC_CUDA(cudaMemcpyAsync(out_M, P->s_M.ptr, R * n_rec * n, cudaMemcpyDeviceToHost, st));
    if (a.out_E) NLMC_CUDA(cudaMemcpyAsync(out_E, P->s_E.ptr, sizeof(double) * R * S, cudaMemcpyDeviceToHost, st));
    NLMC_CUDA(cudaStreamSynchronize(st));
    return NLMC_OK;
}

int nlmc_energy(nlmc_replicas *P, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(P && out_E, "nlmc_energy: NULL argument");
    nlmc_instance *I = P->inst;
    { const int rc_dev = nlmc::instance_device(I); if (rc_dev) return rc_dev; }   // the CSR on the device (uploaded on first use)
    NLMC_CUDA(cudaSetDevice(I->device));
    int rc;
    if ((rc = P->s_E.reserve(sizeof(double) * (size_t)P->R))) return rc;
    energy_kernel<<<(unsigned)P->R, 256, 0, I->stream>>>(I->n, I->row_ptr, I->col, I->val, I->h, P->spins,
                                                         P->s_E.as<double>());
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaMemcpyAsync(out_E, P->s_E.ptr, sizeof(double) * (size_t)P->R, cudaMemcpyDeviceToHost, I->stream));
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    return NLMC_OK;
}

int nlmc_energy_states(nlmc_instance *I, int n_states, const int8_t *states, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(I && n_states >= 0, "nlmc_energy_states: bad arguments");
    if (n_states == 0) return NLMC_OK;
    NLMC_REQUIRE(states && out_E, "nlmc_energy_states: NULL buffer");
    { const int rc_dev = nlmc::instance_device(I); if (rc_dev) return rc_dev; }   // the CSR on the device (uploaded on first use)
    NLMC_CUDA(cudaSetDevice(I->device));
    int8_t *d_states = nullptr;
    double *d_E = nullptr;
    const size_t bytes = (size_t)n_states * (size_t)I->n;
    NLMC_CUDA(cudaMalloc(&d_states, bytes));
    if (cudaMalloc(&d_E, sizeof(double) * (size_t)n_states) != cudaSuccess) {
        cudaFree(d_states);
        set_error("nlmc_energy_states: cudaMalloc failed");
        return NLMC_ERR_CUDA;
    }
    cudaError_t e = cudaMemcpyAsync(d_states, states, bytes, cudaMemcpyHostToDevice, I->stream);
    if (e == cudaSuccess) {
        energy_kernel<<<(unsigned)n_states, 256, 0, I->stream>>>(I->n, I->row_ptr, I->col, I->val, I->h, d_states, d_E);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_E, d_E, sizeof(double) * (size_t)n_states, cudaMemcpyDeviceToHost, I->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(I->stream);
    cudaFree(d_states);
    cudaFree(d_E);
    if (e != cudaSuccess) {
        set_error("nlmc_energy_states: %s", cudaGetErrorString(e));
        return NLMC_ERR_CUDA;
    }
    return NLMC_OK;
}

}  // extern "C"
