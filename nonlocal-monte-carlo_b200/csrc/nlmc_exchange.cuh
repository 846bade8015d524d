// nlmc_exchange.cuh -- K6 for the generic production engines (K2a sparse, K3 dense): replica exchange as a permutation
// of beta labels, entirely on the device (north_star 4).
//
// Replaces the swap block NPT/npt.py:649-680 (pair selection NPT/npt.py:514-533): the rows of an engine are grouped into
// ladders, row = ladder * n_beta + slot; every row carries the index of the temperature it is simulated at.  An accepted
// exchange of the adjacent temperatures (i, i+1) of a ladder swaps the two rows' labels and per-row betas -- no spin
// moves, nothing goes through the host.  One thread per ladder: num_pairs non-overlapping adjacent pairs drawn one after
// the other uniformly from the pairs still available, accepted with min(1, exp((b_next - b_sel) * (E_next - E_sel))).
#pragma once

#include "nlmc_common.cuh"

namespace nlmc {

constexpr int kXMaxBeta = 128;
constexpr int kXRoundLog = 4096;

struct LadderExchange {
    int n_beta = 0, n_ladders = 0;
    double *betas = nullptr;     // [n_beta]
    int32_t *label = nullptr;    // [R] temperature index of each row
    int32_t *slot_of = nullptr;  // [R] row (within its ladder) holding temperature index i: slot_of[ladder*n_beta + i]
    double *E = nullptr;         // [R] energies of the rows
    int32_t *accepted_rounds = nullptr;  // [kXRoundLog]
    uint32_t round = 0;
    bool active() const { return n_beta > 0; }
    void release() {
        void *ptrs[] = {betas, label, slot_of, E, accepted_rounds};
        for (void *p : ptrs) if (p) cudaFree(p);
        betas = nullptr; label = nullptr; slot_of = nullptr; E = nullptr; accepted_rounds = nullptr;
        n_beta = n_ladders = 0;
    }
};

struct PhiloxX {  // Philox4x32-10
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
            const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b;
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

template <typename BetaT>
__global__ void ladder_identity_kernel(int n_beta, int n_rows, const double *betas, int32_t *label, int32_t *slot_of,
                                       BetaT *beta_row) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int s = row % n_beta;
    label[row] = s;
    slot_of[row] = s;
    beta_row[row] = (BetaT)betas[s];
}

template <typename BetaT>
__global__ void ladder_label_swap_kernel(int n_beta, int n_ladders, int num_pairs, const double *__restrict__ betas,
                                         const double *__restrict__ E, int32_t *label, int32_t *slot_of, BetaT *beta_row,
                                         int32_t *accepted_rounds, uint32_t seed_lo, uint32_t seed_hi, uint32_t round,
                                         int ladder_offset) {
    const int ladder = blockIdx.x * blockDim.x + threadIdx.x;
    if (ladder >= n_ladders) return;
    const PhiloxX rng{seed_lo, seed_hi ^ 0x58434847u};
    const size_t base = (size_t)ladder * n_beta;
    uint8_t avail[kXMaxBeta];
    int n_avail = n_beta - 1;
    for (int i = 0; i < n_beta - 1; ++i) avail[i] = 1;
    int acc = 0;
    for (int k = 0; k < num_pairs && n_avail > 0; ++k) {
        const uint4 r = rng((uint32_t)(ladder + ladder_offset), round, (uint32_t)k, 0u);
        int pick = (int)(((unsigned long long)r.x * (unsigned)n_avail) >> 32);
        int i = 0;
        for (;; ++i)
            if (avail[i] && pick-- == 0) break;
        for (int j = max(0, i - 1); j <= min(n_beta - 2, i + 1); ++j)
            if (avail[j]) { avail[j] = 0; --n_avail; }
        const int sa = slot_of[base + i], sb = slot_of[base + i + 1];
        const double x = (betas[i + 1] - betas[i]) * (E[base + sb] - E[base + sa]);
        const double u = ((double)r.y * 4294967296.0 + (double)r.z + 0.5) * (1.0 / 18446744073709551616.0);
        if (u < fmin(1.0, exp(x))) {
            ++acc;
            label[base + sa] = i + 1; label[base + sb] = i;
            slot_of[base + i] = sb; slot_of[base + i + 1] = sa;
            beta_row[base + sa] = (BetaT)betas[i + 1];
            beta_row[base + sb] = (BetaT)betas[i];
        }
    }
    if (acc) atomicAdd(accepted_rounds + (round % kXRoundLog), acc);
}

// rows = n_ladders x n_beta; labels reset to the identity, per-row betas set
template <typename BetaT>
static int exchange_setup(LadderExchange &X, int R, int n_beta, const double *betas, BetaT *beta_row, cudaStream_t st) {
    NLMC_REQUIRE(betas && n_beta >= 1 && n_beta <= kXMaxBeta && R % n_beta == 0,
                 "ladders: n_beta must be in [1, %d] and divide the number of rows (%d)", kXMaxBeta, R);
    X.release();
    X.n_beta = n_beta;
    X.n_ladders = R / n_beta;
    X.round = 0;
    NLMC_CUDA(cudaMalloc(&X.betas, sizeof(double) * (size_t)n_beta));
    NLMC_CUDA(cudaMalloc(&X.label, sizeof(int32_t) * (size_t)R));
    NLMC_CUDA(cudaMalloc(&X.slot_of, sizeof(int32_t) * (size_t)R));
    NLMC_CUDA(cudaMalloc(&X.E, sizeof(double) * (size_t)R));
    NLMC_CUDA(cudaMalloc(&X.accepted_rounds, sizeof(int32_t) * kXRoundLog));
    NLMC_CUDA(cudaMemsetAsync(X.accepted_rounds, 0, sizeof(int32_t) * kXRoundLog, st));
    NLMC_CUDA(cudaMemcpyAsync(X.betas, betas, sizeof(double) * (size_t)n_beta, cudaMemcpyHostToDevice, st));
    ladder_identity_kernel<BetaT><<<(R + 255) / 256, 256, 0, st>>>(n_beta, R, X.betas, X.label, X.slot_of, beta_row);
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaStreamSynchronize(st));  // `betas` is the caller's buffer
    return NLMC_OK;
}

// X.E must hold the energies of the rows (queued on `st` before this call)
template <typename BetaT>
static int exchange_launch(LadderExchange &X, int num_pairs, BetaT *beta_row, unsigned long long seed, int ladder_offset,
                           cudaStream_t st) {
    NLMC_REQUIRE(X.active(), "exchange: declare the ladders first");
    NLMC_CUDA(cudaMemsetAsync(X.accepted_rounds + (X.round % kXRoundLog), 0, sizeof(int32_t), st));
    if (X.n_beta >= 2 && num_pairs > 0) {
        ladder_label_swap_kernel<BetaT><<<(X.n_ladders + 63) / 64, 64, 0, st>>>(
            X.n_beta, X.n_ladders, num_pairs, X.betas, X.E, X.label, X.slot_of, beta_row, X.accepted_rounds, (uint32_t)seed,
            (uint32_t)(seed >> 32), X.round, ladder_offset);
        NLMC_CUDA(cudaGetLastError());
    }
    ++X.round;
    return NLMC_OK;
}

static int exchange_fetch(LadderExchange &X, int R, int32_t *out_labels, int n_rounds, int32_t *out_counts, cudaStream_t st) {
    NLMC_REQUIRE(X.active(), "labels: declare the ladders first");
    NLMC_REQUIRE(n_rounds >= 0 && n_rounds <= kXRoundLog && (n_rounds == 0 || out_counts), "labels: bad n_rounds");
    std::vector<int32_t> log((size_t)kXRoundLog);
    if (out_labels) NLMC_CUDA(cudaMemcpyAsync(out_labels, X.label, sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, st));
    if (n_rounds) NLMC_CUDA(cudaMemcpyAsync(log.data(), X.accepted_rounds, sizeof(int32_t) * kXRoundLog, cudaMemcpyDeviceToHost, st));
    NLMC_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < n_rounds; ++k) {
        const long long r = (long long)X.round - n_rounds + k;
        out_counts[k] = r >= 0 ? log[(size_t)(r % kXRoundLog)] : 0;
    }
    return NLMC_OK;
}

}  // namespace nlmc
