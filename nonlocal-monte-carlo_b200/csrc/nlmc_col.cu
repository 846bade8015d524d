// nlmc_col.cu -- K2a: production heat-bath sweeps for ARBITRARY sparse instances (any J values, any h) with
// graph-coloured parallel updates, one CTA per replica.
//
// Replaces MCMC (NMC/nmc.py:28-91 and copies) in production mode where the bit-packed path does not apply
// (non-lattice graphs such as config C1's random graph, real-valued couplings, fields, NMC phases).  Sites are
// greedily coloured on the host; sites of one colour have no coupling between them, so they are updated in
// parallel (the conditional distribution of each is untouched by the others) and the colours are visited in
// order -- a valid Gibbs sweep with the same single-site rule as the reference:
//     s_i <- +1 with probability 1/(1 + exp(-2 beta f_i)),  f_i = sum_j J_ij s_j + h_i   (== nmc.py:86-87)
// Per replica everything lives in shared memory: spins (int8), the local fields f_i, the NMC phase modes, and the
// CSR itself when it fits (uint16/int32 columns, int32 values).  Fields are INT32 FIXED POINT (scale 2^s with s
// chosen so that max_i(|h_i| + sum_j |J_ij|) * 2^s < 2^30) and are maintained INCREMENTALLY: a flip adds
// 2 J_ij s_i to its neighbours' fields with ATOMS.ADD -- the only shared-memory atomic add the hardware has
// natively (fp64, u64 and fp32 adds compile to CAS spin loops).  Checked in the SASS of this file: every shared-state
// instantiation (kGlobalState = false) has `ATOMS.ADD RZ`, the global-workspace ones `REDG.E.ADD.STRONG.GPU`; in round 1
// the state pointer was a run-time choice, which made it generic and turned every field update into a global atomic.  Integer arithmetic means no
// drift and, for integer J, exact fields and energies.  Each site is served by a group of lanes (a power of two,
// chosen so that one colour fills the CTA): every lane of the group takes the same decision and the lanes split
// the neighbour list of a flip.  A whole batch of sweeps is one launch: per-sweep energies
// E = -1/2 sum_i s_i (f_i + h_i), optional recording of every k-th state (the reference's M[:, ::M_skip]) and
// tracking of the lowest-energy state (m_init = M[:, argmin E], nmc.py:394-395) all happen inside the kernel.
// Random numbers: Philox4x32-10 keyed by (seed; global replica id, site, sweep).
#include <algorithm>
#include <cstdlib>
#include <cmath>

#include "nlmc_common.cuh"
#include "nlmc_exchange.cuh"

struct nlmc_col {
    nlmc_instance *inst = nullptr;
    int n = 0, R = 0, n_colours = 0, replica_offset = 0;
    bool csr_in_smem = false, small_cols = false;
    int32_t *site_order = nullptr;  // [n] sites sorted by colour
    int32_t *colour_ptr = nullptr;  // [n_colours+1]
    uint16_t *col16 = nullptr;      // [nnz] (when n <= 65535)
    int32_t *valfx = nullptr;       // [nnz] J in fixed point (scale 2^fx_shift)
    int8_t *val8 = nullptr;         // [nnz] J itself when every value is an integer in [-127,127] (shifted at use)
    bool int8_vals = false;
    int fx_shift = 0, group = 1;    // fixed-point shift; lanes per site
    int8_t *spins = nullptr;        // [R][n]
    double *beta = nullptr;         // [R]
    uint8_t *modes = nullptr;       // [R][n] 0 normal, 1 hot (beta/temp_x), 2 frozen
    bool modes_on = false;
    double temp_x = 1.0;
    double *bestE = nullptr;        // [R]
    int8_t *bestS = nullptr;        // [R][n]
    uint32_t sweep_counter = 0;
    unsigned long long seed = 0;
    size_t smem_bytes = 0, smem_bytes_nocsr = 0;
    int sm_count = 148;
    int threads = 256;  // CTA size of the sweep kernel (kColThreads or kColThreadsFew)
    bool state_global = false;   // n too large for shared memory: fields / spins / modes of each replica live in global memory
    uint8_t *state_ws = nullptr; // [R][state_stride] workspace of that variant
    size_t state_stride = 0;
    nlmc::LadderExchange xch;    // replica exchange by beta labels (nlmc_col_ladders / nlmc_col_exchange)
    cudaStream_t stream = nullptr;
};

namespace nlmc {

constexpr int kColThreads = 256;       // threads per CTA when many replicas share the GPU
constexpr int kColThreadsFew = 1024;   // ... and when replicas are few: more sites of a colour per pass, shorter sweeps

struct PhiloxC {
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

struct ColArgs {
    int n, nnz, n_colours, n_sweeps, record_every, replica_offset, group;
    uint8_t *state_ws;      // NULL: state in shared memory
    size_t state_stride;
    double scale;  // 2^fx_shift
    const int32_t *rp, *ci;
    const double *val, *h;
    const uint16_t *col16;
    const int32_t *valfx;
    const int8_t *val8;
    int shift;
    const int32_t *site_order, *colour_ptr;
    int8_t *spins;
    const double *beta;
    const double *beta_sched;  // optional [n_sweeps][R]: annealing (beta_run of nmc.py:56-69)
    const uint8_t *modes;
    double temp_x;
    uint32_t seed_lo, seed_hi, sweep0;
    int8_t *out_spins;   // [n_rec][R][n] or null
    double *out_E;       // [n_sweeps][R] or null
    double *bestE;       // [R] or null
    int8_t *bestS;       // [R][n]
    int R;
};

// kGlobalState = false: the replica's state (fields, spins, modes, optionally the CSR) lives in SHARED memory and the
// field updates compile to native shared atomics (ATOMS.ADD); true: instances too large for shared memory keep it in a
// slice of a global workspace -- same code, the atomics go to L2 (ATOM.E.ADD) and reads bypass the incoherent L1.
template <bool kSmemCsr, typename ColT, typename ValT, bool kGlobalState>
__global__ void __launch_bounds__(kColThreadsFew) col_sweep_kernel(ColArgs a) {
    constexpr bool kVal8 = sizeof(ValT) == 1;  // integer couplings stored as int8 and shifted into fixed point at use
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ long long red[kColThreadsFew / 32];
    __shared__ double s_E;
    const int n = a.n, tid = threadIdx.x, r = blockIdx.x, nthr = (int)blockDim.x;
    uint8_t *state = kGlobalState ? a.state_ws + (size_t)blockIdx.x * a.state_stride : sm;
    int32_t *fld = reinterpret_cast<int32_t *>(state);                 // [n] fixed-point local fields (incl. h)
    int32_t *hfx = fld + n;                                            // [n] fixed-point h
    int32_t *rp_s = hfx + n;
    ColT *col_s = reinterpret_cast<ColT *>(rp_s + (kSmemCsr ? n + 1 : 0));
    ValT *val_s = reinterpret_cast<ValT *>(col_s + (kSmemCsr ? a.nnz + (a.nnz & 1) : 0));
    int8_t *spin = reinterpret_cast<int8_t *>(val_s + (kSmemCsr ? a.nnz + ((4 - (a.nnz & 3)) & 3) : 0));
    uint8_t *mode = reinterpret_cast<uint8_t *>(spin + n);

    int8_t *g_spin = a.spins + (size_t)r * n;
    const uint8_t *g_mode = a.modes ? a.modes + (size_t)r * n : nullptr;
    if (kSmemCsr) {
        for (int i = tid; i <= n; i += nthr) rp_s[i] = a.rp[i];
        for (int p = tid; p < a.nnz; p += nthr) {
            val_s[p] = kVal8 ? (ValT)a.val8[p] : (ValT)a.valfx[p];
            col_s[p] = sizeof(ColT) == 2 ? (ColT)a.col16[p] : (ColT)a.ci[p];
        }
    }
    for (int i = tid; i < n; i += nthr) {
        spin[i] = g_spin[i];
        mode[i] = g_mode ? g_mode[i] : 0;
        hfx[i] = __double2int_rn(a.h[i] * a.scale);
    }
    __syncthreads();
    // in the global-workspace variant other threads' atomics and stores land in L2: read around the (incoherent) L1
    constexpr bool gstate = kGlobalState;
    auto ld_fld = [&](int i) -> int { return gstate ? __ldcg(fld + i) : fld[i]; };
    auto ld_spin = [&](int i) -> int { return gstate ? (int)__ldcg(reinterpret_cast<const signed char *>(spin) + i) : (int)spin[i]; };
    auto row_begin = [&](int i) { return kSmemCsr ? rp_s[i] : a.rp[i]; };
    auto col_of = [&](int p) -> int { return kSmemCsr ? (int)col_s[p] : a.ci[p]; };
    auto val_of = [&](int p) -> int {
        if (kVal8) return (int)(kSmemCsr ? (int)val_s[p] : (int)a.val8[p]) << a.shift;
        return kSmemCsr ? (int)val_s[p] : a.valfx[p];
    };
    for (int i = tid; i < n; i += nthr) {  // initial fields, exact integer arithmetic
        int f = hfx[i];
        const int e = row_begin(i + 1);
        for (int p = row_begin(i); p < e; ++p) f += val_of(p) * ld_spin(col_of(p));
        fld[i] = f;
    }
    __syncthreads();

    const float inv_scale = (float)(1.0 / a.scale);
    float m2b = -2.0f * (float)a.beta[r];
    const float inv_tx = (float)(1.0 / a.temp_x);
    const PhiloxC rng{a.seed_lo, a.seed_hi ^ 0x434f4c52u};
    const uint32_t rid = (uint32_t)(a.replica_offset + r);
    const int G = a.group, gl = tid & (G - 1), sites_per_pass = nthr / G;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((tid & 31) & ~(G - 1));
    double best = a.bestE ? a.bestE[r] : 0.0;
    int n_rec = 0;
    for (int s = 0; s < a.n_sweeps; ++s) {
        const uint32_t sweep = a.sweep0 + (uint32_t)s;
        if (a.beta_sched) m2b = -2.0f * (float)a.beta_sched[(size_t)s * a.R + r];
        for (int c = 0; c < a.n_colours; ++c) {
            const int cb = a.colour_ptr[c], ce = a.colour_ptr[c + 1];
            for (int idx = cb + tid / G; idx < ce; idx += sites_per_pass) {
                const int i = a.site_order[idx];
                const int md = mode[i];
                if (md == 2) continue;  // frozen (the reference pins these spins with h = +-1e4, nmc.py:381,400)
                // backbone rows of J and h are divided by temp_x (nmc.py:379-380) <=> beta/temp_x for this site
                const float x = (md == 1 ? m2b * inv_tx : m2b) * ((float)ld_fld(i) * inv_scale);
                const uint4 rnd = rng(rid, (uint32_t)i, sweep, 0u);
                // P(+1) = 1/(1+exp(x)), x = -2 beta f.  The LESS likely state has probability q = 1/(1+exp(|x|)) <= 1/2; it is
                // drawn by comparing a 32-bit uniform with the integer threshold q 2^32 (relative accuracy of q kept down
                // to 2^-32, both tails alike; a 24-bit float uniform rounded to 1.0 forced the spin down once in 2^24 draws)
                const float q = __frcp_rn(1.0f + __expf(fabsf(x)));
                const uint32_t thr = (uint32_t)fminf(q * 4294967296.0f, 4294967040.0f);
                const bool minority = rnd.x < thr;
                const int s_new = (minority != (x < 0.0f)) ? 1 : -1;        // majority state: +1 when x < 0 (field up)
                const int s_old = ld_spin(i);
                __syncwarp(gmask);  // every lane of the group has read spin[i] and fld[i]
                if (s_new != s_old) {
                    if (gl == 0) spin[i] = (int8_t)s_new;
                    const int d = s_new - s_old;
                    const int e = row_begin(i + 1);
                    for (int p = row_begin(i) + gl; p < e; p += G) atomicAdd(&fld[col_of(p)], val_of(p) * d);
                }
            }
            __syncthreads();
        }
        const bool want_E = a.out_E != nullptr || a.bestE != nullptr;
        if (want_E) {  // E = -(m^T J m / 2 + m^T h) = -1/2 sum_i s_i (f_i + h_i), exact in fixed point
            long long part = 0;
            for (int i = tid; i < n; i += nthr) part += (long long)ld_spin(i) * ((long long)ld_fld(i) + (long long)hfx[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if ((tid & 31) == 0) red[tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                long long v = 0;
                for (int w = 0; w < nthr / 32; ++w) v += red[w];
                s_E = -0.5 * (double)v / a.scale;
            }
            __syncthreads();
            const double E = s_E;
            if (a.out_E && tid == 0) a.out_E[(size_t)s * a.R + r] = E;
            if (a.bestE && E < best) {  // strict improvement: the first minimum wins, like np.argmin
                best = E;
                int8_t *dst = a.bestS + (size_t)r * n;
                for (int i = tid; i < n; i += nthr) dst[i] = (int8_t)ld_spin(i);
            }
        }
        if (a.out_spins && a.record_every > 0 && s % a.record_every == 0) {
            int8_t *dst = a.out_spins + ((size_t)n_rec * a.R + r) * n;
            for (int i = tid; i < n; i += nthr) dst[i] = (int8_t)ld_spin(i);
            ++n_rec;
        }
        __syncthreads();  // copies of this sweep's state are done before the next sweep changes it
    }
    for (int i = tid; i < n; i += nthr) g_spin[i] = (int8_t)ld_spin(i);
    if (a.bestE && tid == 0) a.bestE[r] = best;
}

__global__ void col_energy_kernel(int n, const int32_t *__restrict__ rp, const int32_t *__restrict__ ci,
                                  const double *__restrict__ val, const double *__restrict__ h,
                                  const int8_t *__restrict__ spins, double *E) {
    __shared__ double red[8];
    const int8_t *m = spins + (size_t)blockIdx.x * n;
    double part = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double x = 0.0;
        for (int p = rp[i]; p < rp[i + 1]; ++p) x += val[p] * (double)m[ci[p]];
        part += (double)m[i] * (0.5 * x + h[i]);
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) E[blockIdx.x] = -v;
    }
}

__global__ void col_init_kernel(size_t count, int n, int replica_offset, int8_t *spins, uint32_t seed_lo, uint32_t seed_hi) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const PhiloxC rng{seed_lo, seed_hi ^ 0x494e4954u};
    const uint4 x = rng((uint32_t)(replica_offset + i / n), (uint32_t)(i % n), 0u, 7u);
    spins[i] = (x.x & 1u) ? 1 : -1;
}

__global__ void col_fill_kernel(int count, double *p, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

}  // namespace nlmc

extern "C" {

int nlmc_col_destroy(nlmc_col *Cc) {
    if (!Cc) return NLMC_OK;
    cudaSetDevice(Cc->inst->device);
    void *ptrs[] = {Cc->site_order, Cc->colour_ptr, Cc->col16, Cc->valfx, Cc->val8, Cc->spins, Cc->beta, Cc->modes, Cc->bestE, Cc->bestS,
                    Cc->state_ws};
    for (void *p : ptrs) if (p) cudaFree(p);
    Cc->xch.release();
    if (Cc->stream) cudaStreamDestroy(Cc->stream);
    delete Cc;
    return NLMC_OK;
}

int nlmc_col_set_betas(nlmc_col *Cc, const double *betas) {
    NLMC_REQUIRE(Cc && betas, "nlmc_col_set_betas: NULL argument");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    NLMC_CUDA(cudaMemcpyAsync(Cc->beta, betas, sizeof(double) * (size_t)Cc->R, cudaMemcpyHostToDevice, Cc->stream));
    NLMC_CUDA(cudaStreamSynchronize(Cc->stream));
    return NLMC_OK;
}

int nlmc_col_create(nlmc_instance *I, int n_replicas, const double *betas, int replica_offset, unsigned long long seed,
                    nlmc_col **out) {
    using namespace nlmc;
    NLMC_REQUIRE(I && out && betas && n_replicas >= 1, "nlmc_col_create: bad arguments");
    *out = nullptr;
    { const int rc_dev = nlmc::instance_device(I); if (rc_dev) return rc_dev; }   // the CSR on the device (uploaded on first use)
    NLMC_REQUIRE(nlmc::instance_value_symmetric(I), "nlmc_col_create: J must be symmetric (J_ij == J_ji) without repeated entries; "
                                     "the coloured sweep updates the neighbours' fields from the flipped site's own row");
    const int n = I->n, nnz = I->nnz;
    // greedy colouring in site order; a site's colour must differ from all its neighbours' (both directions)
    std::vector<std::vector<int>> adj((size_t)n);
    for (int i = 0; i < n; ++i)
        for (int p = I->h_row_ptr[i]; p < I->h_row_ptr[i + 1]; ++p) {
            const int j = I->h_col[(size_t)p];
            if (j == i || I->h_val[(size_t)p] == 0.0) continue;
            adj[(size_t)i].push_back(j);
            adj[(size_t)j].push_back(i);
        }
    std::vector<int> colour((size_t)n, -1), mark;
    int n_colours = 0;
    for (int i = 0; i < n; ++i) {
        mark.assign((size_t)n_colours + 1, 0);
        for (int j : adj[(size_t)i]) if (colour[(size_t)j] >= 0) mark[(size_t)colour[(size_t)j]] = 1;
        int c = 0;
        while (c < n_colours && mark[(size_t)c]) ++c;
        colour[(size_t)i] = c;
        n_colours = std::max(n_colours, c + 1);
    }
    std::vector<int32_t> order((size_t)n), cptr((size_t)n_colours + 1, 0);
    for (int i = 0; i < n; ++i) ++cptr[(size_t)colour[(size_t)i] + 1];
    for (int c = 0; c < n_colours; ++c) cptr[(size_t)c + 1] += cptr[(size_t)c];
    {
        std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
        for (int i = 0; i < n; ++i) order[(size_t)fill[(size_t)colour[(size_t)i]]++] = i;
    }
    NLMC_CUDA(cudaSetDevice(I->device));
    auto *Cc = new nlmc_col();
    Cc->inst = I;
    Cc->n = n;
    Cc->R = n_replicas;
    Cc->n_colours = n_colours;
    Cc->replica_offset = replica_offset;
    Cc->seed = seed;
    Cc->small_cols = n <= 65535;
    // fixed-point scale: the largest |field| (sum_j |J_ij| + |h_i|) must stay below 2^30
    double max_row = 1e-300;
    int max_colour = 1;
    for (int c = 0; c < n_colours; ++c) max_colour = std::max(max_colour, (int)(cptr[(size_t)c + 1] - cptr[(size_t)c]));
    for (int i = 0; i < n; ++i) {
        double srow = std::fabs(I->h_h[(size_t)i]);
        for (int p = I->h_row_ptr[i]; p < I->h_row_ptr[i + 1]; ++p) srow += std::fabs(I->h_val[(size_t)p]);
        max_row = std::max(max_row, srow);
    }
    Cc->fx_shift = std::max(0, std::min(30, (int)std::floor(std::log2(1073741824.0 / max_row))));
    const size_t base = 2 * sizeof(int32_t) * (size_t)n + 2 * (size_t)n + 64;
    bool all_small_int = nnz > 0;
    for (int p = 0; p < nnz && all_small_int; ++p) {
        const double v = I->h_val[(size_t)p];
        all_small_int = v == std::floor(v) && std::fabs(v) <= 127.0;
    }
    Cc->int8_vals = all_small_int;
    const size_t with_csr = base + sizeof(int32_t) * (size_t)(n + 1) + (all_small_int ? 1 : 4) * ((size_t)nnz + 4) +
                            (Cc->small_cols ? 2 : 4) * ((size_t)nnz + ((size_t)nnz & 1));
    // CSR in shared memory only when it fits AND the replicas are few: with many replicas the small footprint of
    // the L2-resident variant (8 CTAs per SM instead of 1) hides latency better (measured on the C1 graph:
    // 27 vs 41 us per sweep for one replica, but 9e9 vs 2.1e10 attempts/s for 1184 replicas)
    cudaDeviceGetAttribute(&Cc->sm_count, cudaDevAttrMultiProcessorCount, I->device);
    Cc->csr_in_smem = with_csr <= 220 * 1024 && n_replicas <= 2 * Cc->sm_count;
    Cc->smem_bytes = Cc->csr_in_smem ? with_csr : base;
    // few replicas: one big CTA per replica (all sites of a colour in one or two passes); many: small CTAs, more per SM
    Cc->threads = n_replicas <= 2 * Cc->sm_count ? kColThreadsFew : kColThreads;
    Cc->group = 1;
    while (Cc->group < 32 && Cc->group * 2 * max_colour <= Cc->threads) Cc->group *= 2;
    Cc->smem_bytes_nocsr = base;
    if (Cc->smem_bytes > 220 * 1024 || getenv("NLMC_COL_FORCE_GLOBAL")) {  // state does not fit in shared memory: keep it in a global workspace (slower, any n)
        Cc->state_global = true;
        Cc->csr_in_smem = false;
        Cc->state_stride = (base + 255) & ~(size_t)255;
        Cc->smem_bytes = 64;
        Cc->threads = kColThreadsFew;
        Cc->group = 1;
        while (Cc->group < 32 && Cc->group * 2 * max_colour <= Cc->threads) Cc->group *= 2;
    }
    std::vector<int32_t> v32((size_t)std::max(nnz, 1));
    std::vector<uint16_t> c16((size_t)std::max(nnz, 1));
    std::vector<int8_t> v8((size_t)std::max(nnz, 1));
    const double scale = std::ldexp(1.0, Cc->fx_shift);
    for (int p = 0; p < nnz; ++p) {
        if (Cc->int8_vals) v8[(size_t)p] = (int8_t)I->h_val[(size_t)p];
        v32[(size_t)p] = (int32_t)std::llrint(I->h_val[(size_t)p] * scale);
        c16[(size_t)p] = (uint16_t)I->h_col[(size_t)p];
    }
    const size_t rn = (size_t)n_replicas * n;
    bool ok = cudaStreamCreateWithFlags(&Cc->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&Cc->site_order, sizeof(int32_t) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&Cc->colour_ptr, sizeof(int32_t) * cptr.size()) == cudaSuccess &&
              cudaMalloc(&Cc->col16, sizeof(uint16_t) * c16.size()) == cudaSuccess &&
              cudaMalloc(&Cc->valfx, sizeof(int32_t) * v32.size()) == cudaSuccess &&
              cudaMalloc(&Cc->val8, v8.size()) == cudaSuccess &&
              cudaMemcpy(Cc->val8, v8.data(), v8.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMalloc(&Cc->spins, rn) == cudaSuccess && cudaMalloc(&Cc->beta, sizeof(double) * (size_t)n_replicas) == cudaSuccess &&
              cudaMalloc(&Cc->bestE, sizeof(double) * (size_t)n_replicas) == cudaSuccess && cudaMalloc(&Cc->bestS, rn) == cudaSuccess &&
              cudaMemcpy(Cc->site_order, order.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(Cc->colour_ptr, cptr.data(), sizeof(int32_t) * cptr.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(Cc->col16, c16.data(), sizeof(uint16_t) * c16.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(Cc->valfx, v32.data(), sizeof(int32_t) * v32.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              (!Cc->state_global || cudaMalloc(&Cc->state_ws, Cc->state_stride * (size_t)n_replicas) == cudaSuccess);
    if (!ok) {
        set_error("nlmc_col_create: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        nlmc_col_destroy(Cc);
        return NLMC_ERR_CUDA;
    }
    int rc = nlmc_col_set_betas(Cc, betas);
    if (!rc) {
        col_init_kernel<<<(unsigned)((rn + 255) / 256), 256, 0, Cc->stream>>>(rn, n, replica_offset, Cc->spins, (uint32_t)seed,
                                                                             (uint32_t)(seed >> 32));
        col_fill_kernel<<<(n_replicas + 127) / 128, 128, 0, Cc->stream>>>(n_replicas, Cc->bestE, 1e300);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(Cc->stream) != cudaSuccess) {
            set_error("nlmc_col_create: init kernels failed");
            rc = NLMC_ERR_CUDA;
        }
    }
    if (rc) {
        nlmc_col_destroy(Cc);
        return rc;
    }
    *out = Cc;
    return NLMC_OK;
}

int nlmc_col_info(const nlmc_col *Cc, int *n_colours, int *csr_in_smem) {
    NLMC_REQUIRE(Cc, "nlmc_col_info: NULL handle");
    if (n_colours) *n_colours = Cc->n_colours;
    if (csr_in_smem) *csr_in_smem = Cc->csr_in_smem ? 1 : 0;
    return NLMC_OK;
}

int nlmc_col_set_spins(nlmc_col *Cc, const int8_t *spins) {
    NLMC_REQUIRE(Cc && spins, "nlmc_col_set_spins: NULL argument");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    NLMC_CUDA(cudaMemcpyAsync(Cc->spins, spins, (size_t)Cc->R * Cc->n, cudaMemcpyHostToDevice, Cc->stream));
    NLMC_CUDA(cudaStreamSynchronize(Cc->stream));
    return NLMC_OK;
}

int nlmc_col_get_spins(nlmc_col *Cc, int8_t *out) {
    NLMC_REQUIRE(Cc && out, "nlmc_col_get_spins: NULL argument");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    NLMC_CUDA(cudaMemcpyAsync(out, Cc->spins, (size_t)Cc->R * Cc->n, cudaMemcpyDeviceToHost, Cc->stream));
    NLMC_CUDA(cudaStreamSynchronize(Cc->stream));
    return NLMC_OK;
}

int nlmc_col_set_site_modes(nlmc_col *Cc, const uint8_t *modes, double temp_x) {
    NLMC_REQUIRE(Cc, "nlmc_col_set_site_modes: NULL handle");
    NLMC_REQUIRE(!modes || temp_x > 0.0, "nlmc_col_set_site_modes: temp_x must be positive");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    if (!modes) {
        Cc->modes_on = false;
        return NLMC_OK;
    }
    const size_t rn = (size_t)Cc->R * Cc->n;
    if (!Cc->modes) NLMC_CUDA(cudaMalloc(&Cc->modes, rn));
    NLMC_CUDA(cudaMemcpyAsync(Cc->modes, modes, rn, cudaMemcpyHostToDevice, Cc->stream));
    NLMC_CUDA(cudaStreamSynchronize(Cc->stream));
    Cc->modes_on = true;
    Cc->temp_x = temp_x;
    return NLMC_OK;
}

int nlmc_col_best_reset(nlmc_col *Cc) {
    NLMC_REQUIRE(Cc, "nlmc_col_best_reset: NULL handle");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    nlmc::col_fill_kernel<<<(Cc->R + 127) / 128, 128, 0, Cc->stream>>>(Cc->R, Cc->bestE, 1e300);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

int nlmc_col_best_get(nlmc_col *Cc, int8_t *out_spins, double *out_E) {
    NLMC_REQUIRE(Cc, "nlmc_col_best_get: NULL handle");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    if (out_spins) NLMC_CUDA(cudaMemcpyAsync(out_spins, Cc->bestS, (size_t)Cc->R * Cc->n, cudaMemcpyDeviceToHost, Cc->stream));
    if (out_E) NLMC_CUDA(cudaMemcpyAsync(out_E, Cc->bestE, sizeof(double) * (size_t)Cc->R, cudaMemcpyDeviceToHost, Cc->stream));
    NLMC_CUDA(cudaStreamSynchronize(Cc->stream));
    return NLMC_OK;
}

/* n_sweeps sweeps in ONE launch.  out_E [n_sweeps][R] (optional): energy after every sweep.  out_spins
 * [ceil(n_sweeps/record_every)][R][n] (optional): the state after sweeps 0, record_every, ... (the reference's
 * M[:, ::M_skip]).  track_best != 0: keep the lowest-energy state since nlmc_col_best_reset. */
int nlmc_col_sweep(nlmc_col *Cc, int n_sweeps, const double *beta_sched, int record_every, int8_t *out_spins,
                   double *out_E, int track_best) {
    using namespace nlmc;
    NLMC_REQUIRE(Cc && n_sweeps >= 0, "nlmc_col_sweep: bad arguments");
    NLMC_REQUIRE(!out_spins || record_every >= 1, "nlmc_col_sweep: record_every must be >= 1 when recording");
    if (n_sweeps == 0) return NLMC_OK;
    nlmc_instance *I = Cc->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    const size_t R = (size_t)Cc->R, n = (size_t)Cc->n;
    const size_t n_rec = out_spins ? ((size_t)n_sweeps + record_every - 1) / record_every : 0;
    int8_t *d_rec = nullptr;
    double *d_E = nullptr, *d_sched = nullptr;
    if (beta_sched) {
        NLMC_CUDA(cudaMalloc(&d_sched, sizeof(double) * (size_t)n_sweeps * R));
        if (cudaMemcpyAsync(d_sched, beta_sched, sizeof(double) * (size_t)n_sweeps * R, cudaMemcpyHostToDevice, Cc->stream) != cudaSuccess) {
            cudaFree(d_sched);
            set_error("nlmc_col_sweep: copy of the beta schedule failed");
            return NLMC_ERR_CUDA;
        }
    }
    if (out_spins && cudaMalloc(&d_rec, n_rec * R * n) != cudaSuccess) {
        if (d_sched) cudaFree(d_sched);
        set_error("nlmc_col_sweep: cudaMalloc failed");
        return NLMC_ERR_CUDA;
    }
    if (out_E && cudaMalloc(&d_E, sizeof(double) * (size_t)n_sweeps * R) != cudaSuccess) {
        if (d_rec) cudaFree(d_rec);
        if (d_sched) cudaFree(d_sched);
        set_error("nlmc_col_sweep: cudaMalloc failed");
        return NLMC_ERR_CUDA;
    }
    ColArgs a;
    a.n = Cc->n; a.nnz = I->nnz; a.n_colours = Cc->n_colours; a.n_sweeps = n_sweeps; a.record_every = record_every;
    a.replica_offset = Cc->replica_offset;
    a.rp = I->row_ptr; a.ci = I->col; a.val = I->val; a.h = I->h; a.col16 = Cc->col16; a.valfx = Cc->valfx;
    a.state_ws = Cc->state_ws; a.state_stride = Cc->state_stride;
    a.group = Cc->group; a.scale = std::ldexp(1.0, Cc->fx_shift); a.val8 = Cc->val8; a.shift = Cc->fx_shift;
    a.site_order = Cc->site_order; a.colour_ptr = Cc->colour_ptr;
    a.spins = Cc->spins; a.beta = Cc->beta; a.beta_sched = d_sched; a.modes = Cc->modes_on ? Cc->modes : nullptr; a.temp_x = Cc->temp_x;
    a.seed_lo = (uint32_t)Cc->seed; a.seed_hi = (uint32_t)(Cc->seed >> 32); a.sweep0 = Cc->sweep_counter;
    a.out_spins = d_rec; a.out_E = d_E; a.bestE = track_best ? Cc->bestE : nullptr; a.bestS = Cc->bestS; a.R = Cc->R;
    cudaError_t e = cudaSuccess;
    const int smem = (int)Cc->smem_bytes;
#define NLMC_COL_LAUNCH_G(SMEM, COLT, VALT, GST)                                                                       \
    do {                                                                                                              \
        e = cudaFuncSetAttribute(col_sweep_kernel<SMEM, COLT, VALT, GST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e == cudaSuccess) col_sweep_kernel<SMEM, COLT, VALT, GST><<<(unsigned)R, (unsigned)Cc->threads, smem, Cc->stream>>>(a);  \
    } while (0)
#define NLMC_COL_LAUNCH(SMEM, COLT, VALT)                                              \
    do {                                                                              \
        if (Cc->state_ws) NLMC_COL_LAUNCH_G(SMEM, COLT, VALT, true);                  \
        else NLMC_COL_LAUNCH_G(SMEM, COLT, VALT, false);                              \
    } while (0)
    if (Cc->csr_in_smem && Cc->small_cols) {
        if (Cc->int8_vals) NLMC_COL_LAUNCH(true, uint16_t, int8_t); else NLMC_COL_LAUNCH(true, uint16_t, int32_t);
    } else if (Cc->csr_in_smem) {
        if (Cc->int8_vals) NLMC_COL_LAUNCH(true, int32_t, int8_t); else NLMC_COL_LAUNCH(true, int32_t, int32_t);
    } else {
        if (Cc->int8_vals) NLMC_COL_LAUNCH(false, int32_t, int8_t); else NLMC_COL_LAUNCH(false, int32_t, int32_t);
    }
#undef NLMC_COL_LAUNCH
#undef NLMC_COL_LAUNCH_G
    if (e == cudaSuccess) e = cudaGetLastError();
    Cc->sweep_counter += (uint32_t)n_sweeps;
    if (e == cudaSuccess && d_rec) e = cudaMemcpyAsync(out_spins, d_rec, n_rec * R * n, cudaMemcpyDeviceToHost, Cc->stream);
    if (e == cudaSuccess && d_E) e = cudaMemcpyAsync(out_E, d_E, sizeof(double) * (size_t)n_sweeps * R, cudaMemcpyDeviceToHost, Cc->stream);
    if (e == cudaSuccess && (d_rec || d_E || d_sched)) e = cudaStreamSynchronize(Cc->stream);
    if (d_rec) cudaFree(d_rec);
    if (d_E) cudaFree(d_E);
    if (d_sched) cudaFree(d_sched);
    if (e != cudaSuccess) {
        set_error("nlmc_col_sweep: %s", cudaGetErrorString(e));
        return NLMC_ERR_CUDA;
    }
    return NLMC_OK;
}

int nlmc_col_energies(nlmc_col *Cc, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(Cc && out_E, "nlmc_col_energies: NULL argument");
    nlmc_instance *I = Cc->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    double *d_E = nullptr;
    NLMC_CUDA(cudaMalloc(&d_E, sizeof(double) * (size_t)Cc->R));
    col_energy_kernel<<<(unsigned)Cc->R, 256, 0, Cc->stream>>>(Cc->n, I->row_ptr, I->col, I->val, I->h, Cc->spins, d_E);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_E, d_E, sizeof(double) * (size_t)Cc->R, cudaMemcpyDeviceToHost, Cc->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(Cc->stream);
    cudaFree(d_E);
    if (e != cudaSuccess) {
        set_error("nlmc_col_energies: %s", cudaGetErrorString(e));
        return NLMC_ERR_CUDA;
    }
    return NLMC_OK;
}

/* ---- replica exchange by beta labels (nlmc_exchange.cuh) ---- */
int nlmc_col_ladders(nlmc_col *Cc, int n_beta, const double *betas) {
    NLMC_REQUIRE(Cc, "nlmc_col_ladders: NULL handle");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    return nlmc::exchange_setup<double>(Cc->xch, Cc->R, n_beta, betas, Cc->beta, Cc->stream);
}

int nlmc_col_exchange(nlmc_col *Cc, int num_swapping_pairs) {
    using namespace nlmc;
    NLMC_REQUIRE(Cc && Cc->xch.active(), "nlmc_col_exchange: call nlmc_col_ladders first");
    nlmc_instance *I = Cc->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    col_energy_kernel<<<(unsigned)Cc->R, 256, 0, Cc->stream>>>(Cc->n, I->row_ptr, I->col, I->val, I->h, Cc->spins, Cc->xch.E);
    NLMC_CUDA(cudaGetLastError());
    return exchange_launch<double>(Cc->xch, num_swapping_pairs, Cc->beta, Cc->seed, Cc->replica_offset, Cc->stream);
}

int nlmc_col_labels(nlmc_col *Cc, int32_t *out_labels, int n_rounds, int32_t *out_counts) {
    NLMC_REQUIRE(Cc, "nlmc_col_labels: NULL handle");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    return nlmc::exchange_fetch(Cc->xch, Cc->R, out_labels, n_rounds, out_counts, Cc->stream);
}

int nlmc_col_sync(nlmc_col *Cc) {
    NLMC_REQUIRE(Cc, "nlmc_col_sync: NULL handle");
    NLMC_CUDA(cudaSetDevice(Cc->inst->device));
    NLMC_CUDA(cudaStreamSynchronize(Cc->stream));
    return NLMC_OK;
}

}  // extern "C"
