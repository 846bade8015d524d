// nlmc_lbp.cu -- K5: loopy belief propagation of the NMC backbone search on the edges of J.
//
// Replaces LoopyBeliefPropagation (NMC/nmc.py:168-228 == NPT/npt.py:204-264), called once per
// lambda by LBP_convexified (nmc.py:131-161).  The reference keeps dense N x N message matrices;
// off the edges of J they are trivial (u = 0, h_msgs[i,j] = total_i), so the kernel keeps one
// message per stored entry of J plus the common off-edge value tot[i] of each h_msgs row, which
// still enters the reference's convergence maxima (nmc.py:208-209) for rows with an off-diagonal
// zero.  One LBP call = ONE cooperative launch: the Jacobi iterations, the two relative-change
// maxima and the convergence test all stay on the device, separated by grid syncs.
//
// Bit-level fidelity: the additions reproduce numpy's association order (np.sum over a strided
// column is pairwise over all N entries, np.sum(axis=0) is sequential; zeros do not change a
// partial sum so only stored entries are visited).  tanh/arctanh are the restatements of numpy's
// own float64 routines in nlmc_npmath.h (bit-equal to np.tanh / np.arctanh of the AVX-512 numpy
// build the goldens come from): the reference's stopping rule (tolerance = machine epsilon) waits
// for an exact floating-point fixed point, so iteration counts, divergence points and the backbone
// depend on the last bit of those two functions -- see DESIGN.md "LBP parity".
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#include "nlmc_common.cuh"
#include "nlmc_npmath.h"

namespace cg = cooperative_groups;

struct nlmc_lbp {
    nlmc_instance *inst = nullptr;
    int32_t *rev = nullptr;     // [nnz] entry (j,i) for entry (i,j)
    double *u[2] = {nullptr, nullptr};  // [nnz] u messages, double buffered
    double *hm = nullptr;       // [nnz] h messages
    double *tj = nullptr;       // [nnz] tanh(beta J), refreshed at the start of every launch
    double *uin = nullptr;      // [nnz] uin[(i,j)] = u[(j,i)]: the message entry (i,j) gathers, stored where row i reads it
    int32_t *prog = nullptr;    // summation programs: numpy's pairwise order of each column sum, resolved on the host
    int32_t *prog_ptr = nullptr;  // [n+1]
    int warp_rows = 0;          // 1: a warp per row in the gather (rows of >= ~12 entries), 0: a thread per row
    int row_cap = 0;            // longest row (entries) when warp_rows is set: sizes the per-warp shared staging
    size_t smem_bytes = 0;
    double *tot = nullptr;      // [n]   off-edge value of each h_msgs row
    double *eps = nullptr;      // [n]   |h_i| + sum_j |J_ij|   (nmc.py:353)
    double *mstar = nullptr;    // [n]
    double *marg = nullptr;     // [n]
    double *hfield = nullptr;   // [n]   caller-supplied field of nlmc_lbp_run
    double *dense[2] = {nullptr, nullptr};  // [n*n] correlations / J_tilde of nlmc_lbp_byproducts (on demand)
    double *htilde = nullptr;   // [n]
    uint8_t *offedge = nullptr; // [n]   row has an off-diagonal zero
    unsigned long long *red = nullptr;  // [2][4] du, su, dh, sh as ordered bit patterns, double buffered by iteration parity
    int *iter_out = nullptr;
    int cur = 0;
    int grid = 0;
};

namespace nlmc {

// numpy's DOUBLE_pairwise_sum over a dense vector of length n whose only non-zeros are at the
// sorted positions pos[0..cnt) with values v[...].  Explicit-stack post-order walk of numpy's
// recursion; ranges without stored entries are pruned (their sum is 0 and x + 0 == x).
template <typename GetV>
__device__ double pairwise_sparse(int n, const int32_t *__restrict__ pos, int cnt, GetV v) {
    int st_lo[32], st_n[32];
    double st_acc[32];
    int8_t st_stage[32];
    int sp = 0, cur = 0;
    st_lo[0] = 0; st_n[0] = n; st_stage[0] = 0; st_acc[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const int lo = st_lo[sp], len = st_n[sp];
        if (st_stage[sp] == 0) {
            if (cur >= cnt || pos[cur] >= lo + len) {  // no stored entry in range
                ret = 0.0; --sp; continue;
            }
            if (len < 8) {
                double res = 0.0;
                while (cur < cnt && pos[cur] < lo + len) { res = __dadd_rn(res, v(cur)); ++cur; }
                ret = res; --sp; continue;
            }
            if (len <= 128) {
                double r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const int body_end = lo + len - (len % 8);
                while (cur < cnt && pos[cur] < body_end) {
                    const int j = (pos[cur] - lo) & 7;
                    const double x = v(cur);
#pragma unroll
                    for (int q = 0; q < 8; ++q) if (q == j) r[q] = __dadd_rn(r[q], x);
                    ++cur;
                }
                double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                       __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                while (cur < cnt && pos[cur] < lo + len) { res = __dadd_rn(res, v(cur)); ++cur; }
                ret = res; --sp; continue;
            }
            int n2 = len / 2;
            n2 -= n2 % 8;
            st_stage[sp] = 1;
            ++sp;
            st_lo[sp] = lo; st_n[sp] = n2; st_stage[sp] = 0;
            continue;
        }
        if (st_stage[sp] == 1) {  // left child done
            int n2 = len / 2;
            n2 -= n2 % 8;
            st_acc[sp] = ret;
            st_stage[sp] = 2;
            ++sp;
            st_lo[sp] = lo + n2; st_n[sp] = len - n2; st_stage[sp] = 0;
            continue;
        }
        ret = __dadd_rn(st_acc[sp], ret);  // both children done
        --sp;
    }
    return ret;
}

__device__ __forceinline__ double atanh_saturated(double x) {  // nmc.py:230-255 (tanh(19.06) == 1.0)
    const double e = 2.220446049250313e-16;
    x = fmin(fmax(x, -1.0 + e), 1.0 - e);
    return nlmc_np_arctanh(x);
}

__device__ __forceinline__ void atomic_max_nonneg(unsigned long long *addr, double v) {
    // non-negative doubles order like their bit patterns; NaN (0x7ff8...) sorts above everything
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ void block_max2(double &a, double &b, double *sm) {
    a = warp_max(a);
    b = warp_max(b);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { sm[w] = a; sm[32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = lane < nw ? sm[lane] : 0.0;
        b = lane < nw ? sm[32 + lane] : 0.0;
        a = warp_max(a);
        b = warp_max(b);
    }
}

// The order in which numpy's pairwise sum associates the additions depends only on the positions of the non-zeros,
// i.e. on the sparsity pattern of J: it is resolved ONCE on the host into a postfix program per row
// (op >= 0: push the row's op-th incoming message, op == -1: add the two topmost values) and replayed here.
// The per-iteration gather is then a short dependent chain instead of a walk over numpy's recursion tree
// (ncu: 83 % of the kernel's samples were CTAs waiting at the grid barrier for the slowest such walk).
constexpr int kProgStack = 40;
__device__ __forceinline__ double run_sum_program(const int32_t *__restrict__ prog, int len, const double *__restrict__ row) {
    double st[kProgStack];
    int sp = 0;
    for (int k = 0; k < len; ++k) {
        const int op = prog[k];
        if (op >= 0) {
            st[sp++] = row[op];
        } else {
            --sp;
            st[sp - 1] = __dadd_rn(st[sp - 1], st[sp]);
        }
    }
    return sp ? st[0] : 0.0;
}

// host: emit the program of pairwise_sparse(n, pos[0..cnt)) -- same association, zero operands skipped (x + 0 == x)
struct SumProgramBuilder {
    const int32_t *pos;
    int cnt, cur = 0, depth = 0, max_depth = 0;
    std::vector<int32_t> *out;
    void push(int q) { out->push_back(q); max_depth = std::max(max_depth, ++depth); }
    void add() { out->push_back(-1); --depth; }
    bool chain(bool have, int q) {  // res += v(q)
        push(q);
        if (have) add();
        return true;
    }
    bool emit(int lo, int len) {
        if (cur >= cnt || pos[cur] >= lo + len) return false;
        if (len < 8) {
            bool have = false;
            while (cur < cnt && pos[cur] < lo + len) have = chain(have, cur++);
            return have;
        }
        if (len <= 128) {
            const int body_end = lo + len - (len % 8);
            std::vector<int> acc[8];
            while (cur < cnt && pos[cur] < body_end) { acc[(pos[cur] - lo) & 7].push_back(cur); ++cur; }
            auto leaf = [&](int j) { bool have = false; for (int q : acc[j]) have = chain(have, q); return have; };
            auto pair = [&](int j) { const bool a = leaf(j), b = leaf(j + 1); if (a && b) add(); return a || b; };
            auto quad = [&](int j) { const bool a = pair(j), b = pair(j + 2); if (a && b) add(); return a || b; };
            const bool a = quad(0), b = quad(4);
            if (a && b) add();
            bool have = a || b;
            while (cur < cnt && pos[cur] < lo + len) have = chain(have, cur++);
            return have;
        }
        int n2 = len / 2;
        n2 -= n2 % 8;
        const bool a = emit(lo, n2), b = emit(lo + n2, len - n2);
        if (a && b) add();
        return a || b;
    }
};

struct LbpArgs {
    int n, nnz, max_iter, warp_rows, row_cap;
    double beta, lambda, tol;
    const int32_t *rp, *ci, *rev, *prog, *prog_ptr;
    const double *val, *h, *eps, *mstar;
    const double *h_field;  // non-NULL: use this field instead of h + lambda*m_star*eps (nlmc_lbp_run)
    double *u0, *u1, *hm, *uin, *tot, *marg, *tj;
    const uint8_t *offedge;
    unsigned long long *red;
    int *iter_out;  // [0] iteration on exit, [1] index of the buffer holding the final u
};

__global__ void __launch_bounds__(256) lbp_kernel(LbpArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double lbp_dyn[];  // warp-per-row staging (empty in the thread-per-row mode)
    __shared__ double sm[64];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const double inv_beta = __ddiv_rn(1.0, a.beta);
    double *u_old = a.u0, *u_new = a.u1;
    int iteration = 0;
    bool converged = false;
    // two grid-wide syncs per iteration: the reduction slots alternate with the iteration parity, and the slots of
    // the NEXT iteration are cleared during this one's update phase (nobody touches them in between)
    if (tid < 8) a.red[tid] = 0ull;
    for (int p = tid; p < a.nnz; p += nthreads) a.tj[p] = nlmc_np_tanh(__dmul_rn(a.beta, a.val[p]));  // constant over the iterations
    grid.sync();
    for (iteration = 0; iteration < a.max_iter; ++iteration) {
        unsigned long long *red = a.red + 4 * (iteration & 1), *red_next = a.red + 4 * ((iteration + 1) & 1);
        // ---- gather: total_i = hl_i + sum_k u[k,i];  hm[i,j] = total_i - u[j,i]   (nmc.py:200-203)
        // uin holds u[j,i] at entry (i,j), so a row reads its incoming messages contiguously
        double dh = 0.0, sh = 0.0;
        if (a.warp_rows) {
            // a warp per row: the lanes stage the row's incoming messages and its summation program in shared memory
            // with one round of coalesced loads, lane 0 replays the program out of shared memory, then the lanes
            // update the row's h messages in parallel (a thread per row spends ~2 us PER ENTRY in dependent L2 loads
            // when only a few hundred rows exist)
            const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
            double *vals = lbp_dyn + (size_t)wib * a.row_cap;
            int32_t *pr = reinterpret_cast<int32_t *>(lbp_dyn + (size_t)(blockDim.x >> 5) * a.row_cap) + (size_t)wib * 2 * a.row_cap;
            for (int i = tid >> 5; i < a.n; i += nthreads >> 5) {
                const int b = a.rp[i], cnt = a.rp[i + 1] - b;
                const int pp = a.prog_ptr[i], plen = a.prog_ptr[i + 1] - pp;
                for (int q = lane; q < cnt; q += 32) vals[q] = a.uin[b + q];
                for (int k = lane; k < plen; k += 32) pr[k] = a.prog[pp + k];
                __syncwarp();
                double total = 0.0;
                if (lane == 0) {
                    const double hl = a.h_field ? a.h_field[i]
                                                : __dadd_rn(a.h[i], __dmul_rn(__dmul_rn(a.lambda, a.mstar[i]), a.eps[i]));
                    total = __dadd_rn(hl, run_sum_program(pr, plen, vals));
                }
                total = __shfl_sync(0xffffffffu, total, 0);
                for (int q = lane; q < cnt; q += 32) {
                    const double hnew = (a.ci[b + q] == i) ? 0.0 : __dsub_rn(total, vals[q]);
                    const double hold = a.hm[b + q];
                    dh = fmax(dh, fabs(__dsub_rn(hnew, hold)));
                    sh = fmax(sh, __dadd_rn(fabs(hnew), fabs(hold)));
                    a.hm[b + q] = hnew;
                }
                if (lane == 0) {
                    if (a.offedge[i]) {
                        const double told = a.tot[i];
                        dh = fmax(dh, fabs(__dsub_rn(total, told)));
                        sh = fmax(sh, __dadd_rn(fabs(total), fabs(told)));
                    }
                    a.tot[i] = total;
                }
                __syncwarp();
            }
        } else
        for (int i = tid; i < a.n; i += nthreads) {
            const int b = a.rp[i], cnt = a.rp[i + 1] - b;
            const double hl = a.h_field ? a.h_field[i]
                                        : __dadd_rn(a.h[i], __dmul_rn(__dmul_rn(a.lambda, a.mstar[i]), a.eps[i]));  // nmc.py:133-134
            const double *uin_row = a.uin + b;
            const double total = __dadd_rn(hl, run_sum_program(a.prog + a.prog_ptr[i], a.prog_ptr[i + 1] - a.prog_ptr[i], uin_row));
            for (int q = 0; q < cnt; ++q) {
                const double hnew = (a.ci[b + q] == i) ? 0.0 : __dsub_rn(total, uin_row[q]);
                const double hold = a.hm[b + q];
                dh = fmax(dh, fabs(__dsub_rn(hnew, hold)));
                sh = fmax(sh, __dadd_rn(fabs(hnew), fabs(hold)));
                a.hm[b + q] = hnew;
            }
            if (a.offedge[i]) {
                const double told = a.tot[i];
                dh = fmax(dh, fabs(__dsub_rn(total, told)));
                sh = fmax(sh, __dadd_rn(fabs(total), fabs(told)));
            }
            a.tot[i] = total;
        }
        block_max2(dh, sh, sm);
        if (threadIdx.x == 0) { atomic_max_nonneg(red + 2, dh); atomic_max_nonneg(red + 3, sh); }
        grid.sync();
        // ---- update: u = (1/beta) * atanh_sat(tanh(beta J) * tanh(beta h_msgs))       (nmc.py:205)
        if (tid < 4) red_next[tid] = 0ull;
        double du = 0.0, su = 0.0;
        for (int p = tid; p < a.nnz; p += nthreads) {
            const double tj = a.tj[p];
            const double th = nlmc_np_tanh(__dmul_rn(a.beta, a.hm[p]));
            const double un = __dmul_rn(inv_beta, atanh_saturated(__dmul_rn(tj, th)));
            const double uo = u_old[p];
            du = fmax(du, fabs(__dsub_rn(un, uo)));
            su = fmax(su, __dadd_rn(fabs(un), fabs(uo)));
            u_new[p] = un;
            a.uin[a.rev[p]] = un;  // what the transposed entry gathers next
        }
        block_max2(du, su, sm);
        if (threadIdx.x == 0) { atomic_max_nonneg(red + 0, du); atomic_max_nonneg(red + 1, su); }
        grid.sync();
        double *t = u_old; u_old = u_new; u_new = t;  // u_old now holds the newest messages
        const double gdu = __longlong_as_double((long long)red[0]), gsu = __longlong_as_double((long long)red[1]);
        const double gdh = __longlong_as_double((long long)red[2]), gsh = __longlong_as_double((long long)red[3]);
        const double u_change = __ddiv_rn(gdu, gsu), h_change = __ddiv_rn(gdh, gsh);  // 0/0 = nan -> not converged
        converged = (u_change < a.tol) && (h_change < a.tol);  // nmc.py:212
        if (converged) break;
    }
    if (!converged) iteration = a.max_iter - 1;  // python loop variable after exhaustion
    // marginal_i = tanh(beta * (hl_i + sum_k u[k,i])), rows accumulated in order   (nmc.py:216)
    for (int i = tid; i < a.n; i += nthreads) {
        const double hl = a.h_field ? a.h_field[i]
                                    : __dadd_rn(a.h[i], __dmul_rn(__dmul_rn(a.lambda, a.mstar[i]), a.eps[i]));
        double acc = 0.0;
        for (int p = a.rp[i]; p < a.rp[i + 1]; ++p) acc = __dadd_rn(acc, a.uin[p]);
        a.marg[i] = nlmc_np_tanh(__dmul_rn(a.beta, __dadd_rn(hl, acc)));
    }
    if (tid == 0) { a.iter_out[0] = iteration; a.iter_out[1] = (u_old == a.u0) ? 0 : 1; }
}

// u = J * m_star (nmc.py:129), h_msgs = 0 (nmc.py:128)
__global__ void lbp_reset_kernel(int n, int nnz, const int32_t *ci, const int32_t *rev, const double *val,
                                 const double *mstar, double *u, double *uin, double *hm, double *tot) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int p = tid; p < nnz; p += nt) {
        const double v = __dmul_rn(val[p], mstar[ci[p]]);
        u[p] = v;
        uin[rev[p]] = v;
        hm[p] = 0.0;
    }
    for (int i = tid; i < n; i += nt) tot[i] = 0.0;
}

// epsilon_i = |h_i| + np.sum(np.abs(J), axis=1)[i]  (pairwise along the contiguous row, nmc.py:353)
__global__ void lbp_eps_kernel(int n, const int32_t *rp, const int32_t *ci, const double *val, const double *h,
                               double *eps, uint8_t *offedge) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = rp[i], cnt = rp[i + 1] - b;
    const double s = pairwise_sparse(n, ci + b, cnt, [&](int q) { return fabs(val[b + q]); });
    eps[i] = __dadd_rn(fabs(h[i]), s);
    int stored_off = 0;
    for (int q = 0; q < cnt; ++q) stored_off += (ci[b + q] != i);
    offedge[i] = stored_off < n - 1;
}

// By-products of one LBP call (nmc.py:217-226), dense like the reference's.  Off the stored entries of J
// tanh(beta*J) = 0 and h_msgs[i,j] = tot[i], so correlations[i,j] = tanh(beta tot_i) tanh(beta tot_j) / (1 + 1e-10);
// the diagonal is removed (nmc.py:221).  Stored entries are overwritten by lbp_byproducts_edges_kernel.
__global__ void lbp_byproducts_fill_kernel(int n, double beta, const double *tot, double *corr, double *jt) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x, i = blockIdx.x;
    if (j >= n) return;
    const double inv_beta = __ddiv_rn(1.0, beta);
    double c = 0.0;
    if (i != j) {
        const double ti = nlmc_np_tanh(__dmul_rn(beta, tot[i])), tj = nlmc_np_tanh(__dmul_rn(beta, tot[j]));
        c = __ddiv_rn(__dmul_rn(ti, tj), __dadd_rn(1.0, 1e-10));
    }
    if (corr) corr[(size_t)i * n + j] = c;
    if (jt) jt[(size_t)i * n + j] = __dmul_rn(inv_beta, atanh_saturated(c));
}

__global__ void lbp_byproducts_edges_kernel(int n, int nnz, double beta, const int32_t *rp, const int32_t *ci,
                                            const int32_t *rev, const double *val, const double *hm, double *corr,
                                            double *jt) {
    const int i = blockIdx.x;
    const double inv_beta = __ddiv_rn(1.0, beta);
    for (int p = rp[i] + threadIdx.x; p < rp[i + 1]; p += blockDim.x) {
        const int j = ci[p];
        double c = 0.0;
        if (i != j) {
            const double tJ = nlmc_np_tanh(__dmul_rn(beta, val[p]));
            const double th = nlmc_np_tanh(__dmul_rn(beta, hm[p])), tht = nlmc_np_tanh(__dmul_rn(beta, hm[rev[p]]));
            const double num = __dadd_rn(tJ, __dmul_rn(th, tht));
            const double den = __dadd_rn(__dadd_rn(1.0, __dmul_rn(__dmul_rn(tJ, th), tht)), 1e-10);
            c = __ddiv_rn(num, den);
        }
        if (corr) corr[(size_t)i * n + j] = c;
        if (jt) jt[(size_t)i * n + j] = __dmul_rn(inv_beta, atanh_saturated(c));
    }
}

__global__ void lbp_htilde_kernel(int n, double beta, const double *marg, double *ht) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ht[i] = __dmul_rn(__ddiv_rn(1.0, beta), atanh_saturated(marg[i]));
}

// uin[rev[p]] = u[p] (after nlmc_lbp_set_messages)
__global__ void lbp_transpose_kernel(int nnz, const int32_t *rev, const double *u, double *uin) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) uin[rev[p]] = u[p];
}

}  // namespace nlmc

extern "C" {

int nlmc_lbp_destroy(nlmc_lbp *L) {
    if (!L) return NLMC_OK;
    cudaSetDevice(L->inst->device);
    void *ptrs[] = {L->rev, L->prog, L->prog_ptr, L->u[0], L->u[1], L->hm, L->tj, L->uin, L->tot, L->eps, L->mstar, L->marg, L->offedge, L->red, L->iter_out,
                    L->hfield, L->dense[0], L->dense[1], L->htilde};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete L;
    return NLMC_OK;
}

int nlmc_lbp_create(nlmc_instance *I, nlmc_lbp **out) {
    using namespace nlmc;
    NLMC_REQUIRE(I && out, "nlmc_lbp_create: NULL argument");
    *out = nullptr;
    { const int rc_dev = nlmc::instance_device(I); if (rc_dev) return rc_dev; }   // the CSR on the device (uploaded on first use)
    const int n = I->n, nnz = I->nnz;
    // reverse-entry index (requires a symmetric sparsity pattern, as the reference's J is)
    std::vector<int32_t> rev((size_t)std::max(nnz, 1));
    for (int i = 0; i < n; ++i) {
        for (int p = I->h_row_ptr[i]; p < I->h_row_ptr[i + 1]; ++p) {
            const int j = I->h_col[p];
            const int32_t *b = I->h_col.data() + I->h_row_ptr[j], *e = I->h_col.data() + I->h_row_ptr[j + 1];
            const int32_t *it = std::lower_bound(b, e, i);
            if (it == e || *it != i) {  // unsorted rows: fall back to a linear scan
                it = std::find(b, e, i);
                NLMC_REQUIRE(it != e, "nlmc_lbp_create: J must have a symmetric sparsity pattern (entry %d,%d)", i, j);
            }
            rev[(size_t)p] = (int32_t)(it - I->h_col.data());
        }
    }
    // summation programs (rows must be sorted by column, as scipy's csr_matrix(J) delivers them)
    std::vector<int32_t> prog, prog_ptr((size_t)n + 1, 0);
    prog.reserve((size_t)2 * (size_t)std::max(nnz, 1));
    for (int i = 0; i < n; ++i) {
        const int b = I->h_row_ptr[i], cnt = I->h_row_ptr[i + 1] - b;
        const int32_t *cols = I->h_col.data() + b;
        // numpy sums dense rows / columns in index order: epsilon, the gather and the marginals all assume sorted rows
        NLMC_REQUIRE(std::is_sorted(cols, cols + cnt),
                     "nlmc_lbp_create: row %d of J is not sorted by column (scipy: A.sort_indices() on a copy)", i);
        SumProgramBuilder sb{cols, cnt};
        sb.out = &prog;
        sb.emit(0, n);
        NLMC_REQUIRE(sb.max_depth <= kProgStack, "nlmc_lbp_create: summation program of row %d is too deep (%d)", i, sb.max_depth);
        prog_ptr[(size_t)i + 1] = (int32_t)prog.size();
    }
    if (prog.empty()) prog.push_back(0);
    NLMC_CUDA(cudaSetDevice(I->device));
    auto *L = new nlmc_lbp();
    L->inst = I;
    const size_t nz = (size_t)std::max(nnz, 1);
    bool ok = cudaMalloc(&L->rev, sizeof(int32_t) * nz) == cudaSuccess &&
              cudaMalloc(&L->u[0], sizeof(double) * nz) == cudaSuccess &&
              cudaMalloc(&L->u[1], sizeof(double) * nz) == cudaSuccess &&
              cudaMalloc(&L->hm, sizeof(double) * nz) == cudaSuccess &&
              cudaMalloc(&L->uin, sizeof(double) * nz) == cudaSuccess &&
              cudaMalloc(&L->tj, sizeof(double) * nz) == cudaSuccess &&
              cudaMalloc(&L->tot, sizeof(double) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->eps, sizeof(double) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->mstar, sizeof(double) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->marg, sizeof(double) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->hfield, sizeof(double) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->htilde, sizeof(double) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->offedge, (size_t)n) == cudaSuccess &&
              cudaMalloc(&L->red, sizeof(unsigned long long) * 8) == cudaSuccess &&
              cudaMalloc(&L->iter_out, sizeof(int) * 2) == cudaSuccess &&
              cudaMemcpy(L->rev, rev.data(), sizeof(int32_t) * nz, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMalloc(&L->prog, sizeof(int32_t) * prog.size()) == cudaSuccess &&
              cudaMalloc(&L->prog_ptr, sizeof(int32_t) * prog_ptr.size()) == cudaSuccess &&
              cudaMemcpy(L->prog, prog.data(), sizeof(int32_t) * prog.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(L->prog_ptr, prog_ptr.data(), sizeof(int32_t) * prog_ptr.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        set_error("nlmc_lbp_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        nlmc_lbp_destroy(L);
        return NLMC_ERR_CUDA;
    }
    lbp_eps_kernel<<<(n + 127) / 128, 128, 0, I->stream>>>(n, I->row_ptr, I->col, I->val, I->h, L->eps, L->offedge);
    NLMC_CUDA(cudaGetLastError());
    // cooperative grid: as many CTAs as are co-resident, capped by the work
    int dev_sms = 0, per_sm = 0, coop = 0;
    NLMC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, I->device));
    if (!coop) {
        set_error("nlmc_lbp_create: device does not support cooperative launch");
        nlmc_lbp_destroy(L);
        return NLMC_ERR_UNSUPPORTED;
    }
    NLMC_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, I->device));
    // rows of a dozen entries or more get a warp each in the gather, if the staging of the longest row fits
    L->row_cap = I->max_deg;
    L->warp_rows = (nnz >= 12 * n && (size_t)I->max_deg * 128 <= 160 * 1024) ? 1 : 0;
    L->smem_bytes = L->warp_rows ? (size_t)8 * ((size_t)I->max_deg * sizeof(double) + (size_t)2 * I->max_deg * sizeof(int32_t)) : 0;
    if (L->smem_bytes > 48 * 1024 &&
        cudaFuncSetAttribute(lbp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L->smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        L->warp_rows = 0;
        L->smem_bytes = 0;
    }
    NLMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbp_kernel, 256, L->smem_bytes));
    if (per_sm < 1) {
        set_error("nlmc_lbp_create: the LBP kernel does not fit on an SM");
        nlmc_lbp_destroy(L);
        return NLMC_ERR_UNSUPPORTED;
    }
    const int want = std::max(1, (std::max(nnz, n) + 255) / 256);
    L->grid = std::max(1, std::min(want, dev_sms));  // at most one CTA per SM: the grid-wide syncs stay cheap
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    *out = L;
    return NLMC_OK;
}

int nlmc_lbp_epsilon(nlmc_lbp *L, double *out_eps) {
    NLMC_REQUIRE(L && out_eps, "nlmc_lbp_epsilon: NULL argument");
    NLMC_CUDA(cudaSetDevice(L->inst->device));
    NLMC_CUDA(cudaMemcpy(out_eps, L->eps, sizeof(double) * (size_t)L->inst->n, cudaMemcpyDeviceToHost));
    return NLMC_OK;
}

int nlmc_lbp_reset(nlmc_lbp *L, const double *m_star) {
    using namespace nlmc;
    NLMC_REQUIRE(L && m_star, "nlmc_lbp_reset: NULL argument");
    nlmc_instance *I = L->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    NLMC_CUDA(cudaMemcpyAsync(L->mstar, m_star, sizeof(double) * (size_t)I->n, cudaMemcpyHostToDevice, I->stream));
    L->cur = 0;
    lbp_reset_kernel<<<std::max(1, std::min(1024, (std::max(I->nnz, I->n) + 255) / 256)), 256, 0, I->stream>>>(
        I->n, I->nnz, I->col, L->rev, I->val, L->mstar, L->u[0], L->uin, L->hm, L->tot);
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    return NLMC_OK;
}

static int lbp_launch(nlmc_lbp *L, const double *h_field_dev, double lambda, double beta, double tol, int max_iter,
                      double *out_marginal, int *out_iteration) {
    using namespace nlmc;
    nlmc_instance *I = L->inst;
    LbpArgs a;
    a.n = I->n; a.nnz = I->nnz; a.max_iter = max_iter; a.warp_rows = L->warp_rows; a.row_cap = L->row_cap;
    a.beta = beta; a.lambda = lambda; a.tol = tol;
    a.rp = I->row_ptr; a.ci = I->col; a.rev = L->rev; a.prog = L->prog; a.prog_ptr = L->prog_ptr;
    a.val = I->val; a.h = I->h; a.eps = L->eps; a.mstar = L->mstar;
    a.h_field = h_field_dev;
    a.u0 = L->u[L->cur]; a.u1 = L->u[1 - L->cur];
    a.hm = L->hm; a.tj = L->tj; a.uin = L->uin; a.tot = L->tot; a.marg = L->marg; a.offedge = L->offedge;
    a.red = L->red; a.iter_out = L->iter_out;
    void *args[] = {&a};
    NLMC_CUDA(cudaLaunchCooperativeKernel((void *)lbp_kernel, dim3((unsigned)L->grid), dim3(256), args, L->smem_bytes, I->stream));
    int res[2] = {0, 0};
    NLMC_CUDA(cudaMemcpyAsync(res, L->iter_out, sizeof(res), cudaMemcpyDeviceToHost, I->stream));
    if (out_marginal)
        NLMC_CUDA(cudaMemcpyAsync(out_marginal, L->marg, sizeof(double) * (size_t)I->n, cudaMemcpyDeviceToHost, I->stream));
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    *out_iteration = res[0];
    if (res[1] == 1) L->cur = 1 - L->cur;  // the newest messages ended up in the other buffer
    return NLMC_OK;
}

int nlmc_lbp_step(nlmc_lbp *L, double lambda, double beta, double tol, int max_iter, double *out_marginal,
                  int *out_iteration) {
    NLMC_REQUIRE(L && out_iteration, "nlmc_lbp_step: NULL argument");
    NLMC_REQUIRE(max_iter >= 1, "nlmc_lbp_step: max_iterations must be >= 1");
    NLMC_CUDA(cudaSetDevice(L->inst->device));
    return lbp_launch(L, nullptr, lambda, beta, tol, max_iter, out_marginal, out_iteration);
}

int nlmc_lbp_run(nlmc_lbp *L, const double *h_field, double beta, double tol, int max_iter, double *out_marginal,
                 int *out_iteration) {
    NLMC_REQUIRE(L && h_field && out_iteration, "nlmc_lbp_run: NULL argument");
    NLMC_REQUIRE(max_iter >= 1, "nlmc_lbp_run: max_iterations must be >= 1");
    nlmc_instance *I = L->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    NLMC_CUDA(cudaMemcpyAsync(L->hfield, h_field, sizeof(double) * (size_t)I->n, cudaMemcpyHostToDevice, I->stream));
    return lbp_launch(L, L->hfield, 0.0, beta, tol, max_iter, out_marginal, out_iteration);
}

int nlmc_lbp_set_messages(nlmc_lbp *L, const double *h_edge, const double *u_edge, const double *tot) {
    NLMC_REQUIRE(L && h_edge && u_edge && tot, "nlmc_lbp_set_messages: NULL argument");
    nlmc_instance *I = L->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    const size_t nz = sizeof(double) * (size_t)I->nnz;
    NLMC_CUDA(cudaMemcpyAsync(L->hm, h_edge, nz, cudaMemcpyHostToDevice, I->stream));
    NLMC_CUDA(cudaMemcpyAsync(L->u[L->cur], u_edge, nz, cudaMemcpyHostToDevice, I->stream));
    NLMC_CUDA(cudaMemcpyAsync(L->tot, tot, sizeof(double) * (size_t)I->n, cudaMemcpyHostToDevice, I->stream));
    if (I->nnz > 0) {
        nlmc::lbp_transpose_kernel<<<(I->nnz + 255) / 256, 256, 0, I->stream>>>(I->nnz, L->rev, L->u[L->cur], L->uin);
        NLMC_CUDA(cudaGetLastError());
    }
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    return NLMC_OK;
}

int nlmc_lbp_get_messages(nlmc_lbp *L, double *out_h_edge, double *out_u_edge, double *out_tot) {
    NLMC_REQUIRE(L, "nlmc_lbp_get_messages: NULL argument");
    nlmc_instance *I = L->inst;
    NLMC_CUDA(cudaSetDevice(I->device));
    const size_t nz = sizeof(double) * (size_t)I->nnz;
    if (out_h_edge) NLMC_CUDA(cudaMemcpyAsync(out_h_edge, L->hm, nz, cudaMemcpyDeviceToHost, I->stream));
    if (out_u_edge) NLMC_CUDA(cudaMemcpyAsync(out_u_edge, L->u[L->cur], nz, cudaMemcpyDeviceToHost, I->stream));
    if (out_tot) NLMC_CUDA(cudaMemcpyAsync(out_tot, L->tot, sizeof(double) * (size_t)I->n, cudaMemcpyDeviceToHost, I->stream));
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    return NLMC_OK;
}

int nlmc_lbp_byproducts(nlmc_lbp *L, double beta, double *out_corr, double *out_h_tilde, double *out_J_tilde) {
    using namespace nlmc;
    NLMC_REQUIRE(L, "nlmc_lbp_byproducts: NULL argument");
    nlmc_instance *I = L->inst;
    const int n = I->n;
    NLMC_CUDA(cudaSetDevice(I->device));
    const size_t dense_bytes = sizeof(double) * (size_t)n * (size_t)n;
    double *d_corr = nullptr, *d_jt = nullptr;
    if (out_corr) {
        if (!L->dense[0]) NLMC_CUDA(cudaMalloc(&L->dense[0], dense_bytes));
        d_corr = L->dense[0];
    }
    if (out_J_tilde) {
        if (!L->dense[1]) NLMC_CUDA(cudaMalloc(&L->dense[1], dense_bytes));
        d_jt = L->dense[1];
    }
    if ((d_corr || d_jt) && n > 0) {
        lbp_byproducts_fill_kernel<<<dim3((unsigned)n, (unsigned)((n + 255) / 256)), 256, 0, I->stream>>>(
            n, beta, L->tot, d_corr, d_jt);
        NLMC_CUDA(cudaGetLastError());
        lbp_byproducts_edges_kernel<<<(unsigned)n, 64, 0, I->stream>>>(n, I->nnz, beta, I->row_ptr, I->col, L->rev,
                                                                       I->val, L->hm, d_corr, d_jt);
        NLMC_CUDA(cudaGetLastError());
        if (out_corr) NLMC_CUDA(cudaMemcpyAsync(out_corr, d_corr, dense_bytes, cudaMemcpyDeviceToHost, I->stream));
        if (out_J_tilde) NLMC_CUDA(cudaMemcpyAsync(out_J_tilde, d_jt, dense_bytes, cudaMemcpyDeviceToHost, I->stream));
    }
    if (out_h_tilde && n > 0) {
        lbp_htilde_kernel<<<(n + 255) / 256, 256, 0, I->stream>>>(n, beta, L->marg, L->htilde);
        NLMC_CUDA(cudaGetLastError());
        NLMC_CUDA(cudaMemcpyAsync(out_h_tilde, L->htilde, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, I->stream));
    }
    NLMC_CUDA(cudaStreamSynchronize(I->stream));
    return NLMC_OK;
}

}  // extern "C"
