/*
 * nlmc_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).
 *
 * Plain-C restatement of the Monte Carlo hot path of usra-riacs/Nonlocal-Monte-Carlo.  It is
 * the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs) and is never linked, imported or executed by the product path.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the live
 * reference (imported from /root/reference in the build container by oracle/ref_loader.py)
 * and against the golden vectors generated from it (tests/golden/, oracle/make_golden.py).
 *
 * Each function cites the reference file:line it follows.  The one deliberate difference:
 * the reference recomputes the whole vector J.dot(m)+h for every attempt and then reads a
 * single entry x[kk] (NMC/nmc.py:86-87); the restatement computes only that entry, with the
 * same CSR accumulation order scipy's csr_matvec uses (sequential over the stored entries of
 * row kk, starting from 0, then "+ h[kk]"), so the value is bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* np.sign semantics for a finite double: -1, 0, +1 */
static inline int8_t sign_i8(double v) { return (int8_t)((v > 0.0) - (v < 0.0)); }

/*
 * Heat-bath sweeps with an injected random stream.
 * Follows MCMC: NMC/nmc.py:28-91 == NPT/npt.py:47-110 (J,h passed in; anneal handled by the
 * caller through beta_run[]), NPT/apt_preprocessor.py:33-74 == NPT/apt_ICM.py:52-93.
 *
 *   for jj in sweeps:  for kk in permutation(N):                     nmc.py:62,71
 *       x = J.dot(m) + h ;  m[kk] = sign(tanh(beta_run[jj]*x[kk]) - 2*rand() + 1)   nmc.py:86-87
 *       M[:, jj] = m                                                  nmc.py:89
 *
 * n         number of spins
 * rp,ci,val CSR of the matrix the reference passes to MCMC (already row-scaled by the caller
 *           for the NMC backbone phase, nmc.py:379), entries in scipy csr_matrix(J) order
 * h_eff     the h vector the reference passes to MCMC (may hold the +-1e4 freeze, nmc.py:381)
 * beta_run  inverse temperature of each sweep (nmc.py:64-69)
 * perm,u    per sweep the permutation(N) and the N rand() values, in draw order
 * m         spin state in {-1,0,+1}, updated in place
 * M_out     optional [n_sweeps][n] record of the state after each sweep
 * tanh_lut  optional: when lut_w > 0 and the row sum is an exact integer f with |f| <= lut_half and
 *           h_eff[k] == 0, tanh(beta*x) is read from tanh_lut[jj*lut_w + f + lut_half] (values
 *           produced by numpy's own tanh, so the decision is bit-equal to the reference even
 *           where libm's tanh differs from numpy's in the last place).
 */
int nlmc_oracle_mcmc(int n, const int32_t *rp, const int32_t *ci, const double *val,
                     const double *h_eff, int n_sweeps, const double *beta_run,
                     const int32_t *perm, const double *u, int8_t *m, int8_t *M_out,
                     const double *tanh_lut, int lut_half)
{
    const int lut_w = tanh_lut ? 2 * lut_half + 1 : 0;
    for (int jj = 0; jj < n_sweeps; ++jj) {
        const double beta = beta_run[jj];
        const int32_t *pj = perm + (size_t)jj * n;
        const double *uj = u + (size_t)jj * n;
        for (int a = 0; a < n; ++a) {
            const int k = pj[a];
            double x = 0.0;
            for (int p = rp[k]; p < rp[k + 1]; ++p) x += val[p] * (double)m[ci[p]];
            const double rowsum = x;
            x += h_eff[k];
            double t;
            if (lut_w && h_eff[k] == 0.0 && rowsum == floor(rowsum) && fabs(rowsum) <= lut_half)
                t = tanh_lut[(size_t)jj * lut_w + (int)rowsum + lut_half];
            else
                t = tanh(beta * x);
            m[k] = sign_i8(t - 2.0 * uj[a] + 1.0);
        }
        if (M_out) memcpy(M_out + (size_t)jj * n, m, (size_t)n);
    }
    return 0;
}

/*
 * Energies E = -(m^T J m / 2 + m^T h) of recorded states.
 * Follows NMC/nmc.py:386-387,496; NPT/npt.py:40-43,657-658; NPT/apt_preprocessor.py:107-110;
 * NPT/apt_ICM.py:45-49,262-263.  The reference goes through dense BLAS; for +-J instances every
 * partial sum is an exact integer so the order is immaterial (bit-exact); for real-valued J the
 * agreement is to rounding (north_star tolerance 1e-9 relative).
 * M is [n_cols][n] int8.
 */
int nlmc_oracle_energy(int n, const int32_t *rp, const int32_t *ci, const double *val,
                       const double *h, int n_cols, const int8_t *M, double *E)
{
    for (int c = 0; c < n_cols; ++c) {
        const int8_t *m = M + (size_t)c * n;
        double quad = 0.0, lin = 0.0;
        for (int k = 0; k < n; ++k) {
            double x = 0.0;
            for (int p = rp[k]; p < rp[k + 1]; ++p) x += val[p] * (double)m[ci[p]];
            quad += (double)m[k] * x;
            lin += (double)m[k] * h[k];
        }
        E[c] = -(quad / 2.0 + lin);
    }
    return 0;
}

/*
 * Houdayer disagreement clusters.  Follows find_disagreement_clusters, NPT/apt_ICM.py:116-143:
 * connected components of the subgraph induced on {i : s1[i]*s2[i] == -1} with adjacency J != 0,
 * listed in order of their smallest site index (the outer loop visits differing spins in
 * increasing order and starts a cluster at every one not yet visited, apt_ICM.py:122-124).
 * labels[i] = cluster ordinal (0-based, in that order) or -1 where the states agree.
 * Returns the number of clusters.
 */
int nlmc_oracle_disagreement_clusters(int n, const int32_t *rp, const int32_t *ci, const double *val,
                                      const int8_t *s1, const int8_t *s2, int32_t *labels)
{
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int n_clusters = 0;
    for (int i = 0; i < n; ++i) labels[i] = -1;
    for (int s = 0; s < n; ++s) {
        if ((int)s1[s] * (int)s2[s] != -1 || labels[s] >= 0) continue;
        int head = 0, tail = 0;
        queue[tail++] = s;
        labels[s] = n_clusters;
        while (head < tail) {
            const int cur = queue[head++];
            for (int p = rp[cur]; p < rp[cur + 1]; ++p) {
                const int j = ci[p];
                if (val[p] == 0.0) continue; /* dense rows enumerate val != 0, apt_ICM.py:129 */
                if ((int)s1[j] * (int)s2[j] != -1 || labels[j] >= 0) continue;
                labels[j] = n_clusters;
                queue[tail++] = j;
            }
        }
        ++n_clusters;
    }
    free(queue);
    return n_clusters;
}

/*
 * Many independent replicas of nlmc_oracle_mcmc on the host cores (the reference's own
 * parallelism: one task per replica, NPT/npt.py:622-638).  Used by bench.py's CPU baseline.
 * perm/u/m/beta_run are laid out replica-major.  Returns the thread count used.
 */
struct many_job {
    int n_rep, n, n_sweeps;
    const int32_t *rp, *ci, *perm;
    const double *val, *h_eff, *beta_run, *u;
    int8_t *m;
    int next; /* guarded by lock */
    pthread_mutex_t lock;
};

static void *many_worker(void *arg)
{
    struct many_job *j = (struct many_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->lock);
        const int r = j->next++;
        pthread_mutex_unlock(&j->lock);
        if (r >= j->n_rep) break;
        nlmc_oracle_mcmc(j->n, j->rp, j->ci, j->val, j->h_eff, j->n_sweeps,
                         j->beta_run + (size_t)r * j->n_sweeps,
                         j->perm + (size_t)r * j->n_sweeps * j->n, j->u + (size_t)r * j->n_sweeps * j->n,
                         j->m + (size_t)r * j->n, NULL, NULL, 0);
    }
    return NULL;
}

int nlmc_oracle_mcmc_many(int n_rep, int n, const int32_t *rp, const int32_t *ci, const double *val,
                          const double *h_eff, int n_sweeps, const double *beta_run,
                          const int32_t *perm, const double *u, int8_t *m, int n_threads)
{
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > n_rep) n_threads = n_rep;
    if (n_threads < 1) n_threads = 1;
    struct many_job job = {n_rep, n, n_sweeps, rp, ci, perm, val, h_eff, beta_run, u, m, 0,
                           PTHREAD_MUTEX_INITIALIZER};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, many_worker, &job);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    return n_threads;
}

/* ------------------------------------------------------------------------------------------
 * Loopy belief propagation (one lambda step).  Follows LoopyBeliefPropagation,
 * NMC/nmc.py:168-228 == NPT/npt.py:204-264, restated on the edges of J only.
 *
 * The reference holds dense N x N message matrices.  Off the edges they are trivial:
 *   u_msgs[i,j] = atanh_sat(tanh(beta*0) * ...) / beta = 0                      nmc.py:205
 *   h_msgs[i,j] = total_i - u_msgs[j,i] = total_i   (j != i, J_ij == 0)         nmc.py:201-203
 * so the restatement keeps one value per stored entry of J (u[p], hm[p] for entry p = (i,j)),
 * plus tot[i] = the common off-edge value of row i of h_msgs, which still takes part in the
 * reference's convergence maxima (nmc.py:208-209) whenever row i has an off-diagonal zero.
 *
 * Summation orders are numpy's (probed on numpy 2.3.5): np.sum(u_msgs[:, i]) (nmc.py:201) is a
 * pairwise sum over all N entries of the strided column; np.sum(u_msgs, axis=0) (nmc.py:216)
 * accumulates rows sequentially.  Zeros do not change a partial sum, so only the stored
 * entries are visited, in the same association order.
 * ------------------------------------------------------------------------------------------ */

/* numpy DOUBLE_pairwise_sum over a dense vector a[lo..lo+n) whose only non-zeros are the
 * (sorted) positions pos[*cur..] with values v[..]; consumes the entries it covers. */
static double pairwise_sparse(int lo, int n, const int32_t *pos, const double *v, int cnt, int *cur)
{
    if (n < 8) {
        double res = 0.0;
        while (*cur < cnt && pos[*cur] < lo + n) { res += v[*cur]; ++*cur; }
        return res;
    }
    if (n <= 128) {
        double r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int body_end = lo + n - (n % 8);
        while (*cur < cnt && pos[*cur] < body_end) { r[(pos[*cur] - lo) & 7] += v[*cur]; ++*cur; }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        while (*cur < cnt && pos[*cur] < lo + n) { res += v[*cur]; ++*cur; }
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double a = pairwise_sparse(lo, n2, pos, v, cnt, cur);
    const double b = pairwise_sparse(lo + n2, n - n2, pos, v, cnt, cur);
    return a + b;
}

/*
 * Message-gathering half of one LBP iteration (nmc.py:200-203):
 *   total_i = hl[i] + sum_k u[k,i] ;  hm[i,j] = total_i - u[j,i] ;  hm[i,i] = 0 ;  tot[i] = total_i
 * The transcendental half (nmc.py:205) is done by the Python driver with numpy's own tanh/arctanh:
 * the reference's convergence test uses tolerance = machine epsilon, i.e. it waits for an exact
 * floating-point fixed point, and whether one is reached depends on the last bit of those
 * functions (probed: libm vs numpy flips "converged at iteration 15" into "never").
 * rev[p] = index of the stored entry (j,i) for entry p = (i,j).
 */
int nlmc_oracle_lbp_gather(int n, const int32_t *rp, const int32_t *ci, const int32_t *rev,
                           const double *hl, const double *u, double *hm, double *tot)
{
    int maxdeg = 0;
    for (int i = 0; i < n; ++i) if (rp[i + 1] - rp[i] > maxdeg) maxdeg = rp[i + 1] - rp[i];
    double *colv = (double *)malloc(sizeof(double) * (size_t)(maxdeg + 1));
    for (int i = 0; i < n; ++i) {
        const int b = rp[i], cnt = rp[i + 1] - rp[i];
        for (int q = 0; q < cnt; ++q) colv[q] = u[rev[b + q]]; /* u_msgs[k,i], k ascending */
        int cur = 0;
        const double total = hl[i] + pairwise_sparse(0, n, ci + b, colv, cnt, &cur);
        for (int q = 0; q < cnt; ++q) hm[b + q] = (ci[b + q] == i) ? 0.0 : total - colv[q];
        tot[i] = total;
    }
    free(colv);
    return 0;
}

/* column sums of u in row order: acc_i = ((u[k0,i] + u[k1,i]) + ...) (np.sum(u_msgs, axis=0), nmc.py:216) */
int nlmc_oracle_lbp_colsum(int n, const int32_t *rp, const int32_t *rev, const double *u, double *acc)
{
    for (int i = 0; i < n; ++i) {
        double a = 0.0;
        for (int p = rp[i]; p < rp[i + 1]; ++p) a += u[rev[p]];
        acc[i] = a;
    }
    return 0;
}
