// nlmc_icm.cu -- K7: Houdayer iso-cluster identification as a GPU connected-components kernel.
//
// Replaces find_disagreement_clusters (NPT/apt_ICM.py:116-143): the clusters are the connected
// components of the subgraph induced on the sites where the two states disagree
// (s1[i]*s2[i] == -1), adjacency J != 0, and the reference lists them in order of their
// smallest site index (its outer loop visits the differing spins in increasing order,
// apt_ICM.py:122-124).  The kernel labels every differing site with the smallest site index of
// its component (min-label propagation with pointer jumping), then ranks the roots, so
// out_labels[i] is exactly the position of i's cluster in the reference's `clusters` list
// (-1 where the states agree).  One CTA per replica pair; many pairs per launch.
#include "nlmc_common.cuh"

namespace nlmc {

constexpr int kIcmThreads = 512;

__global__ void __launch_bounds__(kIcmThreads) icm_components_kernel(
    int n, const int32_t *__restrict__ rp, const int32_t *__restrict__ ci, const double *__restrict__ val,
    const int8_t *__restrict__ s1_all, const int8_t *__restrict__ s2_all, int32_t *lab_all, int32_t *ord_all,
    int32_t *out_all, int32_t *n_clusters) {
    __shared__ int32_t s_scan[kIcmThreads];
    __shared__ int32_t s_total;
    const int pair = blockIdx.x;
    const int8_t *s1 = s1_all + (size_t)pair * n, *s2 = s2_all + (size_t)pair * n;
    volatile int32_t *lab = lab_all + (size_t)pair * n;
    int32_t *ord = ord_all + (size_t)pair * n;
    int32_t *out = out_all + (size_t)pair * n;
    const int t = threadIdx.x;

    for (int i = t; i < n; i += kIcmThreads) lab[i] = ((int)s1[i] * (int)s2[i] == -1) ? i : -1;
    __syncthreads();

    for (;;) {
        int changed = 0;
        // hook: take the smallest label in the closed neighbourhood
        for (int i = t; i < n; i += kIcmThreads) {
            const int mine = lab[i];
            if (mine < 0) continue;
            int best = mine;
            const int re = rp[i + 1];
            for (int p = rp[i]; p < re; ++p) {
                if (val[p] == 0.0) continue;
                const int lj = lab[ci[p]];
                if (lj >= 0 && lj < best) best = lj;
            }
            if (best < mine) {
                atomicMin((int32_t *)&lab[i], best);
                atomicMin((int32_t *)&lab[mine], best);  // pull the old root down too
                changed = 1;
            }
        }
        __syncthreads();
        // pointer jumping: labels always point to a smaller-or-equal index inside the component
        for (int i = t; i < n; i += kIcmThreads) {
            int l = lab[i];
            if (l < 0) continue;
            int ll = lab[l];
            while (ll != l) {
                l = ll;
                ll = lab[l];
            }
            lab[i] = l;
        }
        if (!__syncthreads_or(changed)) break;
    }

    // rank the roots (sites with lab[i] == i) in increasing index order
    const int chunk = (n + kIcmThreads - 1) / kIcmThreads;
    const int lo = min(n, t * chunk), hi = min(n, lo + chunk);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += (lab[i] == i);
    s_scan[t] = cnt;
    __syncthreads();
    for (int off = 1; off < kIcmThreads; off <<= 1) {  // inclusive Hillis-Steele scan
        const int v = (t >= off) ? s_scan[t - off] : 0;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    int base = s_scan[t] - cnt;
    if (t == kIcmThreads - 1) s_total = s_scan[t];
    for (int i = lo; i < hi; ++i)
        if (lab[i] == i) ord[i] = base++;
    __syncthreads();
    for (int i = t; i < n; i += kIcmThreads) {
        const int l = lab[i];
        out[i] = l >= 0 ? ord[l] : -1;
    }
    if (t == 0) n_clusters[pair] = s_total;
}

}  // namespace nlmc

extern "C" {

int nlmc_icm_clusters(nlmc_instance *I, int n_pairs, const int8_t *s1, const int8_t *s2,
                      int32_t *out_labels, int32_t *out_n_clusters) {
    using namespace nlmc;
    NLMC_REQUIRE(I && n_pairs >= 0, "nlmc_icm_clusters: bad arguments");
    if (n_pairs == 0) return NLMC_OK;
    NLMC_REQUIRE(s1 && s2 && out_labels && out_n_clusters, "nlmc_icm_clusters: NULL buffer");
    { const int rc_dev = nlmc::instance_device(I); if (rc_dev) return rc_dev; }   // the CSR on the device (uploaded on first use)
    NLMC_CUDA(cudaSetDevice(I->device));
    const size_t n = (size_t)I->n, P = (size_t)n_pairs;
    int8_t *d_s = nullptr;
    int32_t *d_i = nullptr;
    NLMC_CUDA(cudaMalloc(&d_s, 2 * P * n));
    if (cudaMalloc(&d_i, sizeof(int32_t) * (3 * P * n + P)) != cudaSuccess) {
        cudaFree(d_s);
        set_error("nlmc_icm_clusters: cudaMalloc failed");
        return NLMC_ERR_CUDA;
    }
    int32_t *d_lab = d_i, *d_ord = d_i + P * n, *d_out = d_i + 2 * P * n, *d_cnt = d_i + 3 * P * n;
    cudaStream_t st = I->stream;
    cudaError_t e = cudaMemcpyAsync(d_s, s1, P * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_s + P * n, s2, P * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        icm_components_kernel<<<(unsigned)P, kIcmThreads, 0, st>>>(I->n, I->row_ptr, I->col, I->val, d_s, d_s + P * n,
                                                                  d_lab, d_ord, d_out, d_cnt);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_labels, d_out, sizeof(int32_t) * P * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_n_clusters, d_cnt, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_s);
    cudaFree(d_i);
    if (e != cudaSuccess) {
        set_error("nlmc_icm_clusters: %s", cudaGetErrorString(e));
        return NLMC_ERR_CUDA;
    }
    return NLMC_OK;
}

}  // extern "C"
