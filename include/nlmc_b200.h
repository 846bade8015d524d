/*
 * nlmc_b200.h -- C ABI of libnlmc_b200.so, the sm_100a CUDA implementation of the Monte Carlo hot
 * path of usra-riacs/Nonlocal-Monte-Carlo.
 *
 * The reference has no FFI: its boundary is the Python class API (NMC.run, NPT.run,
 * APT_preprocessor.run, APT_ICM.run).  The drop-in classes in nonlocal-monte-carlo_b200/nlmc_b200/
 * keep that API and bind the entry points below with ctypes (INTEGRATION.md shows the stub a
 * maintainer of the reference would add).  Each entry point cites the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer argument is HOST memory owned by the caller
 *     unless its name ends in _dev;
 *   - all functions return 0 on success and a negative code on failure; nlmc_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - handles are opaque; one host thread per handle; the library owns device memory and streams;
 *   - spins are int8 in {-1, 0, +1} (np.sign can return 0, NMC/nmc.py:87); J and h are the
 *     NORMALISED values the reference uses inside run() (J / max|J|, NMC/nmc.py:474-476);
 *   - there is no CPU fallback: every call fails with NLMC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NLMC_B200_H
#define NLMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLMC_OK 0
#define NLMC_ERR_ARG (-1)
#define NLMC_ERR_CUDA (-2)
#define NLMC_ERR_UNSUPPORTED (-3)
#define NLMC_ERR_STATE (-4)

typedef struct nlmc_instance nlmc_instance; /* one Ising instance (J in CSR, h) resident in HBM */
typedef struct nlmc_replicas nlmc_replicas; /* R int8 spin configurations of one instance (exact path) */
typedef struct nlmc_msc nlmc_msc;           /* bit-packed replica lattice for the production path  */
typedef struct nlmc_dense nlmc_dense;       /* dense-J production path (tensor-core field contraction) */
typedef struct nlmc_col nlmc_col;           /* sparse production path: graph-coloured updates, one CTA per replica */

const char *nlmc_last_error(void);
int nlmc_version(void);
int nlmc_device_count(void);
/* name / SM count / free+total bytes of a device; any out pointer may be NULL */
int nlmc_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor,
                     uint64_t *free_bytes, uint64_t *total_bytes);

/* host-side format helper: int8 spins -> the float64 arrays the reference's API returns (NMC/nmc.py:52,89), multi-threaded */
int nlmc_host_widen_i8_f64(const int8_t *in, double *out, uint64_t count, int threads);
/* touch every page of a freshly allocated host result buffer (first-touch faults taken while the GPU is busy) */
int nlmc_host_prefault(void *buf, uint64_t bytes, int threads);
/* Device int8 spins -> host float64 in blocks: block b of dev_src (block_elems values each) lands, widened, at
 * host_dst + dst_block[b] * block_elems (dst_block NULL = identity).  The copy goes through a pinned staging buffer of the
 * library in chunks, widened by the host workers while the next chunk is on the wire; `cuda_stream` (NULL = default stream)
 * orders it after the kernels that produced dev_src.  This is how the recorded states of a run -- the reference's M,
 * float64 (R*N) x sweeps, NPT/npt.py:640-644 -- reach the caller, in temperature order when exchanges permuted labels. */
int nlmc_host_fetch_widen_blocks(const int8_t *dev_src, double *host_dst, int n_blocks, uint64_t block_elems,
                                 const int32_t *dst_block, int device, void *cuda_stream);

/* ---- instance ---------------------------------------------------------------------------------
 * Replaces `J = csr_matrix(J)` + `h = asarray(h)` at the top of every MCMC call (NMC/nmc.py:53-54):
 * the CSR is uploaded once, entries in scipy's csr_matrix(J) order. */
int nlmc_instance_create(int n, const int32_t *row_ptr, const int32_t *col, const double *val,
                         const double *h, int device, nlmc_instance **out);
int nlmc_instance_destroy(nlmc_instance *inst);
int nlmc_instance_n(const nlmc_instance *inst);
/* 1 when every stored J value is an integer (then row sums are exact in any order) */
int nlmc_instance_is_integer(const nlmc_instance *inst);
/* 1 when J_ij == J_ji for every stored entry and no row repeats a column: required by the production engines
 * (nlmc_col_create / nlmc_dense_create fail with NLMC_ERR_ARG otherwise); the replay path accepts any J, as the
 * reference does (J.dot(m), NMC/nmc.py:86), and only takes its incremental-field kernel when this holds */
int nlmc_instance_is_symmetric(const nlmc_instance *inst);

/* ---- replicas (exact path) ---------------------------------------------------------------- */
int nlmc_replicas_create(nlmc_instance *inst, int n_replicas, const int8_t *init_spins /*[R][n] or NULL*/,
                         nlmc_replicas **out);
int nlmc_replicas_destroy(nlmc_replicas *reps);
int nlmc_set_spins(nlmc_replicas *reps, int first, int count, const int8_t *spins /*[count][n]*/);
int nlmc_get_spins(nlmc_replicas *reps, int first, int count, int8_t *out /*[count][n]*/);

/* The (J, h) pair the reference hands to MCMC for replica r during an NMC phase
 * (NMC/nmc.py:377-385,398-406; NPT/npt.py:406-414,425,441):
 *   h_eff       the h vector of the phase (h/temp_x on the backbone, +-1e4*m_init on frozen spins), or
 *               NULL for the instance's own h;
 *   row_scaled  row_scaled[k] != 0 <=> row k of J is divided by temp_x (J_c[all_clusters,:] / temp_x),
 *               or NULL for no scaling.
 * The setting persists until changed. */
int nlmc_set_phase(nlmc_replicas *reps, int r, const double *h_eff, const uint8_t *row_scaled, double temp_x);

/* K1 sweep_replay -- MCMC with an injected random stream.
 * Replaces MCMC: NMC/nmc.py:28-91 == NPT/npt.py:47-110, NPT/apt_preprocessor.py:33-74 ==
 * NPT/apt_ICM.py:52-93.  For replica r and sweep s the kernel visits perm[r][s][0..n) in order and
 * sets m[k] = sign(tanh(beta[r][s] * (sum_j J_kj m_j + h_k)) - 2*u[r][s][a] + 1), the row sum
 * accumulated in CSR order exactly as scipy's csr_matvec does.
 *   perm, u     [R][n_sweeps][n]   one np.random.permutation(n) and n np.random.rand() per sweep
 *   beta        [R][n_sweeps]      beta_run of NMC/nmc.py:56-69
 *   tanh_lut    optional [R][n_sweeps][2*lut_half+1]: tanh(beta*f) for integer fields f, computed by the
 *               caller with numpy's tanh; used where the row is unscaled, J is integer and h_eff[k]==0
 *               so that the decision is bit-equal to the reference's
 *   out_M       optional [R][n_sweeps-record_from][n]: state after every sweep >= record_from
 *               (the reference's M[:, jj] = m, NMC/nmc.py:89)
 *   out_E       optional [R][n_sweeps]: E = -(m^T J m/2 + m^T h) after every sweep with the instance's
 *               own J and h (NMC/nmc.py:386-387, NPT/npt.py:40-43, NPT/apt_preprocessor.py:107-110) */
int nlmc_sweep_replay(nlmc_replicas *reps, int n_sweeps, const int32_t *perm, const double *u,
                      const double *beta, const double *tanh_lut, int lut_half,
                      int8_t *out_M, int record_from, double *out_E);

/* K4 energy_csr -- E = -(m^T J m / 2 + m^T h).
 * Replaces the energy loops NMC/nmc.py:386-387,496; NPT/npt.py:31-45,657-658;
 * NPT/apt_preprocessor.py:107-110; NPT/apt_ICM.py:36-50,262-263. */
int nlmc_energy(nlmc_replicas *reps, double *out_E /*[R]*/);
int nlmc_energy_states(nlmc_instance *inst, int n_states, const int8_t *states /*[n_states][n]*/,
                       double *out_E /*[n_states]*/);

/* ---- numpy-equivalent float64 tanh / arctanh ------------------------------------------------------
 * out[i] = np.tanh(x[i]) / np.arctanh(x[i]) evaluated on the device by the functions the LBP and replay
 * kernels use (csrc/nlmc_npmath.h: numpy's simd_tanh_f64 and the SVML routine behind np.arctanh, restated
 * operation for operation).  Replaces the np.tanh / np.arctanh calls of NMC/nmc.py:87,205,216,252. */
int nlmc_np_tanh(const double *x, double *out, int64_t n, int device);
int nlmc_np_arctanh(const double *x, double *out, int64_t n, int device);

/* ---- K5 lbp -- loopy belief propagation of the NMC backbone search -----------------------------
 * Replaces LoopyBeliefPropagation (NMC/nmc.py:168-228 == NPT/npt.py:204-264) as called by
 * LBP_convexified (NMC/nmc.py:93-166): messages live on the stored entries of J and are warm-started
 * from one lambda to the next.  The lambda schedule, the divergence rules (nmc.py:142-161) and
 * find_clusters (nmc.py:257-318) stay on the host (nlmc_b200/nmc_core.py): they are a few scalar
 * comparisons and set operations per call.
 *   nlmc_lbp_create  requires a symmetric sparsity pattern and rows sorted by column (numpy sums the reference's dense
 *                    rows and columns in index order; csr_matrix(dense J) delivers exactly that)
 *   nlmc_lbp_reset   u_msgs = J * m_star, h_msgs = 0                       (nmc.py:128-129)
 *   nlmc_lbp_epsilon epsilon_i = |h_i| + sum_j |J_ij|                      (nmc.py:353)
 *   nlmc_lbp_step    one LoopyBeliefPropagation call with h + lambda*m_star*epsilon (nmc.py:133-139);
 *                    out_marginal[n] = tanh(beta*(h_lambda + sum_k u[k,i])) (nmc.py:216),
 *                    *out_iteration = the reference's `iteration` on exit (max_iter-1 <=> "diverged")
 *   nlmc_lbp_run     the same call with a caller-supplied field h[n] (the public method
 *                    LoopyBeliefPropagation(J, h, beta, h_msgs, u_msgs, tolerance, max_iterations), nmc.py:168)
 *   nlmc_lbp_set_messages / nlmc_lbp_get_messages
 *                    h_msgs / u_msgs on the stored entries of J (CSR entry order) plus tot[i], the common value
 *                    of row i of h_msgs off the stored entries (the reference's dense matrices are exactly this:
 *                    u_msgs = 0 and h_msgs[i,j] = tot[i] off the entries, h_msgs[i,i] = 0)
 *   nlmc_lbp_byproducts
 *                    correlations, h_tilde, J_tilde of the last call (nmc.py:217-226); dense [n][n] row-major like
 *                    the reference's return values; any pointer may be NULL */
typedef struct nlmc_lbp nlmc_lbp;
int nlmc_lbp_create(nlmc_instance *inst, nlmc_lbp **out);
int nlmc_lbp_destroy(nlmc_lbp *lbp);
int nlmc_lbp_epsilon(nlmc_lbp *lbp, double *out_eps /*[n]*/);
int nlmc_lbp_reset(nlmc_lbp *lbp, const double *m_star /*[n]*/);
int nlmc_lbp_step(nlmc_lbp *lbp, double lambda, double beta, double tol, int max_iter,
                  double *out_marginal /*[n] or NULL*/, int *out_iteration);
int nlmc_lbp_run(nlmc_lbp *lbp, const double *h_field /*[n]*/, double beta, double tol, int max_iter,
                 double *out_marginal /*[n] or NULL*/, int *out_iteration);
int nlmc_lbp_set_messages(nlmc_lbp *lbp, const double *h_edge /*[nnz]*/, const double *u_edge /*[nnz]*/,
                          const double *tot /*[n]*/);
int nlmc_lbp_get_messages(nlmc_lbp *lbp, double *out_h_edge /*[nnz] or NULL*/, double *out_u_edge /*[nnz] or NULL*/,
                          double *out_tot /*[n] or NULL*/);
int nlmc_lbp_byproducts(nlmc_lbp *lbp, double beta, double *out_corr /*[n*n] or NULL*/,
                        double *out_h_tilde /*[n] or NULL*/, double *out_J_tilde /*[n*n] or NULL*/);

/* ---- K7 icm_components -- Houdayer iso-cluster identification ---------------------------------
 * Replaces find_disagreement_clusters (NPT/apt_ICM.py:116-143) for n_pairs pairs of states at once:
 * connected components of the subgraph induced on {i : s1[i]*s2[i] == -1}, adjacency J != 0.
 * out_labels[p][i] = position of i's cluster in the reference's `clusters` list (clusters ordered by
 * their smallest site index) or -1 where the states agree; out_n_clusters[p] = len(clusters). */
int nlmc_icm_clusters(nlmc_instance *inst, int n_pairs, const int8_t *s1 /*[n_pairs][n]*/,
                      const int8_t *s2 /*[n_pairs][n]*/, int32_t *out_labels /*[n_pairs][n]*/,
                      int32_t *out_n_clusters /*[n_pairs]*/);

/* ---- K2 / K4' / K6: production path (bit-packed multi-spin coding) ------------------------------
 * For +-J instances with h = 0 and degrees <= 6 (periodic or open lattices, Chimera-like graphs; the 3D EA
 * configs).  n_ladders independent NPT runs ("ladders", one replica per beta each) are packed 32 to a word, all bits
 * of a word at the same beta; n_ladders is rounded up to a multiple of 128 (nlmc_msc_info reports the padded count).
 * Packed states cross this boundary site-major, packed[site][n_words] (word = slot * G + ladder group); on the device
 * they are kept quad-major in colour order.
 *
 *   nlmc_msc_sweep      heat-bath sweeps, graph-coloured parallel updates, Philox4x32-10 randoms.
 *                       Replaces MCMC (NMC/nmc.py:28-91 and copies) for every replica of every ladder;
 *                       same single-site conditional distribution, different (coloured) visiting order.
 *   nlmc_msc_energies   E = -(m^T J m/2) of every replica, exact integers (NPT/npt.py:31-45,657-658).
 *                       out_E [n_beta][n_ladders_padded] or NULL to leave them on the device.
 *   nlmc_msc_round      one swap round of NPT.run (NPT/npt.py:617-680) for all ladders: n_sweeps sweeps,
 *                       energies, then per ladder num_swapping_pairs non-overlapping adjacent pairs
 *                       (NPT/npt.py:514-533) accepted with min(1, exp(dBeta*dE)) (npt.py:668-671) and
 *                       the two configurations exchanged (npt.py:677-678).  out_E (optional) receives the
 *                       energies of the states before the exchange.
 *   nlmc_msc_round_host the same through HOST buffers: packed states in ([n][n_words] uint32, NULL = keep
 *                       the device state), packed states and energies out (either may be NULL).
 *   nlmc_msc_round_host_async  the same without waiting: copies and kernels are queued on the handle's own stream
 *                       and the call returns; the outputs are valid after nlmc_msc_sync.  With pinned host buffers
 *                       and two handles used alternately, one batch's copies overlap the other's sweeps (the
 *                       reference ships m_start to its workers and M back every round, NPT/npt.py:625-644).
 *   nlmc_msc_set/get_spins     one replica (beta_idx, ladder) as int8 +-1.
 *   nlmc_msc_timer_*    CUDA-event timing on the handle's stream (mark 0 = start, 1 = stop). */
/* ladder_offset: global index of this handle's first ladder (a multiple of 128).  Every random stream is
 * keyed by (seed, beta index, GLOBAL ladder index, site, sweep), so a set of ladders evolves identically
 * whether it lives in one handle or is sharded over several handles / GPUs. */
int nlmc_msc_create(nlmc_instance *inst, int n_beta, const double *betas, int n_ladders, int ladder_offset,
                    unsigned long long seed, nlmc_msc **out);
int nlmc_msc_destroy(nlmc_msc *msc);
int nlmc_msc_info(const nlmc_msc *msc, int *n_words, int *n_ladders_padded, int *n_colours, long long *n_bonds);
int nlmc_msc_set_seed(nlmc_msc *msc, unsigned long long seed, unsigned sweep_counter);
/* new inverse temperatures for the existing words (APT_preprocessor walks its ladder one beta at a time,
 * warm-starting from the previous beta's final states, NPT/apt_preprocessor.py:154-166) */
int nlmc_msc_set_betas(nlmc_msc *msc, const double *betas /*[n_beta]*/);
int nlmc_msc_init_random(nlmc_msc *msc, unsigned stream_id);
int nlmc_msc_set_spins(nlmc_msc *msc, int beta_idx, int ladder, const int8_t *spins /*[n]*/);
int nlmc_msc_get_spins(nlmc_msc *msc, int beta_idx, int ladder, int8_t *out /*[n]*/);
int nlmc_msc_set_packed(nlmc_msc *msc, const uint32_t *packed /*[n][n_words]*/);
int nlmc_msc_get_packed(nlmc_msc *msc, uint32_t *packed /*[n][n_words]*/);
int nlmc_msc_sweep(nlmc_msc *msc, int n_sweeps);
int nlmc_msc_energies(nlmc_msc *msc, double *out_E);
/* n_sweeps sweeps, recording after every sweep the state of one ladder (out_M [n_sweeps][n_beta][n], the reference's
 * M[:, jj] = m, NMC/nmc.py:89) and/or the energies of all replicas (out_E [n_sweeps][n_beta][n_ladders_padded],
 * NPT/npt.py:40-43, NPT/apt_preprocessor.py:107-110) on the device; one copy back at the end */
int nlmc_msc_sweep_record(nlmc_msc *msc, int n_sweeps, int ladder, int8_t *out_M, double *out_E);
/* the same with the recorded states laid out as the rows of the reference's M: m_layout 1 = [n_beta][n][n_sweeps] */
int nlmc_msc_sweep_record_layout(nlmc_msc *msc, int n_sweeps, int ladder, int8_t *out_M, double *out_E, int m_layout);
/* the same with the states delivered as float64 rows of M ([n_beta][n][n_sweeps]) through a pinned staging buffer */
int nlmc_msc_sweep_record_f64(nlmc_msc *msc, int n_sweeps, int ladder, double *out_M_f64, double *out_E);
/* the same record left on the DEVICE (no synchronisation): *out_M_dev int8 in the layout m_layout selects, *out_E_dev
 * float64 [n_sweeps][n_beta][n_ladders] (NULL pointer argument = not recorded).  The buffers belong to the handle and
 * stay valid until its next record call; a rank of a sharded ladder all-gathers them from here. */
int nlmc_msc_sweep_record_dev(nlmc_msc *msc, int n_sweeps, int ladder, int m_layout, int8_t **out_M_dev, double **out_E_dev);
int nlmc_msc_round(nlmc_msc *msc, int n_sweeps, int num_swapping_pairs, double *out_E);
int nlmc_msc_round_host(nlmc_msc *msc, const uint32_t *packed_in, int n_sweeps, int num_swapping_pairs,
                        uint32_t *packed_out, double *out_E);
int nlmc_msc_round_host_async(nlmc_msc *msc, const uint32_t *packed_in, int n_sweeps, int num_swapping_pairs,
                              uint32_t *packed_out, double *out_E);
int nlmc_msc_swap_count(nlmc_msc *msc, int *out_accepted, int reset);
/* accepted exchanges of each of the last n_rounds (<= 4096) rounds, oldest first, counted per round on the device
 * (the reference's `count[ii]`, NPT/npt.py:664-680) */
int nlmc_msc_swap_counts(nlmc_msc *msc, int n_rounds, int *out_counts);

/* ---- replica exchange by beta labels, and a ladder's beta range sharded over GPUs (north_star 4) ----
 * Replaces the swap block NPT/npt.py:649-680 in the form SURVEY D4 describes: configurations never move; every
 * (slot, ladder) carries the index of the temperature it is simulated at, and an accepted exchange swaps two labels.
 * A handle owns the slots [slot_begin, slot_begin + slot_count) of a ladder of n_beta_total temperatures.  Random
 * streams are keyed by the GLOBAL slot and ladder indices, so a ladder sharded over several handles / GPUs evolves
 * bit for bit like the single handle that owns all its slots.
 *   slot_count == n_beta_total : self-contained, nlmc_msc_round does sweeps + energies + label exchange on the device.
 *   a block of a sharded ladder: per round  nlmc_msc_sweep -> nlmc_msc_energies_dev (into the caller's all-gather
 *       send buffer) -> [the caller all-gathers the energies of all blocks, 8 bytes per replica, e.g. ncclAllGather]
 *       -> nlmc_msc_exchange_labels (identical Philox-keyed decisions on every rank).
 *   nlmc_msc_set_stream  runs the handle on the caller's CUDA stream (the one the collective is ordered on);
 *   nlmc_msc_get_labels  labels[slot][ladder] (uint8, [n_beta_total][n_ladders_padded]) to reorder outputs to beta order. */
int nlmc_msc_create_labelled(nlmc_instance *inst, int n_beta_total, const double *betas_total, int slot_begin,
                             int slot_count, int n_ladders, int ladder_offset, unsigned long long seed, nlmc_msc **out);
int nlmc_msc_set_stream(nlmc_msc *msc, void *cuda_stream);
int nlmc_msc_energies_dev(nlmc_msc *msc, double *out_E_dev /*[slot_count][n_ladders_padded], device*/);
int nlmc_msc_exchange_labels(nlmc_msc *msc, const double *E_full_dev /*[n_beta_total][n_ladders_padded], device*/,
                             int num_swapping_pairs);
int nlmc_msc_get_labels(nlmc_msc *msc, uint8_t *out_labels);
int nlmc_msc_sync(nlmc_msc *msc);
int nlmc_msc_timer_mark(nlmc_msc *msc, int which);
int nlmc_msc_timer_elapsed_ms(nlmc_msc *msc, float *out_ms);

/* ---- K3: dense-J production path (config C3, Sherrington-Kirkpatrick) ---------------------------
 * The reference evaluates x = J.dot(m) + h for every attempt (NMC/nmc.py:86).  Here all replicas share J, so
 * the fields of a block of 128 sites for ALL replicas are one tensor-core GEMM H_blk = S . J[:, blk]
 * (tcgen05.mma, bf16 spins -- exact -- times J split into n_split bf16 pieces, fp32 accumulation in TMEM);
 * a sweep visits the blocks in order and updates the sites of a block sequentially per replica with the
 * in-block flips corrected on CUDA cores: a fixed-order sequential heat-bath sweep, Philox randoms.
 *   betas [n_replicas]: one inverse temperature per replica (PT exchanges swap these labels);
 *   nlmc_dense_fields    full recompute H = S . J (out_H [R][n], optional) -- 2*R*n^2*n_split flop;
 *   nlmc_dense_energies  E = -(m^T J m/2 + m^T h) from those fields (fp32 products, fp64 accumulation);
 *   nlmc_dense_time_*    CUDA-event timings of the GEMM / of whole sweeps (ms per call). */
int nlmc_dense_create(nlmc_instance *inst, int n_replicas, const double *betas, int n_split,
                      unsigned long long seed, nlmc_dense **out);
int nlmc_dense_destroy(nlmc_dense *d);
int nlmc_dense_set_betas(nlmc_dense *d, const double *betas /*[n_replicas]*/);
int nlmc_dense_set_spins(nlmc_dense *d, const int8_t *spins /*[n_replicas][n]*/);
int nlmc_dense_get_spins(nlmc_dense *d, int8_t *out /*[n_replicas][n]*/);
int nlmc_dense_fields(nlmc_dense *d, float *out_H /*[n_replicas][n] or NULL*/);
int nlmc_dense_sweep(nlmc_dense *d, int n_sweeps);
int nlmc_dense_energies(nlmc_dense *d, double *out_E /*[n_replicas]*/);
/* NMC phases on the dense path (NMC/nmc.py:377-385,398-406): modes [n_replicas][n], 0 = normal, 1 = backbone at
 * beta/temp_x (the reference divides the backbone rows of J and h by temp_x), 2 = frozen (the reference pins
 * the spin with h = +-1e4); NULL switches the modes off. */
int nlmc_dense_set_site_modes(nlmc_dense *d, const uint8_t *modes, double temp_x);
/* m_init = M[:, argmin E] bookkeeping on the device (first minimum wins, nmc.py:394-395) */
int nlmc_dense_best_reset(nlmc_dense *d);
int nlmc_dense_best_update(nlmc_dense *d, double *out_E /*[n_replicas] or NULL*/);
int nlmc_dense_best_get(nlmc_dense *d, int8_t *out_spins /*[n_replicas][n] or NULL*/, double *out_E /*or NULL*/);
int nlmc_dense_sync(nlmc_dense *d);
/* the same for the dense engine (see nlmc_col_ladders) */
int nlmc_dense_ladders(nlmc_dense *dense, int n_beta, const double *betas);
int nlmc_dense_exchange(nlmc_dense *dense, int num_swapping_pairs);
int nlmc_dense_labels(nlmc_dense *dense, int32_t *out_labels /*[R] or NULL*/, int n_rounds, int32_t *out_counts /*[n_rounds]*/);
int nlmc_dense_time_fields(nlmc_dense *d, int repeats, float *out_ms);
int nlmc_dense_time_sweeps(nlmc_dense *d, int n_sweeps, float *out_ms);

/* ---- K2a: sparse production path (any J, any h) -- graph-coloured parallel heat bath ---------------
 * Replaces MCMC (NMC/nmc.py:28-91 and copies) in production mode on arbitrary sparse instances: one CTA per
 * replica with spins, incrementally maintained local fields and (when it fits) the CSR in shared memory -- or, for
 * instances above ~22,000 spins, spins and fields in a global workspace (same kernel, any size);
 * sites of one colour are updated in parallel, colours in order; Philox4x32-10 keyed by
 * (seed; replica_offset + replica, site, sweep).
 *   nlmc_col_sweep  n_sweeps sweeps in ONE launch; beta_sched [n_sweeps][R] (optional) is the annealing schedule
 *                   beta_run (nmc.py:56-69); out_E [n_sweeps][R] (optional) the energy after every sweep
 *                   (nmc.py:386-387); out_spins [ceil(n_sweeps/record_every)][R][n] (optional) the recorded states
 *                   M[:, ::M_skip] (nmc.py:390); track_best keeps m_init = M[:, argmin E] (nmc.py:394-395) on the device.
 *   site modes      0 normal, 1 backbone at beta/temp_x, 2 frozen (NMC phases, nmc.py:377-385,398-406). */
int nlmc_col_create(nlmc_instance *inst, int n_replicas, const double *betas, int replica_offset,
                    unsigned long long seed, nlmc_col **out);
int nlmc_col_destroy(nlmc_col *c);
int nlmc_col_info(const nlmc_col *c, int *n_colours, int *csr_in_smem);
int nlmc_col_set_betas(nlmc_col *c, const double *betas /*[R]*/);
int nlmc_col_set_spins(nlmc_col *c, const int8_t *spins /*[R][n]*/);
int nlmc_col_get_spins(nlmc_col *c, int8_t *out /*[R][n]*/);
int nlmc_col_set_site_modes(nlmc_col *c, const uint8_t *modes /*[R][n] or NULL*/, double temp_x);
int nlmc_col_best_reset(nlmc_col *c);
int nlmc_col_best_get(nlmc_col *c, int8_t *out_spins /*[R][n] or NULL*/, double *out_E /*[R] or NULL*/);
int nlmc_col_sweep(nlmc_col *c, int n_sweeps, const double *beta_sched, int record_every, int8_t *out_spins,
                   double *out_E, int track_best);
int nlmc_col_energies(nlmc_col *c, double *out_E /*[R]*/);
int nlmc_col_sync(nlmc_col *c);
/* Replica exchange of the generic engines as a permutation of beta labels, entirely on the device (replaces the swap
 * block NPT/npt.py:649-680 and the pair selection NPT/npt.py:514-533 in the label form of SURVEY D4): the rows are grouped
 * into ladders, row = ladder * n_beta + slot, slot s starting at betas[s].
 *   *_ladders   declares the grouping, resets the labels to the identity and sets the per-row betas;
 *   *_exchange  energies of all rows (device), num_swapping_pairs non-overlapping adjacent temperature pairs per ladder,
 *               accepted with min(1, exp(dB*dE)); labels and per-row betas updated; queued on the handle's stream;
 *   *_labels    labels[row] = temperature index of the row (NULL = skip) and the accepted exchanges of each of the last
 *               n_rounds rounds (oldest first). */
int nlmc_col_ladders(nlmc_col *col, int n_beta, const double *betas);
int nlmc_col_exchange(nlmc_col *col, int num_swapping_pairs);
int nlmc_col_labels(nlmc_col *col, int32_t *out_labels /*[R] or NULL*/, int n_rounds, int32_t *out_counts /*[n_rounds]*/);

#ifdef __cplusplus
}
#endif
#endif /* NLMC_B200_H */
