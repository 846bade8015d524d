"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

Every fixture stores the instance, the call parameters, the seed given to ``np.random.seed`` /
``random.seed`` immediately before the call, and the reference's outputs.  ``num_cores=1``
everywhere (the only reproducible configuration of the reference).  The
generated files are small (spins are stored as int8) and are committed; the GPU box has no
/root/reference and reads only these files.

Cases are reduced-size versions of BASELINE.json's five configs (same graph families, same
call paths, same keyword arguments) plus element-level vectors and known-answer energies.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != _HERE]  # `oracle` must be the package
sys.path.insert(0, os.path.dirname(_HERE))
from oracle import oracle as O  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EPS = np.finfo(float).eps


def i8(a):
    a = np.asarray(a)
    assert np.all(a == np.round(a)) and np.all(np.abs(a) <= 1)
    return a.astype(np.int8)


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path)} bytes")


class record_backbones:
    """Records the clusters every LBP_convexified call of the reference returns (in call order), also
    from the forked pool worker of NPT.run: the wrapper appends to a file."""

    def __init__(self, cls):
        self.cls = cls

    def __enter__(self):
        import tempfile
        self.path = tempfile.mktemp(suffix=".backbones")
        orig = self.orig = self.cls.LBP_convexified
        path = self.path

        def wrapped(obj, *a, **k):
            out = orig(obj, *a, **k)
            cl = np.concatenate(out[0]).astype(int) if out[0] else np.array([], dtype=int)
            with open(path, "a") as f:
                f.write(" ".join(str(int(v)) for v in cl) + "\n")
            return out

        self.cls.LBP_convexified = wrapped
        return self

    def __exit__(self, *exc):
        self.cls.LBP_convexified = self.orig

    def result(self):
        """(flat indices, sizes) of all recorded backbones"""
        rows = []
        if os.path.exists(self.path):
            with open(self.path) as f:
                rows = [np.array([int(v) for v in line.split()], dtype=np.int64) for line in f]
            os.remove(self.path)
        flat = np.concatenate(rows) if rows else np.array([], dtype=np.int64)
        return flat, np.array([len(r) for r in rows], dtype=np.int64)


def case_mcmc():
    """Element level: MCMC (NMC/nmc.py:28-91) on +-J, Gaussian J with h, annealed and fixed beta."""
    J1, h1 = O.random_pm_graph(48, 0.2, 101)
    rs = np.random.RandomState(102)
    N2 = 24
    J2 = np.zeros((N2, N2))
    iu = np.triu_indices(N2, 1)
    J2[iu] = rs.randn(len(iu[0]))
    J2 += J2.T
    h2 = rs.randn(N2)
    out = {}
    for tag, J, h, beta, anneal, sweeps in (("pm_fixed", J1, h1, 1.1, False, 6), ("pm_anneal", J1, h1, 3.0, True, 9),
                                            ("gauss_fixed", J2, h2, 0.8, False, 5), ("gauss_anneal", J2, h2, 2.0, True, 7)):
        obj = rl.nmc().NMC(J, h)
        rl.seed_all(1000 + sweeps)
        m0 = np.sign(2 * np.random.rand(len(h)) - 1)
        M = obj.MCMC(sweeps, m0.copy(), beta, J, h, anneal=anneal)
        E = np.array([-(M[:, i].T @ J @ M[:, i] / 2 + M[:, i].T @ h) for i in range(sweeps)])
        out.update({f"{tag}_J": J, f"{tag}_h": h, f"{tag}_beta": beta, f"{tag}_anneal": anneal,
                    f"{tag}_sweeps": sweeps, f"{tag}_seed": 1000 + sweeps, f"{tag}_M": i8(M), f"{tag}_E": E})
    save("mcmc_element", **out)


def case_lbp():
    """LBP_convexified (NMC/nmc.py:93-166): marginal of every lambda step and the clusters."""
    J, h = O.random_pm_graph(60, 0.15, 1)
    N = 60
    ref = rl.nmc().NMC(J, h)
    rl.seed_all(3)
    ms = np.sign(2 * np.random.rand(N) - 1)
    epsv = np.abs(h) + np.sum(np.abs(J), axis=1)
    beta = 3.0
    with rl.quiet_tmp_cwd():
        cl, marg_all, _, _, _ = ref.LBP_convexified(3, 0.01, 0.9, ms.copy(), epsv, EPS, 100, 0.9999999, 0.999999, beta)
    # per-lambda iteration counts, replayed through the reference's own LoopyBeliefPropagation
    hm = np.zeros((N, N))
    um = J * ms.reshape(1, -1)
    iters = []
    for lam in marg_all.keys():
        _, _, _, _, it, hm, um = ref.LoopyBeliefPropagation(J, (h + lam * ms * epsv).copy(), beta, hm.copy(), um.copy(), EPS, 100)
        iters.append(it)
    save("lbp", J=J, h=h, m_star=i8(ms), beta=beta, lambdas=np.array(list(marg_all.keys())),
         marginals=np.array([marg_all[k] for k in marg_all.keys()]), iters=np.array(iters),
         clusters=np.concatenate(cl).astype(np.int64), params=np.array([3, 0.01, 0.9, EPS, 100, 0.9999999, 0.999999]))


def case_nmc_run():
    """C1-shaped: NMC.run on a sparse +-1 random graph (README parameters, sweeps reduced)."""
    J, h = O.random_pm_graph(60, 0.15, 1)
    args = (40, 10, 3, 2, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 100, EPS)
    rl.seed_all(21)
    with rl.quiet_tmp_cwd(), record_backbones(rl.nmc().NMC) as rec:
        M, E, mn = rl.nmc().NMC(J, h).run(*args)
    flat, sizes = rec.result()
    save("nmc_run_c1", J=J, h=h, args=np.array(args), seed=21, M=i8(M), E=np.asarray(E), min_energy=mn,
         backbone_flat=flat, backbone_sizes=sizes)
    # the reference unit-test shape: dense Gaussian J with field (NMC/unittests/test_nmc.py:9-17)
    rs = np.random.RandomState(5)
    N = 12
    hg = rs.randn(N)
    Jg = np.zeros((N, N))
    iu = np.triu_indices(N, 1)
    Jg[iu] = rs.randn(len(iu[0]))
    Jg += Jg.T
    args2 = (50, 10, 2, 1, 1, 20, 3, 3, 0.01, 0.9, 0.9999999, 0.999999, 10, EPS)
    rl.seed_all(4)
    with rl.quiet_tmp_cwd(), record_backbones(rl.nmc().NMC) as rec:
        M, E, mn = rl.nmc().NMC(Jg, hg).run(*args2)
    flat, sizes = rec.result()
    save("nmc_run_gauss", J=Jg, h=hg, args=np.array(args2), seed=4, M=i8(M), E=np.asarray(E), min_energy=mn,
         backbone_flat=flat, backbone_sizes=sizes)


NPT_KW = dict(num_cycles=2, full_update_frequency=1, M_skip=1, temp_x=20, global_beta=3, lambda_start=3,
              lambda_end=0.01, lambda_reduction_factor=0.9, threshold_initial=0.9999999,
              threshold_cutoff=0.999999, max_iterations=100, tolerance=EPS)


def case_npt():
    """C2-shaped: APT_preprocessor ladder then NPT.run on 3D +-J EA with doNMC on the coldest replicas."""
    A, h = O.ea3d_pm_j(4, 2)
    J = A.toarray()
    rl.seed_all(13)
    with rl.quiet_tmp_cwd():
        beta, sigma = rl.apt_preprocessor().APT_preprocessor(J.copy(), h.copy()).run(
            num_sweeps_MCMC=30, num_sweeps_read=20, num_rng=5, beta_start=0.5, alpha=1.25, sigma_E_val=1000,
            beta_max=4, use_hash_table=0, num_cores=1)
    save("apt_preprocessor_c2", J=J, h=h, seed=13, args=np.array([30, 20, 5, 0.5, 1.25, 1000, 4]),
         beta=np.array(beta, dtype=np.float64), sigma=np.array(sigma, dtype=np.float64))
    betas = np.array(beta, dtype=np.float64)[:4]
    doNMC = [False, False, True, True]
    kw = dict(num_sweeps_MCMC=60, num_sweeps_read=20, num_swap_attempts=4, num_swapping_pairs=1, **NPT_KW)
    rl.seed_all(12)
    with rl.quiet_tmp_cwd(), record_backbones(rl.npt().NPT) as rec:
        M, E = rl.npt().NPT(J, h).run(betas, 4, doNMC, num_cores=1, **kw)
    flat, sizes = rec.result()
    save("npt_run_c2", J=J, h=h, seed=12, beta_list=betas, doNMC=np.array(doNMC),
         num_sweeps_MCMC=60, num_sweeps_read=20, num_swap_attempts=4, num_swapping_pairs=1, M=i8(M), E=E,
         backbone_flat=flat, backbone_sizes=sizes)


def case_npt_sk():
    """C3-shaped: dense Gaussian SK, NPT with all doNMC False."""
    J, h = O.sk_gaussian(40, 3)
    betas = np.array([0.4, 0.8, 1.2, 1.6, 2.0])
    rl.seed_all(31)
    with rl.quiet_tmp_cwd():
        M, E = rl.npt().NPT(J, h).run(betas, 5, [False] * 5, num_sweeps_MCMC=24, num_sweeps_read=12,
                                      num_swap_attempts=4, num_swapping_pairs=2, num_cores=1)
    save("npt_run_c3", J=J, h=h, seed=31, beta_list=betas, num_sweeps_MCMC=24, num_sweeps_read=12,
         num_swap_attempts=4, num_swapping_pairs=2, M=i8(M), E=E)


def case_icm():
    """C4-shaped: APT_ICM.run on 3D +-J EA (10 sub-replicas per beta, NPT/apt_ICM.py:177)."""
    A, h = O.ea3d_pm_j(4, 4)
    J = A.toarray()
    betas = np.array([0.3, 0.7, 1.1, 1.6])
    out = {}
    for tag, nsm, nsr, nsa, npairs, seed in (("a", 12, 8, 4, 1, 14), ("b", 4, 4, 4, 2, 15)):
        rl.seed_all(seed)
        with rl.quiet_tmp_cwd():
            M, E = rl.apt_icm().APT_ICM(J.copy(), h.copy()).run(betas, 4, num_sweeps_MCMC=nsm, num_sweeps_read=nsr,
                                                                num_swap_attempts=nsa, num_swapping_pairs=npairs)
        out.update({f"{tag}_args": np.array([nsm, nsr, nsa, npairs]), f"{tag}_seed": seed, f"{tag}_M": i8(M), f"{tag}_E": E})
    # element level: clusters of two random states
    rs = np.random.RandomState(77)
    s1 = rs.choice([-1.0, 1.0], size=64)
    s2 = rs.choice([-1.0, 1.0], size=64)
    cl = rl.apt_icm().APT_ICM(J, h).find_disagreement_clusters(s1, s2, J)
    labels = -np.ones(64, dtype=np.int32)
    for k, c in enumerate(cl):
        labels[np.array(c, dtype=int)] = k
    save("apt_icm_c4", J=J, h=h, beta_list=betas, s1=i8(s1), s2=i8(s2), labels=labels, n_clusters=len(cl), **out)


def case_npt_sparse():
    """C5-shaped: NPT.run given a scipy.sparse J (the reference accepts it when no replica does NMC)."""
    A, h = O.ea3d_pm_j(6, 5)
    betas = np.linspace(0.2, 2.0, 6)
    rl.seed_all(51)
    with rl.quiet_tmp_cwd():
        M, E = rl.npt().NPT(A.copy(), h).run(betas, 6, [False] * 6, num_sweeps_MCMC=6, num_sweeps_read=6,
                                             num_swap_attempts=3, num_swapping_pairs=2, num_cores=1)
    save("npt_run_c5", L=6, instance_seed=5, seed=51, beta_list=betas, num_sweeps_MCMC=6, num_sweeps_read=6,
         num_swap_attempts=3, num_swapping_pairs=2, M=i8(M), E=E)


def case_known_answers():
    """Known-answer energies shipped with the reference examples: Wishart
    planted instances, convention J = -J_file, E = -(m^T J m / 2), reported energy * max|J| = gs."""
    base = os.path.join(rl.REFERENCE_ROOT, "NMC", "examples", "wishart_small", "wishart_planting_N_10_alpha_0.50")
    gs = {}
    with open(os.path.join(base, "gs_energies.txt")) as f:
        for line in f:
            if line.strip():
                name, val = line.split()
                gs[name] = float(val)
    Js, Es = [], []
    for inst in (1, 2, 3):
        name = f"wishart_planting_N_10_alpha_0.50_inst_{inst}.txt"
        W = np.zeros((10, 10))
        with open(os.path.join(base, name)) as f:
            for line in f:
                if not line.strip() or line.startswith("#"):
                    continue
                a, b, v = line.split()
                if int(a) != int(b):
                    W[int(a), int(b)] = float(v)
                    W[int(b), int(a)] = float(v)
        Js.append(-W)
        Es.append(gs[name])
    save("known_answers_wishart", J=np.array(Js), gs_energy=np.array(Es))


def case_chimera_known_answer():
    """Chimera droplet instance 001 (128 spins, real-valued J with fields) with its shipped ground state
    (`groundstates_otn2d.txt`: energy and bit string; s = 2b-1, J = -J_file, h = -h_file)."""
    base = os.path.join(rl.REFERENCE_ROOT, "NMC", "examples", "Chimera_droplet_instances", "chimera128_spinglass_power")
    text = open(os.path.join(base, "001.txt")).read()
    gs_line = [l for l in open(os.path.join(base, "groundstates_otn2d.txt")) if l.startswith("001.txt")][0]
    parts = gs_line.split(":")[1].split()
    save("known_answer_chimera128", instance_text=np.array(text), gs_energy=float(parts[0]),
         gs_bits=np.array([int(b) for b in parts[1:]], dtype=np.int8))


def case_public_methods():
    """The classes' element-level public methods besides run(): LoopyBeliefPropagation with all its by-products,
    LBP_convexified's dictionaries, NMC_subroutine with provided clusters (both variants), MCMC_task / NMC_task,
    replica_energy, and the APT classes' MCMC."""
    N = 40
    J, h0 = O.random_pm_graph(N, 0.15, 11)
    rs = np.random.RandomState(12)
    h = 0.1 * rs.randn(N)
    out = {"J": J, "h": h}
    nmc, npt = rl.nmc().NMC(J, h), rl.npt().NPT(J, h)
    # one LBP call from the reference's own initial messages
    ms = np.sign(rs.rand(N) - 0.5)
    epsv = np.abs(h) + np.sum(np.abs(J), axis=1)
    out.update(lbp_field1=h + 0.7 * ms * epsv, lbp_field2=h + 0.5 * ms * epsv)
    res = nmc.LoopyBeliefPropagation(J, h + 0.7 * ms * epsv, 1.5, np.zeros((N, N)), J * ms.reshape(1, -1), 1e-10, 200)
    out.update(lbp_m_star=i8(ms), lbp_beta=1.5, lbp_tol=1e-10, lbp_max_iter=200)
    for k, v in zip(("marg", "corr", "h_tilde", "J_tilde", "iteration", "h_msgs", "u_msgs"), res):
        out["lbp_" + k] = np.asarray(v)
    # a second call warm-started from the first one's messages (row-constant h_msgs off the entries of J)
    res2 = nmc.LoopyBeliefPropagation(J, h + 0.5 * ms * epsv, 1.5, res[5].copy(), res[6].copy(), 1e-10, 200)
    for k, v in zip(("marg", "corr", "h_tilde", "J_tilde", "iteration", "h_msgs", "u_msgs"), res2):
        out["lbp2_" + k] = np.asarray(v)
    # arbitrary dense initial messages, 3 iterations (exercises messages off the entries of J)
    hm0, um0 = rs.randn(N, N), 0.2 * rs.randn(N, N)
    res3 = nmc.LoopyBeliefPropagation(J, h.copy(), 1.5, hm0.copy(), um0.copy(), 1e-10, 3)
    out.update(lbp3_h0=hm0, lbp3_u0=um0)
    for k, v in zip(("marg", "corr", "h_tilde", "J_tilde", "iteration", "h_msgs", "u_msgs"), res3):
        out["lbp3_" + k] = np.asarray(v)
    # LBP_convexified with a loose tolerance (iteration counts insensitive to last-place differences)
    with rl.quiet_tmp_cwd():
        cl, marg_all, mean_all, ht_all, jt_all = nmc.LBP_convexified(2.0, 0.05, 0.8, ms.copy(), epsv, 1e-9, 300, 0.99, 0.9, 2.0)
    lams = list(marg_all.keys())
    out.update(conv_args=np.array([2.0, 0.05, 0.8, 1e-9, 300, 0.99, 0.9, 2.0]), conv_lambdas=np.array(lams),
               conv_marginals=np.array([marg_all[k] for k in lams]), conv_means=np.array([mean_all[k] for k in lams]),
               conv_h_tilde=np.array([ht_all[k] for k in lams]), conv_J_tilde_last=jt_all[lams[-1]],
               conv_clusters_flat=np.concatenate(cl).astype(np.int64) if cl else np.array([], dtype=np.int64),
               conv_cluster_sizes=np.array([len(c) for c in cl], dtype=np.int64))
    # NMC_subroutine with provided clusters (no LBP -> exact), both variants
    clusters = np.array([1, 4, 5, 9, 17, 23, 30, 31])
    sub_args = (3, 6, 2, 2, 1.7, 10.0, 0.5, 0.01, 0.9, 0.9999, 0.999, 50, EPS)
    for tag, obj in (("nmc", nmc), ("npt", npt)):
        rl.seed_all(77)
        with rl.quiet_tmp_cwd():
            M, E, mn, ac = obj.NMC_subroutine(ms.copy(), *sub_args, all_clusters=clusters.copy())
        out.update({f"sub_{tag}_M": i8(M), f"sub_{tag}_E": E, f"sub_{tag}_min": mn, f"sub_{tag}_clusters": np.asarray(ac)})
    out.update(sub_args=np.array(sub_args), sub_clusters=clusters, sub_seed=77)
    # NPT.MCMC_task / NMC_task (backbone of the task recorded) / replica_energy
    betas = np.array([0.4, 0.9, 1.6])
    rl.seed_all(78)
    Mt = npt.MCMC_task(2, 6, ms.copy(), betas)
    out.update(task_seed=78, task_betas=betas, task_M=i8(Mt))
    mn_e, EE1 = npt.replica_energy(Mt, 4)
    out.update(rep_min=mn_e, rep_EE1=EE1)
    rl.seed_all(79)
    with record_backbones(rl.npt().NPT) as rec, rl.quiet_tmp_cwd():
        Mn = npt.NMC_task(ms.copy(), 2, 4, 1, 1, 2.0, 10.0, 2.0, 0.05, 0.8, 0.99, 0.9, 300, 1e-9)
    flat, sizes = rec.result()
    out.update(nmctask_seed=79, nmctask_args=np.array([2, 4, 1, 1, 2.0, 10.0, 2.0, 0.05, 0.8, 0.99, 0.9, 300, 1e-9]),
               nmctask_M=i8(Mn), nmctask_backbone_flat=flat, nmctask_backbone_sizes=sizes)
    # APT classes: MCMC with (N,1) h, MCMC_task, replica_energy
    prep, icm = rl.apt_preprocessor().APT_preprocessor(J, h), rl.apt_icm().APT_ICM(J, h)
    rl.seed_all(80)
    Mp = prep.MCMC(5, ms.copy(), 1.2)
    En, mlast = prep.MCMC_task(ms.copy(), 0.8, 7, 3)
    Mi = icm.MCMC(4, ms.copy(), 0.6)
    with warnings_ignored():
        mn_i, EE_i = icm.replica_energy(Mi, 4)
    out.update(apt_seed=80, prep_M=i8(Mp), prep_task_E=En, prep_task_m=i8(mlast), icm_M=i8(Mi), icm_rep_min=mn_i,
               icm_rep_EE1=EE_i)
    save("public_methods", **out)


def case_dcl_known_answer():
    """DCL (deceptive cluster loops) instance C8/00 with the planted minimum energy of its `_sol.txt` (`min_energy`);
    convention J = -J_file, E = -(m^T J m / 2) (NMC/examples/DCL_example.py:48-54).  The file rounds 1/7 to 0.14286, so the
    energy of the file's couplings is -389.43032 against the stated -389.42857."""
    base = os.path.join(rl.REFERENCE_ROOT, "NMC", "examples", "DCL_instances", "C8")
    text = open(os.path.join(base, "00.txt")).read()
    sol = dict(line.split() for line in open(os.path.join(base, "00_sol.txt")) if line.strip())
    save("known_answer_dcl_c8", instance_text=np.array(text), min_energy=float(sol["min_energy"]),
         n_active=int(sol["nq"]))


def case_wishart36_known_answer():
    """Wishart planted instance N = 36, alpha = 0.50, instance 1 (hard enough that a one-core
    parallel-tempering search needs thousands of sweeps, easy enough that it always gets there), with the planted ground-state
    energy of its `gs_energies.txt`; convention J = -J_file, E = -(m^T J m / 2) (NMC/examples/wishart_example.py:8-60)."""
    base = os.path.join(rl.REFERENCE_ROOT, "NMC", "examples", "wishart_small", "wishart_planting_N_36_alpha_0.50")
    name = "wishart_planting_N_36_alpha_0.50_inst_1.txt"
    gs = dict(line.split() for line in open(os.path.join(base, "gs_energies.txt")) if line.strip())
    save("known_answer_wishart36", instance_text=np.array(open(os.path.join(base, name)).read()),
         gs_energy=float(gs[name]))


def case_contrived_generator():
    """contrived_instance_generator.py (Wishart backbone + trees): adjacency, weights, cross connections, edge removal,
    fields and the instance text, produced by the reference's own functions for one seed."""
    import importlib.util
    import tempfile
    import types
    rl.nmc()  # installs the matplotlib stub
    sys.modules.setdefault("seaborn", types.ModuleType("seaborn"))
    base = os.path.join(rl.REFERENCE_ROOT, "NMC", "examples", "contrived_wishart_backbone")
    spec = importlib.util.spec_from_file_location("cig", os.path.join(base, "contrived_instance_generator.py"))
    cig = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cig)
    Jw, _ = cig.txt_to_A_wishart(os.path.join(base, "wishart_planting_N_10_alpha_0.20",
                                              "wishart_planting_N_10_alpha_0.20_inst_1.txt"))
    Jw = -Jw.toarray()
    b, levels, n_cross, n_remove, seed = 10, 2, 20, 3, 77
    np.random.seed(seed)
    A = cig.generate_adjacency(b, levels)
    A = cig.assign_random_weights(A, b, -1, 1, -10, 10)
    A = cig.add_cross_connections(A, b, n_cross, -1, 1)
    A = cig.remove_random_backbone_edges(A, b, n_remove)
    A[0:b, 0:b] = 10 * Jw / np.max(np.abs(Jw))
    h = (np.random.rand(len(A)) - 0.5) * 2 * 0.2 * 10
    path = os.path.join(tempfile.mkdtemp(), "inst.txt")
    cig.save_to_txt(A, h, path)
    save("contrived_generator", J_backbone=Jw, args=np.array([levels, 0.2, 1, 10, n_cross, 1, n_remove]), seed=seed,
         J=A, h=h, text=np.array(open(path).read()), adjacency_4_1=cig.generate_adjacency(4, 1))


class warnings_ignored:
    def __enter__(self):
        import warnings
        self.c = warnings.catch_warnings()
        self.c.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        return self.c.__exit__(*a)


if __name__ == "__main__":
    if not rl.available():
        sys.exit("reference not mounted; golden vectors can only be generated in the build container")
    import warnings
    warnings.simplefilter("ignore")
    for fn in (case_mcmc, case_lbp, case_nmc_run, case_npt, case_npt_sk, case_icm, case_npt_sparse, case_known_answers,
               case_chimera_known_answer, case_public_methods, case_contrived_generator, case_dcl_known_answer,
               case_wishart36_known_answer):
        fn()
