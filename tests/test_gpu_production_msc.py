"""GPU tests of the production path (bit-packed multi-spin coding, K2/K4'/K6) through the C ABI.

Production mode is statistically -- not bit -- equivalent to the reference, so the checks are:
exactness where the domain offers it (pack/unpack round trip, integer energies vs the oracle, swap
bookkeeping), the single-site conditional distribution of the heat-bath rule, agreement with the
EXACT Boltzmann distribution of small systems (full enumeration), and agreement of per-beta mean
energy and |magnetisation| with samples of the reference algorithm (oracle) within 3 sigma
(north_star: "per-beta mean energy and magnetisation within 3 sigma")."""
import itertools

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nl():
    from nlmc_b200 import _lib, host
    return type("NL", (), dict(lib=_lib, host=host))


def lattice_2d(L, seed):
    rs = np.random.RandomState(seed)
    N = L * L
    idx = np.arange(N)
    x, y = idx % L, idx // L
    rows, cols, vals = [], [], []
    for nb in (((x + 1) % L) + L * y, x + L * ((y + 1) % L)):
        v = rs.choice([-1.0, 1.0], size=N)
        rows += [idx, nb]
        cols += [nb, idx]
        vals += [v, v]
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()
    return A, np.zeros(N)


def exact_mean_energy(A, betas):
    """<E>, <E^2> of the Boltzmann distribution by full enumeration (N <= 16)."""
    from oracle import oracle as O
    n = A.shape[0]
    states = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=np.int8)
    E = O.energy(O.Csr(A), np.zeros(n), states)
    out = []
    for b in betas:
        w = np.exp(-b * (E - E.min()))
        w /= w.sum()
        m1 = (w * E).sum()
        out.append((m1, (w * E * E).sum() - m1 * m1))
    return out


def test_pack_unpack_and_exact_energies(nl):
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(6, 3)
    csr = O.Csr(A)
    prob = nl.host.Problem(A, h)
    betas = np.array([0.3, 0.9, 1.5])
    msc = nl.lib.Msc(prob.inst, betas, 130, seed=5)  # 130 -> padded to 256 ladders
    assert msc.n_ladders == 256 and msc.n_words == 3 * 8 and msc.n_colours == 2 and msc.n_bonds == 3 * 216
    rs = np.random.RandomState(0)
    picks = [(0, 0), (1, 31), (2, 32), (1, 129), (2, 255)]
    states = {k: rs.choice([-1, 1], size=csr.n).astype(np.int8) for k in picks}
    for (b, lad), s in states.items():
        msc.set_spins(b, lad, s)
    for (b, lad), s in states.items():
        assert np.array_equal(msc.get_spins(b, lad), s)
    E = msc.energies()
    assert E.shape == (3, 256)
    for (b, lad), s in states.items():
        assert E[b, lad] == O.energy(csr, h, s)[0]
    # every replica, after some sweeps: device energies == oracle energies of the unpacked states (exact)
    msc.sweep(3)
    E = msc.energies()
    for b in range(3):
        for lad in (0, 17, 100, 255):
            assert E[b, lad] == O.energy(csr, h, msc.get_spins(b, lad))[0]
    # packed round trip through host buffers
    P = msc.get_packed()
    assert P.shape == (216, 24) and P.dtype == np.uint32
    msc.sweep(1)
    msc.set_packed(P)
    assert np.array_equal(msc.get_packed(), P)
    bit = (P[:, 1 * 8 + 129 // 32] >> (129 % 32)) & 1  # word = beta*G + ladder//32
    assert np.array_equal(np.where(bit == 1, 1, -1).astype(np.int8), msc.get_spins(1, 129))


@pytest.mark.parametrize("L,two_colourable", [(5, False), (6, True)])
def test_energies_exact_on_bipartite_and_frustrated_colourings(nl, L, two_colourable):
    """K4' sums the unsatisfied bonds seen from one colour class when the graph is two-colourable (every bond joins
    the two classes) and from all sites otherwise (odd L: the periodic lattice needs more colours): exact both ways."""
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(L, 11)
    csr = O.Csr(A)
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, [0.4, 1.1], 128, seed=2)
    assert (msc.n_colours == 2) == two_colourable
    msc.sweep(4)
    E = msc.energies()
    for b in range(2):
        for lad in (0, 31, 32, 127):
            assert E[b, lad] == O.energy(csr, h, msc.get_spins(b, lad))[0]
    msc.close()


def test_unsupported_instances_fail_loudly(nl):
    from oracle import oracle as O
    J, h = O.sk_gaussian(16, 1)
    with pytest.raises(nl.lib.NlmcError, match="J in"):
        nl.lib.Msc(nl.host.Problem(J, h).inst, [1.0], 1)
    A, h = O.ea3d_pm_j(4, 1)
    with pytest.raises(nl.lib.NlmcError, match="h = 0"):
        nl.lib.Msc(nl.host.Problem(A, np.ones(64)).inst, [1.0], 1)
    J, h = O.random_pm_graph(40, 0.5, 2)
    with pytest.raises(nl.lib.NlmcError, match="degree"):
        nl.lib.Msc(nl.host.Problem(J, h).inst, [1.0], 1)


def test_single_site_conditional_distribution(nl):
    """All 4096 lanes start from the same state S0.  After one sweep the colour-0 sites were drawn from
    P(+1) = 1/(1+exp(-2 beta f(S0))) with f read from S0's (unchanged) colour-1 neighbours."""
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(8, 7)
    csr = O.Csr(A)
    n = csr.n
    betas = np.array([0.25, 0.6, 1.1])
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 1024, seed=99)
    rs = np.random.RandomState(4)
    s0 = rs.choice([-1, 1], size=n).astype(np.int8)
    P = np.where(s0[:, None] > 0, np.uint32(0xffffffff), np.uint32(0)) * np.ones((1, msc.n_words), dtype=np.uint32)
    msc.set_packed(np.ascontiguousarray(P.astype(np.uint32)))
    msc.sweep(1)
    out = msc.get_packed()
    idx = np.arange(n)
    colour0 = ((idx % 8) + (idx // 8) % 8 + idx // 64) % 2 == 0
    f0 = np.asarray(A @ s0.astype(float))  # field from S0 (colour-1 neighbours did not move yet)
    G = msc.n_ladders // 32
    for b, beta in enumerate(betas):
        words = out[:, b * G:(b + 1) * G]
        ups = np.zeros(n)
        for g in range(G):
            ups += np.array([bin(int(v)).count("1") for v in words[:, g]])
        for f in (-6, -4, -2, 0, 2, 4, 6):
            sel = colour0 & (f0 == f)
            trials = sel.sum() * msc.n_ladders
            if trials == 0:
                continue
            p = 1.0 / (1.0 + np.exp(-2 * beta * f))
            sigma = np.sqrt(trials * p * (1 - p))
            assert abs(ups[sel].sum() - trials * p) <= 3.0 * sigma + 1, (beta, f, ups[sel].sum(), trials * p, sigma)


@pytest.mark.parametrize("with_swaps", [False, True])
def test_exact_boltzmann_small_lattice(nl, with_swaps):
    """2D 4x4 periodic +-J (degree 4: exercises the neutral-pair padding).  Long runs over 1024 ladders must
    reproduce the exact <E>(beta) from full enumeration, with and without replica exchange."""
    A, h = lattice_2d(4, 11)
    betas = np.array([0.2, 0.5, 0.9, 1.4])
    exact = exact_mean_energy(A, betas)
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 1024, seed=2024 + with_swaps)
    assert msc.n_colours == 2
    for _ in range(40):  # equilibrate
        msc.round(5, 2 if with_swaps else 0)
    samples = []
    for _ in range(60):
        msc.round(4, 2 if with_swaps else 0)
        samples.append(msc.energies())
    S = np.array(samples)  # [t][beta][ladder]
    if with_swaps:
        assert msc.swap_count() > 0
    for b in range(len(betas)):
        per_ladder = S[:, b, :].mean(axis=0)  # ladders are independent -> honest error bar
        mean, err = per_ladder.mean(), per_ladder.std(ddof=1) / np.sqrt(per_ladder.size)
        assert abs(mean - exact[b][0]) <= 3.0 * err + 1e-9, (betas[b], mean, exact[b][0], err)


def test_exact_boltzmann_three_colour_lattice(nl):
    """2D 3x3 periodic +-J: odd cycles, so the colouring needs more than two classes (one launch per class and the
    all-sites energy path).  <E>(beta) from full enumeration of the 512 states."""
    A, h = lattice_2d(3, 5)
    betas = np.array([0.3, 0.8, 1.5])
    exact = exact_mean_energy(A, betas)
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 1024, seed=77)
    assert msc.n_colours >= 3
    msc.sweep(200)
    samples = []
    for _ in range(60):
        msc.sweep(4)
        samples.append(msc.energies())
    S = np.array(samples)
    for b in range(len(betas)):
        per_ladder = S[:, b, :].mean(axis=0)
        mean, err = per_ladder.mean(), per_ladder.std(ddof=1) / np.sqrt(per_ladder.size)
        assert abs(mean - exact[b][0]) <= 3.0 * err + 1e-9, (betas[b], mean, exact[b][0], err)


def test_swap_bookkeeping_is_exact(nl):
    """After an exchange the energies the swap kernel carried along must equal freshly computed energies of
    the exchanged configurations, and the multiset of configurations of every ladder is conserved."""
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(4, 5)
    betas = np.linspace(0.1, 1.2, 6)
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 128, seed=8)
    msc.sweep(3)
    before = {lad: [msc.get_spins(b, lad).tobytes() for b in range(6)] for lad in (0, 5, 77)}
    E_before = msc.energies()
    msc.round(0, 2)  # no sweeps: energies + exchange only
    accepted = msc.swap_count()
    assert accepted > 0
    E_after = msc.energies()
    assert np.array_equal(np.sort(E_before, axis=0), np.sort(E_after, axis=0))  # energies permuted within ladders
    assert np.count_nonzero(E_before != E_after) > 0
    for lad, confs in before.items():
        after = [msc.get_spins(b, lad).tobytes() for b in range(6)]
        assert sorted(confs) == sorted(after)
    moved = sum(before[lad][b] != msc.get_spins(b, lad).tobytes() for lad in before for b in range(6))
    assert moved % 2 == 0


def test_statistical_equivalence_with_reference_sampler(nl):
    """3D +-J L=6: per-beta <E> and <|m|> of the production kernel vs the reference algorithm (oracle MCMC,
    random-permutation sequential heat bath) -- within 3 sigma of the combined error (north_star)."""
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(6, 21)
    csr = O.Csr(A)
    n = csr.n
    betas = np.array([0.3, 0.6, 0.9])
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 256, seed=31)
    msc.sweep(300)
    acc_E, acc_m = [], []
    for _ in range(20):
        msc.sweep(20)
        acc_E.append(msc.energies())
        P = msc.get_packed()
        G = msc.n_ladders // 32
        mags = np.zeros((len(betas), msc.n_ladders))
        for b in range(len(betas)):
            for g in range(G):
                w = P[:, b * G + g]
                bits = (w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1
                mags[b, g * 32:(g + 1) * 32] = np.abs(2.0 * bits.sum(axis=0) - n) / n
        acc_m.append(mags)
    E_gpu = np.array(acc_E).mean(axis=0)  # [beta][ladder]
    m_gpu = np.array(acc_m).mean(axis=0)
    rs = np.random.RandomState(3)
    chains = 24
    for b, beta in enumerate(betas):
        Es, ms = [], []
        for c in range(chains):
            m0 = rs.choice([-1, 1], size=n).astype(np.int8)
            M, _ = O.mcmc(csr, h, m0, np.full(700, beta), rng=rs)
            tail = M[300::20]
            Es.append(O.energy(csr, h, tail).mean())
            ms.append(np.abs(tail.sum(axis=1)).mean() / n)
        for gpu, ref in ((E_gpu[b], np.array(Es)), (m_gpu[b], np.array(ms))):
            err = np.hypot(gpu.std(ddof=1) / np.sqrt(gpu.size), ref.std(ddof=1) / np.sqrt(ref.size))
            assert abs(gpu.mean() - ref.mean()) <= 3.0 * err, (beta, gpu.mean(), ref.mean(), err)


def test_determinism_and_seed_dependence(nl):
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(4, 2)
    prob = nl.host.Problem(A, h)
    outs = []
    for seed in (7, 7, 8):
        msc = nl.lib.Msc(prob.inst, [0.5, 1.0], 128, seed=seed)
        msc.round(6, 1)
        outs.append(msc.get_packed())
        msc.close()
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


def test_round_host_sync_async_and_device_round_agree(nl):
    """The three ways to run a swap round -- device-resident (nlmc_msc_round), host buffers (nlmc_msc_round_host) and
    host buffers without waiting (nlmc_msc_round_host_async + nlmc_msc_sync, two handles in flight at once) -- give
    the same states and energies for the same seed."""
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(6, 3)
    prob = nl.host.Problem(A, h)
    betas = [0.4, 0.8, 1.2]
    ms = [nl.lib.Msc(prob.inst, betas, 128, seed=21) for _ in range(3)]
    start = ms[0].get_packed()
    shape = ms[0].packed_shape()
    ms[0].round(5, 1)
    ref_state = ms[0].get_packed()
    ref_E = ms[0].energies()
    out1, E1 = np.empty(shape, np.uint32), np.empty((3, 128))
    ms[1].round_host(start.ctypes.data, 5, 1, out1.ctypes.data, E1.ctypes.data)
    out2, E2 = np.empty(shape, np.uint32), np.empty((3, 128))
    other = nl.lib.Msc(prob.inst, betas, 128, seed=99)  # a second batch in flight on its own stream
    out3, E3 = np.empty(shape, np.uint32), np.empty((3, 128))
    ms[2].round_host_async(start.ctypes.data, 5, 1, out2.ctypes.data, E2.ctypes.data)
    other.round_host_async(start.ctypes.data, 5, 1, out3.ctypes.data, E3.ctypes.data)
    ms[2].sync(); other.sync()
    assert np.array_equal(out1, ref_state) and np.array_equal(out2, ref_state)
    assert not np.array_equal(out3, ref_state)
    # energies returned with the states are those of the returned states (after the exchange)
    assert np.array_equal(np.sort(E1, axis=0), np.sort(E2, axis=0))
    assert np.array_equal(E1, ref_E) and np.array_equal(E2, ref_E)
    for m in ms + [other]:
        m.close()


def test_npt_production_mode_api(nl, tmp_cwd):
    """Drop-in NPT in production mode: shapes/dtypes of the reference contract, energies consistent with M."""
    from nlmc_b200 import NPT
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(4, 6)
    betas = np.array([0.3, 0.7, 1.1, 1.5])
    np.random.seed(5)
    obj = NPT(A, h, mode="production")
    obj.num_runs = 3
    M, E = obj.run(betas, 4, [False] * 4, num_sweeps_MCMC=40, num_sweeps_read=20, num_swap_attempts=4,
                   num_swapping_pairs=1)
    assert M.shape == (64 * 4, 10) and M.dtype == np.float64 and E.shape == (4,)
    assert np.all(np.abs(M) == 1)
    csr = O.Csr(A)
    for r in range(4):
        Er = O.energy(csr, h, M[r * 64:(r + 1) * 64, :5].T.astype(np.int8))
        assert E[r] == Er.min()
    assert obj.energies_all_runs.shape == (4, 3)


def test_sharding_invariance(nl):
    """256 ladders in one handle evolve bit-identically to two handles of 128 ladders with ladder offsets 0 and
    128 (same seed): the property that makes N-GPU results identical to 1-GPU results."""
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(6, 9)
    prob = nl.host.Problem(A, h)
    betas = np.array([0.4, 0.8, 1.2])
    whole = nl.lib.Msc(prob.inst, betas, 256, seed=77)
    parts = [nl.lib.Msc(prob.inst, betas, 128, seed=77, ladder_offset=off) for off in (0, 128)]
    for m in [whole] + parts:
        for _ in range(3):
            m.round(4, 1)
    Pw = whole.get_packed().reshape(prob.n, 3, 8)  # [site][beta][group]
    Ew = whole.energies()
    for i, part in enumerate(parts):
        Pp = part.get_packed().reshape(prob.n, 3, 4)
        assert np.array_equal(Pw[:, :, 4 * i:4 * i + 4], Pp)
        assert np.array_equal(Ew[:, 128 * i:128 * (i + 1)], part.energies())
    assert whole.swap_count() == sum(p.swap_count() for p in parts) > 0


def test_npt_production_energy_distribution_matches_reference(nl, tmp_cwd):
    """north_star: 'same best-found energy distribution'.  Whole NPT runs (sweeps + exchanges) on a 64-spin +-J
    lattice: the coldest replica's final energy over 128 production ladders vs 48 runs of the reference algorithm
    (oracle NPT restatement, bit-identical to the reference): means within 3.5 sigma, and the same ground level."""
    import random
    from nlmc_b200 import NPT
    from oracle import oracle as O
    A, h = O.ea3d_pm_j(4, 12)
    J = A.toarray()
    csr = O.Csr(A)
    betas = np.array([0.4, 0.8, 1.2, 1.6])
    kw = dict(num_sweeps_MCMC=120, num_sweeps_read=40, num_swap_attempts=12, num_swapping_pairs=1)
    np.random.seed(3)
    obj = NPT(A, h, mode="production")
    obj.num_runs = 128
    obj.run(betas, 4, [False] * 4, **kw)
    E_gpu = obj.energies_all_runs[-1]          # coldest replica, all 128 ladders, after the last sweep
    E_ref = []
    for s in range(48):
        np.random.seed(1000 + s)
        random.seed(1000 + s)
        M, _ = O.npt_run(J, h, betas, 4, [False] * 4, **kw)
        E_ref.append(O.energy(csr, h, M[3 * 64:, -1].astype(np.int8))[0])
    E_ref = np.array(E_ref)
    err = np.hypot(E_gpu.std(ddof=1) / np.sqrt(E_gpu.size), E_ref.std(ddof=1) / np.sqrt(E_ref.size))
    assert abs(E_gpu.mean() - E_ref.mean()) <= 3.0 * err, (E_gpu.mean(), E_ref.mean(), err)
    assert E_gpu.min() == E_ref.min() or abs(E_gpu.min() - E_ref.min()) <= 4  # both reach the lowest levels
    assert abs(E_gpu.std(ddof=1) - E_ref.std(ddof=1)) <= 0.5 * max(E_gpu.std(ddof=1), E_ref.std(ddof=1)) + 1


def test_full_size_c5_properties(nl):
    """Config C5 at full size (3D +-J EA L = 64: 262,144 spins, 32 betas x 128 ladders = 4096 replicas): properties that
    do not need an oracle run -- bit-sliced energies equal the fp64 energy kernel on unpacked replicas (exact integers),
    a pure exchange step permutes each ladder's energies (checksum of sorted energies unchanged) and keeps the state a
    permutation of itself, and two handles with one seed stay bit-identical (checksum of the packed state)."""
    from nlmc_b200 import instances
    A, h = instances.ea3d_pm_j(64, 5)
    prob = nl.host.Problem(A, h)
    betas = np.linspace(0.2, 2.0, 32)
    a = nl.lib.Msc(prob.inst, betas, 128, seed=1000)
    b = nl.lib.Msc(prob.inst, betas, 128, seed=1000)
    assert a.n_words == 128 and a.n_colours == 2 and a.n_bonds == 3 * 64 ** 3
    for m in (a, b):
        m.round(3, 10)
    Pa = a.get_packed()
    assert np.array_equal(Pa, b.get_packed())                          # determinism at full size
    E = a.energies()
    assert E.shape == (32, 128) and np.all(E == np.round(E)) and np.all(E[-1] < E[0])   # colder is lower
    picks = [(0, 0), (7, 31), (15, 64), (31, 127)]
    S = np.stack([a.get_spins(bi, lad) for bi, lad in picks])
    assert np.array_equal(prob.inst.energy_states(S), np.array([E[bi, lad] for bi, lad in picks]))
    # exchange only: energies are permuted within each ladder, spin content is conserved
    ones_before = int(np.unpackbits(Pa.view(np.uint8)).sum())
    a.round(0, 10)
    E2 = a.energies()
    assert np.array_equal(np.sort(E2, axis=0), np.sort(E, axis=0)) and a.swap_count() > 0
    assert int(np.unpackbits(a.get_packed().view(np.uint8)).sum()) == ones_before
    a.close(); b.close()


@pytest.mark.parametrize("steps,merged", [(6, 0), (4, 1), (5, 1), (8, 1)])
def test_exact_boltzmann_for_every_form_of_the_bernoulli_draw(nl, monkeypatch, steps, merged):
    """The draw g ~ B(q) has tuning knobs (unconditional steps, merged round); every setting must sample the same law.
    The ladder reaches beta = 2.4 (all unconditional steps are zero steps there: thresholds below 2^-13) and beta = 0.05
    (threshold bits set from the second step on), with exchanges, in both the scalar-threshold and the beta-label form."""
    monkeypatch.setenv("NLMC_MSC_STEPS", str(steps))
    monkeypatch.setenv("NLMC_MSC_MERGED", str(merged))
    A, h = lattice_2d(4, 23)
    betas = np.array([0.05, 0.4, 1.0, 2.4])
    exact = exact_mean_energy(A, betas)
    prob = nl.host.Problem(A, h)
    for labelled in (False, True):
        msc = nl.lib.Msc(prob.inst, betas, 1024, seed=77 + steps, labelled=labelled)
        for _ in range(40):
            msc.round(5, 2)
        samples = []
        for _ in range(60):
            msc.round(4, 2)
            E = msc.energies()
            if labelled:  # energies come by slot; put them in temperature order
                lab = msc.labels().astype(np.int64)
                Eb = np.empty_like(E)
                np.put_along_axis(Eb, lab, E, axis=0)
                E = Eb
            samples.append(E)
        S = np.array(samples)
        for b in range(len(betas)):
            per_ladder = S[:, b, :].mean(axis=0)
            mean, err = per_ladder.mean(), per_ladder.std(ddof=1) / np.sqrt(per_ladder.size)
            assert abs(mean - exact[b][0]) <= 3.0 * err + 1e-9, (steps, merged, labelled, betas[b], mean, exact[b][0], err)
        msc.close()


def test_block_fetch_widens_in_the_requested_order(nl):
    """nlmc_host_fetch_widen_blocks: device int8 blocks -> host float64 blocks through the pinned staging buffer, with
    and without a block permutation (how the recorded states of a sharded ladder reach M in temperature order)."""
    import torch
    rs = np.random.RandomState(3)
    for n_blocks, elems in ((1, 1000003), (7, 4099), (32, 65536)):
        src = rs.randint(-1, 2, size=(n_blocks, elems)).astype(np.int8)
        d = torch.from_numpy(src).cuda()
        out = np.full((n_blocks, elems), np.nan)
        nl.lib.fetch_widen_blocks(d.data_ptr(), out, n_blocks, elems)
        torch.cuda.synchronize()
        assert np.array_equal(out, src.astype(np.float64))
        perm = rs.permutation(n_blocks).astype(np.int32)
        out2 = np.full((n_blocks, elems), np.nan)
        nl.lib.fetch_widen_blocks(d.data_ptr(), out2, n_blocks, elems, dst_block=perm)
        torch.cuda.synchronize()
        assert np.array_equal(out2[perm], src.astype(np.float64))


def graph_from_edges(n, edges, seed):
    rs = np.random.RandomState(seed)
    rows, cols, vals = [], [], []
    for (i, j) in edges:
        v = rs.choice([-1.0, 1.0])
        rows += [i, j]
        cols += [j, i]
        vals += [v, v]
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr(), np.zeros(n)


def open_lattice_2d(Lx, Ly, seed):
    """Open boundaries: corner sites have degree 2, edge sites 3, inner sites 4."""
    edges = []
    for y in range(Ly):
        for x in range(Lx):
            i = x + Lx * y
            if x + 1 < Lx:
                edges.append((i, i + 1))
            if y + 1 < Ly:
                edges.append((i, i + Lx))
    return graph_from_edges(Lx * Ly, edges, seed)


def chimera_cell_pair(seed):
    """Two Chimera unit cells (K4,4 each) joined by four couplers: degrees 4 and 5; a path of three extra spins hangs off
    (degrees 1, 2 and a degree-6 hub is made by three more pendant spins)."""
    edges = [(a, 4 + b) for a in range(4) for b in range(4)]                    # cell 0: 0..3 | 4..7
    edges += [(8 + a, 12 + b) for a in range(4) for b in range(4)]              # cell 1: 8..11 | 12..15
    edges += [(4 + b, 12 + b) for b in range(4)]                                # inter-cell couplers: degree 5
    return graph_from_edges(16, edges, seed)


@pytest.mark.parametrize("case", ["open_4x4", "open_3x5", "chimera", "star"])
def test_exact_boltzmann_with_odd_degrees(nl, case):
    """Sites of odd degree (open boundaries, Chimera couplers, pendant spins): fields f = 2c - deg are odd there, with
    their own threshold levels |f| = 1, 3, 5; sites of even and odd degree are separate launch classes.  <E>(beta) must
    equal full enumeration, energies must be exact, in the scalar-threshold and the beta-label form."""
    from oracle import oracle as O
    if case == "open_4x4":
        A, h = open_lattice_2d(4, 4, 3)
    elif case == "open_3x5":
        A, h = open_lattice_2d(3, 5, 4)
    elif case == "chimera":
        A, h = chimera_cell_pair(5)
    else:  # a hub of degree 5 with pendant spins (degree 1) and a tail (degrees 2, 1)
        A, h = graph_from_edges(8, [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (5, 6), (6, 7)], 6)
    deg = np.diff(A.indptr)
    assert np.any(deg & 1)
    betas = np.array([0.1, 0.45, 0.9, 1.7])
    exact = exact_mean_energy(A, betas)
    prob = nl.host.Problem(A, h)
    csr = O.Csr(A)
    for labelled in (False, True):
        msc = nl.lib.Msc(prob.inst, betas, 1024, seed=31 + labelled, labelled=labelled)
        E0 = msc.energies()
        for b, lad in ((0, 0), (1, 33), (3, 1023)):
            assert E0[b, lad] == O.energy(csr, h, msc.get_spins(b, lad))[0]
        for _ in range(40):
            msc.round(5, 2)
        samples = []
        for _ in range(60):
            msc.round(4, 2)
            E = msc.energies()
            if labelled:
                lab = msc.labels().astype(np.int64)
                Eb = np.empty_like(E)
                np.put_along_axis(Eb, lab, E, axis=0)
                E = Eb
            samples.append(E)
        S = np.array(samples)
        for b in range(len(betas)):
            per_ladder = S[:, b, :].mean(axis=0)
            mean, err = per_ladder.mean(), per_ladder.std(ddof=1) / np.sqrt(per_ladder.size)
            assert abs(mean - exact[b][0]) <= 3.0 * err + 1e-9, (case, labelled, betas[b], mean, exact[b][0], err)
        msc.close()


def test_conditional_distribution_on_an_open_lattice(nl):
    """Single-sweep conditional law on a 3D open-boundary lattice (degrees 3..6): colour-0 sites of every degree are drawn
    from P(+1) = 1/(1 + exp(-2 beta f)) with f from the unchanged colour-1 neighbours."""
    L = 6
    edges = []
    for z in range(L):
        for y in range(L):
            for x in range(L):
                i = x + L * (y + L * z)
                if x + 1 < L:
                    edges.append((i, i + 1))
                if y + 1 < L:
                    edges.append((i, i + L))
                if z + 1 < L:
                    edges.append((i, i + L * L))
    A, h = graph_from_edges(L ** 3, edges, 9)
    n = L ** 3
    betas = np.array([0.3, 0.8])
    prob = nl.host.Problem(A, h)
    msc = nl.lib.Msc(prob.inst, betas, 2048, seed=5)
    assert msc.n_colours == 2
    rs = np.random.RandomState(4)
    s0 = rs.choice([-1, 1], size=n).astype(np.int8)
    P = np.where(s0[:, None] > 0, np.uint32(0xffffffff), np.uint32(0)) * np.ones((1, msc.n_words), dtype=np.uint32)
    msc.set_packed(np.ascontiguousarray(P.astype(np.uint32)))
    msc.sweep(1)
    out = msc.get_packed()
    idx = np.arange(n)
    colour0 = ((idx % L) + (idx // L) % L + idx // (L * L)) % 2 == 0
    f0 = np.asarray(A @ s0.astype(float))
    G = msc.n_ladders // 32
    seen = set()
    for b, beta in enumerate(betas):
        ups = np.zeros(n)
        for g in range(G):
            ups += np.array([bin(int(v)).count("1") for v in out[:, b * G + g]])
        for f in range(-6, 7):
            sel = colour0 & (f0 == f)
            trials = sel.sum() * msc.n_ladders
            if trials == 0:
                continue
            seen.add(f)
            p = 1.0 / (1.0 + np.exp(-2 * beta * f))
            sigma = np.sqrt(trials * p * (1 - p))
            assert abs(ups[sel].sum() - trials * p) <= 3.0 * sigma + 1, (beta, f, ups[sel].sum(), trials * p, sigma)
    assert {-5, -3, -1, 1, 3, 5} & seen and {-4, -2, 0, 2, 4} & seen


@pytest.mark.parametrize("labelled", [False, True])
@pytest.mark.parametrize("case", ["periodic", "open"])
def test_cooperative_batch_kernel_equals_launch_chain(nl, monkeypatch, labelled, case):
    """NLMC_MSC_BATCH=1 runs a batch of sweeps as ONE cooperative launch (persistent CTAs, a grid barrier where the launch
    boundary was): same random streams, so the packed state after the batch must equal the launch chain's bit for bit --
    with a ragged last tile (n per colour not a multiple of 256) and with launch classes of both degree parities."""
    from oracle import oracle as O
    if case == "periodic":
        A, h = O.ea3d_pm_j(10, 3)        # 500 sites per colour: one full tile + a ragged one
    else:
        A, h = open_lattice_2d(37, 19, 2)  # degrees 2, 3, 4: even and odd classes in both colours
    prob = nl.host.Problem(A, h)
    betas = np.linspace(0.3, 1.6, 6)
    kw = dict(labelled=True, slot_begin=1, slot_count=4) if labelled else {}
    out = []
    for batch in ("0", "1"):
        monkeypatch.setenv("NLMC_MSC_BATCH", batch)
        m = nl.lib.Msc(prob.inst, betas, 256, seed=11, **kw)
        m.sweep(5)
        m.sweep(1)
        out.append(m.get_packed())
        m.close()
    assert np.array_equal(out[0], out[1])
