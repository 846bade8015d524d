"""TEST INFRASTRUCTURE ONLY: oracle-backed stand-ins for the ctypes classes of nlmc_b200._lib, so that
the HOST logic of the drop-in classes (random-stream bookkeeping, phase set-up, swap rules, output
assembly) can be tested on a machine without a GPU.  Installed by the `fake_device` fixture only;
the product never sees it."""
import numpy as np

from oracle import oracle as O


class FakeInstance:
    def __init__(self, rp, ci, val, h, device=0):
        self.csr = O.Csr.__new__(O.Csr)
        self.csr.n = len(rp) - 1
        self.csr.rp = np.ascontiguousarray(rp, dtype=np.int32)
        self.csr.ci = np.ascontiguousarray(ci, dtype=np.int32)
        self.csr.val = np.ascontiguousarray(val, dtype=np.float64)
        self.csr._rev = None
        self.h = np.asarray(h, dtype=np.float64).reshape(-1)
        self.n = self.csr.n
        self.nnz = len(self.csr.ci)
        self.is_integer = bool(np.all(self.csr.val == np.floor(self.csr.val)))

    def energy_states(self, states):
        return O.energy(self.csr, self.h, states)

    def close(self):
        pass


class FakeReplicas:
    def __init__(self, inst, n_replicas, init_spins=None):
        self.inst, self.R, self.n = inst, int(n_replicas), inst.n
        self.spins = np.ones((self.R, self.n), dtype=np.int8)
        if init_spins is not None:
            self.spins[:] = np.asarray(init_spins, dtype=np.int8).reshape(self.R, self.n)
        self.phase = [(None, None, 1.0)] * self.R

    def set_spins(self, spins, first=0):
        spins = np.asarray(spins, dtype=np.int8).reshape(-1, self.n)
        self.spins[first:first + len(spins)] = spins

    def get_spins(self, first=0, count=None):
        count = self.R - first if count is None else count
        return self.spins[first:first + count].copy()

    def set_phase(self, r, h_eff=None, row_scaled=None, temp_x=1.0):
        self.phase[r] = (None if h_eff is None else np.asarray(h_eff, dtype=np.float64).copy(),
                         None if row_scaled is None else np.asarray(row_scaled).astype(bool), float(temp_x))

    def sweep_replay(self, perm, u, beta, tanh_lut=None, lut_half=0, record_from=0, want_energy=True):
        perm = np.asarray(perm).reshape(self.R, -1, self.n)
        S = perm.shape[1]
        u = np.asarray(u).reshape(self.R, S, self.n)
        beta = np.asarray(beta, dtype=np.float64).reshape(self.R, S)
        M = np.empty((self.R, S - (record_from or 0), self.n), dtype=np.int8)
        E = np.empty((self.R, S))
        csr = self.inst.csr
        rows = csr.row_of
        for r in range(self.R):
            h_eff, scaled, tx = self.phase[r]
            c = csr if scaled is None else csr.with_values(np.where(scaled[rows], csr.val / tx, csr.val))
            Mo, last = O.mcmc(c, self.inst.h if h_eff is None else h_eff, self.spins[r], beta[r], perm=perm[r], u=u[r])
            self.spins[r] = last
            M[r] = Mo[record_from or 0:]
            E[r] = O.energy(csr, self.inst.h, Mo) if S else np.zeros(0)
        return (None if record_from is None else M), (E if want_energy else None)

    def energy(self):
        return O.energy(self.inst.csr, self.inst.h, self.spins)

    def close(self):
        pass


class FakeLbp:
    def __init__(self, inst):
        self.inst = inst
        self.eps = np.abs(inst.h) + O._pairwise_rowsum_abs(inst.csr)

    def epsilon(self):
        return self.eps.copy()

    def reset(self, m_star):
        c = self.inst.csr
        self.ms = np.asarray(m_star, dtype=np.float64).reshape(-1)
        self.u = np.ascontiguousarray(c.val * self.ms[c.ci])
        self.hm = np.zeros_like(self.u)
        self.tot = np.zeros(c.n)

    def step(self, lam, beta, tol, max_iter):
        hl = np.ascontiguousarray(self.inst.h + lam * self.ms * self.eps)
        self.marg, it = O.lbp(self.inst.csr, hl, beta, self.u, self.hm, self.tot, tol, max_iter)
        return self.marg, it

    def run(self, h_field, beta, tol, max_iter):
        hl = np.ascontiguousarray(np.asarray(h_field, dtype=np.float64).reshape(-1))
        self.marg, it = O.lbp(self.inst.csr, hl, beta, self.u, self.hm, self.tot, tol, max_iter)
        return self.marg, it

    def set_messages(self, h_edge, u_edge, tot):
        self.hm = np.array(h_edge, dtype=np.float64)
        self.u = np.array(u_edge, dtype=np.float64)
        self.tot = np.array(tot, dtype=np.float64)

    def get_messages(self):
        return self.hm.copy(), self.u.copy(), self.tot.copy()

    def byproducts(self, beta, want_corr=True, want_J_tilde=True):
        corr, ht, jt = O.lbp_byproducts(self.inst.csr, beta, self.hm, self.tot, self.marg)
        return (corr if want_corr else None), ht, (jt if want_J_tilde else None)

    def close(self):
        pass


def fake_icm_clusters(inst, s1, s2):
    s1 = np.asarray(s1, dtype=np.int8).reshape(-1, inst.n)
    s2 = np.asarray(s2, dtype=np.int8).reshape(-1, inst.n)
    labels = np.empty((len(s1), inst.n), dtype=np.int32)
    counts = np.empty(len(s1), dtype=np.int32)
    for p in range(len(s1)):
        labels[p], counts[p] = O.disagreement_clusters(inst.csr, s1[p], s2[p])
    return labels, counts


def install(monkeypatch):
    from nlmc_b200 import _lib
    monkeypatch.setattr(_lib, "Instance", FakeInstance)
    monkeypatch.setattr(_lib, "Replicas", FakeReplicas)
    monkeypatch.setattr(_lib, "Lbp", FakeLbp)
    monkeypatch.setattr(_lib, "icm_clusters", fake_icm_clusters)


class FakeEngine:
    """Oracle-backed stand-in for the production engines (_lib.Col / _lib.Dense API): R rows, sequential heat-bath
    sweeps from a private RandomState, per-site NMC modes, best-state tracking -- statistically the same sampler, used
    to exercise the HOST logic of production.py without a GPU."""

    def __init__(self, prob, betas, seed):
        self.prob, self.csr, self.h = prob, prob.inst.csr, prob.inst.h
        self.betas = np.asarray(betas, dtype=np.float64).reshape(-1).copy()
        self.R, self.n = len(self.betas), prob.n
        self.rs = np.random.RandomState(seed % (2 ** 31))
        self.spins = self.rs.choice([-1, 1], size=(self.R, self.n)).astype(np.int8)
        self.modes, self.temp_x = None, 1.0
        self.best_E, self.best_S = np.full(self.R, np.inf), self.spins.copy()

    def set_betas(self, betas):
        self.betas = np.asarray(betas, dtype=np.float64).reshape(-1).copy()

    # replica exchange by beta labels (nlmc_col_ladders / _exchange / _labels): rows = ladders x n_beta
    def ladders(self, betas):
        self.lad_betas = np.asarray(betas, dtype=np.float64).reshape(-1).copy()
        nb = len(self.lad_betas)
        assert self.R % nb == 0
        self.lab = np.tile(np.arange(nb, dtype=np.int32), self.R // nb)
        self.betas = self.lad_betas[self.lab].copy()
        self.x_counts = []

    def exchange(self, num_pairs):
        nb = len(self.lad_betas)
        E = self.energies().reshape(-1, nb)
        lab = self.lab.reshape(-1, nb)
        acc = 0
        for l in range(lab.shape[0]):
            avail = list(range(nb - 1))
            for _ in range(num_pairs):
                if not avail:
                    break
                i = avail[self.rs.randint(len(avail))]
                avail = [j for j in avail if abs(j - i) > 1]
                sa, sb = int(np.where(lab[l] == i)[0][0]), int(np.where(lab[l] == i + 1)[0][0])
                x = (self.lad_betas[i + 1] - self.lad_betas[i]) * (E[l, sb] - E[l, sa])
                if self.rs.rand() < min(1.0, np.exp(x)):
                    lab[l, sa], lab[l, sb] = i + 1, i
                    acc += 1
        self.betas = self.lad_betas[self.lab].copy()
        self.x_counts.append(acc)

    def labels(self, n_rounds=0):
        cnt = np.zeros(int(n_rounds), dtype=np.int32)
        tail = self.x_counts[-int(n_rounds):] if n_rounds else []
        if tail:
            cnt[-len(tail):] = tail
        return self.lab.copy(), cnt

    def set_spins(self, spins):
        self.spins = np.asarray(spins, dtype=np.int8).reshape(self.R, self.n).copy()

    def get_spins(self):
        return self.spins.copy()

    def set_site_modes(self, modes, temp_x=1.0):
        self.modes = None if modes is None else np.asarray(modes, dtype=np.uint8).reshape(self.R, self.n).copy()
        self.temp_x = float(temp_x)

    def best_reset(self):
        self.best_E[:] = np.inf

    def best_get(self):
        return self.best_S.copy(), self.best_E.copy()

    def energies(self):
        return O.energy(self.csr, self.h, self.spins)

    def _one_sweep(self, betas):
        c = self.csr
        for r in range(self.R):
            s = self.spins[r]
            for i in self.rs.permutation(self.n):
                md = 0 if self.modes is None else self.modes[r, i]
                if md == 2:
                    continue
                b, e = c.rp[i], c.rp[i + 1]
                f = float(c.val[b:e] @ s[c.ci[b:e]]) + self.h[i]
                beta = betas[r] / self.temp_x if md == 1 else betas[r]
                s[i] = 1 if self.rs.rand() < 1.0 / (1.0 + np.exp(-2.0 * beta * f)) else -1

    def sweep(self, n_sweeps):
        for _ in range(int(n_sweeps)):
            self._one_sweep(self.betas)

    def sweep_record(self, n_sweeps, record_every=1, track_best=False, beta_sched=None, want_states=True, want_energies=True):
        states, E = [], np.empty((n_sweeps, self.R))
        for j in range(int(n_sweeps)):
            self._one_sweep(self.betas if beta_sched is None else np.asarray(beta_sched).reshape(n_sweeps, self.R)[j])
            E[j] = self.energies()
            if track_best:
                better = E[j] < self.best_E
                self.best_E[better], self.best_S[better] = E[j][better], self.spins[better]
            if want_states and j % record_every == 0:
                states.append(self.spins.copy())
        return (np.array(states, dtype=np.int8).reshape(-1, self.R, self.n) if want_states else None), (E if want_energies else None)

    def close(self):
        pass


class FakeMsc:
    """Stand-in for _lib.Msc (bit-packed engine): n_beta x n_ladders replicas as plain int8 arrays, sequential heat-bath
    sweeps, energies, replica exchange with the reference's pair selection -- for the HOST logic of the bit-packed
    production paths (NPT.run without NMC replicas, APT_preprocessor, APT_ICM on lattices)."""

    def __init__(self, inst, betas, n_ladders, seed=0, ladder_offset=0):
        self.inst, self.csr, self.h = inst, inst.csr, inst.h
        self.betas = np.asarray(betas, dtype=np.float64).reshape(-1).copy()
        self.n_beta, self.n = len(self.betas), inst.n
        self.n_ladders_requested = int(n_ladders)
        self.n_ladders = ((int(n_ladders) + 127) // 128) * 128
        self.live = min(self.n_ladders, max(1, int(n_ladders)))   # only the requested ladders are simulated
        self.rs = np.random.RandomState(seed % (2 ** 31))
        self.S = self.rs.choice([-1, 1], size=(self.n_beta, self.live, self.n)).astype(np.int8)
        self.accepted = 0
        self.per_round = []

    def set_betas(self, betas):
        self.betas = np.asarray(betas, dtype=np.float64).reshape(-1).copy()

    def get_spins(self, b, lad):
        return self.S[b, lad].copy()

    def set_spins(self, b, lad, spins):
        self.S[b, lad] = np.asarray(spins, dtype=np.int8)

    def _sweep_once(self):
        c = self.csr
        for b in range(self.n_beta):
            for lad in range(self.live):
                s = self.S[b, lad]
                for i in self.rs.permutation(self.n):
                    lo, hi = c.rp[i], c.rp[i + 1]
                    f = float(c.val[lo:hi] @ s[c.ci[lo:hi]]) + self.h[i]
                    s[i] = 1 if self.rs.rand() < 1.0 / (1.0 + np.exp(-2.0 * self.betas[b] * f)) else -1

    def sweep(self, n_sweeps):
        for _ in range(int(n_sweeps)):
            self._sweep_once()

    def energies(self):
        E = np.zeros((self.n_beta, self.n_ladders))
        E[:, :self.live] = O.energy(self.csr, self.h, self.S.reshape(-1, self.n)).reshape(self.n_beta, self.live)
        return E

    def round(self, n_sweeps, num_pairs, fetch_energies=False):
        self.sweep(n_sweeps)
        E = self.energies()
        before = self.accepted
        for lad in range(self.live):
            avail = list(range(self.n_beta - 1))
            for _ in range(num_pairs):
                if not avail:
                    break
                i = avail[self.rs.randint(len(avail))]
                avail = [j for j in avail if abs(j - i) > 1]
                if self.rs.rand() < min(1.0, np.exp((self.betas[i + 1] - self.betas[i]) * (E[i + 1, lad] - E[i, lad]))):
                    self.S[[i, i + 1], lad] = self.S[[i + 1, i], lad]
                    E[[i, i + 1], lad] = E[[i + 1, i], lad]
                    self.accepted += 1
        self.per_round.append(self.accepted - before)
        return E if fetch_energies else None

    def swap_counts(self, n_rounds):
        out = np.zeros(int(n_rounds), dtype=np.int32)
        tail = self.per_round[-int(n_rounds):] if n_rounds else []
        if tail:
            out[-len(tail):] = tail
        return out

    def swap_count(self, reset=False):
        v = self.accepted
        if reset:
            self.accepted = 0
        return v

    def sweep_record(self, n_sweeps, ladder=0, energies=True, rows_of_M=False):
        Mrec = np.empty((n_sweeps, self.n_beta, self.n), dtype=np.int8) if ladder is not None else None
        Erec = np.empty((n_sweeps, self.n_beta, self.n_ladders)) if energies else None
        for j in range(int(n_sweeps)):
            self._sweep_once()
            if Mrec is not None:
                Mrec[j] = self.S[:, ladder]
            if Erec is not None:
                Erec[j] = self.energies()
        if rows_of_M and Mrec is not None:
            Mrec = np.ascontiguousarray(Mrec.transpose(1, 2, 0))
        return Mrec, Erec

    def sweep_record_f64(self, n_sweeps, ladder=0, out=None):
        Mrec, Erec = self.sweep_record(n_sweeps, ladder=ladder, energies=True, rows_of_M=True)
        if out is not None:
            out.reshape(Mrec.shape)[:] = Mrec
            return out, Erec
        return Mrec.astype(np.float64), Erec

    def close(self):
        pass
