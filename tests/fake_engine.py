"""Test double with CudaEnsemble's interface, backed by the CPU oracle (native-Philox draws + A.1 replay).

Lives in tests/ on purpose: it lets the HOST logic (lazy flush, event order, sharding, all-reduce combination,
file formats, learner loop) run on the CPU-only CI box and under gloo with world_size 2.  It is never importable
from the product package."""
import numpy as np

from oracle import oracle as O

_POT = {"harmonic": O.POT_HARMONIC, "quartic": O.POT_QUARTIC, "double_well": O.POT_DOUBLE_WELL}


class OracleEngine:
    def __init__(self, n_chains, beta, sigma, weight=None, *, seed=1, chain_offset=0, n_chains_total=0,
                 potential="harmonic", rng="philox", arith="fast", device=-1, stream=0):
        self.n_chains, self.chain_offset = int(n_chains), int(chain_offset)
        self.n_chains_total = int(n_chains_total) or self.n_chains
        self.seed, self.beta = int(seed), float(beta)
        self.n_moves = len(sigma)
        self.ens = O.Ensemble(np.zeros(self.n_chains), beta, sigma, weight, potential=_POT[potential])
        self.steps_done = 0
        self.pgmc_samples = 0
        self.gd = np.zeros((16, 5))
        self.launch_count = 0
        self.betas = None

    def set_betas(self, betas):
        self.betas = np.ascontiguousarray(betas, dtype=np.float64)

    def close(self):
        pass

    def set_state(self, x):
        self.ens.x[:] = x
        self.ens.e[:] = self.ens._potential(self.ens.x)

    def init_synthetic(self, seed=None):
        self.set_state(O.init_synthetic(self.seed if seed is None else seed, self.chain_offset, self.n_chains))

    def get_state(self, with_energy=False):
        return (self.ens.x.copy(), self.ens.e.copy()) if with_energy else self.ens.x.copy()

    def set_params(self, k, sigma, log_norm=None):
        if not (sigma > 0 and np.isfinite(sigma)):
            raise ValueError("σ must be finite and > 0")
        self.ens.sigma[k] = sigma

    def get_params(self, k):
        return float(self.ens.sigma[k])

    def sweep(self, K, reduce=False):
        if K:
            uc, z, ua = O.draws_philox(self.seed, self.chain_offset, self.n_chains, self.steps_done, K,
                                       with_cat=self.n_moves > 1)
            self.ens.sweep_replay(uc, z, ua, betas=self.betas)
            self.steps_done += K
            self.launch_count += 1

    def sweep_series(self, Ks, read=True):
        """arianna_sweep_series: len(Ks) store intervals in ONE call (counted as one launch), a record after each."""
        n0 = self.launch_count
        rec = np.empty((len(Ks), 2 + self.n_moves))
        for i, K in enumerate(Ks):
            self.sweep(int(K))
            rec[i] = self.callback_sums()
        self.launch_count = n0 + 1
        self.series_calls = getattr(self, "series_calls", 0) + 1
        return rec

    def callback_sums(self):
        with np.errstate(all="ignore"):
            r = (self.ens.acc / self.ens.tot).sum(axis=1)
        return np.concatenate([[self.ens.e.sum()], r, [float(self.n_chains)]])

    def callbacks(self):
        s = self.callback_sums()
        return s[0] / s[-1], s[1:-1] / s[-1]

    def counters(self):
        return self.ens.acc.sum(axis=1), self.ens.tot.sum(axis=1)

    def pgmc_estimate(self, q_batch, learn_ids):
        n = len(learn_ids) * q_batch
        z = O.draws_pgmc_philox(self.seed, self.chain_offset, self.n_chains, self.pgmc_samples, n)
        self.gd[:len(learn_ids)] += self.ens.pgmc_replay(q_batch, list(learn_ids), z.reshape(len(learn_ids), q_batch, -1))
        self.pgmc_samples += n

    def pgmc_read(self, n):
        return self.gd[:n].copy()

    def pgmc_reset(self):
        self.gd[:] = 0

    def synchronize(self):
        pass
