"""ctypes front-end of the CPU oracle (oracle/arianna_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (montecarlo_b200/) must never import this module.

PARITY UNPINNED for the reference's own arithmetic (see the header of arianna_oracle.c): the reference is pure Julia
and cannot run here.  Its third-party random stream (Julia's Random stdlib: Xoshiro seeding, rand, randn) IS pinned, by
the known answers printed in the Julia manual (tests/test_oracle.py::test_julia_rng_known_answers).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

POT_HARMONIC, POT_QUARTIC, POT_DOUBLE_WELL = 0, 1, 2
OPT_STATIC, OPT_VPG, OPT_BLPG, OPT_BLAPG, OPT_NPG, OPT_ANPG, OPT_BLANPG = range(7)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the recipe in oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp)."""
    srcs = [os.path.join(_HERE, f) for f in ("arianna_oracle.c", "zig_tables_julia.h", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ao_callback_energy.restype = C.c_double
        _lib.ao_lognorm.restype = C.c_double
        _lib.ao_dlogq_dsigma.restype = C.c_double
        _lib.ao_log_proposal_density.restype = C.c_double
        _lib.ao_learning_step.restype = C.c_double
        _lib.ao_xoshiro_next.restype = C.c_uint64
        _lib.ao_xoshiro_rand.restype = C.c_double
        _lib.ao_xoshiro_randn.restype = C.c_double
        _lib.ao_num_threads.restype = C.c_int
        _lib.ao_lognorm_f32.restype = C.c_double
        _lib.ao_callback_energy_f32.restype = C.c_float
    return _lib


def _p(a, ty):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ty))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().ao_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def lognorm(sigma):
    """log((2π)*(σ*σ))/2 per move, the constant of particle_1d.jl:53."""
    return np.array([lib().ao_lognorm(C.c_double(float(s))) for s in np.atleast_1d(sigma)], dtype=np.float64)


def init_synthetic(seed, chain_offset, M):
    x = np.empty(M, dtype=np.float64)
    lib().ao_init_synthetic(C.c_int64(seed), C.c_int64(chain_offset), C.c_int64(M), _p(x, C.c_double))
    return x


def draws_philox(seed, chain_offset, M, t0, K, with_cat=True):
    """Native-mode Metropolis draws as step-major [K][M] arrays (u_cat, z, u_acc)."""
    z = np.empty((K, M), dtype=np.float64)
    ua = np.empty((K, M), dtype=np.float64)
    uc = np.empty((K, M), dtype=np.float64) if with_cat else None
    lib().ao_draws_philox(C.c_int64(seed), C.c_int64(chain_offset), C.c_int64(M), C.c_int64(t0), C.c_int64(K),
                          _p(uc, C.c_double), _p(z, C.c_double), _p(ua, C.c_double))
    return uc, z, ua


def draws_philox_f32(seed, chain_offset, M, t0, K):
    """The engine's native Float32 stream restated with libm (loose: the device uses MUFU Box-Muller): (z, u_acc)."""
    z = np.empty((K, M), dtype=np.float64)
    ua = np.empty((K, M), dtype=np.float64)
    lib().ao_draws_philox_f32(C.c_int64(seed), C.c_int64(chain_offset), C.c_int64(M), C.c_int64(t0), C.c_int64(K),
                              _p(z, C.c_double), _p(ua, C.c_double))
    return z, ua


def draws_pgmc_philox(seed, chain_offset, M, q0, n):
    z = np.empty((n, M), dtype=np.float64)
    lib().ao_draws_pgmc_philox(C.c_int64(seed), C.c_int64(chain_offset), C.c_int64(M), C.c_int64(q0),
                               C.c_int64(n), _p(z, C.c_double))
    return z


class Ensemble:
    """M chains of the particle_1d system + one pool of Gaussian displacement moves, on the CPU.

    Mirrors the state the reference keeps per chain: Particle(x, β, e) (particle_1d.jl:9-16) and
    Move.accepted_calls / total_calls (metropolis.jl:140-147)."""

    def __init__(self, x0, beta, sigma, weight=None, potential=POT_HARMONIC):
        self.x = _f64(x0).copy()
        self.M = self.x.size
        self.pot = int(potential)
        self.e = self._potential(self.x)
        self.beta = float(beta)
        self.sigma = _f64(np.atleast_1d(sigma)).copy()
        self.n_moves = self.sigma.size
        self.weight = _f64(np.atleast_1d(weight if weight is not None else [1.0] * self.n_moves)).copy()
        self.acc = np.zeros((self.n_moves, self.M), dtype=np.int64)
        self.tot = np.zeros((self.n_moves, self.M), dtype=np.int64)

    def _potential(self, x):
        if self.pot == POT_HARMONIC:
            return x * x
        if self.pot == POT_QUARTIC:
            x2 = x * x
            return x2 * x2
        w = x * x - 1.0
        return w * w

    def sweep_replay(self, u_cat, z, u_acc, want_decisions=False, want_alpha=False, betas=None):
        z = _f64(z)
        u_acc = _f64(u_acc)
        K = z.shape[0]
        assert z.shape == (K, self.M) and u_acc.shape == (K, self.M)
        if u_cat is not None:
            u_cat = _f64(u_cat)
            assert u_cat.shape == (K, self.M)
        else:
            assert self.n_moves == 1, "u_cat is required for multi-move pools"
        dec = np.empty((K, self.M), dtype=np.uint8) if want_decisions else None
        mov = np.empty((K, self.M), dtype=np.uint8) if want_decisions else None
        alp = np.empty((K, self.M), dtype=np.float64) if want_alpha else None
        ln = lognorm(self.sigma)
        if betas is None:
            lib().ao_sweep_replay(C.c_int64(self.M), C.c_int64(K), _p(self.x, C.c_double), _p(self.e, C.c_double),
                                  C.c_double(self.beta), C.c_int(self.pot), C.c_int(self.n_moves),
                                  _p(self.sigma, C.c_double), _p(self.weight, C.c_double), _p(ln, C.c_double),
                                  _p(u_cat, C.c_double), _p(z, C.c_double), _p(u_acc, C.c_double),
                                  _p(self.acc, C.c_int64), _p(self.tot, C.c_int64), _p(dec, C.c_uint8),
                                  _p(mov, C.c_uint8), _p(alp, C.c_double))
        else:
            betas = _f64(betas)
            lib().ao_sweep_replay_betas(C.c_int64(self.M), C.c_int64(K), _p(self.x, C.c_double),
                                        _p(self.e, C.c_double), _p(betas, C.c_double), C.c_int(self.pot),
                                        C.c_int(self.n_moves), _p(self.sigma, C.c_double),
                                        _p(self.weight, C.c_double), _p(ln, C.c_double), _p(u_cat, C.c_double),
                                        _p(z, C.c_double), _p(u_acc, C.c_double), _p(self.acc, C.c_int64),
                                        _p(self.tot, C.c_int64), _p(dec, C.c_uint8))
        return dec, mov, alp

    def callback_energy(self):
        return lib().ao_callback_energy(C.c_int64(self.M), _p(self.e, C.c_double))

    def callback_acceptance(self):
        out = np.empty(self.n_moves, dtype=np.float64)
        with np.errstate(all="ignore"):
            lib().ao_callback_acceptance(C.c_int64(self.M), C.c_int(self.n_moves), _p(self.acc, C.c_int64),
                                         _p(self.tot, C.c_int64), _p(out, C.c_double))
        return out

    def pgmc_replay(self, q_batch, learn_ids, z):
        """z: [n_learn][q_batch][M].  Returns summed GradientData records [n_learn][5] = (j, ∇j, ∇logq_f, g, n)."""
        ids = np.ascontiguousarray(learn_ids, dtype=np.int32)
        z = _f64(z)
        assert z.shape == (ids.size, q_batch, self.M)
        gd = np.zeros((ids.size, 5), dtype=np.float64)
        ln = lognorm(self.sigma)
        lib().ao_pgmc_replay(C.c_int64(self.M), C.c_int(q_batch), C.c_int(ids.size), _p(ids, C.c_int32),
                             _p(self.x, C.c_double), _p(self.e, C.c_double), C.c_double(self.beta), C.c_int(self.pot),
                             _p(self.sigma, C.c_double), _p(ln, C.c_double), _p(z, C.c_double), _p(gd, C.c_double))
        return gd

    # --- "Julia-like" generator front-end (CPU baseline / replay-stream manufacture) ---------------------
    def seed_xoshiro(self, seed, chain_offset=0, julia=False):
        """Per-chain generator states for seeds seed + c - 1 (metropolis.jl:262-263).  julia=True: Julia 1.7-1.10's
        own SHA-256 seeding of Xoshiro(n) [EXT, pinned by the Julia manual's known answers]; default: a splitmix64 stand-in."""
        self.states = np.zeros((self.M, 4), dtype=np.uint64)
        fn = lib().ao_xoshiro_seed_chains_julia if julia else lib().ao_xoshiro_seed_chains
        fn(C.c_int64(seed), C.c_int64(chain_offset), C.c_int64(self.M), _p(self.states, C.c_uint64))

    def sweep_xoshiro(self, K):
        ln = lognorm(self.sigma)
        lib().ao_sweep_xoshiro(C.c_int64(self.M), C.c_int64(K), _p(self.x, C.c_double), _p(self.e, C.c_double),
                               C.c_double(self.beta), C.c_int(self.pot), C.c_int(self.n_moves),
                               _p(self.sigma, C.c_double), _p(self.weight, C.c_double), _p(ln, C.c_double),
                               _p(self.states, C.c_uint64), _p(self.acc, C.c_int64), _p(self.tot, C.c_int64))

    def draws_xoshiro(self, K):
        uc = np.empty((K, self.M), dtype=np.float64)
        z = np.empty((K, self.M), dtype=np.float64)
        ua = np.empty((K, self.M), dtype=np.float64)
        lib().ao_draws_xoshiro(C.c_int64(self.M), C.c_int64(K), _p(self.states, C.c_uint64), _p(uc, C.c_double),
                               _p(z, C.c_double), _p(ua, C.c_double))
        return uc, z, ua

    def pgmc_xoshiro(self, q_batch, learn_ids):
        ids = np.ascontiguousarray(learn_ids, dtype=np.int32)
        gd = np.zeros((ids.size, 5), dtype=np.float64)
        ln = lognorm(self.sigma)
        lib().ao_pgmc_xoshiro(C.c_int64(self.M), C.c_int(q_batch), C.c_int(ids.size), _p(ids, C.c_int32),
                              _p(self.x, C.c_double), _p(self.e, C.c_double), C.c_double(self.beta),
                              C.c_int(self.pot), _p(self.sigma, C.c_double), _p(ln, C.c_double),
                              _p(self.states, C.c_uint64), _p(gd, C.c_double))
        return gd


class Ensemble32:
    """M chains of Particle{Float32} with ONE Gaussian displacement move with Float32 parameters (σ = 0.1f0): the
    Float32 instantiation of the reference's generic code (particle_1d.jl:9-16; promotion rules in arianna_oracle.c)."""

    def __init__(self, x0, beta, sigma, potential=POT_HARMONIC):
        self.x = np.ascontiguousarray(x0, dtype=np.float32).copy()
        self.M = self.x.size
        self.pot = int(potential)
        x = self.x
        self.e = (x * x if self.pot == POT_HARMONIC else (x * x) * (x * x) if self.pot == POT_QUARTIC
                  else (x * x - np.float32(1)) * (x * x - np.float32(1))).astype(np.float32)
        self.beta = np.float32(beta)
        self.sigma = np.float32(sigma)
        self.acc = np.zeros(self.M, dtype=np.int64)
        self.tot = 0

    def sweep_replay(self, z, u_acc, want_decisions=False, betas=None):
        z, u_acc = _f64(z), _f64(u_acc)
        K = z.shape[0]
        assert z.shape == (K, self.M) and u_acc.shape == (K, self.M)
        dec = np.empty((K, self.M), dtype=np.uint8) if want_decisions else None
        b = None if betas is None else np.ascontiguousarray(betas, dtype=np.float32)
        lib().ao_sweep_replay_f32(C.c_int64(self.M), C.c_int64(K), _p(self.x, C.c_float), _p(self.e, C.c_float),
                                  C.c_float(self.beta), _p(b, C.c_float), C.c_int(self.pot), C.c_float(self.sigma),
                                  _p(z, C.c_double), _p(u_acc, C.c_double), _p(self.acc, C.c_int64), _p(dec, C.c_uint8))
        self.tot += K
        return dec

    def callback_energy(self):
        return float(lib().ao_callback_energy_f32(C.c_int64(self.M), _p(self.e, C.c_float)))

    def callback_acceptance(self):
        with np.errstate(all="ignore"):
            return float(np.sum(self.acc / np.float64(self.tot)) / self.M) if self.tot else float("nan")


def learning_step(kind, p1, p2, gd_avg, theta):
    g = _f64(gd_avg)
    return lib().ao_learning_step(C.c_int(kind), C.c_double(p1), C.c_double(p2), _p(g, C.c_double),
                                  C.c_double(theta))


def dlogq_dsigma(delta, sigma):
    return lib().ao_dlogq_dsigma(C.c_double(delta), C.c_double(sigma))


def log_proposal_density(delta, sigma):
    return lib().ao_log_proposal_density(C.c_double(delta), C.c_double(sigma))


def xoshiro_next(state):
    s = np.ascontiguousarray(state, dtype=np.uint64)
    r = lib().ao_xoshiro_next(_p(s, C.c_uint64))
    return int(r), s


def xoshiro_seed(seed):
    s = np.zeros(4, dtype=np.uint64)
    lib().ao_xoshiro_seed(C.c_uint64(seed), _p(s, C.c_uint64))
    return s


def xoshiro_seed_julia(seed):
    s = np.zeros(4, dtype=np.uint64)
    lib().ao_xoshiro_seed_julia(C.c_uint64(seed), _p(s, C.c_uint64))
    return s


def sha256(msg: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    buf = (C.c_uint8 * max(1, len(msg))).from_buffer_copy(msg or b"\0")
    lib().ao_sha256(buf, C.c_int64(len(msg)), out)
    return bytes(out)


def xoshiro_randn_stream(state, n):
    s = np.ascontiguousarray(state, dtype=np.uint64).copy()
    return np.array([lib().ao_xoshiro_randn(_p(s, C.c_uint64)) for _ in range(n)])


def ziggurat_tables():
    ki = np.zeros(256, dtype=np.uint64)
    wi = np.zeros(256, dtype=np.float64)
    fi = np.zeros(256, dtype=np.float64)
    lib().ao_ziggurat_tables(_p(ki, C.c_uint64), _p(wi, C.c_double), _p(fi, C.c_double))
    return ki, wi, fi


def num_threads():
    return lib().ao_num_threads()


def set_num_threads(n: int):
    lib().ao_set_num_threads(C.c_int(int(n)))
