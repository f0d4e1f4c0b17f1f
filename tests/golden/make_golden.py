"""Generates tests/golden/*.npz from the CPU oracle (oracle/arianna_oracle.c).

The reference holds no golden vectors for this path and cannot be executed here (pure Julia, no `julia` binary),
so these fixtures pin the ORACLE's behaviour (regression + cross-implementation anchor for the numpy restatement
and the CUDA engine); they are NOT outputs of the reference.  PARITY UNPINNED -- see oracle/arianna_oracle.c.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def replay_case(name, M, K, beta, sigma, weight, pot, seed):
    """Julia-like xoshiro stream -> replay arrays -> oracle sweep; everything needed to re-run is stored."""
    x0 = O.init_synthetic(seed, 0, M)
    gen = O.Ensemble(x0, beta, sigma, weight, potential=pot)
    gen.seed_xoshiro(seed)
    uc, z, ua = gen.draws_xoshiro(K)
    ens = O.Ensemble(x0, beta, sigma, weight, potential=pot)
    dec, mov, alp = ens.sweep_replay(uc, z, ua, want_decisions=True, want_alpha=True)
    np.savez_compressed(os.path.join(HERE, name), x0=x0, beta=beta, sigma=np.array(sigma), weight=np.array(weight),
                        pot=pot, seed=seed, u_cat=uc, z=z, u_acc=ua, x=ens.x, e=ens.e, acc=ens.acc, tot=ens.tot,
                        decisions=np.packbits(dec), moves=mov, energy=ens.callback_energy(),
                        acceptance=ens.callback_acceptance())


def philox_case(name, M, K, beta, sigma, weight, seed, offset):
    x0 = O.init_synthetic(seed, offset, M)
    uc, z, ua = O.draws_philox(seed, offset, M, 0, K, with_cat=True)
    ens = O.Ensemble(x0, beta, sigma, weight)
    ens.sweep_replay(uc, z, ua)
    np.savez_compressed(os.path.join(HERE, name), seed=seed, offset=offset, beta=beta, sigma=np.array(sigma),
                        weight=np.array(weight), K=K, x0=x0, u_cat0=uc[:4], z0=z[:4], u_acc0=ua[:4], x=ens.x,
                        acc=ens.acc, tot=ens.tot)


def pgmc_case(name, M, q_batch, beta, sigma, learn_ids, seed):
    x0 = O.init_synthetic(seed, 0, M)
    z = O.draws_pgmc_philox(seed, 0, M, 0, len(learn_ids) * q_batch).reshape(len(learn_ids), q_batch, M)
    ens = O.Ensemble(x0, beta, sigma, [1.0 / len(sigma)] * len(sigma))
    gd = ens.pgmc_replay(q_batch, learn_ids, z)
    np.savez_compressed(os.path.join(HERE, name), seed=seed, beta=beta, sigma=np.array(sigma),
                        learn_ids=np.array(learn_ids), q_batch=q_batch, x0=x0, z=z, gd=gd, x=ens.x)


if __name__ == "__main__":
    replay_case("replay_single.npz", 96, 64, 2.0, [0.1], [1.0], O.POT_HARMONIC, 42)
    replay_case("replay_multi.npz", 80, 64, 2.0, [0.2] * 7, [0.4] + [0.1] * 6, O.POT_HARMONIC, 42)
    replay_case("replay_doublewell.npz", 64, 48, 1.5, [0.6, 0.05], [0.5, 0.5], O.POT_DOUBLE_WELL, 7)
    philox_case("philox_native.npz", 128, 101, 2.0, [0.1, 0.5], [0.6, 0.4], 42, 1000)
    pgmc_case("pgmc.npz", 64, 10, 2.0, [0.2, 0.2, 0.7], [1, 2], 42)
    print("golden fixtures written to", HERE)
