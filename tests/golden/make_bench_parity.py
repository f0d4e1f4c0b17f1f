"""Writes tests/golden/bench_parity.json: the callback records of bench.py's `parity` mini-run (2^20 chains, seed 42,
β = 2, σ = 0.1, 5 store intervals of 10 Metropolis steps, native Philox stream) computed by the CPU ORACLE
(oracle/arianna_oracle.c).  bench.py compares the all-reduced records of its N-rank GPU run with these
(Σe to 1e-12, the integer Σacc exactly).  `x_bits_checksum_gpu_n1` -- the order-independent checksum of the final
positions of a ONE-GPU run, which an N-rank run must reproduce bit for bit -- cannot come from the oracle (device
and glibc transcendental functions differ in the last ulp): pass the value a 1-GPU `python bench.py` prints in
parity.x_bits_checksum as argv[1] to record it (kept when re-run without an argument).

    python tests/golden/make_bench_parity.py [0x....]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(HERE, "bench_parity.json")
M, SEED, BETA, SIGMA, KS = 1 << 20, 42, 2.0, 0.1, [10, 10, 10, 10, 10]


def main():
    ens = O.Ensemble(O.init_synthetic(SEED, 0, M), BETA, [SIGMA])
    records, acc_sums, done = [], [], 0
    for K in KS:
        _, z, ua = O.draws_philox(SEED, 0, M, done, K, with_cat=False)
        ens.sweep_replay(None, z, ua)
        done += K
        records.append([ens.callback_energy() * M, float(ens.acc[0].sum()) / done, float(M)])
        acc_sums.append(int(ens.acc[0].sum()))
    old = json.load(open(OUT)) if os.path.exists(OUT) else {}
    chk = sys.argv[1] if len(sys.argv) > 1 else old.get("x_bits_checksum_gpu_n1")
    json.dump({"chains": M, "seed": SEED, "beta": BETA, "sigma": SIGMA, "Ks": KS, "records": records,
               "acc_sums": acc_sums, "x_bits_checksum_gpu_n1": chk,
               "source": "records/acc_sums: CPU oracle (tests/golden/make_bench_parity.py); checksum: 1-GPU bench.py run"},
              open(OUT, "w"), indent=1)
    print(open(OUT).read())


if __name__ == "__main__":
    main()
