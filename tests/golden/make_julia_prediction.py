"""A falsifiable PREDICTION of what the real reference writes for BASELINE config 1 -- the reference's own example
script example/particle_1d/harmonic_oscillator/MC_harmonic_oscillator.jl (seed = 42, β = 2, M = 10, 10^5 steps, burn
1000, σ = 0.1, StoreCallbacks on build_schedule(steps, burn, [0, 10])) -- under Julia 1.7 - 1.10.

The reference cannot run here (no Julia), so parity of its arithmetic is unpinned.  But everything that determines its
output is now either restated from its source or pinned by published known answers:
  * the random stream: rng = Xoshiro(42) for the initial condition, rngs = [Xoshiro(42 + c - 1)] for the chains
    (metropolis.jl:262-263); Julia's SHA-256 seeding, xoshiro256++, rand, randn (ziggurat, Julia's literal tables):
    pinned by the Julia manual's known answers (tests/test_oracle.py::test_julia_rng_known_answers).  The ziggurat's
    wedge / tail branches (1 % of the draws) and Distributions' `μ + σ randn` / Categorical scan are restated [EXT];
  * mc_step! / mc_sweep! / the particle_1d methods / both callbacks: restated line by line (oracle/arianna_oracle.c);
  * the file format: println(file, "$(simulation.t) $(callback(simulation))") (src/algorithms.jl:97-102), Julia's
    shortest round-trip Float64 printing.
Julia's `exp` / `log` are not correctly rounded (<= 1 ulp from glibc's); they only enter through `α > u` (a decision
flips only when α and u agree to ~16 digits) and through the host constant log(2πσ²)/2.

So this script runs the ORACLE exactly as that script runs Arianna and records what it gets: the first and last lines
of energy.dat / acceptance.dat, SHA-256 of the complete files, the final positions.  Anyone with Julia runs the
example and compares (julia/tools/check_prediction.jl): equality closes the parity pin for config 1; the first
differing line says where the restatement is off.  tests/test_oracle.py keeps the committed prediction in step with
the oracle, tests/test_gpu_parity.py::test_config1_julia_mode_matches_the_prediction runs the same job on the device
generator (XOSHIRO mode, EXACT arithmetic) against it.

Run from the repo root:  python tests/golden/make_julia_prediction.py [--out DIR]   (DIR: also write the full files)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import oracle_np as N  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "julia_prediction_config1.json")
SEED, BETA, M, STEPS, BURN, SIGMA = 42, 2.0, 10, 10 ** 5, 1000, 0.1


def jl(v):
    """Julia's "$(x)" for a Float64 (shortest round-trip digits; 1.0e-5 / 1.0e6 style exponents)."""
    if v != v:
        return "NaN"
    r = repr(float(v))
    if "e" in r:
        m, ex = r.split("e")
        return f"{m if '.' in m else m + '.0'}e{int(ex)}"
    if v != 0.0 and abs(v) >= 1e6:
        raise ValueError("not needed here")
    return r


def initial_condition():
    """rng = Xoshiro(seed); chains = [System(4rand(rng) - 2, β) for _ in 1:M]   (MC_harmonic_oscillator.jl:10-13)"""
    st = O.xoshiro_seed_julia(SEED)
    x0 = np.empty(M)
    for c in range(M):
        r, st = O.xoshiro_next(st)
        x0[c] = 4.0 * ((r >> 11) * 2.0 ** -53) - 2.0
    return x0


def predict():
    x0 = initial_condition()
    sampletimes = N.build_schedule(STEPS, BURN, [0, 10])
    ens = O.Ensemble(x0, BETA, [SIGMA])
    ens.seed_xoshiro(SEED, julia=True)
    energy = [f"0 {jl(ens.callback_energy())}"]
    accept = ["0 [NaN]"]                                   # 0/0 at t = 0 (store_first, src/algorithms.jl:93)
    done = 0
    for t in sampletimes:
        ens.sweep_xoshiro(t - done)
        done = t
        energy.append(f"{t} {jl(ens.callback_energy())}")
        accept.append(f"{t} [{jl(float(ens.callback_acceptance()[0]))}]")
    return x0, ens, energy, accept


def record(x0, ens, energy, accept):
    def text(lines):
        return "\n".join(lines) + "\n"
    return {
        "what": "PREDICTION (oracle in Julia mode) of the reference's output for MC_harmonic_oscillator.jl; see make_julia_prediction.py",
        "julia": "1.7 - 1.10 (Xoshiro(n) seeding scheme of those versions)",
        "seed": SEED, "beta": BETA, "M": M, "steps": STEPS, "burn": BURN, "sigma": SIGMA,
        "x0": [jl(v) for v in x0], "x0_hex": [float(v).hex() for v in x0],
        "records": len(energy),
        "energy_head": energy[:20], "energy_tail": energy[-5:], "energy_sha256": hashlib.sha256(text(energy).encode()).hexdigest(),
        "acceptance_head": accept[:20], "acceptance_tail": accept[-5:],
        "acceptance_sha256": hashlib.sha256(text(accept).encode()).hexdigest(),
        "x_final_hex": [float(v).hex() for v in ens.x],
        "accepted_calls": [int(v) for v in ens.acc[0]], "total_calls": [int(v) for v in ens.tot[0]],
        "rng_states_final": [[int(w) for w in row] for row in ens.states],
    }


if __name__ == "__main__":
    x0, ens, energy, accept = predict()
    with open(TARGET, "w") as fh:
        json.dump(record(x0, ens, energy, accept), fh, indent=1)
        fh.write("\n")
    print("wrote", TARGET)
    if "--out" in sys.argv:
        d = sys.argv[sys.argv.index("--out") + 1]
        os.makedirs(d, exist_ok=True)
        open(os.path.join(d, "energy.dat"), "w").write("\n".join(energy) + "\n")
        open(os.path.join(d, "acceptance.dat"), "w").write("\n".join(accept) + "\n")
        print("wrote", d)
