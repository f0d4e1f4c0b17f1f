import sys, os, json; sys.path.insert(0, ".")
import torch
import montecarlo_b200 as mb
M4 = 1 << 24
sig, w = [0.2] * 7, [0.4] + [0.1] * 6
with mb.CudaEnsemble(M4, 2.0, sig, w, seed=42, arith="fast") as eng:
    eng.init_synthetic(); eng.sweep(10)
    learn = [1, 2, 3, 4, 5, 6]
    for _ in range(3): eng.pgmc_estimate(10, learn)
    eng.synchronize()
    ts = []
    for _ in range(10):
        eng.pgmc_estimate(10, learn); ts.append(eng.timing()[1])
    print(os.environ.get("ARIANNA_LIB", "default"), "pgmc ms", sum(ts) / len(ts), "evals/s %.4g" % (M4 * 60 / (sum(ts) / len(ts) * 1e-3)))
