import sys; sys.path.insert(0, ".")
import montecarlo_b200 as mb
M4 = 1 << 24
sig, w = [0.2] * 7, [0.4] + [0.1] * 6
with mb.CudaEnsemble(M4, 2.0, sig, w, seed=42, arith="fast") as eng:
    eng.init_synthetic(); eng.sweep(10)
    for _ in range(3):
        eng.pgmc_estimate(10, [1, 2, 3, 4, 5, 6])
    eng.synchronize()
