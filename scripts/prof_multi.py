"""Profiling target: the 7-move pool of pgmc_test.jl:17-25 (σ = 0.2, weights 0.4 + 6 x 0.1), M = 2^24, K = 10 per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlo_b200 as mb

M = 1 << 24
with mb.CudaEnsemble(M, 2.0, [0.2] * 7, [0.4] + [0.1] * 6, seed=42, arith="fast") as eng:
    eng.init_synthetic()
    for _ in range(3):
        eng.sweep(10)
    eng.synchronize()
