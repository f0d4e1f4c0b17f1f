import sys; sys.path.insert(0, ".")
import montecarlo_b200 as mb
with mb.CudaEnsemble(1 << 27, 2.0, [0.1], seed=42, arith="fast", dtype="f32") as eng:
    eng.init_synthetic()
    for _ in range(3):
        eng.sweep(100)
    eng.synchronize()
