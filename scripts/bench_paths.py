"""Secondary measurements of the other §8 paths on one B200 (all through the C ABI): K sweep of the native kernel,
EXACT arithmetic, multi-move pools, replay (HBM-bound), the XOSHIRO device generator, the PGMC estimator (C4) and the
trajectory write-back (C5).  Prints one JSON line per measurement; `scripts/gpu_paths.sh` stores them."""
import json, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montecarlo_b200 as mb

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(eng, fn, reps, warm=2):
    s = eng.torch_stream()
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps):
            fn()
        b.record(s)
        b.synchronize()
    return a.elapsed_time(b) / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    M = 1 << 27
    # ---- native sweep vs K (HBM-bound at K = 1, FP64-side bound beyond) --------------------------------------
    with mb.CudaEnsemble(M, 2.0, [0.1], seed=42, arith="fast") as eng:
        eng.init_synthetic()
        peak = eng.measure_fp64_peak()
        for K in (1, 2, 4, 10, 100, 1000):
            ms = timed(eng, lambda: eng.sweep(K, reduce=True), reps=max(2, 200 // K))
            rate = M * K / (ms * 1e-3)
            emit(path="K1 native FAST single-move", M=M, K=K, ms=ms, chain_steps_per_s=rate,
                 hbm_gbs=24 * M / (ms * 1e-3) / 1e9, hbm_frac=24 * M / (ms * 1e-3) / 1e9 / HBM,
                 fp64_frac=110 * rate / peak)
    with mb.CudaEnsemble(M, 2.0, [0.1], seed=42, arith="exact") as eng:
        eng.init_synthetic()
        ms = timed(eng, lambda: eng.sweep(10, reduce=True), reps=10)
        emit(path="K1 native EXACT single-move", M=M, K=10, ms=ms, chain_steps_per_s=M * 10 / (ms * 1e-3))
    # ---- multi-move pool (pgmc_test pool: 7 moves) --------------------------------------------------------------
    M4 = 1 << 24
    sig, w = [0.2] * 7, [0.4] + [0.1] * 6
    with mb.CudaEnsemble(M4, 2.0, sig, w, seed=42, arith="fast") as eng:
        eng.init_synthetic()
        for K in (1, 10):
            ms = timed(eng, lambda: eng.sweep(K), reps=20)
            emit(path="K1 native FAST 7-move", M=M4, K=K, ms=ms, chain_steps_per_s=M4 * K / (ms * 1e-3),
                 hbm_gbs=(16 + 16 * 7) * M4 / (ms * 1e-3) / 1e9)
        # ---- C4: PGMC estimator, 6 learnable moves x q_batch 10 ------------------------------------------------
        learn = [1, 2, 3, 4, 5, 6]
        ms = timed(eng, lambda: eng.pgmc_estimate(10, learn), reps=10)
        emit(path="K3 PGMC estimator FAST (C4: 6 learnable x q_batch 10)", M=M4, ms=ms,
             trial_evals_per_s=M4 * 60 / (ms * 1e-3), fp64_frac_conv=(12 + 36 + 63) * M4 * 60 / (ms * 1e-3) / peak)
        ms = timed(eng, lambda: (eng.sweep(1), eng.pgmc_estimate(10, learn)), reps=10)
        emit(path="C4 simulation step (Metropolis K=1 + estimator)", M=M4, ms=ms, sim_steps_per_s=1e3 / ms)
    with mb.CudaEnsemble(M4, 2.0, sig, w, seed=42, arith="exact") as eng:
        eng.init_synthetic()
        ms = timed(eng, lambda: eng.pgmc_estimate(10, [1, 2, 3, 4, 5, 6]), reps=5)
        emit(path="K3 PGMC estimator EXACT", M=M4, ms=ms, trial_evals_per_s=M4 * 60 / (ms * 1e-3))
    # ---- replay: draws resident in HBM, 16 B of draws per chain-step ------------------------------------------------
    K = 32
    with mb.CudaEnsemble(M4, 2.0, [0.1], seed=42, arith="exact") as eng:
        eng.init_synthetic()
        with torch.cuda.stream(eng.torch_stream()):
            z = torch.randn((K, M4), dtype=torch.float64, device="cuda")
            ua = torch.rand((K, M4), dtype=torch.float64, device="cuda")
            dec = torch.empty((K, M4), dtype=torch.uint8, device="cuda")
        eng.synchronize(); torch.cuda.synchronize()
        ms = timed(eng, lambda: eng.sweep_replay_device(K, 0, z.data_ptr(), ua.data_ptr(), dec.data_ptr()), reps=5)
        byts = (16 + 1) * K * M4 + 24 * M4
        emit(path="K6 replay EXACT (device draws + decisions out)", M=M4, K=K, ms=ms,
             chain_steps_per_s=M4 * K / (ms * 1e-3), hbm_gbs=byts / (ms * 1e-3) / 1e9, hbm_frac=byts / (ms * 1e-3) / 1e9 / HBM)
        del z, ua, dec
    # ---- XOSHIRO device generator (C2 shape: 2^24 chains) -------------------------------------------------------------
    with mb.CudaEnsemble(M4, 2.0, [0.1], seed=42, rng="xoshiro", arith="exact") as eng:
        eng.init_synthetic()
        st = np.random.default_rng(0).integers(1, 2 ** 63, size=(M4, 4), dtype=np.uint64)
        eng.set_rng_state(st)
        ms = timed(eng, lambda: eng.sweep(100), reps=3, warm=1)
        emit(path="K6 xoshiro256++/ziggurat EXACT (reference generator family on device)", M=M4, K=100, ms=ms,
             chain_steps_per_s=M4 * 100 / (ms * 1e-3), c2_full_seconds=1e4 / 100 * ms * 1e-3)
    # ---- C5: K = 100 sweeps with the x write-back (512 MiB) per store, async into pinned memory -------------------------
    M5 = 1 << 26
    with mb.CudaEnsemble(M5, 2.0, [0.1], seed=42, arith="fast") as eng:
        eng.init_synthetic()
        bufs = [torch.empty(M5, dtype=torch.float64).pin_memory() for _ in range(2)]
        i = [0]

        def step():
            eng.sweep(100, reduce=True)
            eng.get_state_async(bufs[i[0] & 1].data_ptr())
            i[0] += 1
        ms = timed(eng, step, reps=6, warm=1)
        eng.synchronize()
        emit(path="C5 store interval: K=100 sweep + async D2H of x (512 MiB)", M=M5, ms=ms,
             chain_steps_per_s=M5 * 100 / (ms * 1e-3), d2h_gbs=8 * M5 / (ms * 1e-3) / 1e9)
        ms0 = timed(eng, lambda: eng.sweep(100, reduce=True), reps=6, warm=1)
        emit(path="C5 store interval without the write-back", M=M5, ms=ms0, chain_steps_per_s=M5 * 100 / (ms0 * 1e-3))


if __name__ == "__main__":
    main()
