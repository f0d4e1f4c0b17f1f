"""Secondary measurements of the other §8 paths on one B200 (all through the C ABI): K sweep of the native kernel,
EXACT arithmetic, multi-move pools, replay (HBM-bound), the XOSHIRO device generator, the PGMC estimator (C4) and the
trajectory write-back (C5).  Prints one JSON line per measurement; `scripts/gpu_evidence.sh` stores them.  Every line
carries the NVML clock / throttle-reason record sampled DURING its own timed region (bench.py's ClockSampler), CUDA
events on the launching stream after warm-up, inputs larger than L2.  Optional argv: section names (native f32 multi pgmc replay xoshiro c5)."""
import json, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montecarlo_b200 as mb
from bench import ClockSampler

SAMPLER = ClockSampler(0)
LAST_CLOCKS = [None]
ONLY = sys.argv[1:]

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(eng, fn, reps, warm=2):
    s = eng.torch_stream()
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        SAMPLER.start()
        a.record(s)
        for _ in range(reps):
            fn()
        b.record(s)
        b.synchronize()
        LAST_CLOCKS[0] = SAMPLER.stop()
    return a.elapsed_time(b) / reps


def emit(**kw):
    kw["clocks"] = LAST_CLOCKS[0]
    print(json.dumps(kw), flush=True)



def sec_native():
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


def sec_f32():
    # ---- Float32 ensembles (Particle{Float32}): FP32 Box-Muller on the MUFU pipe, x in 4 bytes ---------------------------
    M = 1 << 27
    for arith in ("fast", "exact"):
        with mb.CudaEnsemble(M, 2.0, [0.1], seed=42, arith=arith, dtype="f32") as eng:
            eng.init_synthetic()
            for K in ((1, 10, 100) if arith == "fast" else (10,)):
                ms = timed(eng, lambda: eng.sweep(K, reduce=True), reps=max(3, 100 // K))
                emit(path=f"K1f native {arith.upper()} Float32 single-move", M=M, K=K, ms=ms,
                     chain_steps_per_s=M * K / (ms * 1e-3), hbm_gbs=16 * M / (ms * 1e-3) / 1e9,
                     hbm_frac=16 * M / (ms * 1e-3) / 1e9 / HBM)


M4 = 1 << 24
SIG7, W7 = [0.2] * 7, [0.4] + [0.1] * 6          # pgmc_test.jl:17-25


def sec_multi():
    # ---- multi-move pool (pgmc_test pool: 7 moves) --------------------------------------------------------------
    with mb.CudaEnsemble(M4, 2.0, SIG7, W7, seed=42, arith="fast") as eng:
        eng.init_synthetic()
        for K in (1, 10, 100):
            ms = timed(eng, lambda: eng.sweep(K), reps=20)
            emit(path="K1m native FAST 7-move", M=M4, K=K, ms=ms, chain_steps_per_s=M4 * K / (ms * 1e-3),
                 hbm_gbs=(16 + 16 * 7) * M4 / (ms * 1e-3) / 1e9, hbm_frac=(16 + 16 * 7) * M4 / (ms * 1e-3) / 1e9 / HBM)
        ms = timed(eng, lambda: eng.sweep(10, reduce=True), reps=20)
        emit(path="K1m native FAST 7-move + per-move record", M=M4, K=10, ms=ms, chain_steps_per_s=M4 * 10 / (ms * 1e-3))
        n = eng.series_per_launch
        ms = timed(eng, lambda: eng.sweep_series([10] * n, read=False), reps=5)
        emit(path=f"K1m series: {n} stores of 10 steps per launch, per-move records", M=M4, K=10 * n, ms=ms,
             chain_steps_per_s=M4 * 10 * n / (ms * 1e-3))
    with mb.CudaEnsemble(M4, 2.0, SIG7, W7, seed=42, arith="exact") as eng:
        eng.init_synthetic()
        ms = timed(eng, lambda: eng.sweep(10), reps=10)
        emit(path="K1m native EXACT 7-move", M=M4, K=10, ms=ms, chain_steps_per_s=M4 * 10 / (ms * 1e-3))


def sec_pgmc():
    with mb.CudaEnsemble(M4, 2.0, SIG7, W7, seed=42, arith="fast") as eng:
        eng.init_synthetic()
        peak = eng.measure_fp64_peak()
        # ---- C4: PGMC estimator, 6 learnable moves x q_batch 10 ------------------------------------------------
        learn = [1, 2, 3, 4, 5, 6]
        ms = timed(eng, lambda: eng.pgmc_estimate(10, learn), reps=10)
        emit(path="K3 PGMC estimator FAST (C4: 6 learnable x q_batch 10)", M=M4, ms=ms,
             trial_evals_per_s=M4 * 60 / (ms * 1e-3), fp64_frac_conv=(12 + 36 + 63) * M4 * 60 / (ms * 1e-3) / peak)
        ms = timed(eng, lambda: (eng.sweep(1), eng.pgmc_estimate(10, learn)), reps=10)
        emit(path="C4 simulation step (Metropolis K=1 + estimator)", M=M4, ms=ms, sim_steps_per_s=1e3 / ms)
    with mb.CudaEnsemble(M4, 2.0, SIG7, W7, seed=42, arith="exact") as eng:
        eng.init_synthetic()
        ms = timed(eng, lambda: eng.pgmc_estimate(10, [1, 2, 3, 4, 5, 6]), reps=5)
        emit(path="K3 PGMC estimator EXACT", M=M4, ms=ms, trial_evals_per_s=M4 * 60 / (ms * 1e-3))


def sec_replay():
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
        ms = timed(eng, lambda: eng.sweep_replay_device(K, 0, z.data_ptr(), ua.data_ptr(), 0), reps=5)
        byts = 16 * K * M4 + 24 * M4
        emit(path="K6 replay EXACT (device draws, no decisions)", M=M4, K=K, ms=ms,
             chain_steps_per_s=M4 * K / (ms * 1e-3), hbm_gbs=byts / (ms * 1e-3) / 1e9, hbm_frac=byts / (ms * 1e-3) / 1e9 / HBM)
        del z, ua, dec


def sec_xoshiro():
    # ---- XOSHIRO device generator (C2 shape: 2^24 chains) -------------------------------------------------------------
    with mb.CudaEnsemble(M4, 2.0, [0.1], seed=42, rng="xoshiro", arith="exact") as eng:
        eng.init_synthetic()
        st = np.random.default_rng(0).integers(1, 2 ** 63, size=(M4, 4), dtype=np.uint64)
        eng.set_rng_state(st)
        ms = timed(eng, lambda: eng.sweep(100), reps=3, warm=1)
        emit(path="K6 xoshiro256++/ziggurat EXACT (reference generator family on device)", M=M4, K=100, ms=ms,
             chain_steps_per_s=M4 * 100 / (ms * 1e-3), c2_full_seconds=1e4 / 100 * ms * 1e-3)


def sec_c5():
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


SECTIONS = {"native": sec_native, "f32": sec_f32, "multi": sec_multi, "pgmc": sec_pgmc, "replay": sec_replay, "xoshiro": sec_xoshiro,
            "c5": sec_c5}


def main():
    for name, fn in SECTIONS.items():
        if not ONLY or name in ONLY:
            fn()


if __name__ == "__main__":
    main()
