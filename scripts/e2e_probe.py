"""Probe of arianna_run_host_job on one GPU: wall time, device time of the sweeps and the PCIe view, for several slice
counts, three calls each (the first call of a process pays one-off costs: stream/event creation, allocations)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlo_b200 as mb

M = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 27
K = 20
x_in = torch.empty(M, dtype=torch.float64).pin_memory()
x_out = torch.empty(M, dtype=torch.float64).pin_memory()
with mb.CudaEnsemble(M, 2.0, [0.1], seed=42) as eng:
    eng.init_synthetic()
    eng.get_state_to_ptr(x_in.data_ptr())
    eng.sweep_series([10] * K, read=False); eng.synchronize()
    t0 = time.perf_counter(); eng.sweep_series([10] * K, read=False); eng.synchronize()
    print(json.dumps({"resident_series_wall_ms": 1e3 * (time.perf_counter() - t0), "sweep_ms": eng.timing()[0]}))
    for download in (False, True):
        for slices in (4, 8, 16, 32):
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                eng.run_host_job([10] * K, x_in=x_in.data_ptr(), x_out=x_out.data_ptr() if download else None, n_slices=slices)
                dt = time.perf_counter() - t0
                print(json.dumps({"download": download, "slices": slices, "rep": rep, "wall_ms": round(1e3 * dt, 2),
                                  "sweep_ms": round(eng.timing()[0], 2), "rate": M * 10 * K / dt,
                                  **{k: round(v, 2) for k, v in eng.job_timing().items()}}), flush=True)
