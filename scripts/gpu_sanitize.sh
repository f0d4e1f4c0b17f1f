#!/bin/bash
# compute-sanitizer evidence: memcheck + racecheck + synccheck over a small end-to-end pass of every kernel family.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import montecarlo_b200 as mb
rng = np.random.default_rng(7)
def draws(M, K, with_cat=True):      # replay inputs (any draws do: this script checks memory safety, parity is tests/)
    return (rng.random((K, M)) if with_cat else None), rng.standard_normal((K, M)), rng.random((K, M))
M = 3001
for arith in ("fast", "exact"):
    for sig, w in (([0.1], [1.0]), ([0.2, 0.5, 0.9], [0.5, 0.25, 0.25])):
        with mb.CudaEnsemble(M, 2.0, sig, w, seed=3, arith=arith) as e:
            e.init_synthetic(); e.sweep(7, reduce=True); e.sweep(4); e.callbacks(); e.counters()
            e.sweep_series([4, 3, 6, 1]); e.sweep_series([2] * 20); e.callbacks()      # single-move and multi-move series
            e.pgmc_estimate(3, [0]); e.pgmc_read(1); e.get_state(with_energy=True)
            x0 = e.get_state()
            uc, z, ua = draws(M, 5)
            e.sweep_replay(uc, z, ua, want_decisions=True)                 # odd M: per-thread-load replay kernel
        with mb.CudaEnsemble(M + 1, 2.0, sig, w, seed=3, arith=arith) as e:    # even M: bulk-copy (TMA) replay kernel
            e.init_synthetic()
            uc, z, ua = draws(M + 1, 11)
            e.sweep_replay(uc, z, ua, want_decisions=True)
            e.run_host_job([4, 4, 2], x_in=e.get_state(), x_out=np.empty(M + 1), n_slices=3)
            if len(sig) > 1:
                e.pgmc_estimate(3, [0, 1]); e.pgmc_update_device([1], [("VPG", 0.01, 0.0)]); e.sweep(5); e.get_params(1)
    with mb.CudaEnsemble(M, 2.0, [0.1], seed=3, arith=arith, dtype="f32") as e:   # Float32 ensembles
        e.init_synthetic(); e.sweep(7, reduce=True); e.sweep(4); e.callbacks()
        _, z, ua = draws(M, 5, with_cat=False)
        e.sweep_replay(None, z, ua, want_decisions=True); e.get_state_f32(with_energy=True)
with mb.CudaEnsemble(M, 2.0, [0.1], seed=3, rng="xoshiro", arith="exact") as e:
    e.init_synthetic(); st = np.random.default_rng(0).integers(1, 2**63, size=(M, 4), dtype=np.uint64)
    e.set_rng_state(st); e.sweep(50); e.get_rng_state()
print("sanitize pass ok")
PY
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py 2>&1 | tail -6
  echo "exit=$?"
done 2>&1 | tee gpurun_out/sanitizer.log
