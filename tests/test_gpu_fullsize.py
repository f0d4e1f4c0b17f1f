"""BASELINE.json configs 2, 4 and 5 at their STATED sizes on one B200 (config 3 at 2^27: test_gpu_parity.py
::test_config3_full_size_properties; config 1 verbatim: ::test_config1_verbatim).

The oracle cannot run 2^24..2^26 chains for thousands of steps in test time, so each test runs the engine at full
size and ties it to the oracle through chains that are pure functions of their own inputs: a strided subset of the
chains is re-run by the oracle from the same per-chain inputs (bit-exact / 1e-12 / 1e-10 as the mode allows), and the
full-size ensemble is checked through size-independent properties (3σ analytic averages, conservation of call counts).
"""
import math
import os

import numpy as np
import pytest

import montecarlo_b200 as mb
from montecarlo_b200 import policy_guided as PG
from oracle import oracle as O

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------
# config 2: M = 2^24 chains, 10^4 steps
# ---------------------------------------------------------------------------------------------------------
def test_config2_full_size_device_generator():
    """ALL 2^24 chains x ALL 10^4 steps with the reference's generator family on the device (xoshiro256++ + ziggurat,
    draw order u_cat, z, u_acc per step: metropolis.jl:206, particle_1d.jl:57, metropolis.jl:184).  Every 256th chain
    (2^16 chains) is re-run by the oracle from the same generator state: counters and final generator states must be
    identical (same decisions, same number of raw draws), positions within 1e-12."""
    M, K, seed, beta, sigma = 1 << 24, 10 ** 4, 42, 2.0, 0.1
    x0 = O.init_synthetic(seed, 0, M)
    gen = O.Ensemble(x0, beta, [sigma])
    gen.seed_xoshiro(seed)                                          # per-chain seeds seed + c - 1 (metropolis.jl:262-263)
    st0 = gen.states
    with mb.CudaEnsemble(M, beta, [sigma], seed=seed, rng="xoshiro", arith="exact") as eng:
        eng.set_state(x0)
        eng.set_rng_state(st0)
        eng.set_ziggurat_tables(*O.ziggurat_tables())
        eng.sweep(K, reduce=True)
        me, ma = eng.callbacks()
        x = eng.get_state()
        acc = eng.chain_counters()[0][0]
        st = eng.get_rng_state()
        assert eng.steps_done == K
    idx = np.arange(0, M, 256)
    ref = O.Ensemble(x0[idx], beta, [sigma])
    ref.states = np.ascontiguousarray(st0[idx])
    ref.sweep_xoshiro(K)
    assert np.array_equal(st[idx], ref.states)
    assert np.array_equal(acc[idx].astype(np.int64), ref.acc[0])
    assert np.max(np.abs(x[idx] - ref.x)) < 1e-12
    assert np.mean(x[idx] == ref.x) > 0.5                           # most chains never left the bit-exact fast path
    # the full ensemble: stationary N(0, 1/(2β)) (distribution_test.jl:31-37) and the acceptance of a Gaussian
    # random walk on it, (2/π)·atan(2s/σ); 10^4 steps at τ_int ≈ 60 leave no burn-in bias above 1e-4
    s = 1 / math.sqrt(2 * beta)
    assert abs(x.mean()) < 3 * s / math.sqrt(M) and abs(x.std() - s) < 3 * s / math.sqrt(2 * M)
    assert abs(me - 1 / (2 * beta)) < 3 * math.sqrt(1 / (2 * beta ** 2) / M)
    assert abs(ma[0] - 2 / math.pi * math.atan(2 * s / sigma)) < 1e-3   # cumulative since t = 0: includes the transient


def test_config2_full_width_replay_of_device_draws():
    """Replay mode at config 2's full width: K = 100 steps of draws generated ON THE DEVICE (2 x 13.4 GB resident in
    HBM, never uploaded), decisions written by the kernel; ALL 2^24 chains are compared with the oracle bit for bit
    (decisions, positions, counters), 2^20 chains at a time."""
    import torch
    M, K, beta, sigma = 1 << 24, 100, 2.0, 0.1
    x0 = O.init_synthetic(42, 0, M)
    with mb.CudaEnsemble(M, beta, [sigma], arith="exact") as eng:
        eng.set_state(x0)
        with torch.cuda.stream(eng.torch_stream()):
            g = torch.Generator(device="cuda").manual_seed(1234)
            z = torch.randn((K, M), dtype=torch.float64, device="cuda", generator=g)
            ua = torch.rand((K, M), dtype=torch.float64, device="cuda", generator=g)
            dec = torch.empty((K, M), dtype=torch.uint8, device="cuda")
            eng.sweep_replay_device(K, 0, z.data_ptr(), ua.data_ptr(), dec.data_ptr())
            eng.synchronize()
        x = eng.get_state()
        acc = eng.chain_counters()[0][0]
        step = 1 << 20
        for a in range(0, M, step):
            zh = z[:, a:a + step].contiguous().cpu().numpy()
            uh = ua[:, a:a + step].contiguous().cpu().numpy()
            ref = O.Ensemble(x0[a:a + step], beta, [sigma])
            dref, _, _ = ref.sweep_replay(None, zh, uh, want_decisions=True)
            assert np.array_equal(dec[:, a:a + step].contiguous().cpu().numpy(), dref), a
            assert np.array_equal(x[a:a + step], ref.x), a
            assert np.array_equal(acc[a:a + step].astype(np.int64), ref.acc[0]), a
        del z, ua, dec


# ---------------------------------------------------------------------------------------------------------
# config 4: PolicyGuided MC, pgmc_test.jl:10-52 at M = 2^24
# ---------------------------------------------------------------------------------------------------------
def test_config4_pgmc_full_size(tmp_path):
    """test/pgmc_test.jl:17-35 verbatim -- the 7-move pool, σ₀ = 0.2, the six optimisers with the reference's own
    learning rates, q_batch 10, estimator every step, update every 2nd step after the burn -- on 2^24 chains through
    the driver mirror (steps and burn shortened: 240 / 40).  The first 2^13 chains are then re-run on a second engine
    AND by the oracle under the σ history the full run wrote to parameters/*/parameters.dat: chains bit-identical to
    the full run's, estimator sums within 1e-10 of the oracle's (the reference's own AD tolerance,
    ad_backends_test.jl:31-32) at every update time."""
    M, steps, burn, q = 1 << 24, 240, 40, 10
    seed, beta = 42, 2.0
    chains = mb.ParticleEnsemble(n_chains=M, beta=beta)
    mk = lambda w: mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.2), w)
    pool = (mk(0.4), mk(0.1), mk(0.1), mk(0.1), mk(0.1), mk(0.1), mk(0.1))                  # pgmc_test.jl:17-25
    optimisers = (PG.Static(), PG.VPG(0.001), PG.BLPG(0.001), PG.BLAPG(1e-6, 1e-6), PG.NPG(1e-2, 1e-6),
                  PG.ANPG(1e-6, 1e-6), PG.BLANPG(1e-6, 1e-6))                              # :26
    sampletimes = mb.build_schedule(steps, burn, [0, 10])
    updates = mb.build_schedule(steps, burn, 2)
    algorithm_list = (
        dict(algorithm=mb.Metropolis, pool=pool, seed=seed, parallel=False),
        dict(algorithm=PG.PolicyGradientEstimator, dependencies=(mb.Metropolis,), optimisers=optimisers,
             q_batch_size=q, parallel=True),
        dict(algorithm=PG.PolicyGradientUpdate, dependencies=(PG.PolicyGradientEstimator,), scheduler=updates),
        dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance), scheduler=sampletimes),
        dict(algorithm=mb.StoreParameters, dependencies=(mb.Metropolis,), scheduler=list(range(1, steps + 1))),
    )
    sim = mb.Simulation(chains, algorithm_list, steps, path=str(tmp_path))
    mb.run(sim)
    x_full = chains.x
    acc_full, tot_full = chains.engine.chain_counters()
    assert np.array_equal(tot_full.sum(axis=0), np.full(M, steps))                         # one move per step and chain
    # σ history: line t of parameters/<k>/parameters.dat = σ_k after the events of step t (line 0: t = 0)
    hist = np.stack([np.array([float(l.split("[")[1].rstrip("]\n")) for l in open(tmp_path / "parameters" / str(k + 1) /
                                                                                 "parameters.dat")]) for k in range(7)])
    assert hist.shape == (7, steps + 1) and np.all(hist[:, 0] == 0.2) and np.all(hist[0] == 0.2)   # Static keeps σ₀
    assert np.all(hist[1:, burn - 1] == 0.2) and np.all(hist[1:, burn] != 0.2)             # first update at t = burn
    # the learners move towards the optimum ≈ 1.19 from below (pgmc_test.jl:50 after 10^5 steps; here 100 updates)
    assert np.all(hist[1:, -1] > 0.2) and np.all(hist[1:, -1] < 1.4)
    energies = np.loadtxt(tmp_path / "energy.dat")[:, 1]
    assert abs(energies[-1] - 0.25) < 5e-2                                                  # pgmc_test.jl:45

    # ---- the first 2^13 chains again, on a small engine and on the oracle, under the same σ history ---------------
    n = 1 << 13
    learn = [1, 2, 3, 4, 5, 6]
    weight = [m.weight for m in pool]
    x0 = O.init_synthetic(seed, 0, n)
    ref = O.Ensemble(x0, beta, [0.2] * 7, weight)
    with mb.CudaEnsemble(n, beta, [0.2] * 7, weight, seed=seed, n_chains_total=M, arith="fast") as eng:
        eng.init_synthetic()
        want = np.zeros((6, 5))
        checked = 0
        for t in range(1, steps + 1):
            for k in range(7):                                      # σ in effect during step t = after step t - 1
                eng.set_params(k, hist[k, t - 1])
                ref.sigma[k] = hist[k, t - 1]
            eng.sweep(1)
            uc, z, ua = O.draws_philox(seed, 0, n, t - 1, 1)
            ref.sweep_replay(uc, z, ua)
            eng.pgmc_estimate(q, learn)
            zz = O.draws_pgmc_philox(seed, 0, n, (t - 1) * len(learn) * q, len(learn) * q).reshape(len(learn), q, n)
            xs, es = ref.x.copy(), ref.e.copy()
            want += ref.pgmc_replay(q, learn, zz)
            ref.x[:], ref.e[:] = xs, es                             # FAST arithmetic: the estimator leaves the chains alone
            if t in updates:
                got = eng.pgmc_read(6)
                np.testing.assert_allclose(got, want, rtol=1e-10, err_msg=f"t = {t}")
                eng.pgmc_reset()
                want[:] = 0
                checked += 1
        assert checked == len(updates)
        x_sub = eng.get_state()
        acc_sub, tot_sub = eng.chain_counters()
    assert np.array_equal(x_sub, x_full[:n])                        # bit for bit the chains of the full-size run
    assert np.array_equal(acc_sub, acc_full[:, :n]) and np.array_equal(tot_sub, tot_full[:, :n])
    assert np.max(np.abs(x_sub - ref.x)) < 1e-12
    assert np.array_equal(acc_sub.astype(np.int64), ref.acc) and np.array_equal(tot_sub.astype(np.int64), ref.tot)


# ---------------------------------------------------------------------------------------------------------
# config 5: β sweep {0.5, 1, 2, 4}, M = 2^26 chains, K = 100 fused sweeps with trajectory frames
# ---------------------------------------------------------------------------------------------------------
def test_config5_full_size_beta_sweep(tmp_path):
    """2^26 chains (2^24 per β) as ONE ensemble with per-chain β; StoreTrajectories on build_schedule(steps, burn, 100)
    (512 MiB frames through the device snapshot + copy stream); every β group must sample its own N(0, 1/(2β)) within
    3σ bars (distribution_test.jl:31-37), and a strided subset matches the oracle.  10^5 steps would write 500 GB of
    frames: the schedule is cut at 1500 steps (6 frames)."""
    G = 1 << 24
    bvals = [0.5, 1.0, 2.0, 4.0]
    betas = np.repeat(bvals, G)
    M, steps, burn, sigma, seed = betas.size, 1500, 1000, 0.3, 42
    chains = mb.ParticleEnsemble(n_chains=M, beta=betas)
    pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=sigma), 1.0),)
    sched = mb.build_schedule(steps, burn, 100)
    sim = mb.Simulation(chains, (dict(algorithm=mb.Metropolis, pool=pool, seed=seed),
                                 dict(algorithm=mb.StoreTrajectories, scheduler=sched, store_first=False)),
                        steps, path=str(tmp_path))
    mb.run(sim)
    assert chains.engine.launch_count <= len(sched) + 2            # one fused launch per store interval (K = 1000, 100, ...)
    path = str(tmp_path / "trajectories" / "rank0.bin")
    assert os.path.getsize(path) == len(sched) * (8 + 8 * M)
    frames = np.memmap(path, dtype=np.dtype([("t", "<i8"), ("x", "<f8", (M,))]), mode="r")
    assert list(frames["t"]) == sched
    last = np.array(frames["x"][-1])
    assert np.array_equal(last, chains.x)                           # the last frame is the final state
    for g, b in enumerate(bvals):
        s = 1 / math.sqrt(2 * b)
        xg = last[g * G:(g + 1) * G]
        assert abs(xg.mean()) < 3 * s / math.sqrt(G), (b, xg.mean())
        assert abs(xg.std() - s) < 3 * s / math.sqrt(2 * G), (b, xg.std())
        assert abs((xg * xg).mean() - 1 / (2 * b)) < 3 * math.sqrt(2) * s * s / math.sqrt(G)   # ⟨E⟩ = 1/(2β)
    # every 2^14-th chain against the oracle fed the same counter-based draws (the stream is keyed by the global id)
    idx = np.arange(0, M, 1 << 14)
    x0 = np.concatenate([O.init_synthetic(seed, int(c), 1) for c in idx])
    zs, us = [], []
    for c in idx:
        _, z, ua = O.draws_philox(seed, int(c), 1, 0, steps, with_cat=False)
        zs.append(z[:, 0]); us.append(ua[:, 0])
    ref = O.Ensemble(x0, 2.0, [sigma])
    ref.sweep_replay(None, np.stack(zs, axis=1), np.stack(us, axis=1), betas=betas[idx])
    assert np.max(np.abs(last[idx] - ref.x)) < 1e-12
    del frames
