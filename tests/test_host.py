"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, argument validation,
the Python mirror of Arianna's driver (schedules, event order, lazy fused flush, file formats, learner rules) and
the multi-rank path under gloo with world_size 2.  No kernel runs here (there is no GPU in this container)."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import montecarlo_b200 as mb
from montecarlo_b200 import _lib as L
from montecarlo_b200 import arianna as A
from montecarlo_b200 import policy_guided as PG
from oracle import oracle as O
from oracle import oracle_np as N

from fake_engine import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the C ABI -------------------------------------------------------------------------------------------
def _header_symbols():
    src = open(os.path.join(ROOT, "include", "arianna_cuda.h")).read()
    return sorted(set(re.findall(r"ARIANNA_API[^;(]*?\b(arianna_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_symbols()
    assert len(names) >= 30
    lib = C.CDLL(mb.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/arianna_cuda.h but not exported"
    assert sorted(L.SYMBOLS) == names            # the ctypes table binds exactly the declared surface
    assert L.load().arianna_abi_version() == 2


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", mb.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def _cfg(**kw):
    cfg = L.Config()
    cfg.struct_size = C.sizeof(L.Config)
    cfg.device, cfg.n_chains, cfg.beta, cfg.n_moves = -1, 16, 2.0, 1
    cfg.sigma[0], cfg.weight[0] = 0.1, 1.0
    cfg.arith_mode = L.ARITH_FAST
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _create(cfg):
    h = C.c_void_p()
    rc = L.load().arianna_create(C.byref(cfg), C.byref(h))
    return rc, L.load().arianna_last_error(None).decode()


def test_create_validates_arguments_before_touching_the_device():
    assert _create(_cfg(struct_size=8))[0] == L.ERR_INVALID
    assert _create(_cfg(n_chains=0))[0] == L.ERR_INVALID
    assert _create(_cfg(n_moves=0))[0] == L.ERR_INVALID
    assert _create(_cfg(n_moves=17))[0] == L.ERR_INVALID
    assert _create(_cfg(potential=9))[0] == L.ERR_INVALID
    assert _create(_cfg(rng_mode=7))[0] == L.ERR_INVALID
    c = _cfg()
    c.sigma[0] = -0.1
    assert _create(c)[0] == L.ERR_INVALID
    c = _cfg()
    c.weight[0] = 0.7                              # Categorical([0.7]) throws in the reference [EXT]
    rc, msg = _create(c)
    assert rc == L.ERR_INVALID and "sum to 1" in msg


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc, msg = _create(_cfg())
    assert rc == L.ERR_NO_DEVICE and "no CPU fallback" in msg
    with pytest.raises(mb.AriannaError):
        mb.CudaEnsemble(8, 2.0, [0.1])


def test_null_handle_is_an_error_not_a_crash():
    lib = L.load()
    assert lib.arianna_sweep(None, 1, 0) == L.ERR_INVALID
    assert lib.arianna_destroy(None) == L.OK


# ---- build_schedule (simulation.jl:95-117) -------------------------------------------------------------------
def test_build_schedule_matches_restatement():
    rng = np.random.default_rng(0)
    for _ in range(200):
        steps = int(rng.integers(10, 5000))
        burn = int(rng.integers(0, steps))
        dt = int(rng.integers(1, 200))
        assert A.build_schedule(steps, burn, dt) == N.build_schedule(steps, burn, dt)
        blk = sorted(set(int(v) for v in rng.integers(0, 50, size=3))) or [1]
        if blk[-1] == 0:
            blk.append(5)
        assert A.build_schedule(steps, burn, blk) == N.build_schedule(steps, burn, blk)
    assert A.build_schedule(10 ** 5, 1000, [0, 10])[:3] == [1000, 1010, 1020]
    assert A.build_schedule(1000, 10, 2.0) == N.build_schedule(1000, 10, 2.0)
    assert A.build_schedule(100, 10, 30) == [10, 40, 70, 100]


def test_shard_bounds_partition():
    for n, w in [(10, 1), (10, 2), (10, 3), (2 ** 27, 8), (7, 8)]:
        spans = [A.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2


def test_julia_number_rendering():
    assert A._jl(0.25) == "0.25" and A._jl(1e-5) == "1.0e-5" and A._jl(float("nan")) == "NaN"
    assert A._jl([0.5, float("nan")]) == "[0.5, NaN]" and A._jl(3) == "3"
    # Julia switches to exponent form at 1e6 (Python only at 1e16) and writes "1.0e6", not "1e+06"
    assert A._jl(999999.0) == "999999.0" and A._jl(1e6) == "1.0e6" and A._jl(-2.5e7) == "-2.5e7"
    assert A._jl(123456789.125) == "1.23456789125e8" and A._jl(1e16) == "1.0e16" and A._jl(1.5e300) == "1.5e300"
    assert A._jl(0.0001) == "0.0001" and A._jl(-0.0) == "-0.0" and A._jl(float("-inf")) == "-Inf"


# ---- learner rules -------------------------------------------------------------------------------------------
OPTS = [(PG.VPG(1e-3), O.OPT_VPG, 1e-3, 0), (PG.BLPG(1e-3), O.OPT_BLPG, 1e-3, 0),
        (PG.BLAPG(1e-6, 1e-6), O.OPT_BLAPG, 1e-6, 1e-6), (PG.NPG(1e-2, 1e-6), O.OPT_NPG, 1e-2, 1e-6),
        (PG.ANPG(1e-6, 1e-6), O.OPT_ANPG, 1e-6, 1e-6), (PG.BLANPG(1e-6, 1e-6), O.OPT_BLANPG, 1e-6, 1e-6)]


@pytest.mark.parametrize("opt,kind,p1,p2", OPTS)
def test_learning_step_matches_oracle(opt, kind, p1, p2):
    rec = [3.1, 1.2, -40.0, 210.0, 100.0]          # sums over n = 100 samples
    gd = PG.average(PG.GradientData.from_record(rec))
    th = np.array([0.2])
    PG.learning_step(th, gd, opt)
    want = O.learning_step(kind, p1, p2, np.array(rec[:4]) / 100.0, 0.2)
    assert abs(th[0] - want) < 1e-15


def test_analytic_gradient_matches_reference_tolerance():
    """ad_backends_test.jl:31-32 (atol 1e-10) at its own point δ = 0, σ = 0.2."""
    assert abs(PG.log_proposal_density(0.0, 0.2) - (-math.log(2 * math.pi * 0.04) / 2)) < 1e-10
    assert abs(PG.dlogq_dsigma(0.0, 0.2) + 5.0) < 1e-10
    for d, s in [(0.3, 0.2), (-1.1, 0.7)]:
        assert abs(PG.dlogq_dsigma(d, s) - O.dlogq_dsigma(d, s)) < 1e-12


# ---- the driver loop over a test double (oracle-backed engine) -------------------------------------------------
@pytest.fixture
def fake_engine(monkeypatch):
    monkeypatch.setattr(A, "CudaEnsemble", OracleEngine)


def _mc_setup(tmp_path, M=64, steps=200, burn=50):
    beta, seed = 2.0, 42
    x0 = O.init_synthetic(seed, 0, M)
    chains = mb.ParticleEnsemble(x0, beta)
    pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
    sampletimes = mb.build_schedule(steps, burn, [0, 10])
    algorithm_list = (
        dict(algorithm=mb.Metropolis, pool=pool, seed=seed, parallel=False),
        dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance), scheduler=sampletimes),
        dict(algorithm=mb.StoreTrajectories, scheduler=sampletimes),
        dict(algorithm=mb.StoreLastFrames, scheduler=[steps]),
        dict(algorithm=mb.PrintTimeSteps, scheduler=mb.build_schedule(steps, burn, steps // 10)),
    )
    return x0, chains, sampletimes, mb.Simulation(chains, algorithm_list, steps, path=str(tmp_path))


def test_run_matches_stepwise_oracle_and_fuses_steps(tmp_path, fake_engine):
    M, steps, burn = 64, 200, 50
    x0, chains, sampletimes, sim = _mc_setup(tmp_path, M, steps, burn)
    mb.run(sim)
    # fused flushes: one launch per gap between observation points, not one per step
    assert chains.engine.launch_count == len(sampletimes)
    assert chains.engine.steps_done == steps
    # reference order: callbacks at t see the state after t Metropolis steps (simulation.jl:184-191)
    ref = O.Ensemble(x0, 2.0, [0.1])
    lines = open(tmp_path / "energy.dat").read().split("\n")[:-1]
    assert lines[0].split()[0] == "0"                              # store_first record at t = 0
    assert float(lines[0].split()[1]) == ref.callback_energy()
    acc_lines = open(tmp_path / "acceptance.dat").read().split("\n")[:-1]
    assert acc_lines[0] == "0 [NaN]"                               # 0/0 at t = 0 (metropolis.jl:320)
    done = 0
    for ln, al, t in zip(lines[1:], acc_lines[1:], sampletimes):
        _, z, ua = O.draws_philox(42, 0, M, done, t - done, with_cat=False)
        ref.sweep_replay(None, z, ua)
        done = t
        # summation order differs (device/test-double tree sums vs the reference's sequential sum): ≤ 1e-13 rel.
        assert int(ln.split()[0]) == t and abs(float(ln.split()[1]) / ref.callback_energy() - 1) < 1e-13
        assert al.startswith(f"{t} [") and abs(float(al.split("[")[1][:-1]) / ref.callback_acceptance()[0] - 1) < 1e-13
    # trajectories: reference text layout, 1-based chain directories, "t x" lines (particle_1d.jl:63-66)
    rows = open(tmp_path / "trajectories" / "1" / "trajectory.dat").read().split("\n")[:-1]
    assert len(rows) == 1 + len(sampletimes) and rows[-1] == f"{steps} {A._jl(float(ref.x[0]))}"
    t, x = mb.StoreTrajectories.read_binary(str(tmp_path / "trajectories" / "rank0.bin"), M)
    assert list(t) == [0] + sampletimes and np.array_equal(x[-1], ref.x)
    last = open(tmp_path / "trajectories" / str(M) / "lastframe.dat").read()
    assert last == f"{steps} {A._jl(float(ref.x[-1]))}\n"
    summary = open(tmp_path / "summary.log").read()
    assert "Metropolis" in summary
    # update_summary / finalise_summary (simulation.jl:146-165)
    assert "Report:\n\tSimulation time: " in summary and "\tSimulation size: " in summary and " MB\n" in summary
    assert summary.rstrip().split("\n")[-1].startswith("\tStatus: Completed on ")


def _callbacks_only_setup(path, M=48, steps=400, burn=100, traj_every=None):
    seed = 42
    chains = mb.ParticleEnsemble(O.init_synthetic(seed, 0, M), 2.0)
    pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
    sampletimes = mb.build_schedule(steps, burn, 10)
    algorithm_list = [
        dict(algorithm=mb.Metropolis, pool=pool, seed=seed, parallel=False),
        dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance), scheduler=sampletimes),
        dict(algorithm=mb.PrintTimeSteps, scheduler=mb.build_schedule(steps, burn, steps // 10)),
    ]
    if traj_every:
        algorithm_list.append(dict(algorithm=mb.StoreTrajectories, scheduler=mb.build_schedule(steps, burn, traj_every)))
    return chains, sampletimes, mb.Simulation(chains, tuple(algorithm_list), steps, path=str(path))


def test_lookahead_fuses_callback_only_stores(tmp_path, fake_engine):
    """The schedule is known up front: run() lets the ensemble execute every callback-only store interval up to the
    next event that needs the chains themselves as ONE arianna_sweep_series call.  Output files are byte-identical
    to the one-launch-per-store path; trajectory frames act as barriers and still see the state at THEIR time."""
    for traj_every in (None, 70):
        a, b = tmp_path / f"look{traj_every}", tmp_path / f"plain{traj_every}"
        ca, st, sa = _callbacks_only_setup(a, traj_every=traj_every)
        mb.run(sa)
        cb, _, sb = _callbacks_only_setup(b, traj_every=traj_every)
        sb.lookahead = False
        mb.run(sb)
        assert cb.engine.launch_count >= len(st) and not hasattr(cb.engine, "series_calls")
        for name in ("energy.dat", "acceptance.dat"):
            assert open(a / name).read() == open(b / name).read()
        assert np.array_equal(ca.engine.get_state(), cb.engine.get_state())
        assert ca.engine.steps_done == cb.engine.steps_done == 400
        if traj_every is None:
            assert ca.engine.series_calls == 1 and ca.engine.launch_count == 1     # the whole run in one call
        else:
            ta = mb.StoreTrajectories.read_binary(str(a / "trajectories" / "rank0.bin"), 48)
            tb = mb.StoreTrajectories.read_binary(str(b / "trajectories" / "rank0.bin"), 48)
            assert np.array_equal(ta[0], tb[0]) and np.array_equal(ta[1], tb[1])
            assert 1 < ca.engine.series_calls <= len(ta[0]) + 1 and ca.engine.launch_count < cb.engine.launch_count / 3
    # a bounded look-ahead window chunks the stretch
    ca, st, sa = _callbacks_only_setup(tmp_path / "win")
    ca.max_lookahead = 7
    mb.run(sa)
    assert ca.engine.series_calls == math.ceil(len(st) / 8)
    assert open(tmp_path / "win" / "energy.dat").read() == open(tmp_path / "lookNone" / "energy.dat").read()


def test_lookahead_refuses_unplanned_observation(tmp_path, fake_engine):
    chains, st, sim = _callbacks_only_setup(tmp_path)
    seen = []

    class Peek(A.AriannaAlgorithm):                        # an algorithm run() knows nothing about: a barrier
        def __init__(self, chains, **extras):
            pass

        def make_step(self, simulation):
            seen.append(simulation.chains.x.copy())

    sim2 = mb.Simulation(chains.__class__(O.init_synthetic(42, 0, 48), 2.0), (
        dict(algorithm=mb.Metropolis, pool=(mb.Move(mb.Displacement(0.0), mb.StandardGaussian(),
                                                    mb.ComponentArray(σ=0.1), 1.0),), seed=42),
        dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy,), scheduler=st),
        dict(algorithm=Peek, scheduler=[155, 300]),
    ), 400, path=str(tmp_path / "peek"))
    mb.run(sim2)
    assert len(seen) == 2 and sim2.chains.engine.steps_done == 400
    ref = O.Ensemble(O.init_synthetic(42, 0, 48), 2.0, [0.1])
    _, z, ua = O.draws_philox(42, 0, 48, 0, 155, with_cat=False)
    ref.sweep_replay(None, z, ua)
    assert np.array_equal(seen[0], ref.x)                 # the barrier saw the chains at t = 155, not run ahead
    # an observation that is NOT in the schedule while the device is ahead fails loudly instead of lying
    chains3, st3, sim3 = _callbacks_only_setup(tmp_path / "bad")
    chains3._lookahead = A._make_lookahead(sim3)
    sim3.t = st3[0]
    chains3._advance(st3[0])
    mb.callback_energy(sim3)
    assert chains3._ahead > 0
    with pytest.raises(RuntimeError):
        chains3.x


def test_lookahead_barrier_listed_before_metropolis(tmp_path, fake_engine):
    """An algorithm listed BEFORE Metropolis sees the chains before the Metropolis step of its time
    (simulation.jl:184-191 runs the list in order): a look-ahead stretch must stop short of such a barrier even when a
    callback store falls on the same time.  Also: Metropolis every 2nd step + callbacks every step gives empty store
    intervals (K = 0) inside the stretch."""
    seen = {}

    class Peek(A.AriannaAlgorithm):
        def __init__(self, chains, **extras):
            pass

        def make_step(self, simulation):
            seen[simulation.t] = simulation.chains.x.copy()

    def build(path, lookahead):
        chains = mb.ParticleEnsemble(O.init_synthetic(42, 0, 48), 2.0)
        pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
        sim = mb.Simulation(chains, (
            dict(algorithm=Peek, scheduler=[120, 200, 333]),              # 120 and 200 coincide with stores
            dict(algorithm=mb.Metropolis, pool=pool, seed=42, scheduler=list(range(1, 401, 2)) + [400]),
            dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance),
                 scheduler=list(range(100, 401))),
        ), 400, path=str(path))
        sim.lookahead = lookahead
        return chains, sim

    ca, sa = build(tmp_path / "look", True)
    mb.run(sa)
    seen_a = dict(seen)
    seen.clear()
    cb, sb = build(tmp_path / "plain", False)
    mb.run(sb)
    assert sorted(seen_a) == sorted(seen) == [120, 200, 333]
    for t in seen:
        assert np.array_equal(seen_a[t], seen[t]), t
    for name in ("energy.dat", "acceptance.dat"):
        assert open(tmp_path / "look" / name).read() == open(tmp_path / "plain" / name).read()
    assert np.array_equal(ca.engine.get_state(), cb.engine.get_state())
    assert ca.engine.steps_done == cb.engine.steps_done == 201
    assert ca.engine.series_calls >= 3 and ca.engine.launch_count < cb.engine.launch_count / 10
    # the barrier at t = 120 saw the state after the Metropolis steps scheduled at times < 120: 60 of them
    ref = O.Ensemble(O.init_synthetic(42, 0, 48), 2.0, [0.1])
    _, z, ua = O.draws_philox(42, 0, 48, 0, 60, with_cat=False)
    ref.sweep_replay(None, z, ua)
    assert np.array_equal(seen_a[120], ref.x)


def test_simulation_constructor_contract(tmp_path, fake_engine):
    chains = mb.ParticleEnsemble(np.zeros(4), 2.0)
    pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
    with pytest.raises(AssertionError):                            # scheduler entries must be in 0..steps
        mb.Simulation(chains, (dict(algorithm=mb.Metropolis, pool=pool, scheduler=[5, 500]),), 100, path=str(tmp_path))
    chains = mb.ParticleEnsemble(np.zeros(4), 2.0)
    with pytest.raises(AssertionError):                            # schedulers must be sorted
        mb.Simulation(chains, (dict(algorithm=mb.Metropolis, pool=pool, scheduler=[5, 3]),), 100, path=str(tmp_path))
    with pytest.raises(TypeError):
        mb.Move(mb.Displacement(0.0), object(), mb.ComponentArray(σ=0.1), 1.0)
    # list of Particles (the reference's `chains` vector) is accepted for small ensembles
    sim = mb.Simulation([mb.System(0.5, 2.0), mb.System(-0.5, 2.0)],
                        (dict(algorithm=mb.Metropolis, pool=pool, seed=3),), 10, path=str(tmp_path))
    assert len(sim.chains) == 2 and sim.counters == [0]
    mb.run(sim)
    assert sim.chains.engine.steps_done == 10


def test_pgmc_driver_learns(tmp_path, fake_engine):
    """pgmc_test.jl:10-52 shrunk: 3 moves (Static, VPG, BLANPG); every learner must move σ from 0.2 towards 1.2."""
    M, steps, burn = 1024, 260, 20
    x0 = O.init_synthetic(42, 0, M)
    chains = mb.ParticleEnsemble(x0, 2.0)
    mk = lambda w: mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.2), w)
    pool = (mk(0.4), mk(0.3), mk(0.3))
    optimisers = (PG.Static(), PG.VPG(0.08), PG.BLANPG(4e-3, 1e-6))
    sampletimes = mb.build_schedule(steps, burn, [0, 10])
    algorithm_list = (
        dict(algorithm=mb.Metropolis, pool=pool, seed=42, parallel=False),
        dict(algorithm=PG.PolicyGradientEstimator, dependencies=(mb.Metropolis,), optimisers=optimisers, q_batch_size=4),
        dict(algorithm=PG.PolicyGradientUpdate, dependencies=(PG.PolicyGradientEstimator,),
             scheduler=mb.build_schedule(steps, burn, 2)),
        dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance), scheduler=sampletimes),
        dict(algorithm=mb.StoreParameters, dependencies=(mb.Metropolis,), scheduler=sampletimes),
    )
    sim = mb.Simulation(chains, algorithm_list, steps, path=str(tmp_path))
    mb.run(sim)
    sig = [m.parameters.σ for m in pool]
    assert sig[0] == 0.2 and abs(sig[1] - 1.2) < 0.2 and abs(sig[2] - 1.2) < 0.2, sig
    rows = open(tmp_path / "parameters" / "2" / "parameters.dat").read().split("\n")[:-1]
    assert rows[0] == "0 [0.2]" and rows[-1] == f"{steps} [{sig[1]!r}]"
    e = np.loadtxt(tmp_path / "energy.dat")[:, 1]
    assert abs(e[len(e) // 2:].mean() - 0.25) < 5e-2               # pgmc_test.jl:45
    # estimator accumulates between updates and is reset by the update (update.jl:55)
    assert np.all(chains.engine.gd == 0)


# ---- N > 1: world_size-2 gloo -----------------------------------------------------------------------------------
_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
import montecarlo_b200 as mb
from montecarlo_b200 import arianna as A
from fake_engine import OracleEngine
from oracle import oracle as O
A.CudaEnsemble = OracleEngine
M, steps = 101, 60                      # odd M: ragged shards
x0 = O.init_synthetic(7, 0, M)
chains = mb.ParticleEnsemble(x0, 2.0)
pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.3), 0.5),
        mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.05), 0.5))
sched = mb.build_schedule(steps, 10, 10)
from montecarlo_b200 import policy_guided as PG
sim = mb.Simulation(chains, (dict(algorithm=mb.Metropolis, pool=pool, seed=7),
                             dict(algorithm=PG.PolicyGradientEstimator, dependencies=(mb.Metropolis,),
                                  optimisers=(PG.Static(), PG.VPG(0.05)), q_batch_size=3),
                             dict(algorithm=PG.PolicyGradientUpdate, dependencies=(PG.PolicyGradientEstimator,),
                                  scheduler=mb.build_schedule(steps, 10, 4)),
                             dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance),
                                  scheduler=sched),
                             dict(algorithm=mb.StoreParameters, dependencies=(mb.Metropolis,), scheduler=sched),
                             dict(algorithm=mb.StoreTrajectories, scheduler=[steps], store_first=False)),
                    steps, path={path!r})
mb.run(sim)
np.save(os.path.join({path!r}, f"x_rank{{chains.rank}}.npy"), chains.x)
np.save(os.path.join({path!r}, f"sigma_rank{{chains.rank}}.npy"), np.array([m.parameters.σ for m in pool]))
dist.barrier(); dist.destroy_process_group()
"""


def test_two_rank_gloo_matches_single_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port, path=str(tmp_path)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    # single-process oracle over the WHOLE ensemble, incl. the estimator (every step) and the VPG update (every 4)
    M, steps = 101, 60
    x0 = O.init_synthetic(7, 0, M)
    ref = O.Ensemble(x0, 2.0, [0.3, 0.05], [0.5, 0.5])
    sched = A.build_schedule(steps, 10, 10)
    upd = set(A.build_schedule(steps, 10, 4))
    rows_e = open(tmp_path / "energy.dat").read().split("\n")[:-1]
    rows_a = open(tmp_path / "acceptance.dat").read().split("\n")[:-1]
    gd, q = np.zeros(5), 0
    obs = {}
    for t in range(1, steps + 1):
        uc, z, ua = O.draws_philox(7, 0, M, t - 1, 1)
        ref.sweep_replay(uc, z, ua)
        gd += ref.pgmc_replay(3, [1], O.draws_pgmc_philox(7, 0, M, q, 3).reshape(1, 3, M))[0]
        q += 3
        if t in upd:
            ref.sigma[1] = O.learning_step(O.OPT_VPG, 0.05, 0.0, gd[:4] / gd[4], ref.sigma[1])
            gd[:] = 0
        if t in sched:
            obs[t] = (ref.callback_energy(), ref.callback_acceptance())
    for i, t in enumerate(sched):
        te, ve = rows_e[1 + i].split(" ", 1)
        assert int(te) == t and abs(float(ve) - obs[t][0]) < 1e-12
        got = np.array([float(v.replace("NaN", "nan")) for v in rows_a[1 + i].split(" ", 1)[1].strip("[]").split(", ")])
        np.testing.assert_allclose(got, obs[t][1], rtol=1e-12, equal_nan=True)
    # per-chain results are invariant to the sharding (global chain id keys the stream)
    x = np.concatenate([np.load(tmp_path / f"x_rank{r}.npy") for r in range(2)])
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-12)   # σ differs in the last bits (sum order of the all-reduce)
    sig = [np.load(tmp_path / f"sigma_rank{r}.npy") for r in range(2)]
    assert np.array_equal(sig[0], sig[1])                      # every rank applies the same update
    assert sig[0][0] == 0.3 and abs(sig[0][1] - ref.sigma[1]) < 1e-12 and sig[0][1] != 0.05
    assert os.path.exists(tmp_path / "trajectories" / "rank1.bin")


# ---- bench.py: the reference arm runs on the CPU and prints the contract's JSON line -------------------------
def test_bench_reference_arm_prints_the_contract_line():
    import json
    import subprocess
    import sys
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the arm must still use ALL host cores
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "1", "--ref-log2-chains", "12"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["sample"]["chains"] == 1 << 12 and line["sample"]["full_per_gpu_ensemble"] is False
    assert "2^12 chains" in line["cpu_baseline"]["sample"]
    # the workload description is the one our arm prints for the same flags (the driver compares the two)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ours = bench.workload_config(argparse.Namespace(log2_chains=27, scaling="weak", mc_steps=10), 1)
    assert line["config"] == ours and "stores_per_launch" not in ours
    assert bench.launch_plan(20, 11) == [10, 10] and bench.launch_plan(110, 11) == [11] * 10
    assert bench.launch_plan(25, 11) == [9, 9, 7] and bench.launch_plan(3, 11) == [3] and bench.launch_plan(5, 1) == [1] * 5
    # without an explicit sample size the arm sizes it to the time budget (here: tiny budget -> a bounded sample)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-seconds", "0.05"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    auto = json.loads(out.stdout.strip().splitlines()[-1])
    assert 1 << 10 <= auto["sample"]["chains"] < 1 << 27 and auto["config"] == ours
    assert line["impl"] == "reference" and line["metric"] == "metropolis_chain_steps_per_sec"
    assert line["unit"] == "chain-steps/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "C3" in line["config"]["workload"]
    # under torchrun only rank 0 works: the other ranks exit 0 without output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_emits_strict_json(capsys):
    """bench.py's line must be JSON a strict parser accepts: non-finite floats (the PCIe breakdown of a direction a job
    did not use is NaN) become null at any nesting depth; Python's own json would print the non-standard `NaN`."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    bench.emit({"a": float("nan"), "b": [1.5, float("inf"), {"c": float("-inf"), "d": np.float64("nan")}], "e": 3, "f": "x"})
    out = capsys.readouterr().out.strip()

    def strict(name):
        raise AssertionError(name)
    assert json.loads(out, parse_constant=strict) == {"a": None, "b": [1.5, None, {"c": None, "d": None}], "e": 3, "f": "x"}


def test_bench_refuses_to_run_without_a_gpu():
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


_WORKER_SERIES = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
import montecarlo_b200 as mb
from montecarlo_b200 import arianna as A
from fake_engine import OracleEngine
from oracle import oracle as O
A.CudaEnsemble = OracleEngine
M, steps = 77, 300
chains = mb.ParticleEnsemble(O.init_synthetic(9, 0, M), 2.0)
pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
sim = mb.Simulation(chains, (dict(algorithm=mb.Metropolis, pool=pool, seed=9),
                             dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance),
                                  scheduler=mb.build_schedule(steps, 100, 10)),
                             dict(algorithm=mb.StoreTrajectories, scheduler=[150, steps], store_first=False)),
                    steps, path={path!r})
mb.run(sim)
np.save(os.path.join({path!r}, f"x_rank{{chains.rank}}.npy"), chains.x)
np.save(os.path.join({path!r}, f"calls_rank{{chains.rank}}.npy"), np.array([chains.engine.series_calls, chains.engine.launch_count]))
dist.barrier(); dist.destroy_process_group()
"""


def test_two_rank_gloo_series_lookahead(tmp_path):
    """The look-ahead / series path with the chains sharded over 2 ranks: ONE all-reduce per fused stretch, records
    equal to the single-process oracle's callbacks at every store, trajectory frames (barriers) at their own times."""
    port = 31500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(_WORKER_SERIES.format(root=ROOT, port=port, path=str(tmp_path)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    M, steps = 77, 300
    ref = O.Ensemble(O.init_synthetic(9, 0, M), 2.0, [0.1])
    sched = A.build_schedule(steps, 100, 10)
    rows_e = open(tmp_path / "energy.dat").read().split("\n")[:-1]
    rows_a = open(tmp_path / "acceptance.dat").read().split("\n")[:-1]
    done = 0
    for i, t in enumerate(sched):
        _, z, ua = O.draws_philox(9, 0, M, done, t - done, with_cat=False)
        ref.sweep_replay(None, z, ua)
        done = t
        assert int(rows_e[1 + i].split()[0]) == t and abs(float(rows_e[1 + i].split()[1]) / ref.callback_energy() - 1) < 1e-12
        assert abs(float(rows_a[1 + i].split("[")[1][:-1]) / ref.callback_acceptance()[0] - 1) < 1e-12
    x = np.concatenate([np.load(tmp_path / f"x_rank{r}.npy") for r in range(2)])
    assert np.array_equal(x, ref.x)
    for r in range(2):       # stretches: t=100..150 (barrier at 150), 160..300 (barrier at 300) -> 2 series calls
        calls = np.load(tmp_path / f"calls_rank{r}.npy")
        assert calls[0] == 2 and calls[1] == 2


def test_lookahead_ignores_algorithms_that_only_act_in_finalise(tmp_path, fake_engine):
    """StoreLastFrames with the DEFAULT scheduler (every step) has a no-op make_step: it must not act as a barrier."""
    chains = mb.ParticleEnsemble(O.init_synthetic(1, 0, 20), 2.0)
    pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
    sim = mb.Simulation(chains, (dict(algorithm=mb.Metropolis, pool=pool, seed=1),
                                 dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy,),
                                      scheduler=mb.build_schedule(100, 20, 10)),
                                 dict(algorithm=mb.StoreLastFrames)), 100, path=str(tmp_path))
    mb.run(sim)
    assert chains.engine.series_calls == 1 and chains.engine.steps_done == 100
    assert os.path.exists(tmp_path / "trajectories" / "1" / "lastframe.dat")


# ---- a host written against the C ABI alone (examples/harmonic_oscillator.c) --------------------------------
def _build_c_example(tmp_path):
    exe = str(tmp_path / "harmonic_oscillator")
    libdir = os.path.dirname(mb.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "harmonic_oscillator.c"), "-L", libdir, "-larianna_cuda",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_c_host_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: see tests/test_gpu_parity.py::test_c_host_example")
    out = subprocess.run([exe, "1000", "1200"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "no CPU fallback" in out.stderr


# ---- property tests (hypothesis) of the host-side planning helpers ---------------------------------------------
def test_properties_of_schedules_and_shards():
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 5000), st.integers(0, 200), st.integers(1, 300))
    def linear(steps, burn, dt):
        if burn > steps:
            return
        s = mb.build_schedule(steps, burn, dt)
        assert s == N.build_schedule(steps, burn, dt)                       # the oracle's restatement (simulation.jl:95-97)
        assert s[0] == burn and s[-1] == steps and all(a < b for a, b in zip(s, s[1:]))
        assert all((t - burn) % dt == 0 for t in s[:-1])

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 10 ** 9), st.integers(1, 64))
    def shards(n, world):
        parts = [mb.shard_bounds(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))      # contiguous
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1                         # balanced

    linear()
    shards()


def test_lookahead_plan_properties(tmp_path, fake_engine):
    """For random schedules of a trajectory barrier the look-ahead never crosses a barrier, never skips a store and
    the files equal the one-launch-per-store run."""
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    count = [0]

    @settings(max_examples=12, deadline=None)
    @given(st.integers(1, 40), st.integers(0, 30), st.lists(st.integers(1, 120), min_size=0, max_size=4, unique=True))
    def prop(dt, burn, frames):
        count[0] += 1
        steps = 120
        outs = []
        for look in (True, False):
            d = tmp_path / f"p{count[0]}_{int(look)}"
            chains = mb.ParticleEnsemble(O.init_synthetic(3, 0, 16), 2.0)
            pool = (mb.Move(mb.Displacement(0.0), mb.StandardGaussian(), mb.ComponentArray(σ=0.1), 1.0),)
            algs = [dict(algorithm=mb.Metropolis, pool=pool, seed=3),
                    dict(algorithm=mb.StoreCallbacks, callbacks=(mb.callback_energy, mb.callback_acceptance),
                         scheduler=mb.build_schedule(steps, burn, dt))]
            if frames:
                algs.append(dict(algorithm=mb.StoreTrajectories, scheduler=sorted(frames), store_first=False))
            sim = mb.Simulation(chains, tuple(algs), steps, path=str(d))
            sim.lookahead = look
            mb.run(sim)
            outs.append((open(d / "energy.dat").read(), open(d / "acceptance.dat").read(), chains.engine.get_state(),
                         open(d / "trajectories" / "rank0.bin", "rb").read() if frames else b""))
        assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and outs[0][3] == outs[1][3]
        assert np.array_equal(outs[0][2], outs[1][2])

    prop()
