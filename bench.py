#!/usr/bin/env python
"""bench.py -- Metropolis chain-steps/sec of the fused sweep on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], "C3"): particle_1d harmonic, β = 2, Gaussian displacement σ = 0.1, Float64,
M = 2^27 chains per GPU, StoreCallbacks energy/acceptance every 10 MC steps.  One bench "step" = one store
interval = 10 Metropolis steps over every local chain + its callback record (Σe, Σacc/tot, count).  By default
the engine's preferred number of store intervals (11 on B200) is fused into ONE launch (arianna_sweep_series: chains
stay in registers across the intervals, the records are reduced on the device and all-reduced (N > 1) / copied to
the host together);
`--series 1` is the one-launch-per-store path (arianna_sweep with the reduction fused at its tail).  Chains are
independent, so they shard over ranks with no data-path collective: weak scaling, per-GPU work fixed
(`--scaling strong` keeps the total at 2^27).

  python bench.py [--gpus N --steps K --warmup W]                  # our arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                           # the CPU restatement of the reference path

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "metropolis_chain_steps_per_sec"
UNIT = "chain-steps/s"
FLOPS_PER_CHAIN_STEP = 110.0          # SURVEY.md §8d convention (12 plain + exp 36 + ½(log 44 + sqrt 12 + sincos 70))
BYTES_PER_CHAIN_PER_LAUNCH = 24.0     # x f64 read+write, acc u32 read+write


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=110)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-chains", type=int, default=27, help="chains per GPU (weak) or in total (strong), log2")
    ap.add_argument("--mc-steps", type=int, default=10, help="Metropolis steps per store interval")
    ap.add_argument("--series", type=int, default=0,
                    help="store intervals fused per launch (0 = the engine's preferred count, 1 = one launch per store)")
    ap.add_argument("--slices", type=int, default=8, help="chain slices of the pipelined end-to-end job")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--arith", default="fast", choices=["fast", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-log2-chains", type=int, default=20, help="bounded sample of the workload for the CPU arm")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="time budget of the cpu_baseline leg")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks: NVML sampled in a thread DURING the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of Metropolis.make_step! with parallel=true (OpenMP over chains)
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_rate(log2_chains: int, mc_steps: int, steps: int, warmup: int, budget_s: float = 20.0):
    """chain-steps/s of the reference path on the host cores, on a bounded sample of the workload: 2^log2_chains
    chains, `steps` store intervals of mc_steps Metropolis steps + the energy/acceptance callbacks."""
    from oracle import oracle as O
    M = 1 << log2_chains
    ens = O.Ensemble(O.init_synthetic(42, 0, M), 2.0, [0.1])
    ens.seed_xoshiro(42)
    for _ in range(max(1, warmup)):
        ens.sweep_xoshiro(mc_steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        ens.sweep_xoshiro(mc_steps)
        ens.callback_energy()
        ens.callback_acceptance()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return M * mc_steps * done / dt, done, dt, O.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    rate, done, dt, cores = cpu_reference_rate(args.ref_log2_chains, args.mc_steps, args.steps, args.warmup, 60.0)
    sample = (f"2^{args.ref_log2_chains} chains x {done} store intervals of {args.mc_steps} MC steps "
              f"(C++ restatement of mc_sweep!, xoshiro256++/ziggurat, OpenMP over chains; Julia is absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    m_local = chains_per_rank(args, world)
    return {
        "workload": "C3: particle_1d harmonic beta=2, Gaussian Displacement sigma=0.1, Metropolis + StoreCallbacks "
                    "energy/acceptance every 10 steps (BASELINE.json configs[2])",
        "chains_per_gpu": m_local, "chains_total": m_local * world, "mc_steps_per_store": args.mc_steps,
        "stores_per_launch": args.series, "mc_steps_per_launch": args.mc_steps * args.series,
        "rng": "philox4x32-10 + box-muller (native mode)", "arith": args.arith,
        "l2_policy": "inputs larger than L2 (x + counters = %.0f MiB per GPU vs 126 MB L2)" % (m_local * 12 / 2 ** 20),
        "parallelism": f"chains sharded x{world}, NCCL all-reduce of 3 doubles per store"
                       + (f", {args.series} stores per all-reduce" if args.series > 1 else ""),
    }


def ncu_issue(m_local, mc_steps, series, chain_steps_per_s, sm_count=148, smsp=4, ghz=1.965):
    """Issue-slot view of the same launch: warp instructions per warp-step from the committed ncu capture x the
    measured step rate, against one instruction per SMSP per clock.  (The sweep is issue-bound: an IMAD.WIDE holds
    the dispatch port for ~4.5 cycles, profiles/microbench.)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t["chains"] == m_local and t["mc_steps"] == mc_steps and t.get("series", 1) == series:
            ach = t["warp_inst_per_warp_step"] * chain_steps_per_s / 32.0
            peak = sm_count * smsp * ghz * 1e9
            return {"achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": ach / peak,
                    "warp_inst_per_warp_step": t["warp_inst_per_warp_step"],
                    "ncu_issue_active_pct": t.get("issue_active_pct"), "source": t["source"]}
    except Exception:
        pass
    return None


def ncu_traffic(m_local, mc_steps, series=1):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sweep kernel from the committed `ncu --set full`
    capture (profiles/traffic.json, written by scripts/summarise_profile.py); only valid for the captured shape."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t["chains"] == m_local and t["mc_steps"] == mc_steps and t.get("series", 1) == series:
            return {"bytes_per_launch": t["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": 24 * m_local,
                    "source": t["source"]}
    except Exception:
        pass
    return None


def chains_per_rank(args, world):
    total = 1 << args.log2_chains
    return total if args.scaling == "weak" else total // world


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import montecarlo_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    m_local = chains_per_rank(args, world)
    K, W, S = args.steps, max(3, args.warmup), args.mc_steps
    stream = torch.cuda.Stream(device=local_rank)      # torch owns the stream; the engine launches on it
    eng = mb.CudaEnsemble(m_local, 2.0, [0.1], [1.0], seed=42, chain_offset=rank * m_local,
                          n_chains_total=m_local * world, arith=args.arith, device=local_rank,
                          stream=stream.cuda_stream)
    G = args.series if args.series > 0 else eng.series_per_launch    # store intervals fused per launch
    G = max(1, min(G, 64))
    args.series = G
    sums_host = torch.empty(3 * G, dtype=torch.float64).pin_memory()

    def one_launch(n, timed_events=None):
        """n store intervals: one fused launch (+ the record fold), all-reduce (N > 1), async D2H of the 3n sums."""
        if timed_events is not None:
            timed_events[0].record(stream)
        if G == 1:
            eng.sweep(S, reduce=True)
        else:
            eng.sweep_series([S] * n, read=False)
        if timed_events is not None:
            timed_events[1].record(stream)
        t = eng.callback_sums_tensor() if G == 1 else eng.series_tensor()
        if world > 1:
            t = t.clone()
            dist.all_reduce(t)
        sums_host[:3 * n].copy_(t, non_blocking=True)

    def groups(total):
        return [min(G, total - g) for g in range(0, total, G)]

    with torch.cuda.stream(stream):
        eng.init_synthetic()
        fp64_peak = eng.measure_fp64_peak()
        for n in groups(W):
            one_launch(n)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = eng.launch_count
        sampler = ClockSampler(local_rank)
        sampler.start()
        plan = groups(K)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plan]
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        for n, e in zip(plan, ev):
            one_launch(n, e)
        stop.record(stream)
        stop.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        if world > 1:
            dist.barrier()
        launches = eng.launch_count - launches0
        ms_total = start.elapsed_time(stop)
        # average duration of a FULL launch (G stores) and the MC steps it covers; a ragged last group is left out
        full = [a.elapsed_time(b) for n, (a, b) in zip(plan, ev) if n == min(G, K)]
        kern_ms = float(np.mean(full))
        steps_per_launch = S * min(G, K)
        tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
        energy = float(sums_host[3 * plan[-1] - 3] / sums_host[3 * plan[-1] - 1])

        # ---- e2e: the same job through the C ABI with HOST buffers inside the timed region --------------------
        e2e = None
        if not args.no_e2e:
            x_in = torch.empty(m_local, dtype=torch.float64).pin_memory()
            x_out = torch.empty(m_local, dtype=torch.float64).pin_memory()
            eng.get_state_to_ptr(x_in.data_ptr())
            vals = np.empty(3)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if G == 1:
                eng.set_state_from_ptr(x_in.data_ptr())            # H2D: the job's chains, pinned host -> HBM
                for n in plan:
                    eng.set_params(0, 0.1)                         # the launch's input: policy parameters θ = (σ)
                    eng.sweep(S, reduce=True)
                    if world > 1:
                        t = eng.callback_sums_tensor().clone()
                        dist.all_reduce(t)
                        vals = t.cpu().numpy()                     # D2H: the step's result (3 doubles)
                    else:
                        vals = eng.callback_sums()                 # D2H through arianna_callback_sums
                eng.get_state_to_ptr(x_out.data_ptr())             # D2H: final chains (StoreLastFrames)
            else:
                # the whole job in ONE C-ABI call: chains in, K store intervals, records out, chains out; the library
                # pipelines slices of chains so that the copies overlap the sweeps (arianna_run_host_job)
                eng.set_params(0, 0.1)
                rec = eng.run_host_job([S] * K, x_in=x_in.data_ptr(), x_out=x_out.data_ptr(), n_slices=args.slices,
                                       read=(world == 1))
                if world > 1:
                    t = eng.series_tensor().clone()
                    dist.all_reduce(t)
                    vals = t.cpu().numpy()[-3:]                    # D2H: K records of 3 doubles
                else:
                    vals = rec[-1]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": m_local * world * S * K / dt, "unit": UNIT,
                   "h2d_bytes_per_step": int(8 * m_local / K + 8), "d2h_bytes_per_step": int(8 * m_local / K + 24),
                   "note": ("timed: ONE arianna_run_host_job call = pinned-host x0 -> HBM, K store intervals, records -> "
                            f"host, final x -> pinned host, pipelined over {args.slices} slices of chains" if G > 1 else
                            "timed: pinned-host x0 -> HBM once, per step sigma in + callback sums out, final x -> "
                            "pinned host; one-off copies amortised over the K steps"),
                   "energy": float(vals[0] / vals[2])}

    value = m_local * world * S * K / (ms_total * 1e-3)
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ach_tf = FLOPS_PER_CHAIN_STEP * m_local * steps_per_launch / (kern_ms * 1e-3) / 1e12
        ach_gb = BYTES_PER_CHAIN_PER_LAUNCH * m_local / (kern_ms * 1e-3) / 1e9
        roofline = {
            "bound": "fp64", "kernel": "sweep_philox_kernel<HARMONIC,%s,single-move%s>" % (
                args.arith.upper(), ",series" if G > 1 else ""),
            "achieved": ach_tf, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": ach_tf / (fp64_peak / 1e12),
            "peak_source": "DFMA microbenchmark in this run (arianna_measure_fp64_peak; MEASURED_PEAKS.json has no "
                           "FP64 entry); nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2",
            "note": "issue-bound FP64/integer kernel, no tensor-core work: neither 'hbm' nor 'tensor' applies. achieved = "
                    "110 conventional FP64 flop per chain-step (SURVEY.md 8d, fixed before the build) x steps/s; the "
                    "kernel executes far fewer (tables, FP32 filter), so frac may exceed 1 -- see 'issue' for the "
                    "hardware-side fraction and 'hbm' for the memory view",
            "flops_per_chain_step": FLOPS_PER_CHAIN_STEP, "kernel_ms": kern_ms, "mc_steps_per_launch": steps_per_launch,
            "traffic": ncu_traffic(m_local, S, G),
            "issue": ncu_issue(m_local, S, G, m_local * steps_per_launch / (kern_ms * 1e-3)),
            "hbm": {"achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
                    "bytes_per_chain_step": BYTES_PER_CHAIN_PER_LAUNCH / steps_per_launch,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "mean_energy": energy,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            rate, done, dt, cores = cpu_reference_rate(args.ref_log2_chains, S, 10 ** 9, 2, args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"2^{args.ref_log2_chains} chains x {done} store intervals of {S} MC steps in {dt:.1f} s "
                          "(oracle: C restatement of mc_sweep! with xoshiro256++/ziggurat, OpenMP over chains)"}
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    eng.close()
    del sums_host
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
