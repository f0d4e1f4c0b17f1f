"""Arianna.PolicyGuided mirrored over the device engine (src/PolicyGuided/*.jl).

The per-sample estimator (sample_gradient_data / pgmc_estimate, gradients.jl:93-121) runs in the CUDA kernel with
the analytic ∂σ log q = δ²/σ³ − 1/σ in place of the ForwardDiff / Zygote / Enzyme backends; this module holds the
host-side pieces: the optimiser rules (learning.jl), the estimator/update algorithms and their event order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from .arianna import AriannaAlgorithm, Metropolis, _dist, allreduce_sums

__all__ = ["Static", "VPG", "BLPG", "BLAPG", "NPG", "ANPG", "BLANPG", "GradientData", "average", "learning_step",
           "PolicyGradientEstimator", "PolicyGradientUpdate", "dlogq_dsigma", "log_proposal_density", "reward"]


# -- optimisers (learning.jl:9-164) ------------------------------------------------------------------------
class PolicyGradient:
    """abstract type PolicyGradient (learning.jl:9)."""


@dataclass(frozen=True)
class Static(PolicyGradient):
    pass


@dataclass(frozen=True)
class VPG(PolicyGradient):
    η: float


@dataclass(frozen=True)
class BLPG(PolicyGradient):
    η: float


@dataclass(frozen=True)
class BLAPG(PolicyGradient):
    δ: float
    ϵid: float = 0.0


@dataclass(frozen=True)
class NPG(PolicyGradient):
    η: float
    ϵid: float = 0.0


@dataclass(frozen=True)
class ANPG(PolicyGradient):
    δ: float
    ϵid: float = 0.0


@dataclass(frozen=True)
class BLANPG(PolicyGradient):
    δ: float
    ϵid: float = 0.0


@dataclass
class GradientData:
    """struct GradientData(j, ∇j, ∇logq_forward, g, n) (gradients.jl:41-47); general P on the host."""
    j: float
    dj: np.ndarray            # ∇j, (P,)
    dlogq_forward: np.ndarray  # ∇logq_forward, (P,)
    g: np.ndarray             # (P, P)
    n: int

    @staticmethod
    def zero(P: int = 1) -> "GradientData":  # initialise_gradient_data (gradients.jl:54-61)
        return GradientData(0.0, np.zeros(P), np.zeros(P), np.zeros((P, P)), 0)

    def __add__(self, o: "GradientData") -> "GradientData":  # gradients.jl:68-76
        return GradientData(self.j + o.j, self.dj + o.dj, self.dlogq_forward + o.dlogq_forward, self.g + o.g,
                            self.n + o.n)

    @staticmethod
    def from_record(rec: Sequence[float]) -> "GradientData":
        """From the engine's summed record (j, ∇j, ∇logq_f, g, n), P = 1."""
        return GradientData(float(rec[0]), np.array([rec[1]]), np.array([rec[2]]), np.array([[rec[3]]]),
                            int(round(rec[4])))


def average(gd: GradientData) -> GradientData:  # gradients.jl:83-85  (n == 0 gives NaN/Inf like Julia's x / 0)
    with np.errstate(all="ignore"):
        n = np.float64(gd.n)
        return GradientData(float(np.float64(gd.j) / n), gd.dj / n, gd.dlogq_forward / n, gd.g / n, gd.n)


def learning_step(parameters: np.ndarray, gd: GradientData, opt: PolicyGradient) -> None:
    """learning_step!(parameters, gd, opt): in-place update of θ (learning.jl:32-164)."""
    I = np.eye(parameters.size)
    if isinstance(opt, VPG):                                                       # :32-34
        parameters[:] = parameters + opt.η * gd.dj
    elif isinstance(opt, BLPG):                                                    # :50-52
        parameters[:] = parameters + opt.η * (gd.dj - gd.j * gd.dlogq_forward)
    elif isinstance(opt, BLAPG):                                                   # :76-79
        η = math.sqrt(2 * opt.δ / (float(gd.dj @ gd.dj) + opt.ϵid))
        parameters[:] = parameters + η * (gd.dj - gd.j * gd.dlogq_forward)
    elif isinstance(opt, NPG):                                                     # :103-105
        parameters[:] = parameters + opt.η * np.linalg.inv(gd.g + opt.ϵid * I) @ gd.dj
    elif isinstance(opt, ANPG):                                                    # :130-134
        Finv = np.linalg.inv(gd.g + opt.ϵid * I)
        η = math.sqrt(2 * opt.δ / float(gd.dj @ (Finv @ gd.dj)))
        parameters[:] = parameters + η * Finv @ gd.dj
    elif isinstance(opt, BLANPG):                                                  # :159-164
        Finv = np.linalg.inv(gd.g + opt.ϵid * I)
        bj = gd.dj - gd.j * gd.dlogq_forward
        η = math.sqrt(2 * opt.δ / float(bj @ (Finv @ bj)))
        parameters[:] = parameters + η * Finv @ bj
    elif isinstance(opt, Static):
        pass
    else:
        raise TypeError(f"no learning_step! for {type(opt).__name__}")


# -- scalar helpers mirroring the particle_1d methods (host side; the kernel holds the device versions) --------
def log_proposal_density(delta: float, sigma: float) -> float:
    """-(δ)^2 / (2σ^2) - log(2π * σ^2) / 2 (particle_1d.jl:52-54)."""
    return -(delta * delta) / (2 * (sigma * sigma)) - math.log(2 * math.pi * (sigma * sigma)) / 2


def dlogq_dsigma(delta: float, sigma: float) -> float:
    """Analytic ∂σ log q = δ²/σ³ − 1/σ (replaces withgrad_log_proposal_density!, gradients.jl:28-33)."""
    return (delta * delta) / (sigma * sigma * sigma) - 1.0 / sigma


def reward(delta: float) -> float:
    """reward(action, system) = δ^2 (particle_1d.jl:42-44)."""
    return delta * delta


# -- algorithms ----------------------------------------------------------------------------------------------
class PolicyGradientEstimator(AriannaAlgorithm):
    """PolicyGradientEstimator(chains; dependencies=(Metropolis,), optimisers, q_batch_size=1, ...)
    (estimator.jl:38-147).  make_step = one fused device pass per learnable move, accumulated on the device."""

    def __init__(self, chains, *, dependencies=None, optimisers=None, q_batch_size: int = 1, ad_backend=None,
                 R=None, parallel: bool = False, **extras):
        assert dependencies is not None and len(dependencies) == 1        # estimator.jl:104-105
        assert isinstance(dependencies[0], Metropolis)
        metropolis = dependencies[0]
        self.pool = metropolis.pool
        self.optimisers = tuple(optimisers)
        assert len(self.optimisers) == len(self.pool)                     # :66
        # learnable moves: optimiser is not Static (:72); 0-based here, 1-based in the reference
        self.learn_ids = [k for k, o in enumerate(self.optimisers) if not isinstance(o, Static)]
        self.q_batch_size = int(q_batch_size)
        self.parameters_list = [m.parameters for m in self.pool]
        self.seed = metropolis.seed
        self.chains = chains
        self.steps_since_update = 0

    def make_step(self, simulation):                                      # estimator.jl:111-134
        ch = simulation.chains
        ch.flush()                       # the estimator observes the chains after this step's Metropolis sweep
        ch._push_params()
        ch.engine.pgmc_estimate(self.q_batch_size, self.learn_ids)
        self.steps_since_update += 1

    def gradients_data(self):
        """Accumulated records (summed over all ranks) per learnable move -- gradients_data[k] of the reference."""
        ch = self.chains
        n = len(self.learn_ids)
        dist = _dist()
        if dist is not None and dist.get_world_size() > 1 and dist.get_backend() == "nccl":
            import torch
            with torch.cuda.stream(ch.engine.torch_stream()):
                flat = allreduce_sums(None, ch.engine.pgmc_sums_tensor())
            recs = flat[:5 * n].reshape(n, 5)
        else:
            recs = allreduce_sums(ch.engine.pgmc_read(n).ravel()).reshape(n, 5)
        return [GradientData.from_record(r) for r in recs]

    @property
    def objectives(self):                                                 # objectives[k] = j / n (:131)
        return [gd.j / gd.n if gd.n else float("nan") for gd in self.gradients_data()]

    def write_algorithm(self, io, scheduler):                             # :136-147
        io.write("\tPolicyGradientEstimator\n")
        io.write(f"\t\tCalls: {sum(1 for x in scheduler if 0 < x <= scheduler[-1])}\n")
        io.write(f"\t\tLearnable moves: {[k + 1 for k in self.learn_ids]}\n")
        io.write(f"\t\tQ batch size: {self.q_batch_size}\n\t\tAD backend: analytic (CUDA)\n\t\tSeed: {self.seed}\n")


def _opt_tuple(opt: PolicyGradient):
    """(kind, p1, p2) of an optimiser for arianna_pgmc_update_device."""
    if isinstance(opt, (VPG, BLPG)):
        return type(opt).__name__, opt.η, 0.0
    if isinstance(opt, NPG):
        return "NPG", opt.η, opt.ϵid
    if isinstance(opt, (BLAPG, ANPG, BLANPG)):
        return type(opt).__name__, opt.δ, opt.ϵid
    return "Static", 0.0, 0.0


class PolicyGradientUpdate(AriannaAlgorithm):
    """PolicyGradientUpdate(chains; dependencies=(PolicyGradientEstimator,)) (update.jl:14-67).

    on_device=True (extension): the averaging, the learning_step! of every learnable move and the reset of the
    accumulators run in ONE tiny kernel on the device (arianna_pgmc_update_device), the sweeps read σ from a
    device-resident block, and the host only pulls σ when something asks for it (StoreParameters, the end of the run):
    no all-reduce read-back and no host synchronisation per update."""

    def __init__(self, chains, *, dependencies=None, on_device: bool = False, **extras):
        self.on_device = bool(on_device)
        assert dependencies is not None and len(dependencies) == 1        # update.jl:43-44
        assert isinstance(dependencies[0], PolicyGradientEstimator)
        self.pge = dependencies[0]
        self.optimisers = self.pge.optimisers
        self.learn_ids = self.pge.learn_ids
        self.parameters_list = self.pge.parameters_list

    def make_step(self, simulation):                                      # update.jl:50-57
        ch = simulation.chains
        if self.on_device and hasattr(ch.engine, "pgmc_update_device"):
            dist = _dist()
            if dist is not None and dist.get_world_size() > 1:
                # gradients_data summed over the ranks IN PLACE on the device, then the same update everywhere
                import torch
                with torch.cuda.stream(ch.engine.torch_stream()):
                    dist.all_reduce(ch.engine.pgmc_sums_tensor())
            ch._push_params()
            ch.engine.pgmc_update_device(self.learn_ids, [_opt_tuple(self.optimisers[k]) for k in self.learn_ids])
            ch._params_on_device = True
            self.pge.steps_since_update = 0
            return
        gds = self.pge.gradients_data()
        for k, lid in enumerate(self.learn_ids):
            gd = average(gds[k])
            learning_step(self.parameters_list[lid].data, gd, self.optimisers[lid])
            ch.engine.set_params(lid, self.parameters_list[lid].σ)        # raises if σ left (0, ∞), like Normal(0, σ)
        ch.engine.pgmc_reset()                                            # gradients_data[k] = initialise_gradient_data
        self.pge.steps_since_update = 0

    def write_algorithm(self, io, scheduler):                             # update.jl:59-67
        io.write("\tPolicyGradientUpdate\n")
        io.write(f"\t\tCalls: {sum(1 for x in scheduler if 0 < x <= scheduler[-1])}\n")
        io.write(f"\t\tLearnable moves: {[k + 1 for k in self.learn_ids]}\n\t\tOptimisers:\n")
        for k, opt in enumerate(self.optimisers, 1):
            io.write(f"\t\t\tMove {k}: {opt}\n")
