"""montecarlo_b200 -- B200-native (sm_100a) multi-chain Metropolis engine behind Arianna.jl's API.

Only what the hot path needs lives here:
  csrc/            hand-written CUDA kernels + the C ABI (libarianna_cuda.so, include/arianna_cuda.h)
  engine.py        CudaEnsemble: 1:1 ctypes wrapper over the C ABI
  arianna.py       host-side mirror of Arianna's Simulation / run! / Metropolis / Store* API
  policy_guided.py host-side mirror of Arianna.PolicyGuided (optimisers, estimator / update algorithms)
  julia_rng.py     Julia's Xoshiro(seed) seeding (SHA-256 of the seed's limbs) for the device xoshiro256++ generator

There is no CPU fallback: importing the engine without the built library raises ImportError, and creating an
ensemble without a CUDA device raises AriannaError(ERR_NO_DEVICE).
"""
from . import _lib
from ._build import LIB_PATH, build_library
from ._lib import AriannaError
from .engine import CudaEnsemble, HostBuffer
from .arianna import *  # noqa: F401,F403
from . import julia_rng
from . import policy_guided
from . import policy_guided as PolicyGuided  # `using Arianna.PolicyGuided`

__version__ = "0.1.0"
