"""Builds A/B variants of the sweep tuning knobs into montecarlo_b200/ab/ (git-ignored; travels with gpurun)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecarlo_b200._build import build_library, _HERE
from concurrent.futures import ThreadPoolExecutor

variants = [(m, p) for m in (2, 3, 4) for p in (0, 1)]
def one(v):
    m, p = v
    out = os.path.join(_HERE, "ab", f"lib_m{m}_p{p}.so")
    build_library(defines=(f"ARIANNA_MINB={m}", f"ARIANNA_PIPE={p}"), out=out)
    return out
with ThreadPoolExecutor(8) as ex:
    for o in ex.map(one, variants):
        print(o)
