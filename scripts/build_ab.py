"""Builds A/B variants of the sweep tuning knobs into montecarlo_b200/ab/ (git-ignored; travels with gpurun).
Usage: python scripts/build_ab.py [name=DEF1,DEF2 ...]   (default: the occupancy / software-pipeline grid below)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecarlo_b200._build import build_library, _HERE
from concurrent.futures import ThreadPoolExecutor

variants = {
    "base": (),
    "minb5": ("ARIANNA_MINB=5",),                      # 48 registers, 10 warps per scheduler
    "minb6": ("ARIANNA_MINB=6",),                      # 40 registers, 12 warps per scheduler
    "pipe_minb4": ("ARIANNA_PIPE=1",),
    "pipe_minb3": ("ARIANNA_PIPE=1", "ARIANNA_MINB=3"),
    "minb3": ("ARIANNA_MINB=3",),
}
for a in sys.argv[1:]:
    name, defs = a.split("=", 1)
    variants = {**variants, name: tuple(d for d in defs.split(",") if d)}
os.makedirs(os.path.join(_HERE, "ab"), exist_ok=True)
def one(kv):
    name, defs = kv
    out = os.path.join(_HERE, "ab", f"lib_{name}.so")
    build_library(defines=defs, out=out)
    return out
with ThreadPoolExecutor(8) as ex:
    for o in ex.map(one, variants.items()):
        print(o)
