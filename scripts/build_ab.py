"""Builds A/B variants of the sweep tuning knobs into montecarlo_b200/ab/ (git-ignored; travels with gpurun)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecarlo_b200._build import build_library, _HERE
from concurrent.futures import ThreadPoolExecutor

variants = [(256, 4), (224, 4), (192, 5), (128, 7), (160, 6), (256, 3)]   # (threads per CTA, resident CTAs per SM)
os.makedirs(os.path.join(_HERE, "ab"), exist_ok=True)
def one(v):
    b, m = v
    out = os.path.join(_HERE, "ab", f"lib_b{b}_m{m}.so")
    build_library(defines=(f"ARIANNA_BLOCK={b}", f"ARIANNA_MINB={m}"), out=out)
    return out
with ThreadPoolExecutor(8) as ex:
    for o in ex.map(one, variants):
        print(o)
