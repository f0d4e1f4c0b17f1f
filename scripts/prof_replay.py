"""Profiling target: replay sweep, M = 2^24, K = 32, draws resident in HBM, decisions written."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlo_b200 as mb

M, K = 1 << 24, 32
with mb.CudaEnsemble(M, 2.0, [0.1], seed=42, arith="exact") as eng:
    eng.init_synthetic()
    with torch.cuda.stream(eng.torch_stream()):
        z = torch.randn((K, M), dtype=torch.float64, device="cuda")
        ua = torch.rand((K, M), dtype=torch.float64, device="cuda")
        dec = torch.empty((K, M), dtype=torch.uint8, device="cuda")
    eng.synchronize(); torch.cuda.synchronize()
    for _ in range(3):
        eng.sweep_replay_device(K, 0, z.data_ptr(), ua.data_ptr(), dec.data_ptr())
    eng.synchronize()
