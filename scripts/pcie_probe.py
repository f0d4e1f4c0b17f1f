"""Concurrent host<->device bandwidth per rank (run under torchrun): uploads from default page-locked vs write-combined
memory, downloads, and both directions at once; 1 GiB each, every rank at the same time.  GB/s per rank."""
import json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlo_b200 as mb

rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 27
eng = mb.CudaEnsemble(n, 2.0, [0.1], device=lr)
pinned, wc = mb.HostBuffer(n), mb.HostBuffer(n, write_combined=True)
pinned.array[:] = 1.0
wc.array[:] = 1.0
out = {}

def timed(fn):
    best = 1e9
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return 8 * n / best / 1e9

out["h2d_pinned"] = timed(lambda: eng.set_state_from_ptr(pinned.ptr))
out["h2d_write_combined"] = timed(lambda: eng.set_state_from_ptr(wc.ptr))
out["d2h_pinned"] = timed(lambda: eng.get_state_to_ptr(pinned.ptr))
tp = torch.empty(n, dtype=torch.float64).pin_memory(); dev = torch.empty(n, dtype=torch.float64, device="cuda")
tp2 = torch.empty(n, dtype=torch.float64).pin_memory(); dev2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dev.copy_(tp, non_blocking=True)
    with torch.cuda.stream(s2): tp2.copy_(dev2, non_blocking=True)
    s1.synchronize(); s2.synchronize()
out["duplex_each_direction"] = timed(both)
print(json.dumps({"rank": rank, **{k: round(v, 1) for k, v in out.items()}}), flush=True)
eng.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
