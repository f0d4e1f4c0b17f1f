#!/bin/bash
# One gpurun round trip: parity tests, smoke, bench, launch list.  Usage: gpurun -- bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 100 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench.log
