#!/bin/bash
# Round-2 check B: new multi-move kernel (tests, paths, ncu) + e2e probe.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== paths multi"
timeout 600 python scripts/bench_paths.py multi 2>&1 | tee gpurun_out/paths_multi.jsonl
echo "== e2e probe"
timeout 600 python scripts/e2e_probe.py 2>&1 | tee gpurun_out/e2e_probe.jsonl
echo "== ncu multi-move"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_multi -s 1 -c 1 -f -o gpurun_out/prof_multi2 \
    python scripts/prof_multi.py > gpurun_out/ncu_multi2.log 2>&1
tail -3 gpurun_out/ncu_multi2.log
