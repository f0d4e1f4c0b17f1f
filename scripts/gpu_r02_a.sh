#!/bin/bash
# Round-2 check A: parity tests, bench (driver flags), e2e slice sweep, ncu of the 7-move sweep kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q -x --durations=12 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench (driver flags)"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_s20.log
for sl in 8 16 64; do
  echo "== bench slices $sl"
  timeout 300 python bench.py --steps 20 --warmup 3 --slices $sl --no-cpu-baseline --no-strong --no-parity 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['pcie_rank0'])" | tee -a gpurun_out/bench_slices.log
done
echo "== reference arm (driver flags)"
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) 2>&1 | tail -6 | tee gpurun_out/bench_ref_s20.log
echo "== ncu multi-move"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_philox -s 1 -c 1 -f -o gpurun_out/prof_multi \
    python scripts/prof_multi.py > gpurun_out/ncu_multi.log 2>&1
tail -3 gpurun_out/ncu_multi.log
