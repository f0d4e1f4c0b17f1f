#!/bin/bash
# A/B of the sweep tuning knobs on one GPU: prints value / kernel_ms per variant (series and per-store paths).
mkdir -p gpurun_out
for lib in montecarlo_b200/ab/*.so; do
  for s in 0 1; do
  ARIANNA_LIB=$PWD/$lib python bench.py --steps 66 --warmup 11 --series $s --no-cpu-baseline --no-e2e 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$lib', 'series', d['config']['stores_per_launch'], '%.4g'%d['value'], '%.3f ms'%d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
  done
done 2>&1 | tee gpurun_out/ab.log
