#!/bin/bash
# A/B of the sweep tuning knobs on one GPU: prints value / kernel_ms per variant.
for lib in montecarlo_b200/ab/*.so; do
  ARIANNA_LIB=$PWD/$lib python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$lib', '%.4g'%d['value'], '%.3f ms'%d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/ab.log
