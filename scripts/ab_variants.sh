#!/bin/bash
# A/B of the sweep tuning knobs on one GPU: value / kernel_ms / stores per launch of every library in montecarlo_b200/ab/.
mkdir -p gpurun_out
for lib in montecarlo_b200/ab/*.so; do
  ARIANNA_LIB=$PWD/$lib timeout 300 python bench.py --steps ${AB_STEPS:-44} --warmup 8 --no-cpu-baseline --no-e2e --no-strong --no-parity 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$lib', 'stores/launch', d['engine']['stores_per_launch'], 'value %.4g'%d['value'], 'kernel %.3f ms'%d['roofline']['kernel_ms'], 'per store %.4f ms'%(d['roofline']['kernel_ms']/d['engine']['stores_per_launch']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done 2>&1 | tee gpurun_out/ab.log
