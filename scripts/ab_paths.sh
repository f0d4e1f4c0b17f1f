#!/bin/bash
# A/B of whole builds (montecarlo_b200/ab/*.so): headline store interval + multi-move and PGMC paths
mkdir -p gpurun_out
for lib in montecarlo_b200/ab/*.so; do
  echo "=== $lib"
  ARIANNA_LIB=$PWD/$lib timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e --no-strong --no-parity 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('headline %.4g  per store %.4f ms (%d stores/launch)' % (d['value'], d['roofline']['kernel_ms']/d['engine']['stores_per_launch'], d['engine']['stores_per_launch']))"
  ARIANNA_LIB=$PWD/$lib timeout 300 python scripts/bench_paths.py multi pgmc 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   %-70s %8.4f ms' % (d['path'][:70], d['ms']))"
done 2>&1 | tee gpurun_out/ab_paths.log
