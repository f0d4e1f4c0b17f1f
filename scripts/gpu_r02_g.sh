#!/bin/bash
# quick A/B of the current build: headline (driver flags, no e2e), multi-move and PGMC paths, math + parity subset
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 44 --warmup 8 --no-cpu-baseline --no-e2e --no-strong 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('headline value %.4g kernel %.3f ms per store %.4f parity %s' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms']/d['engine']['stores_per_launch'], d['parity']['ok']), d['parity']['x_bits_checksum'])"
timeout 300 python scripts/bench_paths.py multi pgmc 2>&1 | cut -c1-140
timeout 600 python -m pytest tests/test_gpu_math.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
