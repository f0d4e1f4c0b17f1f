#!/bin/bash
# After a change of the math layer's rounding: re-record the 1-GPU checksum of bench.py's parity mini-run, then the full
# GPU suite and the bench with the driver's flags.
set -u
mkdir -p gpurun_out
chk=$(timeout 300 python bench.py --log2-chains 22 --steps 10 --warmup 3 --no-e2e --no-strong --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.readline())['parity']['x_bits_checksum'])")
echo "checksum $chk"
python tests/golden/make_bench_parity.py "$chk" > /dev/null && cp tests/golden/bench_parity.json gpurun_out/bench_parity.json
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench (driver flags)"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json'))
print('value %.4g e2e %.4g parity %s kernel_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], d['parity']['ok'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
