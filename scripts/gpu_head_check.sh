#!/bin/bash
# HEAD check on one GPU: full parity suite, smoke, compute-sanitizer over every kernel family.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
bash scripts/gpu_sanitize.sh
