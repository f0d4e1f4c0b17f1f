#!/bin/bash
# Round-2 check E: Julia-manual known answers on the device generator, instruction-mix floor microbenchmarks (output
# kept), A/B of occupancy / software-pipeline variants of the headline kernel.
set -u
mkdir -p gpurun_out
echo "== pytest (xoshiro / julia)"
timeout 600 python -m pytest tests -m gpu -q -x -k "xoshiro or julia or config2" 2>&1 | tail -4 | tee gpurun_out/pytest_julia.log
echo "== microbench"
( cd profiles/microbench
  for b in mix mix_nowide; do echo "--- $b"; timeout 120 ./$b; done ) 2>&1 | tee gpurun_out/mix_microbench.txt
echo "== A/B"
bash scripts/ab_variants.sh
