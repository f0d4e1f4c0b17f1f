#!/bin/bash
# ncu evidence for the dominant kernel: launch list (shares) + one full-set capture with source.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 32 --warmup 16 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_philox -s 2 -c 1 -f -o gpurun_out/prof_sweep \
    python bench.py --steps 32 --warmup 16 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
