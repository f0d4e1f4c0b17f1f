#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/bench_paths.py 2>&1 | tee gpurun_out/paths.jsonl
