#!/bin/bash
# Round-2 check C: replay kernel with bulk copies (tests, paths with/without TMA, ncu).
set -u
mkdir -p gpurun_out
echo "== pytest replay"
timeout 1700 python -m pytest tests -m gpu -q -x -k "replay or golden" 2>&1 | tail -8 | tee gpurun_out/pytest_replay.log
echo "== paths replay (TMA)"
timeout 600 python scripts/bench_paths.py replay 2>&1 | tee gpurun_out/paths_replay.jsonl
echo "== paths replay (per-thread loads)"
ARIANNA_REPLAY_TMA=0 timeout 600 python scripts/bench_paths.py replay 2>&1 | tee gpurun_out/paths_replay_old.jsonl
echo "== ncu replay"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_replay -s 1 -c 1 -f -o gpurun_out/prof_replay \
    python scripts/prof_replay.py > gpurun_out/ncu_replay.log 2>&1
tail -3 gpurun_out/ncu_replay.log
