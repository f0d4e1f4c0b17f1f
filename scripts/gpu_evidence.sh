#!/bin/bash
# Round-2 evidence run on ONE GPU: parity tests, smoke, bench (driver flags and defaults, both arms), every secondary
# path with clocks, launch list + ncu captures of the sweep kernels.  Outputs land in gpurun_out/ (copied to profiles/).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== reference arm (driver flags)"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ref_n1.json; cut -c1-200 gpurun_out/bench_ref_n1.json
echo "== bench (driver flags)"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json; cut -c1-200 gpurun_out/bench_n1.json
echo "== bench (defaults)"
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_n1_default.json; cut -c1-200 gpurun_out/bench_n1_default.json
echo "== bench --series 1"
timeout 600 python bench.py --steps 20 --warmup 3 --series 1 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1_series1.json; cut -c1-160 gpurun_out/bench_n1_series1.json
echo "== paths"
timeout 1500 python scripts/bench_paths.py 2>&1 | tee gpurun_out/paths.jsonl | cut -c1-170
echo "== ncu"
bash scripts/gpu_profile.sh
if [ "${SKIP_MULTI:-0}" = "0" ]; then      # (gpurun brings back at most 64 MiB: four .ncu-rep files are too many)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_multi -s 1 -c 1 -f -o gpurun_out/prof_multi_final \
    python scripts/prof_multi.py > gpurun_out/ncu_multi_final.log 2>&1
tail -2 gpurun_out/ncu_multi_final.log
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pgmc_kernel -s 1 -c 1 -f -o gpurun_out/prof_pgmc \
    python scripts/prof_pgmc.py > gpurun_out/ncu_pgmc.log 2>&1
tail -2 gpurun_out/ncu_pgmc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_f32 -s 1 -c 1 -f -o gpurun_out/prof_f32 \
    python scripts/prof_f32.py > gpurun_out/ncu_f32.log 2>&1
tail -2 gpurun_out/ncu_f32.log
