#!/bin/bash
# multi-GPU check: bench at N GPUs (driver flags), both arms, + the 2-GPU tests.  Usage: gpurun --gpus N -- bash scripts/gpu_r02_n.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== pytest two_gpu"
  timeout 900 python -m pytest tests -m gpu -q -x -k "two_gpu" 2>&1 | tail -5 | tee gpurun_out/pytest_2gpu.log
fi
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/bench_n$N.log
echo "== reference N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>&1 | grep -v Warning | tail -2 | tee gpurun_out/bench_ref_n$N.log
