#!/bin/bash
# 8-GPU check: bench (driver flags) with two slice counts, then both arms as the driver runs them.
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
for sl in 16 32; do
echo "== bench N=$N slices $sl"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 --slices $sl 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/bench_n${N}_s$sl.log | \
   python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g strong %.4g parity %s pcie %s kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['strong']['value'], d['parity']['ok'], d['e2e']['pcie_rank0'], d['roofline']['kernel_ms']))"
done
echo "== reference N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/bench_ref_n$N.log | cut -c1-400
