#!/bin/bash
# multi-GPU evidence: both arms at N GPUs with the driver's flags (+ the 2-GPU NCCL tests at N = 2).
# Usage: gpurun --gpus N -- bash scripts/gpu_evidence_n.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" = "2" ] && [ "${SKIP_TESTS:-0}" = "0" ]; then
  echo "== pytest two_gpu"
  timeout 900 python -m pytest tests -m gpu -q -x -k "two_gpu" 2>&1 | tail -3 | tee gpurun_out/pytest_2gpu.log
fi
if [ "${SKIP_REF:-0}" = "0" ]; then
echo "== reference N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 20 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_ref_n$N.json; cut -c1-160 gpurun_out/bench_ref_n$N.json
fi
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json'))
print('value %.4g e2e %.4g e2e+dl %.4g strong %.4g parity %s kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['with_final_state_download']['value'], d['strong']['value'], d['parity']['ok'], d['roofline']['kernel_ms']))
print(json.dumps(d['e2e']['pcie'])[:1800]); print(json.dumps(d['e2e']['with_final_state_download']['pcie'])[:1800])"
