#!/bin/bash
# Reduced multi-GPU run (N = 2, 4 or 8): the bench line, one weak-scaling and one strong-scaling sweep point at 256x256
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29544 bench.py --gpus $N --steps 2 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${N}gpu_r02b.json
python -c "
import json; d = json.load(open('gpurun_out/bench_${N}gpu_r02b.json'))
print('bench', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],1), d['clocks'])"
rm -f gpurun_out/sweep_${N}gpu_r02b.jsonl
timeout 600 $TR --master-port 29548 tools/sweep.py --model dm2 --batches $((64 / N)) --sizes 256 --steps 50 --out gpurun_out/sweep_${N}gpu_r02b.jsonl 2>&1 | grep '^{'
