#!/bin/bash
# Round-2 multi-GPU run on one 8 x B200 box: bench at 8 GPUs, BASELINE.json configs[2] and [3] at their stated scale,
# weak- and strong-scaling sweep points.  Usage: bash tools/gpu/multi8.sh [N]   (N = GPUs, default 8)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
echo "=== bench $N gpus"
timeout 900 $TR --master-port 29544 bench.py --gpus $N --steps 2 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${N}gpu_r02.json
python -c "
import json; d = json.load(open('gpurun_out/bench_${N}gpu_r02.json'))
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],1), d['clocks'])"
echo "=== configs[2]: attack loop, resnet50, 256 images over $N gpus"
timeout 900 $TR --master-port 29545 tools/attack_bench.py --victim resnet50 --images 256 2>&1 | grep '^{' | tail -1 | tee gpurun_out/attack${N}_resnet50_r02.json
echo "=== configs[3]: attack loop, vit_b_16, 256 images x 8 candidates over $N gpus"
timeout 1200 $TR --master-port 29546 tools/attack_bench.py --victim vit_b_16 --images 256 --candidates 8 2>&1 | tail -3 | grep '^{' | tee gpurun_out/attack${N}_vit_r02.json
echo "=== sweep: weak scaling (64 / GPU) and strong scaling (global batch 64) at 256x256"
rm -f gpurun_out/sweep_${N}gpu_r02.jsonl
timeout 900 $TR --master-port 29547 tools/sweep.py --model dm2 --batches 64 --sizes 256 --steps 10,50,100 --streams 2 --out gpurun_out/sweep_${N}gpu_r02.jsonl 2>&1 | grep '^{'
timeout 600 $TR --master-port 29548 tools/sweep.py --model dm2 --batches $((64 / N)) --sizes 256 --steps 50 --out gpurun_out/sweep_${N}gpu_r02.jsonl 2>&1 | grep '^{'
timeout 600 $TR --master-port 29549 tools/sweep.py --model dm2 --batches 64 --sizes 64,128 --steps 50 --streams 2 --out gpurun_out/sweep_${N}gpu_r02.jsonl 2>&1 | grep '^{'
