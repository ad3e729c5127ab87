#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "=== bench 8 gpus"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 2 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_8gpu.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],1), d['clocks'])
"
echo "=== attack loop 8 gpus, resnet50, 256 images"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 tools/attack_bench.py --victim resnet50 --images 256 2>&1 | tail -1 | tee gpurun_out/attack8_resnet50.log
echo "=== attack loop 8 gpus, vit_b_16, 256 images x 8 candidates"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 tools/attack_bench.py --victim vit_b_16 --images 256 --candidates 8 2>&1 | tail -1 | tee gpurun_out/attack8_vit.log
