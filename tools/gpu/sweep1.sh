#!/bin/bash
# configs[4] on one GPU: batch 1..1024 x 64/128/256 x DDIM 10/50/100 with the host-CPU reference column; IDDM throughput
mkdir -p gpurun_out; rm -f gpurun_out/sweep_1gpu_r02.jsonl
python tools/sweep.py --model dm2 --batches 1,8,64,256,1024 --sizes 64,128,256 --steps 10,50,100 --streams 2 --cpu --budget-s 45 --out gpurun_out/sweep_1gpu_r02.jsonl 2>&1 | grep '^{' | cut -c1-330
python tools/iddm_bench.py 2>&1 | grep '^{' | tee gpurun_out/iddm_bench_r02.jsonl
