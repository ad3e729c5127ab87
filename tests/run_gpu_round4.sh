#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -${TAILN:-15} | tee gpurun_out/$name.log; }
run parity_new tests/test_gpu_parity.py -k "ddpm or eta or attack" -s
echo "=== bench 2 gpus"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_2gpu.log
