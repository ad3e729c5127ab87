#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -${TAILN:-12} | tee gpurun_out/$name.log; }
run kernels tests/test_gpu_kernels.py
run parity tests/test_gpu_parity.py
echo "=== bench"; timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_r3.log
