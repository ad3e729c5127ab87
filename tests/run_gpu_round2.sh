#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -${TAILN:-12} | tee gpurun_out/$name.log; }
run k_conv_simt tests/test_gpu_kernels.py -k "conv and simt"
run parity_tf tests/test_gpu_parity.py -k "teacher" -s
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 1500 python bench.py --steps 1 --warmup 1 2>&1 | tail -5 | tee gpurun_out/bench_first.log
