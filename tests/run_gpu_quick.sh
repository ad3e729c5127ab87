#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest "$@" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -${TAILN:-6} | tee gpurun_out/$name.log; }
run kernels tests/test_gpu_kernels.py -k "${KSEL:-stem or groupnorm}"
run parity tests/test_gpu_parity.py -k "forward or config1"
run iddm tests/test_gpu_iddm.py
echo "=== bench"; timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'conv frac', round(d['roofline']['frac'],3), 'path frac', round(d['whole_path_tensor_frac'],3), d['clocks'])
for k, v in d['forward_breakdown'].items(): print(' ', k, v)
" | tee gpurun_out/bench_quick.log
