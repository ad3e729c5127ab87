#!/bin/bash
# End-of-round verification: the GPU suite, smoke(), the default bench line and the reference arm.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/final_gpu_suite.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
echo "=== bench (defaults)"; timeout 1500 python bench.py 2>&1 | tail -1 | tee gpurun_out/final_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu_baseline', d.get('cpu_baseline')); print('clocks', d['clocks'])
print('whole_path_tensor_frac', d['whole_path_tensor_frac'])
for k, v in d['forward_breakdown'].items(): print(' ', k, v)
"
echo "=== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/final_bench_reference.json | cut -c1-400
