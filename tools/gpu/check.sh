#!/bin/bash
# Iteration check on the GPU box: the whole GPU suite (all failures, not just the first), smoke(), a short bench.
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --durations=8 -s > gpurun_out/suite_full.log 2>&1
grep -E "passed|failed|error" gpurun_out/suite_full.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/suite_full.log | head -40
echo "=== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
if [ "$1" != "nobench" ]; then
echo "=== bench"; timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_quick.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('roofline', d['roofline'])
for k, v in d['forward_breakdown'].items(): print(' ', k, v)
"
fi
