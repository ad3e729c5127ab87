#!/bin/bash
# A/B on ONE box (clocks differ by +-15 % between boxes): round-1-like settings vs the round-2 defaults
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-torch-gpu --profile-repeats 10"
show='import sys, json
d = json.loads(sys.stdin.read())
print(sys.argv[1], {k: round(d[k], 3) for k in ("value", "ms_per_step")}, "clocks", d["clocks"]["sm_mhz"], "conv", d["roofline"]["achieved_min_median_max"], {k: v["ms"] for k, v in d["forward_breakdown"].items()})'
for v in "$@"; do
  case $v in
    r1like) $B --wide-prenorm 0 --gemm-operands bf16 --graph-scope step 2>&1 | tail -1 | tee gpurun_out/ab_r1like.json | python -c "$show" r1-like;;
    default) $B 2>&1 | tail -1 | tee gpurun_out/ab_default.json | python -c "$show" default;;
    narrow) $B --wide-prenorm 0 2>&1 | tail -1 | tee gpurun_out/ab_narrow.json | python -c "$show" narrow-fp16;;
    bf16wide) $B --gemm-operands bf16 2>&1 | tail -1 | tee gpurun_out/ab_bf16wide.json | python -c "$show" bf16-wide;;
  esac
done
