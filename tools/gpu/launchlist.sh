#!/bin/bash
# ncu launch list (B200_PROFILING.md recipe, durations only) over a window wide enough to hold a whole DDIM step
# of both sub-batch streams; tools/ncu_summarize.py cuts it at the k_ddim_step launches.
mkdir -p gpurun_out
TAG=${TAG:-r01c}
CMD="python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -c 200 gpurun_out/plain.log; wc -l gpurun_out/launches_$TAG.csv
